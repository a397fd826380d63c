// emu_kernels.cpp -- TEST INFRASTRUCTURE ONLY: runs the kernels AND launchers of datasketch_b200/csrc/minhash_kernels.cu
// and codec_kernels.cu on host threads (tests/emu/cuda_emu.h) so the CPU test-suite can check their logic against the oracle.
//   g++ -std=c++17 -O1 -pthread -DDSK_EMU -Itests/emu -shared -fPIC tests/emu/emu_kernels.cpp
#include "cuda_emu.h"

thread_local uint3e threadIdx, blockIdx, gridDim, blockDim;
thread_local EmuWarp *emu_warp = nullptr;
thread_local int emu_lane = 0;
thread_local EmuCta *emu_cta = nullptr;

#include "../../datasketch_b200/csrc/minhash_kernels.cu"
#include "../../datasketch_b200/csrc/codec_kernels.cu"

#include <vector>

// a / b: the uint64 permutation parameters; init (optional): [n_docs or 1][k] u32/u64 running signatures.
// rescan = 0 goes through the product launcher (launch_minhash_bulk: P selection, slices, unit size, counters);
// rescan = 1 instantiates the opt-in re-scan variant directly (the launcher reads DSK_RESCAN once per process).
extern "C" int emu_minhash_bulk(const void *tokens, int token_is_u64, const int64_t *offsets, int64_t n_docs,
                                const uint64_t *a, const uint64_t *b, int k, int mode, int rescan,
                                const void *init, int64_t init_stride, int init_is_u64, void *out, int out_is_u64,
                                int docs_per_unit, int grid_x) {
    const int P = k <= 32 ? 1 : k <= 64 ? 2 : k <= 128 ? 4 : 8;
    const int slices = (k + 32 * P - 1) / (32 * P);
    const int kpad = (k + 255) / 256 * 256;
    std::vector<uint32_t> tab((size_t)4 * kpad);
    for (int i = 0; i < kpad; ++i) {  // padding slots repeat real permutations, like dsk_perm_create
        const int s = i % k;
        tab[i] = (uint32_t)a[s]; tab[kpad + i] = (uint32_t)(a[s] >> 32);
        tab[2 * kpad + i] = (uint32_t)b[s]; tab[3 * kpad + i] = (uint32_t)(b[s] >> 32);
    }
    std::vector<unsigned> counters(64, 0u);
    dsk::BulkParams prm{};
    prm.tokens = tokens; prm.offsets = offsets; prm.n_docs = n_docs; prm.n_tokens = offsets[n_docs];
    prm.a_lo = tab.data(); prm.a_hi = tab.data() + kpad; prm.b_lo = tab.data() + 2 * kpad; prm.b_hi = tab.data() + 3 * kpad;
    prm.k = k; prm.init = init; prm.init_stride = init_stride; prm.init_is_u64 = init_is_u64;
    prm.out = out; prm.out_is_u64 = out_is_u64; prm.work_counter = counters.data();
    prm.docs_per_unit = docs_per_unit; prm.n_peers = 0; prm.peer_row_offset = 0;
    if (token_is_u64 && mode != dsk::MODE_EXACT) return -1;
    if (!rescan || P < 4)   // (the variant exists for P = 4 and 8 only)  sm_count = grid_x / 4 CTAs per SM -> the launcher's grid is grid_x (per slice) for small inputs
        return dsk::launch_minhash_bulk(prm, mode, token_is_u64, grid_x, nullptr);
    if (mode != dsk::MODE_TWO_PHASE) return -1;
    dim3 grid((unsigned)grid_x, (unsigned)slices);
    if (P == 4) emu_launch(dsk::minhash_bulk_kernel<4, dsk::MODE_TWO_PHASE, uint32_t, 4, 1>, grid, dsk::kWarps * 32, 0, prm);
    else emu_launch(dsk::minhash_bulk_kernel<8, dsk::MODE_TWO_PHASE, uint32_t, 4, 1>, grid, dsk::kWarps * 32, 0, prm);
    return 0;
}

extern "C" int emu_seg_min(const uint32_t *part, const int64_t *seg, int64_t n_docs, int k, const void *init,
                           int64_t init_stride, int init_is_u64, void *out, int out_is_u64) {
    return dsk::launch_seg_min(part, seg, n_docs, k, init, init_stride, init_is_u64, out, out_is_u64, 1, nullptr);
}
extern "C" int emu_sig_merge_min(const uint32_t *x, const uint32_t *y, int64_t n, uint32_t *out) {
    return dsk::launch_sig_merge_min(x, y, n, out, 1, nullptr);
}
extern "C" int emu_lean_pack(const void *sig, int sig_is_u64, int64_t n, int k, int64_t seed, int big_endian, uint8_t *rec,
                             int sm_count) {
    return dsk::launch_lean_pack(sig, sig_is_u64, n, k, seed, big_endian, rec, sm_count, nullptr);
}
extern "C" int emu_lean_unpack(const uint8_t *rec, int64_t n, int k, int64_t seed, int big_endian, void *sig, int sig_is_u64,
                               int *status, int sm_count) {
    return dsk::launch_lean_unpack(rec, n, k, seed, big_endian, sig, sig_is_u64, status, sm_count, nullptr);
}
extern "C" int emu_band_keys(const uint32_t *sig, int64_t n, int k, int b, int r, uint8_t *out) {
    return dsk::launch_band_keys_be(sig, n, k, b, r, out, 1, nullptr);
}
extern "C" int emu_band_fingerprints(const uint32_t *sig, int64_t n, int k, int b, int r, uint64_t *out) {
    return dsk::launch_band_fingerprints(sig, n, k, b, r, out, 1, nullptr);
}
extern "C" int emu_bbit_pack(const uint32_t *sig, int64_t n, int k, int b, int slot, uint64_t *out) {
    return dsk::launch_bbit_pack(sig, n, k, b, slot, out, 1, nullptr);
}
extern "C" int emu_bbit_unpack(const uint64_t *blocks, int64_t n, int k, int slot, uint32_t *sig) {
    return dsk::launch_bbit_unpack(blocks, n, k, slot, sig, 1, nullptr);
}

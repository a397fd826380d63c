// emu_minhash.cpp -- TEST INFRASTRUCTURE ONLY: runs datasketch_b200/csrc/minhash_kernels.cu's kernel template on
// host threads (tests/emu/cuda_emu.h) so the CPU test-suite can check its logic against the oracle.
//   g++ -std=c++17 -O1 -pthread -DDSK_EMU -Itests/emu -x c++ -shared -fPIC tests/emu/emu_minhash.cpp
#include "cuda_emu.h"

thread_local uint3e threadIdx, blockIdx, gridDim, blockDim;
thread_local EmuWarp *emu_warp = nullptr;
thread_local int emu_lane = 0;

#include "../../datasketch_b200/csrc/minhash_kernels.cu"

#include <thread>
#include <vector>

namespace {

using KernelFn = void (*)(const dsk::BulkParams);

struct LaneArgs {
    KernelFn fn;
    const dsk::BulkParams *prm;
    EmuWarp *warp;
    unsigned tid, bx, by, gx, gy;
};

void lane_main(LaneArgs a) {
    threadIdx = {a.tid, 0, 0};
    blockIdx = {a.bx, a.by, 0};
    gridDim = {a.gx, a.gy, 1};
    blockDim = {dsk::kWarps * 32, 1, 1};
    emu_warp = a.warp;
    emu_lane = (int)(a.tid & 31);
    a.fn(*a.prm);
}

// one CTA at a time (the kernel's __shared__ arrays are static here), kWarps * 32 host threads per CTA
void run_grid(KernelFn fn, const dsk::BulkParams &prm, unsigned gx, unsigned gy) {
    for (unsigned by = 0; by < gy; ++by)
        for (unsigned bx = 0; bx < gx; ++bx) {
            std::vector<EmuWarp> warps(dsk::kWarps);
            for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < dsk::kWarps * 32; ++t)
                th.emplace_back(lane_main, LaneArgs{fn, &prm, &warps[t >> 5], t, bx, by, gx, gy});
            for (auto &x : th) x.join();
            for (auto &w : warps) pthread_barrier_destroy(&w.bar);
        }
}

template <int MODE, typename TokT, int RESCAN>
KernelFn pick(int k) {
    if (k <= 32) return dsk::minhash_bulk_kernel<1, MODE, TokT, 4, 0>;
    if (k <= 64) return dsk::minhash_bulk_kernel<2, MODE, TokT, 4, 0>;
    if (k <= 128) return dsk::minhash_bulk_kernel<4, MODE, TokT, 4, RESCAN>;
    return dsk::minhash_bulk_kernel<8, MODE, TokT, 4, RESCAN>;
}

}  // namespace

// a / b: the uint64 permutation parameters; init (optional): [n_docs or 1][k] u32/u64 running signatures.
// Returns 0, or -1 for an unsupported combination (fast modes need u32 tokens).
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
    std::vector<unsigned> counters(slices, 0u);
    dsk::BulkParams prm{};
    prm.tokens = tokens; prm.offsets = offsets; prm.n_docs = n_docs; prm.n_tokens = offsets[n_docs];
    prm.a_lo = tab.data(); prm.a_hi = tab.data() + kpad; prm.b_lo = tab.data() + 2 * kpad; prm.b_hi = tab.data() + 3 * kpad;
    prm.k = k; prm.init = init; prm.init_stride = init_stride; prm.init_is_u64 = init_is_u64;
    prm.out = out; prm.out_is_u64 = out_is_u64; prm.work_counter = counters.data();
    prm.docs_per_unit = docs_per_unit; prm.n_peers = 0; prm.peer_row_offset = 0;
    KernelFn fn = nullptr;
    if (token_is_u64) {
        if (mode != dsk::MODE_EXACT) return -1;
        fn = pick<dsk::MODE_EXACT, uint64_t, 0>(k);
    } else if (mode == dsk::MODE_TWO_PHASE) {
        fn = rescan ? pick<dsk::MODE_TWO_PHASE, uint32_t, 1>(k) : pick<dsk::MODE_TWO_PHASE, uint32_t, 0>(k);
    } else if (mode == dsk::MODE_DIRECT) {
        fn = pick<dsk::MODE_DIRECT, uint32_t, 0>(k);
    } else {
        fn = pick<dsk::MODE_EXACT, uint32_t, 0>(k);
    }
    run_grid(fn, prm, (unsigned)grid_x, (unsigned)slices);
    return 0;
}

// emu_capi.cpp -- TEST INFRASTRUCTURE ONLY: the whole library (every kernel file + the C-ABI layer dsk_api.cu) compiled
// with g++ on the emulation shim, exporting the same C-ABI as libdsk_b200.so.  "Device" memory is host memory, the one
// emulated device has two SMs, every stream operation is synchronous.  The CPU test-suite loads it with ctypes to drive
// the C-ABI itself -- argument validation and error codes, the permutation analysis, the host pipeline of
// dsk_minhash_bulk_host (slicing, long-document splitting, running-state merge) -- and, swapped in for the real library
// inside a test fixture, the Python layer's host-buffer paths.  The product never loads this file's output.
//   g++ -std=c++17 -O1 -pthread -ffp-contract=off -DDSK_EMU -Itests/emu -shared -fPIC tests/emu/emu_capi.cpp
#include "cuda_emu.h"

thread_local uint3e threadIdx, blockIdx, gridDim, blockDim;
thread_local EmuWarp *emu_warp = nullptr;
thread_local int emu_lane = 0;
thread_local EmuCta *emu_cta = nullptr;

#include "../../datasketch_b200/csrc/minhash_kernels.cu"
#include "../../datasketch_b200/csrc/signature_kernel.cu"
#include "../../datasketch_b200/csrc/codec_kernels.cu"
#include "../../datasketch_b200/csrc/lsh_kernels.cu"
#include "../../datasketch_b200/csrc/jaccard_kernels.cu"
#include "../../datasketch_b200/csrc/sha1_kernels.cu"
#include "../../datasketch_b200/csrc/hash_kernels.cu"
#include "../../datasketch_b200/csrc/wmh_kernels.cu"
#include "../../datasketch_b200/csrc/bloom_kernels.cu"
#include "../../datasketch_b200/csrc/dsk_api.cu"

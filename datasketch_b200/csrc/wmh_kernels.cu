// wmh_kernels.cu -- Weighted MinHash (Ioffe ICWS) sampling on sm_100a.
//
// Replaces the per-sample Python loop of WeightedMinHashGenerator.minhash
// (datasketch/weighted_minhash.py:147-158) for batches of vectors:
//   vlog = log(v)                      (zeros -> NaN and are skipped, :148-152)
//   t    = floor(vlog / r_i + beta_i)
//   ln_y = (t - beta_i) * r_i
//   ln_a = (ln_c_i - ln_y) - r_i
//   k    = first index of the NaN-skipping minimum of ln_a;  out[i] = (k, int(t[k]))
// MANY = true evaluates the reference's experimental minhash_many instead (weighted_minhash.py:221-224):
//   ln_y = ((t - beta_i) + 1) * r_i;  ln_a = ln_c_i - ln_y      (same algebra, different float32 rounding)
// and an all-zero row is reported in status[] (the caller returns None for it, :228 / :241-245).
// Every float32 operation is a separate IEEE round-to-nearest op in exactly that order
// (__fdiv_rn/__fadd_rn/__fsub_rn/__fmul_rn: no FMA contraction, no fast-math), so all steps after
// the logarithm are bit-identical to numpy; the logarithm is log() in double rounded to float32
// (numpy's SIMD float32 log is not correctly rounded, so this one step agrees to <= 1 ulp; see
// DESIGN.md and tests/test_wmh_gpu.py for the measured effect on (k, t): none on the fixtures).
//
// Layout: parameters are stored transposed, [dim][ss_pad], so the threads of a CTA (one per
// sample) read consecutive addresses; they are 3 * ss * dim * 4 B (6.3 MB at dim 4096, ss 128) and
// stay L2-resident.  A CTA register-blocks kVec input vectors against each parameter load.
#include "dsk_common.cuh"

namespace dsk {

constexpr int kWmhThreads = 128;  // samples per CTA (one thread each)
constexpr int kVec = 8;           // vectors per CTA pass
constexpr int kTileD = 256;       // dims staged per tile

struct WmhParams {
    const float *rs_t, *lncs_t, *betas_t;  // [dim][ss_pad]
    int ss, ss_pad, dim;
    const float *v;                        // [n][dim]
    int64_t n;
    int64_t *out;                          // [n][ss][2]
    int32_t *status;                       // [n]: 1 = all-zero input (weighted_minhash.py:149-150)
    int input_log;                         // v already holds ln(weight) as float32 (NaN where the weight is zero)
};

// ln_a for one (sample, dim): each float32 operation rounded separately, in the reference's order
template <bool MANY>
__device__ __forceinline__ float wmh_ln_a(float t, float be, float r, float lc) {
    if constexpr (MANY) {
        const float ln_y = __fmul_rn(__fadd_rn(__fsub_rn(t, be), 1.0f), r);
        return __fsub_rn(lc, ln_y);
    } else {
        const float ln_y = __fmul_rn(__fsub_rn(t, be), r);
        return __fsub_rn(__fsub_rn(lc, ln_y), r);
    }
}

template <bool MANY>
__global__ void __launch_bounds__(kWmhThreads) wmh_kernel(const WmhParams p) {
    __shared__ __align__(16) float s_vlog[kTileD][kVec];  // [d][u]: 2 x LDS.128 per dim
    const int tid = threadIdx.x;
    const int s = blockIdx.y * kWmhThreads + tid;         // this thread's sample
    const bool s_ok = s < p.ss;
    const int64_t groups = (p.n + kVec - 1) / kVec;

    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t u0 = g * kVec;
        float best[kVec], bt[kVec];
        int bk[kVec];
#pragma unroll
        for (int u = 0; u < kVec; ++u) { best[u] = __int_as_float(0x7f800000); bk[u] = -1; bt[u] = 0.f; }

        for (int d0 = 0; d0 < p.dim; d0 += kTileD) {
            const int nd = min(kTileD, p.dim - d0);
            __syncthreads();  // previous tile fully consumed
            for (int idx = tid; idx < kVec * kTileD; idx += kWmhThreads) {
                const int u = idx / kTileD, dd = idx - u * kTileD;
                float lg = __int_as_float(0x7fc00000);  // NaN: padding and zero weights are skipped
                if (dd < nd && u0 + u < p.n) {
                    const float x = __ldg(p.v + (u0 + u) * (int64_t)p.dim + d0 + dd);
                    // host-computed logarithms (numpy's own float32 log: bit-identical to the reference by
                    // construction) or, opt-in, the device's correctly rounded one (<= 1 ulp from numpy's)
                    if (p.input_log) lg = x;
                    else if (x != 0.f) lg = (float)log((double)x);
                }
                s_vlog[dd][u] = lg;
            }
            __syncthreads();
            if (s_ok) {
                const float *rp = p.rs_t + (int64_t)d0 * p.ss_pad + s;
                const float *cp = p.lncs_t + (int64_t)d0 * p.ss_pad + s;
                const float *bp = p.betas_t + (int64_t)d0 * p.ss_pad + s;
                // kUnrollD dims per trip: all 3*kUnrollD parameter loads (L2-resident, ~500 cycles) are issued
                // before the first evaluation, so their latency overlaps kUnrollD*kVec evaluations
                constexpr int kUnrollD = 4;
                int dd = 0;
                for (; dd + kUnrollD <= nd; dd += kUnrollD) {
                    float r[kUnrollD], lc[kUnrollD], be[kUnrollD];
#pragma unroll
                    for (int q = 0; q < kUnrollD; ++q) {
                        r[q] = __ldg(rp + (int64_t)(dd + q) * p.ss_pad);
                        lc[q] = __ldg(cp + (int64_t)(dd + q) * p.ss_pad);
                        be[q] = __ldg(bp + (int64_t)(dd + q) * p.ss_pad);
                    }
#pragma unroll
                    for (int q = 0; q < kUnrollD; ++q) {
                        const float4 x0 = *reinterpret_cast<const float4 *>(&s_vlog[dd + q][0]);
                        const float4 x1 = *reinterpret_cast<const float4 *>(&s_vlog[dd + q][4]);
                        const float xs[kVec] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                        for (int u = 0; u < kVec; ++u) {
                            const float t = floorf(__fadd_rn(__fdiv_rn(xs[u], r[q]), be[q]));
                            const float ln_a = wmh_ln_a<MANY>(t, be[q], r[q], lc[q]);
                            if (ln_a < best[u] || (bk[u] < 0 && ln_a == ln_a)) {  // NaN never wins; first index kept on ties
                                best[u] = ln_a; bk[u] = d0 + dd + q; bt[u] = t;
                            }
                        }
                    }
                }
                for (; dd < nd; ++dd) {
                    const float r = __ldg(rp + (int64_t)dd * p.ss_pad);
                    const float lc = __ldg(cp + (int64_t)dd * p.ss_pad);
                    const float be = __ldg(bp + (int64_t)dd * p.ss_pad);
                    const float4 x0 = *reinterpret_cast<const float4 *>(&s_vlog[dd][0]);
                    const float4 x1 = *reinterpret_cast<const float4 *>(&s_vlog[dd][4]);
                    const float xs[kVec] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int u = 0; u < kVec; ++u) {
                        const float t = floorf(__fadd_rn(__fdiv_rn(xs[u], r), be));
                        const float ln_a = wmh_ln_a<MANY>(t, be, r, lc);
                        if (ln_a < best[u] || (bk[u] < 0 && ln_a == ln_a)) {
                            best[u] = ln_a; bk[u] = d0 + dd; bt[u] = t;
                        }
                    }
                }
            }
        }
        if (s_ok) {
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                if (u0 + u < p.n) {
                    int64_t *o = p.out + ((u0 + u) * (int64_t)p.ss + s) * 2;
                    o[0] = bk[u] < 0 ? 0 : bk[u];
                    o[1] = bk[u] < 0 ? 0 : (int64_t)bt[u];
                    if (s == 0) p.status[u0 + u] = bk[u] < 0 ? 1 : 0;
                }
            }
        }
    }
}

// [ss][dim] -> [dim][ss_pad] (one-time, at generator creation)
__global__ void wmh_transpose_kernel(const float *src, int ss, int dim, int ss_pad, float *dst) {
    const int64_t total = (int64_t)dim * ss_pad, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int d = (int)(e / ss_pad), s = (int)(e - (int64_t)d * ss_pad);
        dst[e] = s < ss ? src[(int64_t)s * dim + d] : 1.0f;
    }
}

cudaError_t launch_wmh_transpose(const float *src, int ss, int dim, int ss_pad, float *dst, cudaStream_t s) {
    DSK_LAUNCH(wmh_transpose_kernel, 256, 256, 0, s, src, ss, dim, ss_pad, dst);
    return cudaGetLastError();
}

cudaError_t launch_wmh(const float *rs_t, const float *lncs_t, const float *betas_t, int ss, int ss_pad, int dim,
                       const float *v, int64_t n, int64_t *out, int32_t *status, int many, int input_log, int sm_count,
                       cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    WmhParams p;
    p.input_log = input_log;
    p.rs_t = rs_t; p.lncs_t = lncs_t; p.betas_t = betas_t;
    p.ss = ss; p.ss_pad = ss_pad; p.dim = dim; p.v = v; p.n = n; p.out = out; p.status = status;
    const int64_t groups = (n + kVec - 1) / kVec;
    const int slices = (ss + kWmhThreads - 1) / kWmhThreads;
    int64_t gx = groups;
    const int64_t cap = (int64_t)sm_count * 8;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)slices);
    if (many) DSK_LAUNCH((wmh_kernel<true>), grid, kWmhThreads, 0, s, p);
    else DSK_LAUNCH((wmh_kernel<false>), grid, kWmhThreads, 0, s, p);
    return cudaGetLastError();
}

}  // namespace dsk

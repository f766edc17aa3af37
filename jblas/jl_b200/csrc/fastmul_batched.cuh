// fastmul_batched.cuh -- fastmul!-class small products, batched: D_b = A_b * X_b for b = 0..batch-1.
//
// the only regime where the reference publishes a number is a single tiny product
// (jmul! 16x32x14 Float64 in 128 ns, test/runtests.jl:110-122).  One such product cannot amortise a GPU launch, so the
// B200 form of `fastmul!` (src/kernels.jl:202-208: small static matrices, fully unrolled, row remainder masked) is a
// BATCH of independent small products in one launch.  The work is HBM-bound (16x32x14 f64: 9.5 KB moved for 14 kflop,
// 1.5 flop/B), so the design goal is to keep HBM saturated, not the FMA pipes:
//   * persistent CTAs, grid-stride over groups of G products, two shared-memory stages filled with cp.async
//     (element granularity: any alignment, any stride), so the loads of group g+1 fly while group g is multiplied;
//   * each thread owns a 2x2 block of one product's D (two rows x two columns), which makes shared-memory traffic
//     (3 loads per 4 FMAs) cheaper than the HBM time; X columns are stored with an odd pitch so the column lanes of a
//     warp fall into distinct banks;
//   * every element is the reference chain: -0.0 start, fma over ascending k -- bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace jb {

template <typename T>
__global__ void __launch_bounds__(256)
fastmul_batched_kernel(T* __restrict__ D, const T* __restrict__ A, const T* __restrict__ X, int M, int N, int P, int64_t batch,
                       int64_t strideD, int64_t strideA, int64_t strideX, int G, int xpitch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int a_elems = M * N, x_elems = xpitch * P;
    const int slot = (a_elems + x_elems + 1) & ~1;  // one product: A (M x N dense) then X (N x P, column pitch xpitch); even
                                                    // element count keeps every A copy 2-element aligned for paired loads
    const int stage_elems = G * slot;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int mb = (M + 1) >> 1, pb = (P + 1) >> 1, bpp = mb * pb;  // 2x2 output blocks per product
    const int64_t ngroups = (batch + G - 1) / G;

    // No integer divisions on the copy path: products are walked by warps, elements by lanes.
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const bool tiny = (a_elems < 256) || (N < 16);
    auto load_group = [&](int64_t grp, int stage) {
        T* st = smem + (size_t)stage * stage_elems;
        const int64_t p0 = grp * G;
        const int g_here = (int)min((int64_t)G, batch - p0);
        if (tiny) {  // very small matrices: flat element loops keep every lane busy (the divisions are cheap next to HBM time)
            for (int e = tid; e < g_here * a_elems; e += nthr) {
                const int g = e / a_elems, i = e - g * a_elems;
                cp_async_elem<T>(smem_u32(st + g * slot + i), A + (p0 + g) * strideA + i, true);
            }
            for (int e = tid; e < g_here * N * P; e += nthr) {
                const int g = e / (N * P), i = e - g * (N * P);
                const int c = i / N, k = i - c * N;
                cp_async_elem<T>(smem_u32(st + g * slot + a_elems + c * xpitch + k), X + (p0 + g) * strideX + i, true);
            }
            return;
        }
        for (int g = 0; g < g_here; ++g) {
            const T* ga = A + (p0 + g) * strideA;
            T* sa = st + g * slot;
            for (int i = tid; i < a_elems; i += nthr) cp_async_elem<T>(smem_u32(sa + i), ga + i, true);
            const T* gx = X + (p0 + g) * strideX;
            T* sx = sa + a_elems;
            for (int c = warp; c < P; c += nwarps)
                for (int k = lane; k < N; k += 32) cp_async_elem<T>(smem_u32(sx + c * xpitch + k), gx + c * N + k, true);
        }
    };

    int64_t grp = blockIdx.x;
    if (grp < ngroups) load_group(grp, 0);
    cp_async_commit();
    int stage = 0;
    for (; grp < ngroups; grp += gridDim.x, stage ^= 1) {
        const int64_t nxt = grp + gridDim.x;
        if (nxt < ngroups) load_group(nxt, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const T* st = smem + (size_t)stage * stage_elems;
        const int64_t p0 = grp * G;
        const int g_here = (int)min((int64_t)G, batch - p0);
        for (int w = tid; w < g_here * bpp; w += nthr) {
            const int g = w / bpp, blk = w - g * bpp;
            const int rb = blk % mb, cb = blk / mb;
            const int r0 = 2 * rb, c0 = 2 * cb;
            const int r1 = min(r0 + 1, M - 1), c1 = min(c0 + 1, P - 1);  // clamped duplicates on odd edges (not stored)
            const T* sa = st + g * slot;
            const T* sx = sa + a_elems;
            T d00 = T(-0.0), d10 = T(-0.0), d01 = T(-0.0), d11 = T(-0.0);  // fma(a, b, -0.0) == a*b: the plain first product
            if ((M & 1) == 0) {  // rows (r0, r0+1) are one aligned pair: a single paired shared-memory load
                struct alignas(2 * sizeof(T)) Pair { T lo, hi; };
#pragma unroll 4
                for (int k = 0; k < N; ++k) {
                    const Pair a = *reinterpret_cast<const Pair*>(sa + k * M + r0);
                    const T x0 = sx[c0 * xpitch + k], x1 = sx[c1 * xpitch + k];
                    d00 = fma_t(a.lo, x0, d00);
                    d10 = fma_t(a.hi, x0, d10);
                    d01 = fma_t(a.lo, x1, d01);
                    d11 = fma_t(a.hi, x1, d11);
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < N; ++k) {
                    const T a0 = sa[k * M + r0], a1 = sa[k * M + r1];
                    const T x0 = sx[c0 * xpitch + k], x1 = sx[c1 * xpitch + k];
                    d00 = fma_t(a0, x0, d00);
                    d10 = fma_t(a1, x0, d10);
                    d01 = fma_t(a0, x1, d01);
                    d11 = fma_t(a1, x1, d11);
                }
            }
            T* dp = D + (p0 + g) * strideD;
            dp[(size_t)c0 * M + r0] = d00;
            if (r0 + 1 < M) dp[(size_t)c0 * M + r0 + 1] = d10;
            if (c0 + 1 < P) {
                dp[(size_t)(c0 + 1) * M + r0] = d01;
                if (r0 + 1 < M) dp[(size_t)(c0 + 1) * M + r0 + 1] = d11;
            }
        }
        __syncthreads();  // the stage is free for the load issued in the next iteration
    }
    cp_async_wait<0>();
}

}  // namespace jb

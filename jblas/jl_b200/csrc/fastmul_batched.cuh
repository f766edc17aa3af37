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
#include "gemm_dmma.cuh"
#include "gemm_dmma_tma.cuh"

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

// ---- Float64, products up to 32 x N x 32: tensor pipe, fragments straight from HBM ---------------------------------------
// The SIMT kernel above spends ~450 shared-memory wavefronts per 16x32x14 product (3 loads per 4 FMAs) against ~60 to stage it:
// shared-memory bandwidth, not HBM, bounds it (53 % of the HBM peak).  mma.sync.m8n8k4.f64 needs one 8-byte element per lane
// per 8x4 / 4x8 fragment, and in column-major storage those fragments are sector-aligned already:
//   A fragment (mi, k4): lane (g, t) reads A[mi*8 + g, k4*4 + t]  -> 4 runs of 64 contiguous bytes (g), one per column t
//   X fragment (k4, ni): lane (g, t) reads X[k4*4 + t, ni*8 + g]  -> 8 runs of 32 contiguous bytes (t), one per column g
// so ONE WARP owns one product (or one 32 x 32 block of it) and loads every fragment of a 32-deep k chunk with independent 8-byte global loads (for
// 16x32x14 that is all 32 loads of the product in flight at once, 8 KB per warp), runs the DMMAs, and stores the accumulator
// fragments (64-byte runs).  No shared memory, no barriers; latency is covered by the other resident warps.
// DMMA is bit-identical to the sequential chain on B200 (tests/test_gemm_gpu.py), so results equal the SIMT kernel's.
// Larger products (M or P > 32) are cut into 32 x 32 blocks of D, one warp each; the warps of one product run side by side,
// so the A row-block / X column-block a sibling already fetched comes from L1/L2.  Tiny products (M, P <= 8, N <= 16) move
// too few bytes per warp to cover the HBM latency, so a warp takes U = 4 consecutive products per iteration.
template <int MI, int NI, int KC, int U, bool ONE_BLOCK>
__global__ void __launch_bounds__(128, (MI * NI >= 12 ? 3 : (U > 1 ? 6 : 1)))  // 3 CTAs/SM = 170 registers: the 4x4 grid otherwise takes 186
                                                                               // and drops to 2; the tiny-product variant is latency-bound:
                                                                               // 6 CTAs/SM (ncu: 152 registers, 17 % warps active without)
fastmul_batched_dmma_kernel(double* __restrict__ D, const double* __restrict__ A, const double* __restrict__ X, int M, int N, int P,
                            int64_t batch, int64_t strideD, int64_t strideA, int64_t strideX, int mblocks, int pblocks)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nblk = ONE_BLOCK ? 1 : mblocks * pblocks;  // ONE_BLOCK (M, P <= 32): no 64-bit division per product
    const int64_t items = (batch * nblk + U - 1) / U;  // a work item = U consecutive (product, block) pairs
    for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < items; item += nwarps) {
        const double* __restrict__ a[U];
        const double* __restrict__ x[U];
        int64_t prod[U];
        int r0[U], c0[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t w = item * U + u;
            live[u] = w < batch * nblk;
            if constexpr (ONE_BLOCK) {
                prod[u] = live[u] ? w : 0;
                r0[u] = c0[u] = 0;
            } else {
                prod[u] = live[u] ? w / nblk : 0;
                const int blk = live[u] ? (int)(w - prod[u] * nblk) : 0;
                r0[u] = (blk % mblocks) * 32;
                c0[u] = (blk / mblocks) * 32;
            }
            a[u] = A + prod[u] * strideA;
            x[u] = X + prod[u] * strideX;
        }
        double acc[U][MI][NI][2];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) acc[u][mi][ni][0] = acc[u][mi][ni][1] = -0.0;  // fma(a, b, -0.0) == a*b
        for (int k0 = 0; k0 < N; k0 += 4 * KC) {
            double af[U][KC][MI], bf[U][KC][NI];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    const int k = k0 + 4 * kk + t;
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        const int r = r0[u] + mi * 8 + g;
                        af[u][kk][mi] = (live[u] && k < N && r < M) ? __ldg(a[u] + (size_t)k * M + r) : 0.0;
                    }
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) {
                        const int c = c0[u] + ni * 8 + g;
                        // k >= N: (+0) * (-0.0) = -0.0 and c + (-0.0) == c for every c, so a padded step changes nothing
                        bf[u][kk][ni] = (k < N) ? ((live[u] && c < P) ? __ldg(x[u] + (size_t)c * N + k) : 0.0) : -0.0;
                    }
                }
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                if (k0 + 4 * kk < N) {  // warp-uniform
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                            for (int ni = 0; ni < NI; ++ni)
                                dmma884(acc[u][mi][ni][0], acc[u][mi][ni][1], af[u][kk][mi], bf[u][kk][ni]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            double* __restrict__ d = D + prod[u] * strideD;
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = c0[u] + ni * 8 + 2 * t + c;
                    if (col >= P) continue;
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        const int r = r0[u] + mi * 8 + g;
                        if (r < M) d[(size_t)col * M + r] = acc[u][mi][ni][c];
                    }
                }
        }
    }
}

// ---- Float32, products up to 32 x N x 16: FFMA2 with warp-private staging -----------------------------------------------------
// There is no exact FP32 tensor path, and the generic kernel above is bound by its shared-memory traffic and CTA barriers
// (16x32x14: 2.0 TB/s, 31 % of HBM).  Here LPP = 8 (M <= 16) or 16 (M <= 32) lanes own one product, so a warp multiplies 4 or 2
// products side by side with NO CTA-wide synchronisation: each warp double-buffers its own products in shared memory with TMA bulk
// copies (cp.async.bulk, one per matrix: A and X exactly as they lie in HBM, no transposition; one lane issues them, an mbarrier
// per stage counts the bytes) and synchronises with __syncwarp only.  (A first version staged with per-lane 16-byte cp.async:
// 4.1 TB/s -- the copies, not HBM, were the limit.)
//   lane l of a product holds rows (2l, 2l+1) of ALL columns: PC packed accumulators (fma.rn.f32x2: one instruction = both rows);
//   per 4 k:  4 LDS.64 (its two rows of A[:, k..k+3]) + PC LDS.128 (X[k..k+3, c], the same address for the whole product =
//             broadcast) feed 4*PC FFMA2 -- 18 loads per 56 FFMA2 for P = 14, against 3 loads per 4 FMAs above;
//   product slots are skewed by 16 floats so the two products of a half-warp sit in different banks.
// Every element is the reference chain (ascending k, start -0.0): bit-identical to the oracle.
// Preconditions (checked by the host): M even, N % 4 == 0, strides % 4 == 0, 16-byte aligned bases.
template <int LPP, int PC>
__global__ void __launch_bounds__(64)
fastmul_batched_f32_warp_kernel(float* __restrict__ D, const float* __restrict__ A, const float* __restrict__ X, int M, int N, int P,
                                int64_t batch, int64_t strideD, int64_t strideA, int64_t strideX, int slot_floats)
{
    constexpr int PPW = 32 / LPP;  // products per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
    const int q = lane / LPP, l = lane % LPP;
    float* wbase = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 2 * PPW * slot_floats;  // [stage][slot][A | X]
    const int a_floats = M * N, x_floats = N * P;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_cta;
    const int64_t items = (batch + PPW - 1) / PPW;
    // one mbarrier per (warp, stage), after the data of all warps
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)warps_per_cta * 2 * PPW * slot_floats) + warp * 2;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();

    auto load = [&](int64_t item, int stage) {  // lane 0: arm the stage's barrier with the byte count, then one bulk copy per matrix
        if (lane != 0) return;
        float* st = wbase + (size_t)stage * PPW * slot_floats;
        const int64_t b0 = item * PPW;
        const int live = (int)min((int64_t)PPW, batch - b0);
        mbar_expect_tx(&bars[stage], (uint32_t)(live * (a_floats + x_floats) * 4));
        for (int pq = 0; pq < live; ++pq) {
            float* sp = st + pq * slot_floats;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(sp)),
                         "l"(A + (b0 + pq) * strideA), "r"(a_floats * 4), "r"(smem_u32(&bars[stage]))
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             smem_u32(sp + a_floats)),
                         "l"(X + (b0 + pq) * strideX), "r"(x_floats * 4), "r"(smem_u32(&bars[stage]))
                         : "memory");
        }
    };

    int64_t item = (int64_t)blockIdx.x * warps_per_cta + warp;
    if (item < items) load(item, 0);
    int stage = 0;
    uint32_t parity = 0;  // bit s = phase parity of stage s
    for (; item < items; item += nwarps, stage ^= 1) {
        const int64_t nxt = item + nwarps;
        if (nxt < items) load(nxt, stage ^ 1);
        mbar_wait(&bars[stage], (parity >> stage) & 1u);
        parity ^= 1u << stage;
        const float* sa = wbase + ((size_t)stage * PPW + q) * slot_floats;
        const float* sx = sa + a_floats;
        const int r0 = min(2 * l, M - 2);  // lanes beyond the last row pair recompute it (never stored): no out-of-slot reads
        uint64_t acc[PC];
#pragma unroll
        for (int c = 0; c < PC; ++c) acc[c] = 0x8000000080000000ull;  // (-0.0f, -0.0f): fma(a, b, -0) == a*b
#pragma unroll 2
        for (int k = 0; k < N; k += 4) {
            uint64_t a[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const uint64_t*>(sa + (k + kk) * M + r0);
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                const float4 xv = *reinterpret_cast<const float4*>(sx + c * N + k);
                uint64_t b0, b1, b2, b3;
                asm("mov.b64 %0, {%1, %1};\n" : "=l"(b0) : "f"(xv.x));
                asm("mov.b64 %0, {%1, %1};\n" : "=l"(b1) : "f"(xv.y));
                asm("mov.b64 %0, {%1, %1};\n" : "=l"(b2) : "f"(xv.z));
                asm("mov.b64 %0, {%1, %1};\n" : "=l"(b3) : "f"(xv.w));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[c]) : "l"(a[0]), "l"(b0));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[c]) : "l"(a[1]), "l"(b1));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[c]) : "l"(a[2]), "l"(b2));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[c]) : "l"(a[3]), "l"(b3));
            }
        }
        const int64_t b = item * PPW + q;
        if (b < batch && 2 * l < M) {  // M is even: rows r0, r0+1 exist together
            float* dp = D + b * strideD + r0;
#pragma unroll
            for (int c = 0; c < PC; ++c)
                if (c < P) *reinterpret_cast<uint64_t*>(dp + (size_t)c * M) = acc[c];
        }
        __syncwarp();  // every lane is done with this stage before lane 0 lets the next bulk copies overwrite it
    }
}

}  // namespace jb

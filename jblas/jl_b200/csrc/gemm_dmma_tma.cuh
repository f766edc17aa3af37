// gemm_dmma_tma.cuh -- FP64 tensor-core GEMM, warp-specialised TMA producer / DMMA consumers, persistent CTAs.
//
// This is the main FP64 kernel.  It replaces, in one design, the three things north_star names:
//   * "packing buffers / cache blocking" (the reference has only software prefetch, src/memory_management.jl:181-277,
//     and an uncalled planner, :78-140)  ->  TMA (cp.async.bulk.tensor) stages A and X tiles into a multi-stage
//     shared-memory ring; no thread spends issue slots on copies;
//   * the register-tile SIMD micro-kernel (src/gemm.jl:149-170, src/kernels.jl:212-275)  ->  mma.sync.m8n8k4.f64
//     (SASS DMMA.8x8x4), 8x4 tiles per warp, accumulators in registers (FP64 has no tcgen05/TMEM kind);
//   * the two tile loops of jmul! (src/gemm.jl:313)  ->  a persistent grid (one CTA per SM) walking a rasterised
//     tile list, so the producer runs ahead across tile boundaries and the pipeline never drains.
//
// Roles (384 threads = 3 warpgroups): warps 0..7 consume (2 x 4 warp grid, 64 x 32 per warp), warp 8 lane 0 produces;
// setmaxnreg moves the producer warpgroup's registers to the consumers.
// Synchronisation is mbarrier-only in the main loop: full[s] (TMA complete_tx) and empty[s] (one arrive per
// consumer warp); no CTA-wide barrier after start-up.
//
// Shared-memory layout = what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index XOR row&7):
//   A sub-tile: 8 boxes (16 m x 16 k), box = 16 rows (k) of 128 B (16 doubles of m)
//   X sub-tile: 1 box  (16 k x 128 n),     128 rows (n) of 128 B (16 doubles of k)
// DMMA fragment loads are 8-byte LDS; a half-warp (16 lanes) must hit 16 distinct 8-byte slots of the 128-byte
// bank window.  With the hardware swizzle that holds if the LOGICAL rows/columns of an MMA tile are permuted:
//   A: MMA row g of sub-tile `sub` of a 16-row box is physical row  pi(g) = (g&1) + 8*((g>>1)&1) + 2*(g>>2) + 4*sub
//   X: MMA column g of an 8-column group is physical column          sg(g) = 2*(g&3) + (g>>2)
// (derivation in DESIGN.md "bank-conflict-free fragments"); ncu confirms 0 shared-memory bank conflicts.
// The accumulator fragment is un-permuted on the way out with the same two maps.
//
// Numerics: ascending k, 4 at a time, starting from -0.0 (or the old D when ACC); contract = the reference
// tolerance 2*K*eps*(|A||X|); measured against the oracle in tests/test_gemm_gpu.py.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_dmma.cuh"

namespace jb {

// ---- mbarrier / TMA primitives (inline PTX) -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

template <int KSUB_, int STAGES_>
struct DmmaTmaCfg {
    static constexpr int BM = 128, BN = 128, KSUB = KSUB_, BK = 16 * KSUB_, STAGES = STAGES_;
    static constexpr int CONSUMER_WARPS = 8;
    // 3 warpgroups: two consumer warpgroups + one producer warpgroup (only its first lane works).  A 9th warp alone
    // would put 3 warps on one SM sub-partition and cap EVERY thread at 168 registers; with whole warpgroups the
    // producer donates its registers to the consumers via setmaxnreg (40 vs 232 per thread: 32*(232+232+40) <= 16384).
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr int A_SUB_BYTES = BM * 16 * 8;  // 8 boxes x 2 KiB
    static constexpr int B_SUB_BYTES = BN * 16 * 8;  // 1 box of 16 KiB
    static constexpr int SUB_BYTES = A_SUB_BYTES + B_SUB_BYTES;
    static constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + 1024;  // +align slack
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX,
                     double* __restrict__ D, int M, int N, int K, int64_t ldd, int tiles_m, int tiles_n, int group_m)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, KSUB = Cfg::KSUB, STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    // the 128B swizzle is a function of address bits 4..9: tile bases must be 1024-byte aligned
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int num_tiles = tiles_m * tiles_n;
    const int KT = (K + BK - 1) / BK;

    if (warp >= Cfg::CONSUMER_WARPS) {
        // ===================== producer warpgroup: one thread issues every TMA of the CTA =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Cfg::PRODUCER_REGS));
        if (warp == Cfg::CONSUMER_WARPS && lane == 0) {
            tma_prefetch_desc(&mapA);
            tma_prefetch_desc(&mapX);
            int s = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int tm, tn;
                raster(tile, tiles_m, tiles_n, group_m, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&empty[s], phase ^ 1);
                    mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    unsigned char* st = tiles + (size_t)s * Cfg::STAGE_BYTES;
#pragma unroll
                    for (int sub = 0; sub < KSUB; ++sub) {
                        const int k0 = kt * BK + sub * 16;
                        unsigned char* sa = st + sub * Cfg::SUB_BYTES;
#pragma unroll
                        for (int mo = 0; mo < BM / 16; ++mo) tma_load_2d(sa + mo * 2048, &mapA, &full[s], m0 + mo * 16, k0);
                        tma_load_2d(sa + Cfg::A_SUB_BYTES, &mapX, &full[s], k0, n0);
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===================== consumers: 2 warpgroups = 8 warps, 64 x 32 each =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, t = lane & 3;
    // A fragment: physical row inside a 16-row box for MMA row g of sub-tile 0 (sub-tile 1 adds 4): chunk/half form
    const int a_chunk = ((g >> 1) & 1) * 4 + (g >> 2);  // logical 16-byte chunk (sub-tile 1: +2)
    const int a_half = g & 1;
    // X fragment: physical column inside an 8-column group for MMA column g
    const int sg_g = 2 * (g & 3) + (g >> 2);
    // accumulator columns 2t, 2t+1 -> physical columns sg(2t), sg(2t+1)
    const int sg_c0 = 2 * ((2 * t) & 3) + ((2 * t) >> 2);
    const int sg_c1 = 2 * ((2 * t + 1) & 3) + ((2 * t + 1) >> 2);
    const int pi_g = (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2);  // + 4*sub
    // Thread-constant swizzled byte offsets; everything else in a fragment address is a compile-time immediate.
    //   A: row r = 4*k4 + t, chunk = (a_chunk | p<<1) ^ (r & 7) = a_chunk ^ t ^ (p<<1) ^ (q<<2),  p = mi&1, q = k4&1
    //   X: chunk = (2*k4 + (t>>1)) ^ sg(g)
    const uint32_t tiles_u32 = smem_u32(tiles);
    uint32_t offA[2][2], offB[4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            offA[p][q] = (uint32_t)((wm * 4) * 2048 + t * 128 + (((a_chunk ^ t) ^ (p << 1) ^ (q << 2)) << 4) + a_half * 8);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
        offB[k4] = (uint32_t)((wn * 32 + sg_g) * 128 + (((k4 * 2 + (t >> 1)) ^ sg_g) << 4) + (t & 1) * 8);

    const bool k_tail = (K % BK) != 0;
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        raster(tile, tiles_m, tiles_n, group_m, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;

        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if constexpr (ACC) {
                        int gm = m0 + wm * 64 + (mi >> 1) * 16 + pi_g + 4 * (mi & 1);
                        int gn = n0 + wn * 32 + ni * 8 + (c ? sg_c1 : sg_c0);
                        acc[mi][ni][c] = (gm < M && gn < N) ? D[(size_t)gn * ldd + gm] : 0.0;
                    } else {
                        acc[mi][ni][c] = -0.0;
                    }
                }

        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(&full[s], phase);
            const uint32_t st = tiles_u32 + (uint32_t)s * Cfg::STAGE_BYTES;
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    // address = stage base + one of 8 thread-constant swizzled offsets + a compile-time immediate
                    double a[8], b[4];
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi)
                        a[mi] = lds_f64(st + offA[mi & 1][k4 & 1] + (sub * Cfg::SUB_BYTES + (mi >> 1) * 2048 + k4 * 512));
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                        b[ni] = lds_f64(st + offB[k4] + (sub * Cfg::SUB_BYTES + Cfg::A_SUB_BYTES + ni * 1024));
                    // K tail: TMA zero-fills k >= K in both operands; (+0)*(+0) added to a -0.0 accumulator would give
                    // +0.0.  Feeding -0.0 on the X side makes the padded product -0.0, and c + (-0.0) == c for every c.
                    if (k_tail && kt == KT - 1 && (kt * BK + sub * 16 + k4 * 4 + t) >= K) {
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) b[ni] = -0.0;
                    }
#pragma unroll
                    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }

        // ---- store (overwrite, column-major; src/gemm.jl:3-11), un-permuting rows/columns ----
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gn = n0 + wn * 32 + ni * 8 + (c ? sg_c1 : sg_c0);
                if (gn >= N) continue;
                double* dcol = D + (size_t)gn * ldd;
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    const int gm = m0 + wm * 64 + (mi >> 1) * 16 + pi_g + 4 * (mi & 1);
                    if (gm < M) dcol[gm] = acc[mi][ni][c];
                }
            }
    }
}

}  // namespace jb

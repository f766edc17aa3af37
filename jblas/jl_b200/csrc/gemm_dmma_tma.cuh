// gemm_dmma_tma.cuh -- FP64 tensor-core GEMM, warp-specialised TMA producer / DMMA consumers, persistent CTAs.
//
// This is the main FP64 kernel family.  It replaces, in one design, the three things north_star names:
//   * "packing buffers / cache blocking" (the reference has only software prefetch, src/memory_management.jl:181-277,
//     and an uncalled planner, :78-140)  ->  TMA (cp.async.bulk.tensor) stages A and X tiles into a multi-stage
//     shared-memory ring; no thread spends issue slots on copies;
//   * the register-tile SIMD micro-kernel (src/gemm.jl:149-170, src/kernels.jl:212-275)  ->  mma.sync.m8n8k4.f64
//     (SASS DMMA.8x8x4), MI x NI tiles per warp, accumulators in registers (FP64 has no tcgen05/TMEM kind);
//   * the two tile loops of jmul! (src/gemm.jl:313)  ->  a persistent grid (one or two CTAs per SM) walking a rasterised
//     tile list, so the producer runs ahead across tile boundaries and the pipeline never drains.  The list is handed
//     out DYNAMICALLY: a CTA's first tile is blockIdx.x, every later one comes from a global atomic counter that the
//     producer thread fetches one tile ahead and passes to the consumers through the pipeline stage (stage_tile[s]).
//     A statically strided list leaves SMs idle in the last round (and, with two CTAs per SM, can put both long
//     lists on one SM: 2048^3 ran at 8 tile-times instead of 7.25).  The counter pair resets itself: the last CTA to
//     draw a past-the-end ticket zeroes it for the next launch that uses the slot (capi.cu hands out slots in a ring).
//
// Roles: WARPS_M x WARPS_N consumer warps (4 or 8 = 1 or 2 warpgroups) + one producer warpgroup whose first lane
// issues every TMA of the CTA.  With 8 consumer warps, setmaxnreg moves the idle producer warpgroup's registers to the
// consumers (a lone 9th warp would put 3 warps on one SM sub-partition and cap EVERY thread at 168 registers).
// Synchronisation is mbarrier-only in the main loop: full[s] (TMA complete_tx) and empty[s] (one arrive per
// consumer warp); no CTA-wide barrier after start-up.
//
// Shared-memory layout = what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index XOR row&7):
//   A sub-tile: BM/16 boxes (16 m x 16 k), box = 16 rows (k) of 128 B (16 doubles of m)
//   X sub-tile: 1 box (16 k x BN n),       BN rows (n) of 128 B (16 doubles of k)
// DMMA fragment loads are 8-byte LDS; a half-warp (16 lanes) must hit 16 distinct 8-byte slots of the 128-byte
// bank window.  With the hardware swizzle that holds if the LOGICAL rows/columns of an MMA tile are permuted:
//   A: MMA row g of sub-tile `sub` of a 16-row box is physical row  pi(g) = (g&1) + 8*((g>>1)&1) + 2*(g>>2) + 4*sub
//   X: MMA column g of an 8-column group is physical column          sg(g) = 2*(g&3) + (g>>2)
// (derivation in DESIGN.md "bank-conflict-free fragments"); ncu: 0.05 % of shared wavefronts conflict.
// The accumulator fragment is un-permuted on the way out with the same two maps.
//
// Numerics: ascending k, 4 at a time, starting from -0.0 (or the old D when ACC).  Contract = the reference tolerance
// 2*K*eps*(|A||X|); measured on B200: bit-identical to the sequential fma chain (tests/test_gemm_gpu.py).
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
// Same, with an L2 eviction-priority hint (the 64-bit policy encodings createpolicy.fractional produces for fraction 1.0).
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull, kL2EvictFirst = 0x12F0000000000000ull, kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// 3-D form (third coordinate = the product of a batch; fastmul_batched on the same pipeline)
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
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

// CTA tile = (WARPS_M * MI * 8) x (WARPS_N * NI * 8); warp tile = MI x NI DMMA tiles (MI even: a 16-row box = 2 tiles).
template <int WARPS_M_, int WARPS_N_, int MI_, int NI_, int KSUB_, int STAGES_, int MINB_ = 1>
struct DmmaTmaCfg {
    static constexpr int MIN_BLOCKS = MINB_;  // CTAs per SM: 2 lets one CTA's epilogue overlap the other's main loop
    static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, MI = MI_, NI = NI_;
    static constexpr int BM = WARPS_M * MI * 8, BN = WARPS_N * NI * 8, KSUB = KSUB_, BK = 16 * KSUB_, STAGES = STAGES_;
    static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
    static_assert(MI % 2 == 0 && (CONSUMER_WARPS == 4 || CONSUMER_WARPS == 8) && BN <= 256, "unsupported tile");
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr bool REALLOC_REGS = CONSUMER_WARPS == 8 && MINB_ == 1;  // 384 threads: 168 regs at launch -> 232 / 40
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr int A_SUB_BYTES = BM * 128;  // BM/16 boxes x 2 KiB
    static constexpr int B_SUB_BYTES = BN * 128;  // 1 box
    static constexpr int SUB_BYTES = A_SUB_BYTES + B_SUB_BYTES;
    static constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
    static_assert(SUB_BYTES % 1024 == 0 && A_SUB_BYTES % 1024 == 0, "swizzle needs 1024-byte aligned tiles");
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + STAGES * sizeof(int) + 1024;  // +align slack
};

// One pipeline stage of MMAs.  TAIL = this is the last k-tile and K is not a multiple of BK: TMA zero-filled k >= K in
// both operands, and (+0)*(+0) added to a -0.0 accumulator would give +0.0; feeding -0.0 on the X side makes the
// padded product -0.0, and c + (-0.0) == c for every c.  Kept out of the steady-state loop (template) on purpose.
template <typename Cfg, bool TAIL>
__device__ __forceinline__ void dmma_consume_stage(double (&acc)[Cfg::MI][Cfg::NI][2], uint32_t st, const uint32_t (&offA)[2][2],
                                                   const uint32_t (&offB)[4], int k_stage0, int K, int t)
{
    // Fragments are double-buffered in registers: the loads of step i+1 are issued before the MMAs of step i, so a
    // warp that is alone on its SM sub-partition (4-warp tiles) does not expose the shared-memory latency.
    // address = stage base + one of 8 thread-constant swizzled offsets + a compile-time immediate
    constexpr int STEPS = Cfg::KSUB * 4;
    double a[2][Cfg::MI], b[2][Cfg::NI];
    auto load = [&](int step, int buf) {
        const int sub = step >> 2, k4 = step & 3;
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; ++mi)
            a[buf][mi] = lds_f64(st + offA[mi & 1][k4 & 1] + (sub * Cfg::SUB_BYTES + (mi >> 1) * 2048 + k4 * 512));
#pragma unroll
        for (int ni = 0; ni < Cfg::NI; ++ni)
            b[buf][ni] = lds_f64(st + offB[k4] + (sub * Cfg::SUB_BYTES + Cfg::A_SUB_BYTES + ni * 1024));
        if constexpr (TAIL) {
            if (k_stage0 + step * 4 + t >= K) {
#pragma unroll
                for (int ni = 0; ni < Cfg::NI; ++ni) b[buf][ni] = -0.0;
            }
        }
    };
    load(0, 0);
#pragma unroll
    for (int step = 0; step < STEPS; ++step) {
        if (step + 1 < STEPS) load(step + 1, (step + 1) & 1);
#pragma unroll
        for (int mi = 0; mi < Cfg::MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < Cfg::NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[step & 1][mi], b[step & 1][ni]);
    }
}

// 8-byte cp.async whose completion counts as one arrival on an mbarrier (the RAGGED producers below)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// BATCHED: the tile list runs over `batch` independent products (3-D tensor maps, third coordinate = product; D of product b
// starts batch_stride_d elements after that of product b-1) -- the fastmul!-class batch for products too big for one warp.
// RAGGED: operands whose base or leading dimension breaks the 16-byte alignment TMA needs (M = 1023 doubles per column): the
// whole producer warpgroup stages the tiles with element-wise 8-byte cp.async INTO THE SAME 128B-SWIZZLED LAYOUT the TMA
// boxes would have, so the consumers, the pipeline and the arithmetic are unchanged and no re-aligned copy of A and X is
// made (the reference's analogue: the masked row remainder of fastmul!, src/kernels.jl:59-75,101-120).  Out-of-range
// elements are zero-filled like TMA does.  Tiles are handed out by static stride (the ticket would have to be broadcast).
template <typename Cfg, bool ACC, bool BATCHED = false, bool RAGGED = false>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_BLOCKS)
gemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX,
                     double* D, int M, int N, int K, int64_t ldd, int tiles_m, int tiles_n, int group_m,
                     uint64_t l2_policy_a, uint64_t l2_policy_x, int* __restrict__ tile_ctr, const double* Cin, int64_t ldc,
                     int batch = 1, int64_t batch_stride_d = 0, const double* __restrict__ Araw = nullptr, const double* __restrict__ Xraw = nullptr,
                     int64_t lda = 0, int64_t ldx = 0)
{
    static_assert(!(RAGGED && BATCHED) && (!RAGGED || Cfg::MI * Cfg::NI <= 16), "ragged staging: single products, warp tiles that fit 168 registers");
    constexpr bool REALLOC = Cfg::REALLOC_REGS && !RAGGED;  // the ragged producers do address arithmetic: they keep their registers
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, KSUB = Cfg::KSUB, STAGES = Cfg::STAGES;
    constexpr int MI = Cfg::MI, NI = Cfg::NI;
    extern __shared__ unsigned char smem_raw[];
    // the 128B swizzle is a function of address bits 4..9: tile bases must be 1024-byte aligned
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    volatile int* stage_tile = reinterpret_cast<volatile int*>(empty + STAGES);  // tile a stage belongs to; -1 = no more work

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Programmatic dependent launch (host: launch_pdl).  launch_dependents lets the NEXT kernel of the stream be scheduled while
    // this one runs (its CTAs take the SMs as ours retire and do their barrier set-up early); wait holds THIS kernel until every
    // earlier kernel of the stream has completed and flushed -- before the first global access (tile counter, TMA, C).  Both
    // are no-ops for plain launches.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], RAGGED ? 129 : 1);  // RAGGED: one cp.async-completion arrival per producer thread + one release-arrive
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    const int tiles_per_product = tiles_m * tiles_n;
    const int num_tiles = tiles_per_product * (BATCHED ? batch : 1);
    const int KT = (K + BK - 1) / BK;

    if (warp >= Cfg::CONSUMER_WARPS) {
        // ===================== producer warpgroup: one thread issues every TMA of the CTA =====================
        if constexpr (REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Cfg::PRODUCER_REGS));
        if constexpr (RAGGED) {
            const int ptid = tid - Cfg::CONSUMER_WARPS * 32;  // 0..127
            // Everything that does not change from stage to stage is computed ONCE: the producers share their issue slots with
            // the DMMA consumers, and recomputing 40 index splits and 64-bit addresses per stage cost 13 % of the kernel.
            // A sub-tile: BM/16 boxes of 16 rows (k) x 128 B (16 m); element (m, kk): chunk (m%16)/2 ^ (kk & 7), half m & 1
            // X sub-tile: BN rows (n) x 128 B (16 k);                element (kk, n): chunk kk/2 ^ (n & 7),      half kk & 1
            constexpr int NA = (BM * 16) / 128, NX = (BN * 16) / 128;
            uint32_t dstA[NA], dstX[NX];   // shared-memory byte offsets inside a sub-tile
            int64_t srcA[NA], srcX[NX];    // element offsets from the tile / k origin (row / column clamped into the matrix, per tile)
            int mlA[NA], kkA[NA], nlX[NX], kkX[NX];
#pragma unroll
            for (int it = 0; it < NA; ++it) {
                const int idx = ptid + it * 128, ml = idx % BM, kk = idx / BM;
                mlA[it] = ml; kkA[it] = kk;
                dstA[it] = (uint32_t)((ml >> 4) * 2048 + kk * 128 + (((((ml & 15) >> 1) ^ (kk & 7))) << 4) + (ml & 1) * 8);
            }
#pragma unroll
            for (int it = 0; it < NX; ++it) {
                const int idx = ptid + it * 128, kk = idx & 15, nl = idx >> 4;
                nlX[it] = nl; kkX[it] = kk;
                dstX[it] = (uint32_t)(Cfg::A_SUB_BYTES + nl * 128 + ((((kk >> 1) ^ (nl & 7))) << 4) + (kk & 1) * 8);
            }
            int s = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int tm, tn;
                raster(tile, tiles_m, tiles_n, group_m, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                // Per tile: which of this thread's rows / columns exist, and source offsets whose row / column is clamped into the
                // matrix -- every address formed below is valid, an element outside the matrix is a copy of ZERO bytes (zero fill,
                // like TMA).  The per-copy work in the k loop is then one 64-bit add, the size from the bit mask and the copy: these
                // instructions share the issue slots with the DMMA consumers (ncu: tensor pipe 88 % against 97 % with TMA boxes).
                uint32_t okA = 0, okX = 0;
                const int mlast = M - 1 - m0, nlast = N - 1 - n0;
#pragma unroll
                for (int it = 0; it < NA; ++it) {
                    okA |= (uint32_t)(mlA[it] <= mlast) << it;
                    srcA[it] = (int64_t)kkA[it] * lda + min(mlA[it], mlast);
                }
#pragma unroll
                for (int it = 0; it < NX; ++it) {
                    okX |= (uint32_t)(nlX[it] <= nlast) << it;
                    srcX[it] = (int64_t)min(nlX[it], nlast) * ldx + kkX[it];
                }
                const bool whole = (m0 + BM <= M) && (n0 + BN <= N);  // no size arithmetic at all
                const double* tileA = Araw + m0;
                const double* tileX = Xraw + (size_t)n0 * ldx;
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&empty[s], phase ^ 1);
                    const uint32_t st = smem_u32(tiles) + (uint32_t)s * Cfg::STAGE_BYTES;
#pragma unroll
                    for (int sub = 0; sub < KSUB; ++sub) {
                        const int k0 = kt * BK + sub * 16;
                        const uint32_t sa = st + sub * Cfg::SUB_BYTES;
                        const double* pa = tileA + (size_t)k0 * lda;
                        const double* px = tileX + k0;
                        if (k0 + 16 <= K) {  // every k of the sub-tile exists (all but the last k-tile)
                            if (whole) {
#pragma unroll
                                for (int it = 0; it < NA; ++it) cp_async8(sa + dstA[it], pa + srcA[it], 8);
#pragma unroll
                                for (int it = 0; it < NX; ++it) cp_async8(sa + dstX[it], px + srcX[it], 8);
                            } else {
#pragma unroll
                                for (int it = 0; it < NA; ++it) cp_async8(sa + dstA[it], pa + srcA[it], (int)((okA >> it) & 1u) << 3);
#pragma unroll
                                for (int it = 0; it < NX; ++it) cp_async8(sa + dstX[it], px + srcX[it], (int)((okX >> it) & 1u) << 3);
                            }
                        } else {  // K tail: out-of-range k is zero-filled too, and its address is never formed
#pragma unroll
                            for (int it = 0; it < NA; ++it) {
                                const bool ok = ((okA >> it) & 1u) && k0 + kkA[it] < K;
                                cp_async8(sa + dstA[it], ok ? pa + srcA[it] : Araw, ok ? 8 : 0);
                            }
#pragma unroll
                            for (int it = 0; it < NX; ++it) {
                                const bool ok = ((okX >> it) & 1u) && k0 + kkX[it] < K;
                                cp_async8(sa + dstX[it], ok ? px + srcX[it] : Xraw, ok ? 8 : 0);
                            }
                        }
                    }
                    cp_async_mbar_arrive_noinc(&full[s]);  // fires when this thread's copies have landed
                    if (ptid == 0) {
                        stage_tile[s] = tile;
                        mbar_arrive(&full[s]);  // release: orders the store above before the consumers' read
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
            mbar_wait(&empty[s], phase ^ 1);  // end marker: a stage that carries no data
            if (ptid == 0) stage_tile[s] = -1;
            mbar_arrive(&full[s]);
            if (ptid == 0) mbar_arrive(&full[s]);  // 129 arrivals complete the phase
            return;
        }
        if (warp == Cfg::CONSUMER_WARPS && lane == 0) {
            tma_prefetch_desc(&mapA);
            tma_prefetch_desc(&mapX);
            int s = 0;
            uint32_t phase = 0;
            int tile = blockIdx.x;  // grid <= num_tiles: every CTA owns at least this one
            while (tile < num_tiles) {
                // ticket for the NEXT tile, drawn now so the atomic's round trip hides behind this tile's k loop
                // (tile_ctr == nullptr: static stride, kept for A/B measurements -- JBLAS_B200_STATIC_TILES=1)
                const int next = tile_ctr ? atomicAdd(tile_ctr, 1) + (int)gridDim.x : tile + (int)gridDim.x;
                int tm, tn;
                const int prod = BATCHED ? tile / tiles_per_product : 0;
                raster(BATCHED ? tile - prod * tiles_per_product : tile, tiles_m, tiles_n, group_m, tm, tn);
                const int m0 = tm * BM, n0 = tn * BN;
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&empty[s], phase ^ 1);
                    stage_tile[s] = tile;  // ordered before the consumers' read by the release-arrive below
                    mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    unsigned char* st = tiles + (size_t)s * Cfg::STAGE_BYTES;
#pragma unroll
                    for (int sub = 0; sub < KSUB; ++sub) {
                        const int k0 = kt * BK + sub * 16;
                        unsigned char* sa = st + sub * Cfg::SUB_BYTES;
#pragma unroll
                        // L2 eviction priorities per operand are chosen by the host (capi.cu: launch_dmma_tma).
                        for (int mo = 0; mo < BM / 16; ++mo) {
                            if constexpr (BATCHED)
                                tma_load_3d_hint(sa + mo * 2048, &mapA, &full[s], m0 + mo * 16, k0, prod, l2_policy_a);
                            else
                                tma_load_2d_hint(sa + mo * 2048, &mapA, &full[s], m0 + mo * 16, k0, l2_policy_a);
                        }
                        if constexpr (BATCHED)
                            tma_load_3d_hint(sa + Cfg::A_SUB_BYTES, &mapX, &full[s], k0, n0, prod, l2_policy_x);
                        else
                            tma_load_2d_hint(sa + Cfg::A_SUB_BYTES, &mapX, &full[s], k0, n0, l2_policy_x);
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                tile = next;
            }
            // end marker: one more stage that carries no data
            mbar_wait(&empty[s], phase ^ 1);
            stage_tile[s] = -1;
            mbar_arrive(&full[s]);
            // every CTA draws exactly one past-the-end ticket; whoever reports last knows nobody will draw again
            if (tile_ctr && atomicAdd(tile_ctr + 1, 1) == (int)gridDim.x - 1) {
                tile_ctr[0] = 0;
                tile_ctr[1] = 0;
            }
        }
        return;
    }

    // ===================== consumers: WARPS_M x WARPS_N warps, (MI*8) x (NI*8) each =====================
    if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
    const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
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
    const int wrow0 = wm * MI * 8, wcol0 = wn * NI * 8;            // warp origin inside the CTA tile
    // Thread-constant swizzled byte offsets; everything else in a fragment address is a compile-time immediate.
    //   A: row r = 4*k4 + t, chunk = (a_chunk | p<<1) ^ (r & 7) = a_chunk ^ t ^ (p<<1) ^ (q<<2),  p = mi&1, q = k4&1
    //   X: chunk = (2*k4 + (t>>1)) ^ sg(g)
    const uint32_t tiles_u32 = smem_u32(tiles);
    uint32_t offA[2][2], offB[4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q)
            offA[p][q] = (uint32_t)((wrow0 / 16) * 2048 + t * 128 + (((a_chunk ^ t) ^ (p << 1) ^ (q << 2)) << 4) + a_half * 8);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
        offB[k4] = (uint32_t)((wcol0 + sg_g) * 128 + (((k4 * 2 + (t >> 1)) ^ sg_g) << 4) + (t & 1) * 8);

    const bool k_tail = (K % BK) != 0;
    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&full[s], phase);  // first stage of the next tile (waited on again, trivially, at kt = 0)
        const int tile = stage_tile[s];
        if (tile < 0) break;
        int tm, tn;
        const int prod = BATCHED ? tile / tiles_per_product : 0;
        raster(BATCHED ? tile - prod * tiles_per_product : tile, tiles_m, tiles_n, group_m, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        double* const Dp = D + (BATCHED ? (int64_t)prod * batch_stride_d : 0);

        double acc[MI][NI][2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if constexpr (ACC) {
                        int gm = m0 + wrow0 + (mi >> 1) * 16 + pi_g + 4 * (mi & 1);
                        int gn = n0 + wcol0 + ni * 8 + (c ? sg_c1 : sg_c0);
                        acc[mi][ni][c] = (gm < M && gn < N) ? Cin[(size_t)gn * ldc + gm] : 0.0;
                    } else {
                        acc[mi][ni][c] = -0.0;
                    }
                }

        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(&full[s], phase);
            const uint32_t st = tiles_u32 + (uint32_t)s * Cfg::STAGE_BYTES;
            if (k_tail && kt == KT - 1)
                dmma_consume_stage<Cfg, true>(acc, st, offA, offB, kt * BK, K, t);
            else
                dmma_consume_stage<Cfg, false>(acc, st, offA, offB, kt * BK, K, t);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }

        // ---- store (overwrite, column-major; src/gemm.jl:3-11), un-permuting rows/columns ----
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gn = n0 + wcol0 + ni * 8 + (c ? sg_c1 : sg_c0);
                if (gn >= N) continue;
                double* dcol = Dp + (size_t)gn * ldd;
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    const int gm = m0 + wrow0 + (mi >> 1) * 16 + pi_g + 4 * (mi & 1);
                    if (gm < M) dcol[gm] = acc[mi][ni][c];
                }
            }
    }
}

}  // namespace jb

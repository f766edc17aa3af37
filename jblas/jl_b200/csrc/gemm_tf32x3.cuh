// gemm_tf32x3.cuh -- opt-in Float32 GEMM on the 5th-generation tensor cores: 3xTF32 split precision,
// tcgen05.mma (kind::tf32) with the accumulator in TMEM, TMA-fed shared-memory ring, persistent CTAs.
//
// north_star: "FP32 runs on an exact SIMT path plus an opt-in 3xTF32 tcgen05/TMEM split-precision path".
// The exact path (gemm_simt_f32x2.cuh) reproduces the reference's chain bit for bit; this one trades that for
// tensor-core throughput under a STATED looser bound:
//       |D - D_oracle32|_ij <= (2*K*2^-23 + 2^-18) * (|A||X|)_ij
// Method: every operand is split a = hi + lo with hi = tf32(a) (round-to-nearest, cvt.rna.tf32.f32) and
// lo = tf32(a - hi) in a pre-pass (split_tf32_kernel); per k-block three MMAs accumulate into ONE FP32 TMEM
// accumulator:  lo*hi, hi*lo, hi*hi  (the lo*lo term, ~2^-22 relative, is dropped).  Credited flops stay 2*M*N*K.
//
// Kernel anatomy (256 threads, one CTA per SM, persistent over a rasterised tile list; tile = 128 x BN):
//   warp 0 lane 0 : TMA producer -- 4 tensor maps (A^T_hi, A^T_lo as 32(k) x 128(m) boxes; X_hi, X_lo as 32(k) x BN
//                   boxes; all K-major), CU_TENSOR_MAP_SWIZZLE_128B, completes on full[s]
//   warp 1 lane 0 : MMA issuer   -- 3 x 4 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) per stage from
//                   shared-memory descriptors; tcgen05.commit frees the stage (empty[s]) and publishes the
//                   accumulator (tmem_full[a])
//   warp 2        : TMEM allocator (512 columns = two accumulators: the epilogue of tile i overlaps the MMAs of i+1)
//   warps 4..7    : epilogue     -- tcgen05.ld 32x32b.x32 (lane = row m, 32 consecutive columns n), coalesced column-
//                   major stores (a warp writes 128 contiguous bytes per column), then tmem_empty[a]
//
// Shared-memory operand layout = the canonical K-major UMMA layout that TMA's 128B swizzle produces directly:
//   rows (m for A^T, n for X) of 128 B = 32 k; 8 rows = one 1024-byte swizzle atom
//   -> descriptor SBO = 1024 B (next 8 rows), start address +32 B per UMMA_K (8 TF32)
// A is column-major in HBM (M contiguous), which would be an "MN-major" operand; MN-major TF32 needs the special
// 128B-swizzle/32B-atom layout, so the split pre-pass writes A TRANSPOSED instead (K contiguous) at no extra traffic.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_dmma_tma.cuh"

namespace jb {

// a = hi + lo, both exactly representable in TF32: hi = rna(a), lo = rna(a - hi).  Non-finite a travels in hi only.  A FINITE a
// within half a TF32 ulp of FLT_MAX would round UP to +-Inf (and a - Inf, Inf*x + (-Inf)*x would poison the sum with NaN):
// there hi is truncated toward zero instead (still TF32-exact, |a - hi| < 1 ulp) and lo carries the rest.
__device__ __forceinline__ void split_tf32(float a, float& hf, float& lf)
{
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(h) : "f"(a));
    hf = __uint_as_float(h);
    if (isfinite(a) && !isfinite(hf)) hf = __uint_as_float(__float_as_uint(a) & 0xffffe000u);  // round toward zero: stays finite
    float rest = a - hf;  // exact: hi holds the leading bits of a
    if (!isfinite(a)) rest = 0.f;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(l) : "f"(rest));
    lf = __uint_as_float(l);
}

// ---- operand split pre-pass: src (rows x cols, ld) -> hi, lo (rows x cols, ld2), both exactly representable in TF32 ----
__global__ void split_tf32_kernel(const float* __restrict__ src, int64_t lds, int rows, int cols, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t ld2)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int c = blockIdx.y; c < cols; c += gridDim.y) {
        float hf, lf;
        split_tf32(src[(size_t)c * lds + r], hf, lf);
        hi[(size_t)c * ld2 + r] = hf;
        lo[(size_t)c * ld2 + r] = lf;
    }
}

// Same split, but the output is TRANSPOSED: src is rows x cols (column-major, rows contiguous), hi/lo are written as
// cols x rows (cols contiguous).  Used for A: MN-major TF32 operands need the special 128B/32B-atom layout, whereas a
// K-major A^T uses exactly the same plain 128B-swizzle layout as X.  32x32 tiles through shared memory keep both the
// global read (along rows) and the global writes (along cols) coalesced.
__global__ void split_tf32_transpose_kernel(const float* __restrict__ src, int64_t lds, int rows, int cols, float* __restrict__ hi,
                                            float* __restrict__ lo, int64_t ld2)
{
    __shared__ float th[32][33], tl[32][33];
    const int r0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;  // block (32, 8)
    for (int c0 = blockIdx.y * 32; c0 < cols; c0 += gridDim.y * 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + tx, c = c0 + ty + 8 * j;
            float hf = 0.f, lf = 0.f;
            if (r < rows && c < cols) {
                split_tf32(src[(size_t)c * lds + r], hf, lf);
            }
            th[ty + 8 * j][tx] = hf;  // [c][r]
            tl[ty + 8 * j][tx] = lf;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx, r = r0 + ty + 8 * j;  // output element (c, r): c contiguous
            if (r < rows && c < cols) {
                hi[(size_t)r * ld2 + c] = th[tx][ty + 8 * j];
                lo[(size_t)r * ld2 + c] = tl[tx][ty + 8 * j];
            }
        }
        __syncthreads();
    }
}

// ---- tcgen05 primitives (inline PTX) -----------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// UMMA shared-memory matrix descriptor (sm_100 format, cute/arch/mma_sm100_desc.hpp documents the bit fields):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int BN_, int STAGES_>
struct Tf32x3Cfg {
    static constexpr int BM = 128, BN = BN_, BK = 32, STAGES = STAGES_;
    static constexpr int THREADS = 256;
    static constexpr int A_BYTES = BM * BK * 4;  // one of hi/lo: 4 m-chunks x (32 k-rows x 128 B)
    static constexpr int B_BYTES = BN * BK * 4;  // one of hi/lo: BN rows x 128 B
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int TMEM_COLS = 512;  // two accumulators of BN (<= 256) columns; power of two >= 32
    static_assert(BN % 32 == 0 && BN <= 256 && 2 * BN <= TMEM_COLS, "unsupported BN");
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
    // instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), A and B K-major (bits 15,16 = 0), N>>3 at 17, M>>4 at 24
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo,
                   float* D, int M, int N, int K, int64_t ldd, int tiles_m, int tiles_n, int group_m, const float* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    const int num_tiles = tiles_m * tiles_n;
    const int KT = (K + BK - 1) / BK;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapXhi); tma_prefetch_desc(&mapXlo);
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
                    const int k0 = kt * BK;
                    tma_load_2d(st, &mapAhi, &full[s], k0, m0);
                    tma_load_2d(st + Cfg::A_BYTES, &mapAlo, &full[s], k0, m0);
                    tma_load_2d(st + 2 * Cfg::A_BYTES, &mapXhi, &full[s], k0, n0);
                    tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapXlo, &full[s], k0, n0);
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t tiles_u32 = smem_u32(tiles);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&full[s], phase);
                    tc_fence_after();
                    const uint32_t st = tiles_u32 + (uint32_t)s * Cfg::STAGE_BYTES;
                    const uint32_t a_hi = st, a_lo = st + Cfg::A_BYTES;
                    const uint32_t b_hi = st + 2 * Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        // both operands K-major: next UMMA_K = next 8 k inside the 128-byte row (+32 B)
                        const uint64_t dah = umma_smem_desc(a_hi + ks * 32, 16, 1024);
                        const uint64_t dal = umma_smem_desc(a_lo + ks * 32, 16, 1024);
                        const uint64_t dbh = umma_smem_desc(b_hi + ks * 32, 16, 1024);
                        const uint64_t dbl = umma_smem_desc(b_lo + ks * 32, 16, 1024);
                        umma_tf32(d_tmem, dal, dbh, Cfg::IDESC, (kt | ks) ? 1u : 0u);  // small terms first
                        umma_tf32(d_tmem, dah, dbl, Cfg::IDESC, 1u);
                        umma_tf32(d_tmem, dah, dbh, Cfg::IDESC, 1u);
                    }
                    umma_commit(&empty[s]);  // the stage may be overwritten once these MMAs have read it
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int tm, tn;
            raster(tile, tiles_m, tiles_n, group_m, tm, tn);
            const int gm = tm * BM + q * 32 + lane;  // this thread's row
            const int n0 = tn * BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                if (gm < M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int gn = n0 + c * 32 + j;
                        if (gn < N) {
                            float* p = D + (size_t)gn * ldd + gm;
                            float v = __uint_as_float(r[j]);
                            if constexpr (ACC) v += Cin[(size_t)gn * ldc + gm];
                            *p = v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =====================================================================================================================
// CTA-pair variant (cta_group::2): the two CTAs of a cluster -- two SMs of one TPC -- multiply ONE 256 x 256 tile of D.
//
// Why: the single-CTA kernel above moves (128 + 256) x 32 x 4 x 2 = 96 KiB per 32-k stage per SM through L2 and shared memory
// for 3 x 4 MMAs of 128 x 256 x 8: 64 MMA-flop per byte.  At 8192^3 that is 51.5 GB of tile traffic in 4.1 ms = 12.5 TB/s
// out of L2 and 96 B/clk of shared-memory reads per SM, with only two 96 KiB stages to hide it (tensor pipe 81 % active,
// profiles/r1_ncu_tf32x3_8192.txt).  A pair shares the X operand: each CTA stages its 128 rows of A^T (hi, lo) and HALF of the
// 256 X columns (hi, lo) = 64 KiB per stage, tcgen05.mma.cta_group::2 reads A from both CTAs (M = 256) and the two X halves
// from the two shared memories (N = 256), and every CTA's TMEM receives its own 128 rows x 256 columns of D.  L2 and
// shared-memory traffic per flop drop by a third (96 flop per byte) and three stages fit.
//
// Protocol (the leader is cluster rank 0):
//   * TMA producers run in BOTH CTAs (each fills its own shared memory), but every box completes its bytes on the LEADER's
//     full[s] (cp.async.bulk.tensor...cta_group::2 with the barrier address mapped into rank 0); the leader's producer arms
//     full[s] with the bytes of both CTAs.
//   * only the leader issues MMAs; tcgen05.commit...multicast::cluster arrives on empty[s] / tmem_full[a] of BOTH CTAs.
//   * the epilogue warps of both CTAs drain their own TMEM and arrive on the LEADER's tmem_empty[a] (count 8, remote arrive
//     from the peer).
//   * TMEM allocation and release are pair-wide (cta_group::2), bracketed by cluster barriers; a producer leaves only after
//     every commit aimed at its CTA has landed.
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank)
{
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA box into THIS CTA's shared memory, bytes completed on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int STAGES_>
struct Tf32x3PairCfg {
    static constexpr int BM = 256, BN = 256, BK = 32, STAGES = STAGES_;  // tile of the PAIR; each CTA: 128 rows of A, 128 columns of X
    static constexpr int THREADS = 256;
    static constexpr int A_BYTES = 128 * BK * 4;  // one of hi/lo
    static constexpr int B_BYTES = 128 * BK * 4;  // one of hi/lo, this CTA's half of the columns
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // per CTA
    static constexpr int TMEM_COLS = 512;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + (2 * STAGES + 4) * sizeof(uint64_t) + 16 + 1024;
    // c = F32, a = b = TF32, both K-major, N = 256, M = 256 (the pair's)
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <typename Cfg, bool ACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                        const __grid_constant__ CUtensorMap mapXhi, const __grid_constant__ CUtensorMap mapXlo,
                        float* D, int M, int N, int K, int64_t ldd, int tiles_m, int tiles_n, int group_m, const float* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    // the dynamic shared-memory window starts at the same offset in both CTAs, so the aligned tile base does too
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2], used in the leader only
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);   // the leader's producer arms it; both CTAs' boxes complete bytes on the leader's
            mbar_init(&empty[s], 1);  // one multicast commit per use
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_base_smem, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised and its TMEM is allocated before anything is aimed at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    const int num_tiles = tiles_m * tiles_n;
    const int KT = (K + BK - 1) / BK;
    const int first_tile = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapXhi); tma_prefetch_desc(&mapXlo);
            int s = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                int tm, tn;
                raster(tile, tiles_m, tiles_n, group_m, tm, tn);
                const int m0 = tm * BM + (int)rank * 128, n0 = tn * BN + (int)rank * 128;
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&empty[s], phase ^ 1);
                    if (leader) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
                    const uint32_t bar = mapa_u32(&full[s], 0);
                    unsigned char* st = tiles + (size_t)s * Cfg::STAGE_BYTES;
                    const int k0 = kt * BK;
                    tma_load_2d_pair(st, &mapAhi, bar, k0, m0);
                    tma_load_2d_pair(st + Cfg::A_BYTES, &mapAlo, bar, k0, m0);
                    tma_load_2d_pair(st + 2 * Cfg::A_BYTES, &mapXhi, bar, k0, n0);
                    tma_load_2d_pair(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapXlo, bar, k0, n0);
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
            }
            // tail: every commit aimed at this CTA's empty barriers has landed before the CTA may leave
            for (int i = 0; i < STAGES; ++i) {
                mbar_wait(&empty[s], phase ^ 1);
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the leader) =====================
        if (leader && lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t tiles_u32 = smem_u32(tiles);
            for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
                mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);  // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&full[s], phase);
                    tc_fence_after();
                    const uint32_t st = tiles_u32 + (uint32_t)s * Cfg::STAGE_BYTES;
                    const uint32_t a_hi = st, a_lo = st + Cfg::A_BYTES;
                    const uint32_t b_hi = st + 2 * Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint64_t dah = umma_smem_desc(a_hi + ks * 32, 16, 1024);
                        const uint64_t dal = umma_smem_desc(a_lo + ks * 32, 16, 1024);
                        const uint64_t dbh = umma_smem_desc(b_hi + ks * 32, 16, 1024);
                        const uint64_t dbl = umma_smem_desc(b_lo + ks * 32, 16, 1024);
                        umma_tf32_pair(d_tmem, dal, dbh, Cfg::IDESC, (kt | ks) ? 1u : 0u);  // small terms first
                        umma_tf32_pair(d_tmem, dah, dbl, Cfg::IDESC, 1u);
                        umma_tf32_pair(d_tmem, dah, dbh, Cfg::IDESC, 1u);
                    }
                    umma_commit_pair(&empty[s]);
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                umma_commit_pair(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs): own TMEM -> registers -> global =====================
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t empty0 = mapa_u32(&tmem_empty[0], 0), empty1 = mapa_u32(&tmem_empty[1], 0);
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            int tm, tn;
            raster(tile, tiles_m, tiles_n, group_m, tm, tn);
            const int gm = tm * BM + (int)rank * 128 + q * 32 + lane;  // this thread's row
            const int n0 = tn * BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                if (gm < M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int gn = n0 + c * 32 + j;
                        if (gn < N) {
                            float* p = D + (size_t)gn * ldd + gm;
                            float v = __uint_as_float(r[j]);
                            if constexpr (ACC) v += Cin[(size_t)gn * ldc + gm];
                            *p = v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? empty1 : empty0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the pair leaves together: no MMA, commit or remote arrive is still aimed at a CTA that has gone
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace jb

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

}  // namespace jb

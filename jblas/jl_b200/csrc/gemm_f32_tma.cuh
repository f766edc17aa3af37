// gemm_f32_tma.cuh -- exact Float32 GEMM, warp-specialised: TMA producer / FFMA2 consumers, persistent CTAs, 8 x 16 thread tile.
//
// Why a third Float32 kernel.  The cp.async FFMA2 kernel (gemm_simt_f32x2.cuh) stops at 54 TFLOP/s, and the ablation probes
//  say why: with an 8 x 8 thread tile the micro-kernel needs 12 shared-memory wavefronts per warp per k -- 75 %
// of the shared-memory pipe at full FMA rate -- and every thread also spends issue slots on cp.async address arithmetic and
// waits at a CTA barrier per k-tile.  This kernel removes all three:
//   * thread tile 8 rows x 16 columns (128 packed accumulators): 16 wavefronts per 128 FMAs instead of 12 per 64 -> 50 % of
//     the shared-memory pipe at full rate (probe: 64 TFLOP/s against 62 for 8 x 8);
//   * operands arrive by TMA (one elected thread of a producer warpgroup issues two boxes per stage), so the consumers issue
//     nothing but LDS and FFMA2;
//   * full/empty mbarriers per stage: no CTA-wide barrier in the k loop; setmaxnreg hands the producer warpgroup's registers
//     to the consumers (232 per thread);
//   * persistent grid over the rasterised tile list with the same self-resetting dynamic tile counter as the FP64 kernel.
// Arithmetic is unchanged: every element is the reference chain  d = A[i,1]*X[1,j]; d = fma(A[i,n], X[n,j], d)  in ascending
// n (src/gemm.jl:86,165), each half of an fma.rn.f32x2 is a correctly rounded fma, accumulators start at -0.0 (or C), the K
// tail runs a bounded loop (TMA's zero padding is never multiplied) -> BIT-IDENTICAL to the oracle.
//
// Shared memory per stage (BK = 32): A box {128 m, 32 k}, no swizzle: sA[k][m], a quarter-warp's LDS.128 covers 128 contiguous
// bytes; X box {32 k, 256 n} with the hardware 128B swizzle: sX[n][k] in 128-byte rows, 16-byte chunk c stored at c ^ (n & 7),
// so the four column lanes of a warp (n = .. + ty) read four different chunks -> conflict-free without padding.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_dmma_tma.cuh"
#include "gemm_simt_f32x2.cuh"

namespace jb {

template <int STAGES_>
struct F32TmaCfg {
    static constexpr int BM = 128, BN = 256, BK = 32, STAGES = STAGES_;
    static constexpr int CONSUMER_WARPS = 8, THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr int A_BYTES = BM * BK * 4, X_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + X_BYTES;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * sizeof(uint64_t) + STAGES * sizeof(int) + 1024;
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_f32_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX, float* D, int M, int N, int K,
                    int64_t ldd, int tiles_m, int tiles_n, int group_m, int* __restrict__ tile_ctr, const float* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    volatile int* stage_tile = reinterpret_cast<volatile int*>(empty + STAGES);

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
            int tile = blockIdx.x;
            while (tile < num_tiles) {
                const int next = tile_ctr ? atomicAdd(tile_ctr, 1) + (int)gridDim.x : tile + (int)gridDim.x;
                int tm, tn;
                raster(tile, tiles_m, tiles_n, group_m, tm, tn);
                for (int kt = 0; kt < KT; ++kt) {
                    mbar_wait(&empty[s], phase ^ 1);
                    stage_tile[s] = tile;
                    mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    unsigned char* st = tiles + (size_t)s * Cfg::STAGE_BYTES;
                    tma_load_2d(st, &mapA, &full[s], tm * BM, kt * BK);
                    tma_load_2d(st + Cfg::A_BYTES, &mapX, &full[s], kt * BK, tn * BN);
                    if (++s == STAGES) { s = 0; phase ^= 1; }
                }
                tile = next;
            }
            mbar_wait(&empty[s], phase ^ 1);
            stage_tile[s] = -1;  // end marker
            mbar_arrive(&full[s]);
            if (tile_ctr && atomicAdd(tile_ctr + 1, 1) == (int)gridDim.x - 1) {
                tile_ctr[0] = 0;
                tile_ctr[1] = 0;
            }
        }
        return;
    }

    // ===================== consumers: 2 x 4 warps of 64 x 64, thread tile 8 rows x 16 columns =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
    constexpr int NJ = 16, NP = 4;  // columns per thread, row pairs per thread
    const int wm = warp & 1, wn = warp >> 1;
    const int tx = lane & 7, ty = lane >> 3;
    // thread rows: wm*64 + i*32 + tx*4 + v (i < 2, v < 4);  thread columns: wn*64 + j*4 + ty (j < 16)
    const int row_base = wm * 64 + tx * 4;
    const int col_base = wn * 64 + ty;
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t offA = (uint32_t)(row_base * 4);                        // + k*512 (+128 for the second row group)
    const uint32_t offX = (uint32_t)(Cfg::A_BYTES + col_base * 128);       // + j*512 + ((c ^ key) << 4) + (k & 3)*4
    const bool d_vec_ok = (ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);

    int s = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&full[s], phase);
        const int tile = stage_tile[s];
        if (tile < 0) break;
        int tm, tn;
        raster(tile, tiles_m, tiles_n, group_m, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        const bool interior = (m0 + BM <= M) && (n0 + BN <= N);

        uint64_t acc[NJ][NP];  // [column j][row pair i*2 + h: rows i*32 + tx*4 + 2h, +1]
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int gn = n0 + col_base + j * 4;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if constexpr (ACC) {
                    const int gm = m0 + row_base + (p >> 1) * 32 + (p & 1) * 2;
                    const float lo = (gm < M && gn < N) ? Cin[(size_t)gn * ldc + gm] : 0.f;
                    const float hi = (gm + 1 < M && gn < N) ? Cin[(size_t)gn * ldc + gm + 1] : 0.f;
                    acc[j][p] = pack_f32x2(lo, hi);
                } else {
                    acc[j][p] = 0x8000000080000000ull;  // (-0.0f, -0.0f): fma(a, b, -0) == a*b exactly
                }
            }
        }

        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(&full[s], phase);
            const uint32_t st = tiles_u32 + (uint32_t)s * Cfg::STAGE_BYTES;
            const uint32_t pA = st + offA, pX = st + offX;
            const int kmax = min(BK, K - kt * BK);
            auto step = [&](const ulonglong2 (&a)[2], const float (&bs)[NJ]) {
                // (b, b) pairs fold into FFMA2's scalar-broadcast operand form; the A PAIR is the operand held in the reuse
                // cache (16 consecutive FFMA2 share it): 3 fresh registers per FFMA2 instead of 4
                uint64_t bb[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bb[j] = pack_f32x2(bs[j], bs[j]);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const uint64_t ap = (p & 1) ? a[p >> 1].y : a[p >> 1].x;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) ffma2(acc[j][p], ap, bb[j]);
                }
            };
            if (kmax == BK) {
#pragma unroll
                for (int c = 0; c < BK / 4; ++c) {  // one 16-byte chunk of k per swizzle step
                    const uint32_t o0 = (uint32_t)((c ^ ty) << 4), o1 = o0 ^ 64u;  // chunk ^ (n & 7), n & 7 = (j & 1)*4 + ty
#pragma unroll
                    for (int half = 0; half < 2; ++half) {  // k = 4c + 2*half, +1
                        float2 b[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
                            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(b[j].x), "=f"(b[j].y) : "r"(pX + j * 512 + ((j & 1) ? o1 : o0) + half * 8));
#pragma unroll
                        for (int kv = 0; kv < 2; ++kv) {
                            const int k = 4 * c + 2 * half + kv;
                            ulonglong2 a[2];
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];\n" : "=l"(a[i].x), "=l"(a[i].y) : "r"(pA + k * 512 + i * 128));
                            float bs[NJ];
#pragma unroll
                            for (int j = 0; j < NJ; ++j) bs[j] = kv ? b[j].y : b[j].x;
                            step(a, bs);
                        }
                    }
                }
            } else {  // K tail: bounded loop, TMA's zero padding is never multiplied
                for (int k = 0; k < kmax; ++k) {
                    ulonglong2 a[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];\n" : "=l"(a[i].x), "=l"(a[i].y) : "r"(pA + k * 512 + i * 128));
                    const uint32_t o0 = (uint32_t)((((k >> 2) ^ ty) << 4) + (k & 3) * 4), o1 = o0 ^ 64u;
                    float bs[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(bs[j]) : "r"(pX + j * 512 + ((j & 1) ? o1 : o0)));
                    step(a, bs);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == STAGES) { s = 0; phase ^= 1; }
        }

        // ---- store (src/gemm.jl:3-11: plain overwrite, column-major) ----
        if (interior && d_vec_ok) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float* dcol = D + (size_t)(n0 + col_base + j * 4) * ldd + m0 + row_base;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    ulonglong2 o;
                    o.x = acc[j][2 * i];
                    o.y = acc[j][2 * i + 1];
                    *reinterpret_cast<ulonglong2*>(dcol + i * 32) = o;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int gn = n0 + col_base + j * 4;
                if (gn >= N) continue;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int gm = m0 + row_base + (p >> 1) * 32 + (p & 1) * 2;
                    float lo, hi;
                    unpack_f32x2(acc[j][p], lo, hi);
                    if (gm < M) D[(size_t)gn * ldd + gm] = lo;
                    if (gm + 1 < M) D[(size_t)gn * ldd + gm + 1] = hi;
                }
            }
        }
    }
}

}  // namespace jb

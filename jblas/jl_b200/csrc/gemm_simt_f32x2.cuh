// gemm_simt_f32x2.cuh -- exact Float32 SIMT GEMM on Blackwell's packed FFMA2 (PTX fma.rn.f32x2), D = A*X or D += A*X.
//
// Why a second Float32 kernel: the scalar-FFMA kernel (gemm_simt.cuh) is bound by warp-scheduler issue slots and
// register-file bank conflicts, not by the FMA pipe (ncu: fma pipe 59 %, dispatch stalls; profiles/).  sm_100 adds
// fma.rn.f32x2 (SASS FFMA2): ONE instruction performs TWO independent IEEE fused multiply-adds on 64-bit register
// pairs.  Half the issue slots for the same flops, and even/odd register pairs are bank-balanced by construction.
// Each half is a correctly rounded fma, so every output element is still the reference chain
//     d = A[i,1]*X[1,j];  d = fma(A[i,n], X[n,j], d)   (n ascending; src/gemm.jl:86,165)
// and the result stays BIT-IDENTICAL to the oracle.
//
// Pairing: an accumulator pair is two consecutive ROWS (m, m+1) of one column.  A pairs come straight out of the
// 16-byte shared-memory loads (m is contiguous in sA[k][m]); the X scalar of the column is duplicated into a
// register pair once per k and reused by the four row-pairs of that column.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tile_loader.cuh"

namespace jb {

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(uint64_t& d, uint64_t a, uint64_t b)
{
    // volatile: keeps the issue order written below (A pair reused over 8 consecutive FFMA2); ptxas otherwise regroups
    // the sequence around the broadcast scalar, which costs one more fresh register read per instruction
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(d) : "l"(a), "l"(b));
}

// Small-problem tiles.  The chain forbids split-K, so the only way to put a 256^3..1024^3 product on all 148 SMs is
// more, smaller tiles: thread tile (4*RI) x NJ instead of 8 x 8, warp tile (32*RI) x (4*NJ), WM x WN warps per CTA.
template <int WM, int WN, int RI_, int NJ_, int BK_, int STAGES_, int MINB>
struct F32x2Cfg {
    static constexpr int WARPS_M = WM, RI = RI_, NJ = NJ_;
    static constexpr int BM = WM * 32 * RI_, BN = WN * 4 * NJ_, BK = BK_, STAGES = STAGES_;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int VEC = 4;
    static constexpr int LDA = BM;
    static constexpr int LDB = BK + VEC;
    static constexpr int STAGE_ELEMS = BK * LDA + BN * LDB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(float);
    static constexpr int MIN_BLOCKS = MINB;
};

// Tiling contract (SimtCfg<float,...> has RI = 2, NJ = 8: warp 64 x 32, thread 8 x 8):
//   thread rows  wm*32*RI + i*32 + tx*4 + v   (i < RI, v < 4)      thread cols  wn*4*NJ + j*4 + ty   (j < NJ)
template <typename Cfg, bool ALIGNED, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_BLOCKS)
gemm_simt_f32x2_kernel(float* D, const float* __restrict__ A, const float* __restrict__ X, int M, int N, int K,
                       int64_t ldd, int64_t lda, int64_t ldx, int tiles_m, int tiles_n, int group_m, const float* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
    constexpr int LDA = Cfg::LDA, LDB = Cfg::LDB, THREADS = Cfg::THREADS;
    static_assert(Cfg::VEC == 4, "Float32 only");
    // programmatic dependent launch (capi.cu: launch_pdl): no-ops for plain launches; the wait precedes every global access
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);

    int tm, tn;
    raster(blockIdx.x, tiles_m, tiles_n, group_m, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int RI = Cfg::RI, NJ = Cfg::NJ, NP = 2 * Cfg::RI;
    const int wm = warp % Cfg::WARPS_M, wn = warp / Cfg::WARPS_M;
    const int tx = lane & 7, ty = lane >> 3;
    const int row_base = wm * 32 * RI + tx * 4;
    const int col_base = wn * 4 * NJ + ty;

    uint64_t acc[NJ][NP];  // [column j][row pair: rows (i*32 + 2h, +1) for pair index i*2 + h]
    const bool d_vec_ok = (ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    const bool interior = (m0 + BM <= M) && (n0 + BN <= N);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int gn = n0 + col_base + j * 4;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if constexpr (ACC) {
                const int gm = m0 + row_base + (p >> 1) * 32 + (p & 1) * 2;
                float lo = (gm < M && gn < N) ? Cin[(size_t)gn * ldc + gm] : 0.f;
                float hi = (gm + 1 < M && gn < N) ? Cin[(size_t)gn * ldc + gm + 1] : 0.f;
                acc[j][p] = pack_f32x2(lo, hi);
            } else {
                acc[j][p] = 0x8000000080000000ull;  // (-0.0f, -0.0f): fma(a, b, -0) == a*b exactly
            }
        }
    }

    const int KT = (K + BK - 1) / BK;
    auto stageA = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS; };
    auto stageB = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS + BK * LDA; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT)
            load_stage<float, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(s), stageB(s), A, X, lda, ldx, M, N, K, m0, n0,
                                                                       s * BK, tid);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT)
                load_stage<float, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(nk % STAGES), stageB(nk % STAGES), A, X,
                                                                           lda, ldx, M, N, K, m0, n0, nk * BK, tid);
            cp_async_commit();
        }
        const float* sA = stageA(kt % STAGES) + row_base;
        const float* sB = stageB(kt % STAGES) + col_base * LDB;
        const int kmax = min(BK, K - kt * BK);
        if (kmax == BK) {
#pragma unroll
            for (int kc = 0; kc < BK; kc += 2) {
                float2 b[NJ];  // X[k, k+1] of the thread's 8 columns
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const float2*>(sB + j * 4 * LDB + kc);
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {
                    ulonglong2 a[RI];  // rows (0,1),(2,3) and (32,33),(34,35) of the thread, as packed pairs
#pragma unroll
                    for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const ulonglong2*>(sA + (kc + kv) * LDA + i * 32);
                    // (b, b) pairs fold into FFMA2's scalar-broadcast operand form (SASS `Rb.F32`), no MOVs are emitted.
                    // Loop order: the A PAIR is the operand held in the reuse cache (8 consecutive FFMA2 share it), the
                    // fresh operands per FFMA2 are one scalar + one accumulator pair = 3 registers instead of 4.
                    uint64_t bb[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const float bs = kv ? b[j].y : b[j].x;
                        bb[j] = pack_f32x2(bs, bs);
                    }
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const uint64_t ap = (p & 1) ? a[p >> 1].y : a[p >> 1].x;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) ffma2(acc[j][p], ap, bb[j]);
                    }
                }
            }
        } else {  // K tail: bounded loop, no padded multiplies
            for (int k = 0; k < kmax; ++k) {
                ulonglong2 a[RI];
#pragma unroll
                for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const ulonglong2*>(sA + k * LDA + i * 32);
                uint64_t bb[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float bs = sB[j * 4 * LDB + k];
                    bb[j] = pack_f32x2(bs, bs);
                }
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const uint64_t ap = (p & 1) ? a[p >> 1].y : a[p >> 1].x;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) ffma2(acc[j][p], ap, bb[j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- store (src/gemm.jl:3-11: plain overwrite, column-major) ----
    if (interior && d_vec_ok) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float* dcol = D + (size_t)(n0 + col_base + j * 4) * ldd + m0 + row_base;
#pragma unroll
            for (int i = 0; i < RI; ++i) {
#ifdef F32X2_STORE64
                *reinterpret_cast<uint64_t*>(dcol + i * 32) = acc[j][2 * i];
                *reinterpret_cast<uint64_t*>(dcol + i * 32 + 2) = acc[j][2 * i + 1];
#else
                ulonglong2 o;
                o.x = acc[j][2 * i];
                o.y = acc[j][2 * i + 1];
                *reinterpret_cast<ulonglong2*>(dcol + i * 32) = o;
#endif
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

}  // namespace jb

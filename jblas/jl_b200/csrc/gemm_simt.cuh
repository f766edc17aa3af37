// gemm_simt.cuh -- exact SIMT GEMM (DFMA for double, FFMA for float), D = A*X or D += A*X.
//
// Replaces the reference's register-tile micro-kernels (src/kernels.jl:212-275 kernel!/initkernel!, and the
// tile body of jmul!, src/gemm.jl:313-336).  The reference's 40x5 AVX-512 register tile becomes an 8x8
// per-thread register tile; a warp owns 64x32, a CTA WM x WN warps.
//
// BIT-EXACTNESS CONTRACT.  Every output element is one chain, sequential and ascending in k, of fused
// multiply-adds -- exactly the reference's  d = A[i,1]*X[1,j]; d = fma(A[i,n], X[n,j], d)  (src/gemm.jl:86,165).
//   * accumulators start at -0.0: fma(a, b, -0.0) == a*b for every a, b (including signed zeros, Inf, NaN), so
//     the first step reproduces the reference's plain rounded product (initialize_block) without a special case;
//   * ACC (kernel! semantics, src/kernels.jl:226) starts from an existing matrix instead: Cin == D is the reference's
//     D += A*X, any other Cin is the planned fused form D = A*X + C (src/memory_management.jl:72-76);
//   * no split-K, no tree reduction; the K tail runs a bounded loop instead of multiplying zero padding (so a
//     -0.0 result is not turned into +0.0 by fma(0, 0, -0.0)).
// The result therefore matches oracle/oracle_gemm.c bit for bit on finite inputs.
#pragma once
#include "common.cuh"
#include "tile_loader.cuh"

namespace jb {

template <typename T, int WM, int WN, int BK_, int STAGES_, int MINB>
struct SimtCfg {
    static constexpr int BM = WM * 64, BN = WN * 32, BK = BK_, STAGES = STAGES_;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int VEC = 16 / (int)sizeof(T);  // elements per 16-byte shared-memory load
    static constexpr int NI = 8 / VEC;               // 16-byte A loads per thread per k
    static constexpr int WARPS_M = WM, RI = NI, NJ = 8;  // (gemm_simt_f32x2.cuh reads its thread tile from these)
    static constexpr int LDA = BM;                   // sA[k][m]: reads are contiguous in m -> conflict-free
    static constexpr int LDB = BK + VEC;             // sB[n][k]: +16 B pitch puts the 4 n-lanes in distinct banks
    static constexpr int STAGE_ELEMS = BK * LDA + BN * LDB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(T);
    static constexpr int MIN_BLOCKS = MINB;
};

template <typename T, int VEC>
struct alignas(16) Vec16 {
    T v[VEC];
};

template <typename T, typename Cfg, bool ALIGNED, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_BLOCKS)
gemm_simt_kernel(T* D, const T* __restrict__ A, const T* __restrict__ X, int M, int N, int K, int64_t ldd,
                 int64_t lda, int64_t ldx, int tiles_m, int tiles_n, int group_m, const T* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, VEC = Cfg::VEC, NI = Cfg::NI;
    constexpr int LDA = Cfg::LDA, LDB = Cfg::LDB, THREADS = Cfg::THREADS;
    using V = Vec16<T, VEC>;
    // programmatic dependent launch (capi.cu: launch_pdl): no-ops for plain launches; the wait precedes every global access
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);

    int tm, tn;
    raster(blockIdx.x, tiles_m, tiles_n, group_m, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (BM / 64), wn = warp / (BM / 64);
    const int tx = lane & 7, ty = lane >> 3;  // 8 lanes along m, 4 along n
    // thread rows : m0 + wm*64 + i*(8*VEC) + tx*VEC + v   (i < NI, v < VEC)
    // thread cols : n0 + wn*32 + j*4 + ty                 (j < 8)
    const int row_base = wm * 64 + tx * VEC;
    const int col_base = wn * 32 + ty;

    T acc[8][8];  // [j][i*VEC+v]
    const bool d_vec_ok = (ldd % VEC == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    const bool interior = (m0 + BM <= M) && (n0 + BN <= N);
    if constexpr (ACC) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gn = n0 + col_base + j * 4;
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    int gm = m0 + row_base + i * 8 * VEC + v;
                    acc[j][i * VEC + v] = (gm < M && gn < N) ? Cin[(size_t)gn * ldc + gm] : T(0);
                }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[j][r] = T(-0.0);
    }

    const int KT = (K + BK - 1) / BK;
    auto stageA = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS; };
    auto stageB = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS + BK * LDA; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT)
            load_stage<T, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(s), stageB(s), A, X, lda, ldx, M, N, K, m0, n0,
                                                                   s * BK, tid);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT)
                load_stage<T, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(nk % STAGES), stageB(nk % STAGES), A, X, lda,
                                                                       ldx, M, N, K, m0, n0, nk * BK, tid);
            cp_async_commit();
        }
        const T* sA = stageA(kt % STAGES) + row_base;
        const T* sB = stageB(kt % STAGES) + col_base * LDB;
        const int kmax = min(BK, K - kt * BK);
        if (kmax == BK) {
#pragma unroll
            for (int kc = 0; kc < BK; kc += VEC) {
                V b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const V*>(sB + j * 4 * LDB + kc);
#pragma unroll
                for (int kv = 0; kv < VEC; ++kv) {
                    V a[NI];
#pragma unroll
                    for (int i = 0; i < NI; ++i) a[i] = *reinterpret_cast<const V*>(sA + (kc + kv) * LDA + i * 8 * VEC);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < NI; ++i)
#pragma unroll
                            for (int v = 0; v < VEC; ++v)
                                acc[j][i * VEC + v] = fma_t(a[i].v[v], b[j].v[kv], acc[j][i * VEC + v]);
                }
            }
        } else {  // K tail: bounded loop, no padded multiplies
            for (int k = 0; k < kmax; ++k) {
                T a[8], b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = sB[j * 4 * LDB + k];
#pragma unroll
                for (int i = 0; i < NI; ++i)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) a[i * VEC + v] = sA[k * LDA + i * 8 * VEC + v];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[j][r] = fma_t(a[r], b[j], acc[j][r]);
            }
        }
    }
    cp_async_wait<0>();

    // ---- store (src/gemm.jl:3-11: plain overwrite of the tile, column-major, leading dimension ldd) ----
    if (interior && d_vec_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            T* dcol = D + (size_t)(n0 + col_base + j * 4) * ldd + m0 + row_base;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                V o;
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = acc[j][i * VEC + v];
                *reinterpret_cast<V*>(dcol + i * 8 * VEC) = o;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gn = n0 + col_base + j * 4;
            if (gn >= N) continue;
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    int gm = m0 + row_base + i * 8 * VEC + v;
                    if (gm < M) D[(size_t)gn * ldd + gm] = acc[j][i * VEC + v];
                }
        }
    }
}

}  // namespace jb

// gemm_dmma.cuh -- FP64 tensor-core GEMM (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), D = A*X or D += A*X.
//
// The tensor-pipe alternative to gemm_simt.cuh for double precision (north_star: "FP64 runs on the DMMA
// tensor-core path ... keeping whichever ncu shows is faster per shape").  FP64 has no tcgen05 kind, so the
// legacy warp-level mma.sync is the only tensor path on sm_100a; nvcc lowers m16n8k16.f64 to 8 x DMMA.8x8x4
// anyway, so m8n8k4 is issued directly to keep the k order explicit.
//
// Numerics: every element is still accumulated in ascending k (4 at a time), starting from -0.0 (or the old
// D when ACC).  Whether the 4 products inside one DMMA are chained with single rounding like 4 sequential
// FMAs is a hardware property; tests/test_gemm_gpu.py measures it against the oracle.  The contract for this
// kernel is the reference tolerance 2*K*eps*(|A||X|); the SIMT kernel is the bit-exact one.
//
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64): g = lane/4, t = lane%4
//   A (8x4, row)  a0 = A[g][t]          B (4x8, col)  b0 = B[t][g]          C (8x8)  c0,c1 = C[g][2t], C[g][2t+1]
// Shared-memory pitches make both fragment loads conflict-free per half-warp (16 lanes x 8 B = 32 banks):
//   sA[k][m], pitch LDA = BM+4  : offset t*LDA + g  -> (4t + g) mod 16 distinct
//   sB[n][k], pitch LDB = BK+4  : offset g*LDB + t  -> (4g + t) mod 16 distinct
#pragma once
#include "common.cuh"
#include "tile_loader.cuh"

namespace jb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// same, accumulator written from a separate start value (first k-step of a chain: no register initialisation pass)
__device__ __forceinline__ void dmma884_from(double& d0, double& d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int WM, int WN, int BK_, int STAGES_>
struct DmmaCfg {
    static constexpr int BM = WM * 64, BN = WN * 32, BK = BK_, STAGES = STAGES_;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int LDA = BM + 4;
    static constexpr int LDB = BK + 4;
    static constexpr int STAGE_ELEMS = BK * LDA + BN * LDB;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
};

template <typename Cfg, bool ALIGNED, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_dmma_kernel(double* D, const double* __restrict__ A, const double* __restrict__ X, int M, int N, int K,
                 int64_t ldd, int64_t lda, int64_t ldx, int tiles_m, int tiles_n, int group_m, const double* Cin, int64_t ldc)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
    constexpr int LDA = Cfg::LDA, LDB = Cfg::LDB, THREADS = Cfg::THREADS;
    // programmatic dependent launch (capi.cu: launch_pdl): no-ops for plain launches; the wait precedes every global access
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);

    int tm, tn;
    raster(blockIdx.x, tiles_m, tiles_n, group_m, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (BM / 64), wn = warp / (BM / 64);
    const int g = lane >> 2, t = lane & 3;

    double acc[8][4][2];  // [m-tile][n-tile][c0,c1]
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if constexpr (ACC) {
                    int gm = m0 + wm * 64 + mi * 8 + g, gn = n0 + wn * 32 + ni * 8 + 2 * t + c;
                    acc[mi][ni][c] = (gm < M && gn < N) ? Cin[(size_t)gn * ldc + gm] : 0.0;
                } else {
                    acc[mi][ni][c] = -0.0;
                }
            }

    const int KT = (K + BK - 1) / BK;
    const bool k_tail = (K % BK) != 0;
    auto stageA = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS; };
    auto stageB = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS + BK * LDA; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT)
            load_stage<double, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(s), stageB(s), A, X, lda, ldx, M, N, K, m0,
                                                                        n0, s * BK, tid);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT)
                load_stage<double, BM, BN, BK, LDA, LDB, THREADS, ALIGNED>(stageA(nk % STAGES), stageB(nk % STAGES), A, X,
                                                                            lda, ldx, M, N, K, m0, n0, nk * BK, tid);
            cp_async_commit();
        }
        // Zero padding beyond K (cp.async zero-fill) is multiplied here; the X side of padded k is replaced by -0.0
        // below so that the padded product is -0.0 and leaves every accumulator (including -0.0) unchanged.
        const double* sA = stageA(kt % STAGES) + t * LDA + wm * 64 + g;
        const double* sB = stageB(kt % STAGES) + (wn * 32 + g) * LDB + t;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = sA[k4 * LDA + mi * 8];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = sB[ni * 8 * LDB + k4];
            if (k_tail && kt == KT - 1 && (kt * BK + k4 + t) >= K) {  // padded k: make the product -0.0 (c + -0.0 == c)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = -0.0;
            }
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int gn = n0 + wn * 32 + ni * 8 + 2 * t + c;
            if (gn >= N) continue;
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                int gm = m0 + wm * 64 + mi * 8 + g;
                if (gm < M) D[(size_t)gn * ldd + gm] = acc[mi][ni][c];
            }
        }
}

}  // namespace jb

/*
 * oracle_gemm.c -- CPU ORACLE for the jBLAS.jl `jmul!` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's shared object.  The product (jblas/jl_b200, libjblas_b200.so) never does.
 *
 * PARITY UNPINNED: the reference (JuliaBLAS/jBLAS.jl) ships no golden vectors, no @test, and
 * cannot be executed here (no Julia; its SIMD dependencies SIMDPirates 0.1.0 /
 * VectorizationBase 0.1.0 are un-vendored path-dev checkouts, Manifest.toml:79-83,120-124).
 * This file is therefore a RESTATEMENT of the reference's arithmetic, cross-checked against
 *   (a) oracle/structural_jmul.py, a line-by-line emulation of the code `jmul!` generates
 *       (tile loops, pointer arithmetic, k chunking) evaluated in exact rational arithmetic, and
 *   (b) numpy/OpenBLAS A@X within 2*K*eps*(|A||X|)  (the author's own informal oracle,
 *       test/runtests.jl:139-140).
 *
 * Arithmetic restated (all citations relative to /root/reference):
 *   first term     d  = A[i,1] * X[1,j]            src/gemm.jl:76-89   (initialize_block: `*`)
 *                                                   src/kernels.jl:260  (initkernel!: evmul)
 *   accumulation   d  = fma(A[i,n], X[n,j], d)      src/gemm.jl:154-168 (fma_increment_block)
 *                                                   src/kernels.jl:231,266 (vmuladd)
 *   k order        n = 2..N strictly ascending      src/gemm.jl:299-304,319-333
 *   store          D[i,j] = d  (overwrite)          src/gemm.jl:3-11    (store_block)
 *   accumulate=1   d starts from D[i,j], n = 1..N   src/kernels.jl:225-236 (kernel!)
 * Every element's chain is independent of the register tile, so any loop order that keeps each
 * chain sequential in k with a fused multiply-add is bit-identical to the reference interior.
 * Unlike the reference (src/gemm.jl:266-267,313: remainder rows/cols are never computed) the
 * oracle computes every element -- BASELINE.json's ragged config requires it.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IB 32 /* rows per register block (4 zmm of doubles)  */
#define JB 4  /* columns per register block                  */

/* ------------------------------------------------------------------------------------------ */
/* double                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* Generic (any ib x jb) block: scalar chain, used on the ragged border. */
static void edge_f64(double *D, const double *A, const double *X, int64_t K, int64_t ldd, int64_t lda, int64_t ldx,
                     int ib, int jb, int accumulate)
{
    for (int j = 0; j < jb; ++j)
        for (int i = 0; i < ib; ++i) {
            double d;
            int64_t k = 0;
            if (accumulate) d = D[i + j * ldd];          /* kernel!: load D first (src/kernels.jl:226) */
            else { d = A[i] * X[j * ldx]; k = 1; }       /* plain rounded product for k = 1 (src/gemm.jl:86) */
            for (; k < K; ++k) d = fma(A[i + k * lda], X[k + j * ldx], d);
            D[i + j * ldd] = d;
        }
}

/* Full IB x JB block: same chains, i vectorised by the compiler (lanes are independent elements, so the lane
 * width cannot change any value).  Instantiated per ISA with #pragma GCC target and chosen at run time, because
 * the prebuilt .so travels to a GPU box whose host CPU may differ from the build container's. */
#define DEFINE_BLOCK(NAME, T, FMA)                                                                                 \
    static void NAME(T *restrict D, const T *restrict A, const T *restrict X, int64_t K, int64_t ldd, int64_t lda, \
                     int64_t ldx, int accumulate)                                                                  \
    {                                                                                                              \
        T acc[JB][IB];                                                                                             \
        for (int j = 0; j < JB; ++j) {                                                                             \
            if (accumulate) { /* kernel!: load D first (src/kernels.jl:226) */                                     \
                for (int i = 0; i < IB; ++i) acc[j][i] = D[i + j * ldd];                                           \
            } else { /* plain rounded product for k = 1 (src/gemm.jl:86) */                                        \
                T x = X[j * ldx];                                                                                  \
                for (int i = 0; i < IB; ++i) acc[j][i] = A[i] * x;                                                 \
            }                                                                                                      \
        }                                                                                                          \
        for (int64_t k = accumulate ? 0 : 1; k < K; ++k) { /* ascending k, fused multiply-add (src/gemm.jl:165) */ \
            const T *a = A + k * lda;                                                                              \
            for (int j = 0; j < JB; ++j) {                                                                         \
                T x = X[k + j * ldx];                                                                              \
                for (int i = 0; i < IB; ++i) acc[j][i] = FMA(a[i], x, acc[j][i]);                                  \
            }                                                                                                      \
        }                                                                                                          \
        for (int j = 0; j < JB; ++j)                                                                               \
            for (int i = 0; i < IB; ++i) D[i + j * ldd] = acc[j][i];                                               \
    }

#pragma GCC push_options
#pragma GCC target("avx512f,avx512vl,avx512dq,avx2,fma")
DEFINE_BLOCK(block_f64_avx512, double, __builtin_fma)
DEFINE_BLOCK(block_f32_avx512, float, __builtin_fmaf)
#pragma GCC pop_options
#pragma GCC push_options
#pragma GCC target("avx2,fma")
DEFINE_BLOCK(block_f64_avx2, double, __builtin_fma)
DEFINE_BLOCK(block_f32_avx2, float, __builtin_fmaf)
#pragma GCC pop_options
DEFINE_BLOCK(block_f64_generic, double, fma) /* libm fma: slow but still a true fused op */
DEFINE_BLOCK(block_f32_generic, float, fmaf)

typedef void (*block_f64_fn)(double *restrict, const double *restrict, const double *restrict, int64_t, int64_t,
                             int64_t, int64_t, int);
typedef void (*block_f32_fn)(float *restrict, const float *restrict, const float *restrict, int64_t, int64_t, int64_t,
                             int64_t, int);
static int isa_level(void)
{
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq"))
        return 512;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return 256;
    return 0;
}

static int gemm_f64(double *D, const double *A, const double *X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                    int64_t lda, int64_t ldx, int accumulate)
{
    if (M < 0 || N < 0 || K < 0) return -1;
    if (M == 0 || N == 0) return 0;
    if (ldd < M || lda < M || ldx < (K > 0 ? K : 1)) return -2;
    if (K == 0) { /* empty contraction: jmul! would index X[1,j] out of bounds; we define D = 0 (accumulate: D) */
        if (!accumulate)
            for (int64_t j = 0; j < N; ++j) memset(D + j * ldd, 0, (size_t)M * sizeof(double));
        return 0;
    }
    const int64_t njb = (N + JB - 1) / JB, nib = (M + IB - 1) / IB;
    const int isa = isa_level();
    block_f64_fn block_f64 = isa == 512 ? block_f64_avx512 : (isa == 256 ? block_f64_avx2 : block_f64_generic);
    /* bi outer: the IB x K slab of A stays cache-resident while the column blocks stream past */
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t bi = 0; bi < nib; ++bi)
        for (int64_t bj = 0; bj < njb; ++bj) {
            int64_t j0 = bj * JB, i0 = bi * IB;
            int jb = (int)(N - j0 < JB ? N - j0 : JB), ib = (int)(M - i0 < IB ? M - i0 : IB);
            if (ib == IB && jb == JB)
                block_f64(D + i0 + j0 * ldd, A + i0, X + j0 * ldx, K, ldd, lda, ldx, accumulate);
            else
                edge_f64(D + i0 + j0 * ldd, A + i0, X + j0 * ldx, K, ldd, lda, ldx, ib, jb, accumulate);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* float (the reference is generic in T: sizeof(T) everywhere, src/kernel_structure.jl:76-78)  */
/* ------------------------------------------------------------------------------------------ */
static void edge_f32(float *D, const float *A, const float *X, int64_t K, int64_t ldd, int64_t lda, int64_t ldx, int ib,
                     int jb, int accumulate)
{
    for (int j = 0; j < jb; ++j)
        for (int i = 0; i < ib; ++i) {
            float d;
            int64_t k = 0;
            if (accumulate) d = D[i + j * ldd];
            else { d = A[i] * X[j * ldx]; k = 1; }
            for (; k < K; ++k) d = fmaf(A[i + k * lda], X[k + j * ldx], d);
            D[i + j * ldd] = d;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* exported                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* D = A*X (accumulate=0, jmul!/initkernel!) or D += A*X (accumulate=1, kernel!), column-major. */
int oracle_gemm_f64(double *D, const double *A, const double *X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                    int64_t lda, int64_t ldx, int accumulate)
{
    return gemm_f64(D, A, X, M, K, N, ldd, lda, ldx, accumulate);
}

/* B = |A|*|X| with the same chain -- the (|A||B|)_ij factor of the 2*K*eps*(|A||B|) acceptance bound.
 * A and X are overwritten-free: absolute values are taken into scratch copies supplied by the caller
 * (absA: M*K, absX: K*N doubles, dense). */
int oracle_absgemm_f64(double *B, const double *A, const double *X, int64_t M, int64_t K, int64_t N, int64_t ldb,
                       int64_t lda, int64_t ldx, double *absA, double *absX)
{
    for (int64_t k = 0; k < K; ++k)
        for (int64_t i = 0; i < M; ++i) absA[i + k * M] = fabs(A[i + k * lda]);
    for (int64_t j = 0; j < N; ++j)
        for (int64_t k = 0; k < K; ++k) absX[k + j * K] = fabs(X[k + j * ldx]);
    return gemm_f64(B, absA, absX, M, K, N, ldb, M > 0 ? M : 1, K > 0 ? K : 1, 0);
}

int oracle_gemm_f32(float *D, const float *A, const float *X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                    int64_t lda, int64_t ldx, int accumulate)
{
    if (M < 0 || N < 0 || K < 0) return -1;
    if (M == 0 || N == 0) return 0;
    if (ldd < M || lda < M || ldx < (K > 0 ? K : 1)) return -2;
    if (K == 0) {
        if (!accumulate)
            for (int64_t j = 0; j < N; ++j) memset(D + j * ldd, 0, (size_t)M * sizeof(float));
        return 0;
    }
    const int64_t njb = (N + JB - 1) / JB, nib = (M + IB - 1) / IB;
    const int isa = isa_level();
    block_f32_fn block_f32 = isa == 512 ? block_f32_avx512 : (isa == 256 ? block_f32_avx2 : block_f32_generic);
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t bi = 0; bi < nib; ++bi)
        for (int64_t bj = 0; bj < njb; ++bj) {
            int64_t j0 = bj * JB, i0 = bi * IB;
            int jb = (int)(N - j0 < JB ? N - j0 : JB), ib = (int)(M - i0 < IB ? M - i0 : IB);
            if (ib == IB && jb == JB)
                block_f32(D + i0 + j0 * ldd, A + i0, X + j0 * ldx, K, ldd, lda, ldx, accumulate);
            else
                edge_f32(D + i0 + j0 * ldd, A + i0, X + j0 * ldx, K, ldd, lda, ldx, ib, jb, accumulate);
        }
    return 0;
}

/* Sampled oracle for sizes where the full product is out of reach on a CPU (32768^3): computes the chain
 * only for the listed (row, col) pairs.  Legitimate because each element's chain is tile-independent. */
int oracle_gemm_f64_sampled(double *out, const double *A, const double *X, int64_t K, int64_t lda, int64_t ldx,
                            const int64_t *rows, const int64_t *cols, int64_t nsamples)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsamples; ++s) {
        const double *a = A + rows[s];
        const double *x = X + cols[s] * ldx;
        double d = a[0] * x[0];
        for (int64_t k = 1; k < K; ++k) d = fma(a[k * lda], x[k], d);
        out[s] = d;
    }
    return 0;
}

int oracle_gemm_f32_sampled(float *out, const float *A, const float *X, int64_t K, int64_t lda, int64_t ldx,
                            const int64_t *rows, const int64_t *cols, int64_t nsamples)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsamples; ++s) {
        const float *a = A + rows[s];
        const float *x = X + cols[s] * ldx;
        float d = a[0] * x[0];
        for (int64_t k = 1; k < K; ++k) d = fmaf(a[k * lda], x[k], d);
        out[s] = d;
    }
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

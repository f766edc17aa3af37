/*
 * jmul_baseline.c -- CPU BASELINE: a faithful restatement of the LOOP NEST that jBLAS.jl's
 * `@generated jmul!` emits (not just its arithmetic).  TEST / BENCH INFRASTRUCTURE ONLY: loaded by
 * tests/ and by bench.py's cpu_baseline / --impl reference leg, never by the product.
 *
 * "restatement (Julia unavailable)": Julia is not installed in this environment, so this file is the
 * timed stand-in for the reference.  PARITY UNPINNED (see oracle_gemm.c header).
 *
 * What is restated (citations relative to /root/reference):
 *   tile shape        pick_kernel_size(T)                          src/kernel_structure.jl:76-99
 *                     REGISTER_SIZE/REGISTER_COUNT from the CPU    deps/build.jl:3-34
 *                       AVX-512: 64 B x 32 regs -> F64 40x5, F32 80x5
 *                       AVX2   : 32 B x 16 regs -> F64 12x4, F32 24x4
 *   loop nest         for cc in col tiles (outer), rc in row tiles src/gemm.jl:313
 *   init              D_r_c = A_r * X_c   (k = 1)                  src/gemm.jl:71-91, :318
 *   k loop            n = 2..Nr, then Nd chunks of cache_length    src/gemm.jl:299-304, :319-333
 *   fma block         row_loads loads of A[:,n], cols broadcasts,
 *                     row_loads*cols fma                           src/gemm.jl:149-170
 *   prefetch          A at +Aprefetch_freq columns every k, X once
 *                     per cache line, D tile before store          src/gemm.jl:314-335,
 *                                                                  src/memory_management.jl:181-277
 *   store             vstore! rows x cols, ld = M                  src/gemm.jl:3-11, :334
 *   edges             row_remainder/col_remainder are computed and NEVER used: rows M-M%rows+1:M and
 *                     cols P-P%cols+1:P of D are left untouched    src/gemm.jl:266-267,313,340-345
 *   threading         none (single thread)                         grep Threads src/ -> nothing
 * `fill_edges=1` additionally computes the skipped remainder with the scalar chain (NOT in the
 * reference) so the output can be compared in full; it is excluded from any timing claim.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* pick_kernel_size, restated: src/kernel_structure.jl:76-99 */
void jmul_pick_kernel_size(int t_size, int register_size, int register_count, int *vector_length, int *rows,
                           int *cols)
{
    int epr = register_size / t_size;
    int cache_line = epr;
    int max_total = epr * register_count;
    int num_cache_lines = (max_total + cache_line - 1) / cache_line;
    int prev_rows = 0, prev_cols = 0;
    double prev_ratio = -INFINITY;
    for (int a_loads = 1; a_loads <= num_cache_lines; ++a_loads) {
        int num_rows = a_loads * epr;
        int num_cols = (register_count - a_loads - 1) / a_loads;
        int length_D = num_rows * num_cols;
        int num_loads = num_cols + a_loads;
        double next_ratio = (double)length_D / (double)num_loads;
        if (next_ratio < prev_ratio) break;
        prev_ratio = next_ratio;
        prev_rows = num_rows;
        prev_cols = num_cols;
    }
    *vector_length = epr;
    *rows = prev_rows;
    *cols = prev_cols;
}

/* k-range split of src/gemm.jl:299-304 (pure scheduling; numerics unaffected) */
static void split_k(int64_t K, int cache_length, int64_t *Nr, int64_t *Nd)
{
    int64_t nd = (K - 1) / cache_length, nr = (K - 1) % cache_length;
    nr += 1;
    if (nr < cache_length / 3 && nd > 0) {
        nr += cache_length;
        nd -= 1;
    }
    *Nr = nr;
    *Nd = nd;
}

#define CACHELINE_SIZE 64

#if defined(__x86_64__)
#include <immintrin.h>
#endif

/* One instantiation of the generated jmul! body for (T, VL, ROW_LOADS, COLS) under a target ISA. */
#define DEFINE_JMUL_INTR(NAME, T, VEC, VL, ROW_LOADS, COLS, TARGET, LOADU, STOREU, SET1, MUL, FMADD)              \
    __attribute__((target(TARGET))) static void NAME(T *pD, const T *pA, const T *pX, int64_t M, int64_t N,      \
                                                     int64_t P, int64_t cc_lo, int64_t cc_hi, int Apf)          \
    { /* jBLAS dims: D is MxP, A is MxN, X is NxP; N is the contraction dim (src/gemm.jl:244) */                 \
        enum { ROWS = VL * ROW_LOADS, CL = CACHELINE_SIZE / sizeof(T) };                                          \
        (void)P;                                                                                                  \
        const int64_t row_chunks = M / ROWS;                                                                      \
        int64_t Nr, Nd;                                                                                           \
        split_k(N, CL, &Nr, &Nd);                                                                                 \
        for (int64_t cc = cc_lo; cc < cc_hi; ++cc)                                                                \
            for (int64_t rc = 0; rc < row_chunks; ++rc) {                                                         \
                VEC acc[COLS][ROW_LOADS], a[ROW_LOADS];                                                           \
                const T *Ablk = pA + rc * ROWS;                                                                   \
                const T *Xblk = pX + cc * COLS * N;                                                               \
                /* prefetch_load_Xi / prefetch_load_Ai (src/gemm.jl:314-315) */                                   \
                for (int c = 0; c < COLS; ++c) __builtin_prefetch(Xblk + c * N, 0, 3);                            \
                for (int r = 0; r < ROW_LOADS; ++r) __builtin_prefetch(Ablk + r * VL, 0, 3);                      \
                /* init block: plain rounded product, k = 1 (src/gemm.jl:76-89) */                               \
                for (int r = 0; r < ROW_LOADS; ++r) a[r] = LOADU(Ablk + r * VL);                                  \
                for (int c = 0; c < COLS; ++c) {                                                                  \
                    VEC x = SET1(Xblk[0 + c * N]);                                                                \
                    for (int r = 0; r < ROW_LOADS; ++r) acc[c][r] = MUL(a[r], x);                                 \
                }                                                                                                 \
                /* n = 2..Nr (1-based), then Nd chunks of one cache line of X (src/gemm.jl:319-333) */            \
                int64_t n = 1;                                                                                    \
                for (; n < Nr; ++n) {                                                                             \
                    const T *An = Ablk + n * M;                                                                   \
                    __builtin_prefetch(An + (int64_t)Apf * M, 0, 3);                                              \
                    for (int r = 0; r < ROW_LOADS; ++r) a[r] = LOADU(An + r * VL);                                \
                    for (int c = 0; c < COLS; ++c) {                                                              \
                        VEC x = SET1(Xblk[n + c * N]);                                                            \
                        for (int r = 0; r < ROW_LOADS; ++r) acc[c][r] = FMADD(a[r], x, acc[c][r]);                \
                    }                                                                                             \
                }                                                                                                 \
                for (int64_t nd = 0; nd < Nd; ++nd) {                                                             \
                    for (int c = 0; c < COLS; ++c) __builtin_prefetch(Xblk + n + CL * 7 + c * N, 0, 3);           \
                    for (int q = 0; q < CL; ++q, ++n) {                                                           \
                        const T *An = Ablk + n * M;                                                               \
                        __builtin_prefetch(An + (int64_t)Apf * M, 0, 3);                                          \
                        for (int r = 0; r < ROW_LOADS; ++r) a[r] = LOADU(An + r * VL);                            \
                        for (int c = 0; c < COLS; ++c) {                                                          \
                            VEC x = SET1(Xblk[n + c * N]);                                                        \
                            for (int r = 0; r < ROW_LOADS; ++r) acc[c][r] = FMADD(a[r], x, acc[c][r]);            \
                        }                                                                                         \
                    }                                                                                             \
                }                                                                                                 \
                /* prefetch_storage + store block (src/gemm.jl:334-335, :3-11) */                                 \
                T *Dblk = pD + rc * ROWS + cc * COLS * M;                                                         \
                for (int c = 0; c < COLS; ++c)                                                                    \
                    for (int r = 0; r < ROW_LOADS; ++r) STOREU(Dblk + r * VL + c * M, acc[c][r]);                 \
            }                                                                                                     \
    }

/* AVX-512: REGISTER_SIZE=64, REGISTER_COUNT=32 -> F64 (8,40,5), F32 (16,80,5) */
DEFINE_JMUL_INTR(jmul_f64_avx512, double, __m512d, 8, 5, 5, "avx512f", _mm512_loadu_pd, _mm512_storeu_pd,
                 _mm512_set1_pd, _mm512_mul_pd, _mm512_fmadd_pd)
DEFINE_JMUL_INTR(jmul_f32_avx512, float, __m512, 16, 5, 5, "avx512f", _mm512_loadu_ps, _mm512_storeu_ps,
                 _mm512_set1_ps, _mm512_mul_ps, _mm512_fmadd_ps)
/* AVX2: REGISTER_SIZE=32, REGISTER_COUNT=16 -> F64 (4,12,4), F32 (8,24,4) */
DEFINE_JMUL_INTR(jmul_f64_avx2, double, __m256d, 4, 3, 4, "avx2,fma", _mm256_loadu_pd, _mm256_storeu_pd,
                 _mm256_set1_pd, _mm256_mul_pd, _mm256_fmadd_pd)
DEFINE_JMUL_INTR(jmul_f32_avx2, float, __m256, 8, 3, 4, "avx2,fma", _mm256_loadu_ps, _mm256_storeu_ps,
                 _mm256_set1_ps, _mm256_mul_ps, _mm256_fmadd_ps)

static int have_avx512(void) { return __builtin_cpu_supports("avx512f"); }
static int have_avx2(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }

/* Which machine constants deps/build.jl would have written on this host: 512, 256, or 0 (unsupported). */
int jmul_baseline_isa(void) { return have_avx512() ? 512 : (have_avx2() ? 256 : 0); }

/* Tile the restatement uses on this host for element size t_size. Returns 0 on success. */
int jmul_baseline_tile(int t_size, int *vector_length, int *rows, int *cols)
{
    int isa = jmul_baseline_isa();
    if (!isa) return -1; /* deps/build.jl:15,23 throws on unsupported CPUs */
    jmul_pick_kernel_size(t_size, isa / 8, isa == 512 ? 32 : 16, vector_length, rows, cols);
    return 0;
}

static void edges_f64(double *D, const double *A, const double *X, int64_t M, int64_t K, int64_t N, int64_t mfull,
                      int64_t nfull)
{
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = (j < nfull ? mfull : 0); i < M; ++i) {
            double d = A[i] * X[j * K];
            for (int64_t k = 1; k < K; ++k) d = fma(A[i + k * M], X[k + j * K], d);
            D[i + j * M] = d;
        }
}
static void edges_f32(float *D, const float *A, const float *X, int64_t M, int64_t K, int64_t N, int64_t mfull,
                      int64_t nfull)
{
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = (j < nfull ? mfull : 0); i < M; ++i) {
            float d = A[i] * X[j * K];
            for (int64_t k = 1; k < K; ++k) d = fmaf(A[i + k * M], X[k + j * K], d);
            D[i + j * M] = d;
        }
}

/*
 * jmul!(D, A, X) restated.  BLAS naming at this boundary: D is MxN, A is MxK, X is KxN, dense column-major
 * with leading dimension = row count (MMatrix storage).
 *   col_tile_lo/hi : run only column tiles [lo, hi) of the outer `cc` loop (hi<0 => all); used to time a
 *                    bounded slab of a huge problem -- legitimate since `cc` tiles are independent.
 *   nthreads       : 1 = the reference (single-threaded).  >1 = OpenMP over column tiles, NOT in the reference.
 *   fill_edges     : 0 = the reference (edges untouched); 1 = also compute them (not in the reference).
 *   covered[2]     : out, rows/cols of D the tile loops cover (M - M%rows, N - N%cols).
 */
#define DEFINE_DRIVER(NAME, T, K512, K256, EDGES)                                                                  \
    int NAME(T *D, const T *A, const T *X, int64_t M, int64_t K, int64_t N, int64_t col_tile_lo,                 \
             int64_t col_tile_hi, int nthreads, int fill_edges, int64_t *covered)                               \
    {                                                                                                             \
        int vl, rows, cols;                                                                                       \
        if (M < 0 || N < 0 || K < 1) return -1;                                                                   \
        if (jmul_baseline_tile((int)sizeof(T), &vl, &rows, &cols)) return -3;                                     \
        int isa = jmul_baseline_isa();                                                                            \
        int64_t col_chunks = N / cols;                                                                            \
        if (col_tile_hi < 0 || col_tile_hi > col_chunks) col_tile_hi = col_chunks;                                \
        if (col_tile_lo < 0) col_tile_lo = 0;                                                                     \
        if (covered) {                                                                                            \
            covered[0] = M - M % rows;                                                                            \
            covered[1] = (col_tile_hi - col_tile_lo) * cols;                                                      \
        }                                                                                                         \
        const int Apf = 7; /* default Aprefetch_freq, src/gemm.jl:245 */                                          \
        if (nthreads <= 1) {                                                                                      \
            if (isa == 512)                                                                                       \
                K512(D, A, X, M, K, N, col_tile_lo, col_tile_hi, Apf);                                            \
            else                                                                                                  \
                K256(D, A, X, M, K, N, col_tile_lo, col_tile_hi, Apf);                                            \
        } else {                                                                                                  \
            _Pragma("omp parallel for schedule(dynamic, 4) num_threads(nthreads)") for (int64_t cc = col_tile_lo; \
                                                                                         cc < col_tile_hi; ++cc)  \
            {                                                                                                     \
                if (isa == 512)                                                                                   \
                    K512(D, A, X, M, K, N, cc, cc + 1, Apf);                                                      \
                else                                                                                              \
                    K256(D, A, X, M, K, N, cc, cc + 1, Apf);                                                      \
            }                                                                                                     \
        }                                                                                                         \
        if (fill_edges) EDGES(D, A, X, M, K, N, M - M % rows, col_chunks * cols);                                 \
        return 0;                                                                                                 \
    }

DEFINE_DRIVER(jmul_baseline_f64, double, jmul_f64_avx512, jmul_f64_avx2, edges_f64)
DEFINE_DRIVER(jmul_baseline_f32, float, jmul_f32_avx512, jmul_f32_avx2, edges_f32)

int jmul_baseline_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* fastmul! (src/kernels.jl:43-130, 202-208), restated for the batched-products baseline        */
/* ------------------------------------------------------------------------------------------ */
/* The generated fastmul! keeps ALL of D in registers: Q = M/W row vectors x P columns of accumulators; the first contraction
 * step is a plain product (mulinit, src/kernels.jl:3-23), every later step loads the Q vectors of A[:, n], broadcasts X[n, p]
 * for each column p and does vmuladd (:91-100); D is stored once at the end (:112-117).  Restated here for the shapes whose
 * accumulators fit the register file (Q*P <= 28 zmm / 12 ymm), M a multiple of the vector length; other shapes run the plain
 * chain loops (same arithmetic, no claim about speed).  One thread, like the reference. */
#define DEFINE_FASTMUL(NAME, VEC, VL, QMAX, PMAX, TARGET, LOADU, STOREU, SET1, MUL, FMADD)                          \
    __attribute__((target(TARGET))) static int NAME(double *D, const double *A, const double *X, int64_t M,        \
                                                    int64_t N, int64_t P)                                          \
    {                                                                                                             \
        if (M % VL || M / VL > QMAX || P > PMAX || (M / VL) * P > QMAX * PMAX || N < 1) return 0;                 \
        const int Q = (int)(M / VL);                                                                              \
        VEC acc[QMAX][PMAX], a[QMAX];                                                                             \
        for (int q = 0; q < Q; ++q) a[q] = LOADU(A + q * VL);                                                     \
        for (int p = 0; p < P; ++p) {                                                                             \
            VEC x = SET1(X[p * N]);                                                                               \
            for (int q = 0; q < Q; ++q) acc[q][p] = MUL(a[q], x);                                                 \
        }                                                                                                         \
        for (int64_t n = 1; n < N; ++n) {                                                                         \
            for (int q = 0; q < Q; ++q) a[q] = LOADU(A + n * M + q * VL);                                         \
            for (int p = 0; p < P; ++p) {                                                                         \
                VEC x = SET1(X[n + p * N]);                                                                       \
                for (int q = 0; q < Q; ++q) acc[q][p] = FMADD(a[q], x, acc[q][p]);                                \
            }                                                                                                     \
        }                                                                                                         \
        for (int p = 0; p < P; ++p)                                                                               \
            for (int q = 0; q < Q; ++q) STOREU(D + p * M + q * VL, acc[q][p]);                                    \
        return 1;                                                                                                 \
    }

DEFINE_FASTMUL(fastmul_f64_avx512, __m512d, 8, 2, 14, "avx512f", _mm512_loadu_pd, _mm512_storeu_pd, _mm512_set1_pd, _mm512_mul_pd,
               _mm512_fmadd_pd)
DEFINE_FASTMUL(fastmul_f64_avx2, __m256d, 4, 2, 6, "avx2,fma", _mm256_loadu_pd, _mm256_storeu_pd, _mm256_set1_pd, _mm256_mul_pd,
               _mm256_fmadd_pd)

/* Specialised copy for the reference's published shape (16x32x14: Q = 2, P = 14, 28 zmm accumulators) with compile-time
 * trip counts, so the compiler keeps the accumulators in registers exactly as the generated Julia code does. */
__attribute__((target("avx512f"))) static void fastmul_f64_16x32x14_avx512(double *D, const double *A, const double *X)
{
    __m512d acc[2][14], a0 = _mm512_loadu_pd(A), a1 = _mm512_loadu_pd(A + 8);
#pragma GCC unroll 14
    for (int p = 0; p < 14; ++p) {
        __m512d x = _mm512_set1_pd(X[p * 32]);
        acc[0][p] = _mm512_mul_pd(a0, x);
        acc[1][p] = _mm512_mul_pd(a1, x);
    }
    for (int n = 1; n < 32; ++n) {
        a0 = _mm512_loadu_pd(A + n * 16);
        a1 = _mm512_loadu_pd(A + n * 16 + 8);
#pragma GCC unroll 14
        for (int p = 0; p < 14; ++p) {
            __m512d x = _mm512_set1_pd(X[n + p * 32]);
            acc[0][p] = _mm512_fmadd_pd(a0, x, acc[0][p]);
            acc[1][p] = _mm512_fmadd_pd(a1, x, acc[1][p]);
        }
    }
#pragma GCC unroll 14
    for (int p = 0; p < 14; ++p) {
        _mm512_storeu_pd(D + p * 16, acc[0][p]);
        _mm512_storeu_pd(D + p * 16 + 8, acc[1][p]);
    }
}

/* `batch` independent products D_b(MxP) = A_b(MxN) * X_b(NxP), dense column-major, matrix b at b*stride; jBLAS dimension names. */
int fastmul_baseline_batched_f64(double *D, const double *A, const double *X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                 int64_t strideD, int64_t strideA, int64_t strideX)
{
    if (M < 0 || N < 1 || P < 0 || batch < 0) return -1;
    const int isa = jmul_baseline_isa();
    for (int64_t b = 0; b < batch; ++b) {
        double *d = D + b * strideD;
        const double *a = A + b * strideA, *x = X + b * strideX;
        if (isa == 512 && M == 16 && N == 32 && P == 14) {
            fastmul_f64_16x32x14_avx512(d, a, x);
            continue;
        }
        if ((isa == 512 && fastmul_f64_avx512(d, a, x, M, N, P)) || (isa == 256 && fastmul_f64_avx2(d, a, x, M, N, P))) continue;
        for (int64_t p = 0; p < P; ++p)
            for (int64_t i = 0; i < M; ++i) {
                double v = a[i] * x[p * N];
                for (int64_t n = 1; n < N; ++n) v = fma(a[i + n * M], x[n + p * N], v);
                d[i + p * M] = v;
            }
    }
    return 0;
}

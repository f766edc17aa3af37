// aux_kernels.cuh -- input generation (mrandn) and pipe-rate probes.
#pragma once
#include "common.cuh"

namespace jb {

// ---------------------------------------------------------------------------------------------------------
// mrandn (src/randmat.jl:5-14): the reference fills an MMatrix with randn() from Julia's global, unseeded
// RNG.  Here: counter-based Philox4x32-10 keyed by (seed), Box-Muller in double precision; a Float32 matrix
// receives the double draw rounded to float, which is what `x[i] = randn()` does for an MMatrix{..,Float32}.
// Element i depends only on (seed, i): any launch geometry and any sharding of a matrix across GPUs
// produce the same values.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename T>
__global__ void randn_fill_kernel(T* __restrict__ out, int64_t first, int64_t n, uint64_t seed)
{
    // out[e - first] = stream element e for e in [first, first + n); element e is half (e & 1) of pair e >> 1
    const int64_t p_lo = first >> 1, p_hi = (first + n - 1) >> 1;
    for (int64_t p = p_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= p_hi;
         p += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10((uint32_t)p, (uint32_t)(p >> 32), 0x6a424c41u /* "jBLA" */, 0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        // two uniforms in (0,1) with 53 and 52 random bits
        double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        double u2 = ((double)((((uint64_t)r[2] << 32) | r[3]) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
        double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        const int64_t e0 = 2 * p - first, e1 = e0 + 1;
        if (e0 >= 0 && e0 < n) out[e0] = (T)(rad * c);
        if (e1 >= 0 && e1 < n) out[e1] = (T)(rad * s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pipe-rate probes: register-only loops that measure what the FP64 / FP32 / DMMA pipes sustain on this part
// at the clocks it actually holds.  These are the compute roofline denominators (MEASURED_PEAKS.json only has
// HBM and bf16 figures).
// ---------------------------------------------------------------------------------------------------------
template <int CHAINS>
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters, double a, double b)
{
    double c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) probe_ffma_kernel(float* out, int iters, float a, float b)
{
    float c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fmaf(a, c[i], b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 12345.678f) out[0] = s;
}

template <int TILES>
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters, double a, double b)
{
    double c0[TILES], c1[TILES];
#pragma unroll
    for (int i = 0; i < TILES; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
    double av = a + threadIdx.x, bv = b - threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

// ---- probes with the GEMM's own operand pattern (8x8 outer product per step: two fresh register operands per FMA,
//      one reused), which is what the register file must sustain in a real micro-kernel ----
__global__ void __launch_bounds__(256) probe_dmma_tile_kernel(double* out, int iters, double a0, double b0)
{
    double acc[8][4][2], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = b0 - 1e-3 * (threadIdx.x + j);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                             : "d"(a[i]), "d"(b[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 12345.678) out[0] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) probe_fma_tile_kernel(T* out, int iters, T a0, T b0)
{
    T acc[8][8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = a0 + T(1e-3) * T(threadIdx.x + i);
        b[i] = b0 - T(1e-3) * T(threadIdx.x + i);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[j][r] = T(0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[j][r] = fma_t(a[r], b[j], acc[j][r]);
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < 8; ++r) s += acc[j][r];
    if (s == T(12345.678)) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_ffma2_tile_kernel(float* out, int iters, float a0, float b0)
{
    uint64_t acc[8][4], a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo = a0 + 1e-3f * (threadIdx.x + i), hi = a0 - 1e-3f * (threadIdx.x + i);
        asm("mov.b64 %0, {%1, %2};\n" : "=l"(a[i]) : "f"(lo), "f"(hi));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = b0 - 1e-3f * (threadIdx.x + j);
        asm("mov.b64 %0, {%1, %1};\n" : "=l"(b[j]) : "f"(v));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[j][r] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)  // the A pair is the reused operand, the broadcast scalar and the accumulator are fresh
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[j][r]) : "l"(a[r]), "l"(b[j]));
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) s ^= acc[j][r];
    if (s == 0x123456789abcdefull) out[0] = 1.f;
}

#ifdef JBLAS_B200_TUNING_PROBES  // development aid (tools/probe_ffma2.py): `python -m jblas.jl_b200.build --probes`; not in the shipped library
// The exact-FP32 kernel's inner loop (gemm_simt_f32x2.cuh) on a resident shared-memory tile, without any global traffic:
// what the FFMA2 stream sustains WITH its shared-memory operand loads.  MODE bits: 1 = __syncthreads per 32-deep k-tile,
// 2 = A through four LDS.64 instead of two LDS.128, 4 = no X loads (registers reused), 8 = no A loads.
template <int MODE, int NJ = 8, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) probe_ffma2_lds_kernel(float* out, int iters)
{
    constexpr int BM = 128, BN = 16 * NJ, BK = (NJ > 8 ? 16 : 32), LDA = BM, LDB = BK + 4;
    __shared__ __align__(16) float sAb[BK * LDA];
    __shared__ __align__(16) float sBb[BN * LDB];
    for (int i = threadIdx.x; i < BK * LDA; i += 256) sAb[i] = 1e-3f * (float)((i * 7 + 3) % 13);
    for (int i = threadIdx.x; i < BN * LDB; i += 256) sBb[i] = 1e-3f * (float)((i * 5 + 1) % 11);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // MODE bit 16: lanes of a quarter-warp = 2 row groups x 4 column groups (instead of 8 row groups of one column group)
    // MODE bit 32: quarter q = 4 row groups x 2 column groups chosen so that the two quarters of each half-warp read DISJOINT
    //              chunks of both operands (an LDS.128 is served per half-warp; duplicates across its two quarters cost a wavefront)
    const int wm = warp % 2, wn = warp / 2, qq = lane >> 3, ll = lane & 7;
    const int tx = (MODE & 32) ? ((ll & 3) + 4 * (qq & 1)) : ((MODE & 16) ? (lane >> 2) : (lane & 7));
    const int ty = (MODE & 32) ? ((ll >> 2) + 2 * ((qq & 1) ^ (qq >> 1))) : ((MODE & 16) ? (lane & 3) : (lane >> 3));
    const float* sA = sAb + wm * 64 + tx * 4;
    const float* sB = sBb + (wn * 4 * NJ + ty) * LDB;
    uint64_t acc[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[j][p] = 0ull;
    float2 b[NJ];
    ulonglong2 a[2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = make_float2(1e-3f * j, 2e-3f * j);
    a[0] = make_ulonglong2(0x3a83126f3a83126full, 0x3a83126f3a83126full);
    a[1] = a[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kc = 0; kc < BK; kc += 2) {
            if constexpr (!(MODE & 4)) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const float2*>(sB + j * 4 * LDB + kc);
            }
#pragma unroll
            for (int kv = 0; kv < 2; ++kv) {
                if constexpr (!(MODE & 8)) {
                    if constexpr (MODE & 2) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            a[i].x = *reinterpret_cast<const uint64_t*>(sA + (kc + kv) * LDA + i * 32);
                            a[i].y = *reinterpret_cast<const uint64_t*>(sA + (kc + kv) * LDA + i * 32 + 2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const ulonglong2*>(sA + (kc + kv) * LDA + i * 32);
                    }
                }
                uint64_t bb[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float bs = kv ? b[j].y : b[j].x;
                    asm("mov.b64 %0, {%1, %1};\n" : "=l"(bb[j]) : "f"(bs));
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const uint64_t ap = (p & 1) ? a[p >> 1].y : a[p >> 1].x;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;\n" : "+l"(acc[j][p]) : "l"(ap), "l"(bb[j]));
                }
            }
        }
        if constexpr (MODE & 1) __syncthreads();
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) s ^= acc[j][p];
    if (s == 0x123456789abcdefull) out[0] = 1.f;
}

#endif  // JBLAS_B200_TUNING_PROBES

}  // namespace jb

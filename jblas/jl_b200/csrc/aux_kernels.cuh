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

}  // namespace jb

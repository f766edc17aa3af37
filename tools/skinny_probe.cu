// skinny_probe.cu -- development harness for gemm_skinny.cuh: variants timed on cold operands (rotating sets > L2), bit-compared
// with a plain fma-chain kernel.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/skinny_probe tools/skinny_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include "../jblas/jl_b200/csrc/gemm_skinny.cuh"
using namespace jb;

__global__ void chain_ref(double* D, const double* A, const double* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    int m = i % M, n = i / M;
    double d = -0.0;
    for (int k = 0; k < K; ++k) d = fma(A[(size_t)k * lda + m], X[(size_t)n * ldx + k], d);
    D[(size_t)n * ldd + m] = d;
}
__global__ void fill(double* p, size_t n, uint64_t seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull ^ seed;
        x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29;
        p[i] = ((double)(x >> 11) / 9007199254740992.0 - 0.5) * 4.0;
    }
}

// register-only DMMA loop: ACCS independent accumulator pairs per warp, `warps` warps per CTA, one CTA per SM
template <int ACCS>
__global__ void dmma_rate_kernel(double* out, int iters, double a0, double b0)
{
    double acc[ACCS][2];
#pragma unroll
    for (int i = 0; i < ACCS; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACCS; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ACCS; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}
// the skinny kernel's k-step in registers only: acc[2][8], two A values, eight B values; ORDER 0 = ni outer / mi inner, 1 = mi outer / ni inner
template <int ORDER, int NIX>
__global__ void dmma_pattern_kernel(double* out, int iters, double a0, double b0)
{
    double acc[2][NIX][2];
#pragma unroll
    for (int i = 0; i < NIX; ++i) acc[0][i][0] = acc[0][i][1] = acc[1][i][0] = acc[1][i][1] = 0.0;
    double ax = a0 + threadIdx.x * 1e-9, ay = a0 - threadIdx.x * 1e-9, b[NIX];
#pragma unroll
    for (int i = 0; i < NIX; ++i) b[i] = b0 + i * 1e-7;
    for (int it = 0; it < iters; ++it) {
        if (ORDER == 0) {
#pragma unroll
            for (int i = 0; i < NIX; ++i) { dmma884(acc[0][i][0], acc[0][i][1], ax, b[i]); dmma884(acc[1][i][0], acc[1][i][1], ay, b[i]); }
        } else {
#pragma unroll
            for (int i = 0; i < NIX; ++i) dmma884(acc[0][i][0], acc[0][i][1], ax, b[i]);
#pragma unroll
            for (int i = 0; i < NIX; ++i) dmma884(acc[1][i][0], acc[1][i][1], ay, b[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NIX; ++i) s += acc[0][i][0] + acc[0][i][1] + acc[1][i][0] + acc[1][i][1];
    if (s == 123.456) out[0] = s;
}
template <int ORDER, int NIX>
static void dmma_pattern(int warps, int sms, double* out)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_pattern_kernel<ORDER, NIX><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * 2 * NIX * iters * (double)sms * warps;
    printf("dmma pattern: order %d, 2 x %d tiles, %2d warps/SM: %6.2f TFLOP/s\n", ORDER, NIX, warps, flops / (best * 1e-3) / 1e12);
    fflush(stdout);
}

template <int ACCS>
static void dmma_rate(int warps, int sms, double* out)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_rate_kernel<ACCS><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * ACCS * iters * (double)sms * warps;
    printf("dmma rate: %2d accumulators/warp, %2d warps/SM: %6.2f TFLOP/s\n", ACCS, warps, flops / (best * 1e-3) / 1e12);
    fflush(stdout);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(const double* A, int M, int K, int64_t lda, int KC)
{
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        enc = (EncodeTiledFn)fn;
    }
    CUtensorMap m;
    cuuint64_t gdim[4] = {(cuuint64_t)M, 2, 4, (cuuint64_t)(K / 8)};
    cuuint64_t gstr[3] = {(cuuint64_t)(4 * lda * 8), (cuuint64_t)(lda * 8), (cuuint64_t)(8 * lda * 8)};
    cuuint32_t box[4] = {16, 2, 4, (cuuint32_t)(KC / 8)};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

static CUtensorMap make_map_x(const double* X, int K, int N, int64_t ldx, int BN)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    CUtensorMap m;
    cuuint64_t gdim[3] = {4, (cuuint64_t)N, (cuuint64_t)(K / 4)};
    cuuint64_t gstr[2] = {(cuuint64_t)(ldx * 8), 32};
    cuuint32_t box[3] = {4, (cuuint32_t)BN, (cuuint32_t)(K / 4)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)X, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)r); exit(1); }
    return m;
}

template <typename Cfg>
static void run(const char* name, int M, int N, int K, std::vector<double*>& As, double* X, std::vector<double*>& Ds, double* Dref, int ctas)
{
    const size_t smem = Cfg::smem(K);
    auto kern = gemm_skinny_f64_kernel<Cfg, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, smem);
    const int R = (int)As.size();
    std::vector<CUtensorMap> maps(R);
    for (int i = 0; i < R; ++i) maps[i] = make_map(As[i], M, K, M, Cfg::KC);
    const CUtensorMap mapX = make_map_x(X, K, N, K, Cfg::BN);
    for (int i = 0; i < R; ++i) kern<<<ctas, Cfg::THREADS, smem>>>(maps[i], mapX, Ds[i], M, N, K, M, nullptr, 0, nullptr);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 120;
    float best = 1e9f, tot = 0;
    for (int outer = 0; outer < 3; ++outer) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) kern<<<ctas, Cfg::THREADS, smem>>>(maps[i % R], mapX, Ds[i % R], M, N, K, M, nullptr, 0, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        tot = ms / reps; if (tot < best) best = tot;
    }
    if (getenv("SKINNY_TRACE")) {
        const int nw = ctas * Cfg::WARPS;
        unsigned long long* tr;
        cudaMalloc(&tr, (size_t)nw * 12 * 8);
        cudaMemset(tr, 0, (size_t)nw * 12 * 8);
        kern<<<ctas, Cfg::THREADS, smem>>>(maps[1], mapX, Ds[1], M, N, K, M, nullptr, 0, tr);
        cudaDeviceSynchronize();
        std::vector<unsigned long long> ht((size_t)nw * 12);
        cudaMemcpy(ht.data(), tr, ht.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int w = 0; w < nw; ++w) if (ht[w * 12] && ht[w * 12] < t0) t0 = ht[w * 12];
        const char* names[12] = {"entry", "X staged", "first box", "block 0 stored", "block 1 stored", "block 2 stored", "block 3 stored", "block 4", "block 5", "block 6", "block 7", "block 8"};
        for (int sl = 0; sl < 12; ++sl) {
            double mn = 1e18, mx = 0, sum = 0; int cnt = 0;
            for (int w = 0; w < nw; ++w) { unsigned long long v = ht[w * 12 + sl]; if (!v) continue; double d = (double)(v - t0) / 1e3; mn = d < mn ? d : mn; mx = d > mx ? d : mx; sum += d; ++cnt; }
            if (cnt) printf("    trace %-16s warps %5d  min %6.2f  avg %6.2f  max %6.2f us\n", names[sl], cnt, mn, sum / cnt, mx);
        }
        cudaFree(tr);
    }
    std::vector<double> h((size_t)M * N), r((size_t)M * N);
    cudaMemcpy(h.data(), Ds[0], h.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r.data(), Dref, r.size() * 8, cudaMemcpyDeviceToHost);
    const bool same = memcmp(h.data(), r.data(), h.size() * 8) == 0;
    const double us = best * 1e3, bytes = ((double)M * K + (double)K * N + (double)M * N) * 8, flops = 2.0 * M * N * K;
    fflush(stdout);
    printf("%-28s ctas %4d regs %3d occ %d smem %6zu  %7.2f us  %6.2f TFLOP/s  %5.2f TB/s  %s\n", name, ctas, fa.numRegs, occ, smem, us, flops / us / 1e6,
           bytes / us / 1e6, same ? "bit-identical" : "MISMATCH");
    fflush(stdout);
}

int main(int argc, char** argv)
{
    int M = argc > 1 ? atoi(argv[1]) : 65536, N = argc > 2 ? atoi(argv[2]) : 64, K = argc > 3 ? atoi(argv[3]) : 64;
    const int only = argc > 4 ? atoi(argv[4]) : -1;
    int idx = 0;
    const int R = 6;
    std::vector<double*> As(R), Ds(R);
    double *X, *Dref;
    for (int i = 0; i < R; ++i) {
        cudaMalloc(&As[i], (size_t)M * K * 8); cudaMalloc(&Ds[i], (size_t)M * N * 8);
        fill<<<1024, 256>>>(As[i], (size_t)M * K, 11 + i);
        cudaMemset(Ds[i], 0xff, (size_t)M * N * 8);
    }
    cudaMalloc(&X, (size_t)K * N * 8); cudaMalloc(&Dref, (size_t)M * N * 8);
    fill<<<64, 256>>>(X, (size_t)K * N, 99);
    chain_ref<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(Dref, As[0], X, M, N, K, M, M, K);
    cudaDeviceSynchronize();
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (only == -2) {
        for (int w : {4, 8, 16}) { dmma_pattern<0, 8>(w, sms, Dref); dmma_pattern<1, 8>(w, sms, Dref); dmma_pattern<0, 4>(w, sms, Dref); dmma_pattern<1, 4>(w, sms, Dref); }
        for (int w : {4, 8}) { dmma_rate<4>(w, sms, Dref); dmma_rate<8>(w, sms, Dref); dmma_rate<16>(w, sms, Dref); dmma_rate<32>(w, sms, Dref); }
        return 0;
    }
    printf("M %d N %d K %d, %d SMs, %d rotating sets (%.0f MB)\n", M, N, K, sms, R, R * ((double)M * K + (double)M * N) * 8 / 1e6);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 0>>("w8 kc64 plain halves", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 1>>("w8 kc64 stagger", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 2>>("w8 kc64 late 2nd box", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 3>>("w8 kc64 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 32, 3>>("w8 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 16, 16, 3>>("w16 kc16 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 16, 32, 3>>("w16 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 12, 32, 3>>("w12 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    return 0;
}

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

// The skinny kernel's k-step with operands re-loaded from (resident) shared memory every step, no global traffic: which part of
// the instruction mix costs tensor-pipe time?  MI2 = pairs of 8-row tiles (one LDS.128 of A each), NI column tiles;
// BMODE 0: one LDS.64 per B fragment, 1: one LDS.128 per two B fragments; NALU extra integer instructions per step.
template <int MI2, int NI, int BMODE, int NALU>
__global__ void __launch_bounds__(256, 1) dmma_lds_kernel(double* out, int iters)
{
    extern __shared__ __align__(16) unsigned char sm[];
    double* sd = reinterpret_cast<double*>(sm);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sd[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t baseA = smem_u32(sd) + lane * 16, baseB = smem_u32(sd) + 8192 + lane * (BMODE ? 16 : 8);
    double acc[2 * MI2][NI][2];
#pragma unroll
    for (int m = 0; m < 2 * MI2; ++m)
#pragma unroll
        for (int n = 0; n < NI; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    double2 a[2][MI2];
    double b[2][NI];
    unsigned alu = threadIdx.x;
    unsigned alu4[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    auto load = [&](int it, int which) {
        const uint32_t off = (uint32_t)(it & 7) * 512;
#pragma unroll
        for (int m = 0; m < MI2; ++m)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a[which][m].x), "=d"(a[which][m].y) : "r"(baseA + off + m * 4096));
        if (BMODE == 0) {
#pragma unroll
            for (int n = 0; n < NI; ++n) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(b[which][n]) : "r"(baseB + off + n * 256));
        } else {
#pragma unroll
            for (int n = 0; n < NI; n += 2)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(b[which][n]), "=d"(b[which][n + 1]) : "r"(baseB + off + n * 256));
        }
    };
    auto mma = [&](int which) {
#pragma unroll
        for (int n = 0; n < NI; ++n)
#pragma unroll
            for (int m = 0; m < MI2; ++m) {
                dmma884(acc[2 * m][n][0], acc[2 * m][n][1], a[which][m].x, b[which][n]);
                dmma884(acc[2 * m + 1][n][0], acc[2 * m + 1][n][1], a[which][m].y, b[which][n]);
            }
        if (NALU > 0) {
#pragma unroll
            for (int i = 0; i < NALU; ++i) asm volatile("xor.b32 %0, %0, %1;\n" : "+r"(alu) : "r"(0x9E3779B9u + i));
        } else {  // negative: independent integer multiply-adds (IMAD: the FMA-side integer pipe), 4 chains
#pragma unroll
            for (int i = 0; i < -NALU; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;\n" : "+r"(alu4[i & 3]) : "r"(0x9E3779B9u + i), "r"(i + 1));
        }
    };
    load(0, 0);
    for (int it = 0; it < iters; it += 2) {
        load(it + 1, 1);
        mma(0);
        load(it + 2, 0);
        mma(1);
    }
    double sacc = 0;
#pragma unroll
    for (int m = 0; m < 2 * MI2; ++m)
#pragma unroll
        for (int n = 0; n < NI; ++n) sacc += acc[m][n][0] + acc[m][n][1];
    if (sacc == 123.456 || alu == 0x12345678u || (alu4[0] ^ alu4[1] ^ alu4[2] ^ alu4[3]) == 0x12345678u) out[0] = sacc;
}
template <int MI2, int NI, int BMODE, int NALU>
static void dmma_lds(int warps, int sms, double* out)
{
    const int iters = 4000;
    cudaFuncSetAttribute(dmma_lds_kernel<MI2, NI, BMODE, NALU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_lds_kernel<MI2, NI, BMODE, NALU><<<sms, warps * 32, 65536>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, dmma_lds_kernel<MI2, NI, BMODE, NALU>);
    const double flops = 2.0 * 256 * 2 * MI2 * NI * iters * (double)sms * warps;
    printf("dmma+lds: %d x %d tiles, B via %s, %2d extra ALU/step, %2d warps/SM, %3d regs: %6.2f TFLOP/s  (%s)\n", 2 * MI2, NI, BMODE ? "LDS.128" : "LDS.64 ", NALU, warps,
           fa.numRegs, flops / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

// Item boundaries in the synthetic loop: every 16 k-steps of 8 DMMAs (one 16 x 32 item of K = 64) the warp does what the real
// kernel does between items.  MODE 0: nothing (continuous stream); 1: SERIAL boundary -- lane 0 draws a ticket from a shared
// counter (atomicAdd + shfl), a dependent shared-memory read stands in for the mbarrier wait, the first fragments are loaded with
// their latency exposed, the finished accumulators are stored; 2: the same work, but the ticket/"wait"/first fragment loads are
// issued one k-step BEFORE the boundary (nothing is waited for at the boundary itself), stores deferred by two steps.
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) dmma_item_kernel(double* out, int items, double* scratch)
{
    extern __shared__ __align__(16) unsigned char sm[];
    double* sd = reinterpret_cast<double*>(sm);
    __shared__ int ctr;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sd[i] = 1.0 + 1e-9 * i;
    if (threadIdx.x == 0) ctr = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t baseA = smem_u32(sd) + lane * 16, baseB = smem_u32(sd) + 8192 + lane * 8;
    constexpr int NI = 4;
    double acc[2][2][NI][2];
    double2 a[2];
    double b[2][NI];
    auto load = [&](int it, int which, uint32_t extra) {
        const uint32_t off = (uint32_t)(it & 7) * 512 + extra;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a[which].x), "=d"(a[which].y) : "r"(baseA + off));
#pragma unroll
        for (int n = 0; n < NI; ++n) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(b[which][n]) : "r"(baseB + off + n * 256));
    };
    auto mma = [&](double (&c)[2][NI][2], int which) {
#pragma unroll
        for (int n = 0; n < NI; ++n) {
            dmma884(c[0][n][0], c[0][n][1], a[which].x, b[which][n]);
            dmma884(c[1][n][0], c[1][n][1], a[which].y, b[which][n]);
        }
    };
    auto mma_first = [&](double (&c)[2][NI][2], int which) {
        const double nz = -0.0;
#pragma unroll
        for (int n = 0; n < NI; ++n) {
            dmma884_from(c[0][n][0], c[0][n][1], a[which].x, b[which][n], nz, nz);
            dmma884_from(c[1][n][0], c[1][n][1], a[which].y, b[which][n], nz, nz);
        }
    };
    double* myout = scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    auto store = [&](double (&c)[2][NI][2], int item) {
        double* p = myout + (size_t)(item & 3) * gridDim.x * blockDim.x * 2 * 8;
#pragma unroll
        for (int n = 0; n < NI; ++n) {
            *reinterpret_cast<double2*>(p + (size_t)n * 2 * gridDim.x * blockDim.x * 2) = make_double2(c[0][n][0], c[1][n][0]);
            *reinterpret_cast<double2*>(p + (size_t)(n * 2 + 1) * gridDim.x * blockDim.x * 2) = make_double2(c[0][n][1], c[1][n][1]);
        }
    };
    auto ticket = [&]() -> uint32_t {
        int v = 0;
        if (lane == 0) v = atomicAdd(&ctr, 1);
        v = __shfl_sync(0xffffffffu, v, 0);
        uint32_t w;  // dependent shared-memory round trip: stands in for the mbarrier try_wait of the next box
        asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(w) : "r"(smem_u32(sd) + ((uint32_t)v & 63u) * 4));
        return (w & 1u) * 0u + ((uint32_t)v & 1u) * 16u * 0u;
    };
    if (MODE == 0) {
        mma_first(acc[0], 0);
        load(0, 0, 0);
        for (int it = 0; it < items * 16; it += 2) {
            load(it + 1, 1, 0);
            mma(acc[0], 0);
            load(it + 2, 0, 0);
            mma(acc[0], 1);
        }
        store(acc[0], 0);
    } else if (MODE == 1) {
        for (int item = 0; item < items; ++item) {
            const uint32_t extra = ticket();
            load(0, 0, extra);
            load(1, 1, extra);
            mma_first(acc[0], 0);
            load(2, 0, extra);
            mma(acc[0], 1);
            for (int it = 2; it < 16; it += 2) {
                load(it + 1, 1, extra);
                mma(acc[0], 0);
                if (it + 2 < 16) load(it + 2, 0, extra);
                mma(acc[0], 1);
            }
            store(acc[0], item);
        }
    } else {
        uint32_t extra = ticket();
        load(0, 0, extra);
        for (int item = 0; item < items; item += 2) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                double(&c)[2][NI][2] = acc[half];
                double(&pc)[2][NI][2] = acc[half ^ 1];
                load(1, 1, extra);
                mma_first(c, 0);
                load(2, 0, extra);
                mma(c, 1);
                if (item + half > 0) store(pc, item + half - 1);
                uint32_t next_extra = extra;
                for (int it = 2; it < 16; it += 2) {
                    load(it + 1, 1, extra);
                    mma(c, 0);
                    if (it == 12) next_extra = ticket();  // the next item's ticket and "wait": two k-steps before the boundary
                    if (it + 2 < 16) load(it + 2, 0, extra);
                    else load(0, 0, next_extra);          // first fragments of the NEXT item, one k-step ahead as everywhere else
                    mma(c, 1);
                }
                extra = next_extra;
            }
        }
        store(acc[1], items - 1);
    }
    double sacc = 0;
#pragma unroll
    for (int n = 0; n < NI; ++n) sacc += acc[0][0][n][0] + acc[1][1][n][1];
    if (sacc == 123.456) out[0] = sacc;
}
template <int MODE, int WARPS>
static void dmma_item(int sms, double* out, double* scratch)
{
    const int items = 256;
    cudaFuncSetAttribute(dmma_item_kernel<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_item_kernel<MODE, WARPS><<<sms, WARPS * 32, 65536>>>(out, items, scratch);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, dmma_item_kernel<MODE, WARPS>);
    const double flops = 2.0 * 256 * 8 * 16 * items * (double)sms * WARPS;
    printf("dmma items of 16 k-steps x 8 DMMA, boundary mode %d (%s), %2d warps/SM, %3d regs: %6.2f TFLOP/s  (%s)\n", MODE,
           MODE == 0 ? "none" : MODE == 1 ? "serial" : "pipelined across the boundary", WARPS, fa.numRegs, flops / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
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
    size_t smem = Cfg::smem(K);
    if (getenv("SKINNY_PAD")) smem += (size_t)atoi(getenv("SKINNY_PAD"));  // e.g. force one CTA per SM for a 4-warp configuration
    if (smem > 232448) { printf("%-28s needs %zu bytes of shared memory: skipped\n", name, smem); return; }
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
        const char* names[12] = {"entry", "X staged", "first box", "item 0 multiplied", "item 1", "item 2", "item 3", "item 4", "item 5", "item 6", "item 7", "item 8"};
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

template <typename Cfg, bool TEAM_KERNEL = false>
static void run_xreg(const char* name, int M, int N, int K, std::vector<double*>& As, double* X, std::vector<double*>& Ds, double* Dref, int ctas)
{
    if (K != Cfg::K) { printf("%-28s needs K = %d: skipped\n", name, Cfg::K); return; }
    const size_t smem = Cfg::SMEM;
    if (smem > 232448) { printf("%-28s needs %zu bytes of shared memory: skipped\n", name, smem); return; }
    void (*kern)(const CUtensorMap, const double*, int64_t, double*, int, int, int64_t, const double*, int64_t);
    if constexpr (TEAM_KERNEL) kern = gemm_skinny_team_f64_kernel<Cfg, false>;
    else kern = gemm_skinny_xreg_f64_kernel<Cfg, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int R = (int)As.size();
    std::vector<CUtensorMap> maps(R);
    for (int i = 0; i < R; ++i) maps[i] = make_map(As[i], M, K, M, Cfg::K);
    for (int i = 0; i < R; ++i) { cudaMemset(Ds[i], 0xff, (size_t)M * N * 8); kern<<<ctas, Cfg::THREADS, smem>>>(maps[i], X, K, Ds[i], M, N, M, nullptr, 0); }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 120;
    float best = 1e9f;
    const bool pdl = getenv("SKINNY_PDL") != nullptr;  // programmatic dependent launch between the back-to-back launches
    for (int outer = 0; outer < 3; ++outer) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) {
            if (pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, kern, maps[i % R], (const double*)X, (int64_t)K, Ds[i % R], M, N, (int64_t)M, (const double*)nullptr, (int64_t)0);
            } else {
                kern<<<ctas, Cfg::THREADS, smem>>>(maps[i % R], X, K, Ds[i % R], M, N, M, nullptr, 0);
            }
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms / reps < best) best = ms / reps;
    }
    std::vector<double> h((size_t)M * N), r((size_t)M * N);
    cudaMemcpy(h.data(), Ds[0], h.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r.data(), Dref, r.size() * 8, cudaMemcpyDeviceToHost);
    const bool same = memcmp(h.data(), r.data(), h.size() * 8) == 0;
    const double us = best * 1e3, bytes = ((double)M * K + (double)K * N + (double)M * N) * 8, flops = 2.0 * M * N * K;
    printf("%-28s ctas %4d regs %3d smem %6zu  %7.2f us  %6.2f TFLOP/s  %5.2f TB/s  %s\n", name, ctas, fa.numRegs, smem, us, flops / us / 1e6,
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
    if (only == -4) {
        double* scratch;
        cudaMalloc(&scratch, (size_t)sms * 512 * 2 * 8 * 4 * 8 + 4096);
        dmma_item<0, 8>(sms, Dref, scratch); dmma_item<1, 8>(sms, Dref, scratch); dmma_item<2, 8>(sms, Dref, scratch);
        dmma_item<0, 12>(sms, Dref, scratch); dmma_item<1, 12>(sms, Dref, scratch); dmma_item<2, 12>(sms, Dref, scratch);
        dmma_item<0, 16>(sms, Dref, scratch); dmma_item<1, 16>(sms, Dref, scratch); dmma_item<2, 16>(sms, Dref, scratch);
        return 0;
    }
    if (only == -3) {
        for (int w : {8, 12}) {
            dmma_lds<1, 8, 0, 0>(w, sms, Dref);
            dmma_lds<1, 8, 1, 0>(w, sms, Dref);
            dmma_lds<1, 8, 0, 30>(w, sms, Dref);
            dmma_lds<1, 8, 0, -16>(w, sms, Dref);
            dmma_lds<1, 8, 0, -48>(w, sms, Dref);
            dmma_lds<1, 8, 0, -96>(w, sms, Dref);
            dmma_lds<1, 4, 0, -24>(w, sms, Dref);
            dmma_lds<1, 4, 0, -48>(w, sms, Dref);
            dmma_lds<2, 8, 0, 0>(w, sms, Dref);
            dmma_lds<2, 8, 1, 0>(w, sms, Dref);
            dmma_lds<2, 4, 0, 0>(w, sms, Dref);
            dmma_lds<1, 4, 0, 0>(w, sms, Dref);
        }
        return 0;
    }
    if (only == -2) {
        for (int w : {4, 8, 16}) { dmma_pattern<0, 8>(w, sms, Dref); dmma_pattern<1, 8>(w, sms, Dref); dmma_pattern<0, 4>(w, sms, Dref); dmma_pattern<1, 4>(w, sms, Dref); }
        for (int w : {4, 8}) { dmma_rate<4>(w, sms, Dref); dmma_rate<8>(w, sms, Dref); dmma_rate<16>(w, sms, Dref); dmma_rate<32>(w, sms, Dref); }
        return 0;
    }
    printf("M %d N %d K %d, %d SMs, %d rotating sets (%.0f MB)\n", M, N, K, sms, R, R * ((double)M * K + (double)M * N) * 8 / 1e6);
    if (only == -5 || only == 100) run_xreg<SkinnyRegCfg<16, 8, 2>>("xreg k64 w8 nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 101) run_xreg<SkinnyRegCfg<16, 8, 3>>("xreg k64 w8 nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 102) run_xreg<SkinnyRegCfg<16, 12, 2>>("xreg k64 w12 nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 103) run_xreg<SkinnyRegCfg<16, 10, 2>>("xreg k64 w10 nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 104) run_xreg<SkinnyRegCfg<16, 4, 2>>("xreg k64 w4 nbuf2 (lone warps)", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 105) run_xreg<SkinnyRegCfg<8, 12, 2>>("xreg k32 w12 nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 106) run_xreg<SkinnyRegCfg<16, 16, 2, 2>>("xreg k64 w16 quarters", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 107) run_xreg<SkinnyRegCfg<16, 12, 2, 2>>("xreg k64 w12 quarters", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 108) run_xreg<SkinnyRegCfg<16, 8, 2, 2>>("xreg k64 w8 quarters", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 109) run_xreg<SkinnyRegCfg<16, 8, 3, 2>>("xreg k64 w8 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 110) run_xreg<SkinnyRegCfg<16, 12, 2, 2>>("xreg k64 w12 quarters", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 111) run_xreg<SkinnyRegCfg<16, 4, 2, 2>>("xreg k64 w4 quarters (lone)", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 120) run_xreg<SkinnyTeamCfg<16, 16, 3, 2>, true>("team k64 w16 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 121) run_xreg<SkinnyTeamCfg<16, 16, 2, 2>, true>("team k64 w16 quarters nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 122) run_xreg<SkinnyTeamCfg<16, 12, 3, 2>, true>("team k64 w12 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 123) run_xreg<SkinnyTeamCfg<16, 8, 3, 2>, true>("team k64 w8 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 124) run_xreg<SkinnyTeamCfg<16, 8, 3, 4>, true>("team k64 w8 halves nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 125) run_xreg<SkinnyTeamCfg<16, 16, 4, 2>, true>("team k64 w16 quarters nbuf4", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 126) run_xreg<SkinnyTeamCfg<8, 16, 3, 2>, true>("team k32 w16 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 127) run_xreg<SkinnyTeamCfg<32, 8, 3, 2>, true>("team k128 w8 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 128) run_xreg<SkinnyTeamCfg<32, 12, 2, 2>, true>("team k128 w12 quarters nbuf2", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 129) run_xreg<SkinnyTeamCfg<24, 12, 3, 2>, true>("team k96 w12 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 130) run_xreg<SkinnyTeamCfg<24, 8, 3, 2>, true>("team k96 w8 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 131) run_xreg<SkinnyTeamCfg<4, 16, 3, 2>, true>("team k16 w16 quarters nbuf3", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 132) run_xreg<SkinnyTeamCfg<16, 16, 3, 2, 32>, true>("team k64 w16 quarters BN32", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 133) run_xreg<SkinnyTeamCfg<16, 16, 3, 1, 32>, true>("team k64 w16 eighths BN32", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 134) run_xreg<SkinnyTeamCfg<8, 16, 3, 2, 32>, true>("team k32 w16 quarters BN32", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 135) run_xreg<SkinnyTeamCfg<16, 16, 3, 1, 16>, true>("team k64 w16 eighths BN16", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only == 136) run_xreg<SkinnyTeamCfg<8, 16, 3, 1, 16>, true>("team k32 w16 eighths BN16", M, N, K, As, X, Ds, Dref, sms);
    if (only == -5 || only >= 100) return 0;
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 12, 32, 3>>("w12 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 3>>("w8 kc64 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 32, 3>>("w8 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 0>>("w8 kc64 plain halves", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<4, 12, 32, 3>>("ni4 w12 kc32 stagger+late", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 3 + 16>>("w8 kc64 stagger+late+antiphase", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 64, 2 + 16>>("w8 kc64 late+antiphase", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 12, 32, 3 + 16>>("w12 kc32 stagger+late+antiphase", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 12, 32, 2 + 16>>("w12 kc32 late+antiphase", M, N, K, As, X, Ds, Dref, sms);
    if (only < 0 || only == idx++) run<SkinnyCfg<8, 8, 32, 2 + 16>>("w8 kc32 late+antiphase", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 4, 64, 2>>("w4 kc64 late (SKINNY_PAD=40000: one warp per sub-partition)", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 4, 64, 2 + 12>>("w4 kc64 late NO STORES, NO A", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 12, 32, 3 + 4>>("w12 kc32 NO STORES", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 12, 32, 3 + 8>>("w12 kc32 NO A BOXES", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 12, 32, 3 + 12>>("w12 kc32 NO STORES, NO A", M, N, K, As, X, Ds, Dref, sms);
    if (only == idx++) run<SkinnyCfg<8, 8, 64, 3 + 12>>("w8 kc64 NO STORES, NO A", M, N, K, As, X, Ds, Dref, sms);
    return 0;
}

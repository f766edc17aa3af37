// call_overhead.cu -- host cost per call of the device-pointer C-ABI entries on small products (what a Julia ccall pays beyond the kernel).
//   nvcc -O2 -std=c++17 -o tools/call_overhead tools/call_overhead.cu -ldl   (loads jblas/jl_b200/libjblas_b200.so with dlopen)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef int (*init_fn)(int);
typedef int (*gemm_fn)(double*, const double*, const double*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, void*);
typedef const char* (*err_fn)(void);

int main(int argc, char** argv)
{
    void* h = dlopen(argc > 1 ? argv[1] : "jblas/jl_b200/libjblas_b200.so", RTLD_NOW);
    if (!h) { printf("dlopen failed: %s\n", dlerror()); return 1; }
    init_fn init = (init_fn)dlsym(h, "jblas_b200_init");
    gemm_fn gemm = (gemm_fn)dlsym(h, "jblas_b200_gemm_f64_dev");
    err_fn err = (err_fn)dlsym(h, "jblas_b200_last_error");
    if (init(0)) { printf("init: %s\n", err()); return 1; }
    cudaStream_t s;
    cudaStreamCreate(&s);
    const int shapes[][3] = {{16, 32, 14}, {64, 64, 64}, {256, 256, 256}, {512, 512, 512}, {1024, 1024, 1024}, {16384, 64, 64}, {1023, 4097, 777}};
    for (auto& sh : shapes) {
        const int64_t M = sh[0], K = sh[1], N = sh[2];
        double *A, *X, *D;
        cudaMalloc(&A, M * K * 8); cudaMalloc(&X, K * N * 8); cudaMalloc(&D, M * N * 8);
        cudaMemset(A, 0, M * K * 8); cudaMemset(X, 0, K * N * 8);
        for (int i = 0; i < 20; ++i) if (gemm(D, A, X, M, K, N, M, M, K, 0, 0, s)) { printf("gemm: %s\n", err()); return 1; }
        cudaStreamSynchronize(s);
        const int reps = 2000;
        // (1) host cost: time to ISSUE reps calls (the stream absorbs them asynchronously as long as the queue is not full)
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) gemm(D, A, X, M, K, N, M, M, K, 0, 0, s);
        auto t1 = std::chrono::steady_clock::now();
        cudaEventRecord(e1, s);
        cudaStreamSynchronize(s);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%5lld x %5lld x %5lld: host issue %6.2f us per call, stream time %7.2f us per call\n", (long long)M, (long long)N, (long long)K,
               std::chrono::duration<double, std::micro>(t1 - t0).count() / reps, ms * 1e3 / reps);
        cudaFree(A); cudaFree(X); cudaFree(D);
    }
    return 0;
}

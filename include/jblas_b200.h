/*
 * jblas_b200.h -- C ABI of libjblas_b200.so: the B200 (sm_100a) replacement for ONE hot path of
 * JuliaBLAS/jBLAS.jl, the dense matrix multiply  D = A*X  into a preallocated column-major D
 * (`jmul!`, which BASELINE.json calls `gemm!`), with its tile-level siblings kernel!/initkernel!/fastmul!.
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, no C++/torch types.  A Julia
 * maintainer binds these with `ccall` (julia/jBLASB200.jl); the only FFI precedent in the
 * reference has exactly this calling shape:
 *     ccall((:sym, lib), Cvoid, (Ptr{Float64},Ptr{Float64},Ptr{Float64}), D, A, X)     test/runtests.jl:97-101
 *
 * Conventions (all from the reference, citations relative to /root/reference):
 *   - dense column-major storage, element (i,j) at p[i + j*ld]        src/gemm.jl:3-11, :309-311 (MMatrix)
 *   - output first: (D, A, X)                                         src/gemm.jl:244
 *   - D is overwritten (jmul!/initkernel!/fastmul!) or accumulated (kernel!)   src/kernels.jl:226 vs :260
 *   - per element:  d = A[i,1]*X[1,j];  d = fma(A[i,n], X[n,j], d), n ascending   src/gemm.jl:86,165,319-333
 *   - caller owns all three matrices; the callee allocates nothing the caller sees   test/runtests.jl:113-114
 * Unlike the reference, remainder rows/columns ARE computed (the reference silently skips them,
 * src/gemm.jl:266-267,313) and sizes/leading dimensions are validated instead of being undefined behaviour.
 *
 * Dimension names: the BLAS-style entry points use (M, K, N): D is MxN, A is MxK, X is KxN.
 * The jBLAS-style entry points (jblas_b200_jmul_*, _kernel_*, ...) keep the reference's names
 * (M, N, P): D is MxP, A is MxN, X is NxP, N is the contraction dimension (src/gemm.jl:244).
 *
 * Return value: 0 on success, negative JBLAS_B200_E* on error; jblas_b200_last_error() returns a
 * thread-local message.  There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef JBLAS_B200_H
#define JBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JBLAS_B200_VERSION 100 /* 0.1.0, as Project.toml:4 */

/* error codes */
#define JBLAS_B200_OK 0
#define JBLAS_B200_EINVAL (-1)       /* bad dimension / leading dimension / null pointer / selector     */
#define JBLAS_B200_ECUDA (-2)        /* a CUDA runtime call failed (message has the CUDA error string)  */
#define JBLAS_B200_ENOTINIT (-3)     /* jblas_b200_init has not succeeded in this process                */
#define JBLAS_B200_EUNSUPPORTED (-4) /* kernel selector not available for this shape / dtype            */
#define JBLAS_B200_ENOMEM (-5)       /* device allocation failed                                         */

/* FP64 kernel selector (north_star: DMMA tensor path vs SIMT, faster one kept per shape) */
#define JBLAS_B200_F64_AUTO 0
#define JBLAS_B200_F64_DMMA 1 /* mma.sync m8n8k4 f64 (DMMA.8x8x4); tolerance contract 2*K*eps*(|A||X|) */
#define JBLAS_B200_F64_SIMT 2 /* DFMA chain; bit-identical to the reference chain                       */
/* FP32 mode selector */
#define JBLAS_B200_F32_EXACT 0  /* FFMA chain; bit-identical to the reference chain in Float32          */
#define JBLAS_B200_F32_3XTF32 1 /* opt-in split-precision tcgen05/TMEM path, looser stated bound        */

/* dtype tags for the generic helpers */
#define JBLAS_B200_DT_F64 0
#define JBLAS_B200_DT_F32 1

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* Bind the calling process to CUDA device `device` (one process per GPU) and create the context
 * (streams, events, staging buffers).  Idempotent for the same device.  Every entry makes the bound device current
 * on the calling thread; device pointers must live on it (JBLAS_B200_EINVAL otherwise). */
int jblas_b200_init(int device);
int jblas_b200_shutdown(void);
int jblas_b200_version(void);
const char* jblas_b200_last_error(void);
/* Number of CUDA devices visible, or a negative error code. */
int jblas_b200_device_count(void);

/* ---- the hot path on HOST pointers: the literal drop-in for jmul!/gemm! ----------------------------
 * Replaces: jmul!(D, A, X)  src/gemm.jl:244-348.  Synchronous like the Julia call: stages A and X to the
 * device (pipelined over column panels), runs the kernel, copies D back.  With accumulate != 0 the old D
 * is uploaded first and D += A*X is computed (kernel! semantics, src/kernels.jl:212-241). */
int jblas_b200_gemm_f64(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                        int64_t lda, int64_t ldx, int accumulate, int kernel);
int jblas_b200_gemm_f32(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                        int64_t lda, int64_t ldx, int accumulate, int mode);

/* ---- the same call on SEVERAL GPUs of this process (single-node multi-GPU mode) --------
 * Replaces: jmul!(D, A, X)  src/gemm.jl:244-246 with its outer column-tile loop (src/gemm.jl:313) cut across GPUs:
 * GPU g of `ngpus` (devices 0..ngpus-1) owns one contiguous column block of X and D and needs all of A.  Host pointers,
 * synchronous, same arguments as jblas_b200_gemm_*.  Every K panel of A crosses PCIe once, one slice per GPU over that
 * GPU's own link, and reaches the other GPUs over NVLink (copy engines, peer access); X blocks go up and D blocks come
 * back over each GPU's link while the panels are multiplied.  Per element the k order is that of a single launch, so
 * the result equals the one-GPU result bit for bit.  jblas_b200_mgpu_init(ngpus) creates the per-GPU contexts and
 * enables peer access (ngpus <= 0: every visible GPU; returns the count); the gemm entries call it on demand.
 * jblas_b200_time_last_ms() reports the wall-clock time of the call. */
int jblas_b200_mgpu_init(int ngpus);
int jblas_b200_mgpu_gemm_f64(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                             int64_t lda, int64_t ldx, int accumulate, int kernel, int ngpus);
int jblas_b200_mgpu_gemm_f32(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                             int64_t lda, int64_t ldx, int accumulate, int mode, int ngpus);

/* jBLAS-named forms on dense MMatrix-style storage (leading dimension = row count).
 * jmul!   src/gemm.jl:244      D(MxP) = A(MxN) * X(NxP)
 * fastmul! src/kernels.jl:202  same contract for small matrices; any M (the reference masks the row remainder,
 *                              src/kernels.jl:59-75) */
int jblas_b200_jmul_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P);
int jblas_b200_jmul_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P);
int jblas_b200_fastmul_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P);
int jblas_b200_fastmul_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P);
/* kernel!(pD, pA, pX, Kernel{Mk,Pk,stride_AD,stride_X,N})   src/kernels.jl:239   D += A*X
 * initkernel!(...)                                          src/kernels.jl:273   D  = A*X
 * stride_AD is the column stride (in elements) of BOTH A and D, stride_X that of X (src/kernels.jl:213-215). */
int jblas_b200_kernel_f64(double* pD, const double* pA, const double* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                          int64_t stride_X, int64_t N);
int jblas_b200_initkernel_f64(double* pD, const double* pA, const double* pX, int64_t Mk, int64_t Pk,
                              int64_t stride_AD, int64_t stride_X, int64_t N);
int jblas_b200_kernel_f32(float* pD, const float* pA, const float* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                          int64_t stride_X, int64_t N);
int jblas_b200_initkernel_f32(float* pD, const float* pA, const float* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                              int64_t stride_X, int64_t N);

/* ---- the hot path on DEVICE pointers (matrices resident in HBM) -----------------------------------
 * Same contract, asynchronous on `stream` (a cudaStream_t; NULL = the CUDA default stream).  These are what
 * the multi-GPU driver and the benchmarks call; D, A, X must not alias. */
int jblas_b200_gemm_f64_dev(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                            int64_t lda, int64_t ldx, int accumulate, int kernel, void* stream);
int jblas_b200_gemm_f32_dev(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                            int64_t lda, int64_t ldx, int accumulate, int mode, void* stream);

/* ---- the fused forms the reference planned ---------------------------------------------
 * src/memory_management.jl:72-76: "add support for how many copies of each of the matrices we have, so that we can
 * perform calculations such as  D = A*X + C  or  D = A*(X + C)  in one step" (the D_count/A_count/X_count keywords of
 * blocking_structure, :78).  The reference never wrote them; defined here in its own chain arithmetic:
 *   gemm_plus_c   : D = A*X + C, C is M x N.  Every element's chain STARTS from C[i,j] (d = C[i,j]; d = fma(A[i,k],
 *                   X[k,j], d), k ascending) -- exactly kernel!'s accumulate (src/kernels.jl:226) with the start value
 *                   read from C instead of D.  C may alias D (then it is kernel!); no extra pass over D.
 *   gemm_x_plus_c : D = A*(X + C), C is K x N.  X + C is rounded once per element, then the jmul! chain.
 * `_dev`: device pointers, asynchronous on `stream`.  Without `_dev`: host pointers, synchronous. */
int jblas_b200_gemm_plus_c_f64_dev(double* D, const double* A, const double* X, const double* C, int64_t M, int64_t K,
                                   int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel, void* stream);
int jblas_b200_gemm_plus_c_f32_dev(float* D, const float* A, const float* X, const float* C, int64_t M, int64_t K,
                                   int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int mode, void* stream);
int jblas_b200_gemm_x_plus_c_f64_dev(double* D, const double* A, const double* X, const double* C, int64_t M, int64_t K,
                                     int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel, void* stream);
int jblas_b200_gemm_x_plus_c_f32_dev(float* D, const float* A, const float* X, const float* C, int64_t M, int64_t K,
                                     int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int mode, void* stream);
int jblas_b200_gemm_plus_c_f64(double* D, const double* A, const double* X, const double* C, int64_t M, int64_t K, int64_t N,
                               int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel);
int jblas_b200_gemm_plus_c_f32(float* D, const float* A, const float* X, const float* C, int64_t M, int64_t K, int64_t N,
                               int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int mode);
int jblas_b200_gemm_x_plus_c_f64(double* D, const double* A, const double* X, const double* C, int64_t M, int64_t K, int64_t N,
                                 int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel);
int jblas_b200_gemm_x_plus_c_f32(float* D, const float* A, const float* X, const float* C, int64_t M, int64_t K, int64_t N,
                                 int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int mode);

/* ---- fastmul!-class BATCHED small products on device pointers ---------------------------
 * `batch` independent products D_b = A_b * X_b in one launch; jBLAS names: D MxP, A MxN, X NxP, every matrix
 * dense column-major, matrix b starts stride_* elements after matrix b-1.  The single-product fastmul!
 * (src/kernels.jl:202-208) cannot amortise a GPU launch; this is its B200 form.  Same chain per element as jmul!
 * (bit-identical to the reference arithmetic).  HBM-bound by design. */
int jblas_b200_fastmul_batched_f64_dev(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P,
                                       int64_t batch, int64_t strideD, int64_t strideA, int64_t strideX, void* stream);
int jblas_b200_fastmul_batched_f32_dev(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P,
                                       int64_t batch, int64_t strideD, int64_t strideA, int64_t strideX, void* stream);
/* Same on HOST pointers (what a Julia ccall on an Array{T,3} or a vector of MMatrix storage hits): synchronous; the batch is
 * streamed through the device in chunks with H2D, the kernel and D2H overlapped.  Gaps between matrices (stride > matrix size)
 * are neither read nor written. */
int jblas_b200_fastmul_batched_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                   int64_t strideD, int64_t strideA, int64_t strideX);
int jblas_b200_fastmul_batched_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                   int64_t strideD, int64_t strideA, int64_t strideX);

/* ---- device memory / transfers (the shim owns no caller memory; these are conveniences) ------------ */
int jblas_b200_alloc(void** dptr, size_t bytes);
int jblas_b200_free(void* dptr);
int jblas_b200_h2d(void* dst_dev, const void* src_host, size_t bytes);
int jblas_b200_d2h(void* dst_host, const void* src_dev, size_t bytes);
/* Pin / unpin a caller-owned host buffer so the host-pointer entry points can DMA at full PCIe rate. */
int jblas_b200_host_register(void* host, size_t bytes);
int jblas_b200_host_unregister(void* host);
int jblas_b200_stream_sync(void* stream);

/* ---- peer memory for the multi-GPU mode (one process per GPU on one node) ----------------
 * The exchange step of the column-sharded product is "A becomes visible on every GPU".  Instead of a collective whose
 * kernels take SMs from the GEMM, the owner EXPORTS its A allocation, every other process MAPS it (CUDA IPC) and pulls
 * K panels with the copy engines over NVLink (jblas_b200_copy_async on a side stream), overlapped with the multiplies.
 *   ipc_export : handle (64 bytes) + byte offset of `dptr` inside its allocation, to be sent to the peers by any means
 *   ipc_open   : maps the peer allocation into this process; *peer_ptr addresses the same bytes as the exporter's dptr
 *   ipc_close  : unmaps (pass the pointer ipc_open returned)
 *   copy_async : device-to-device copy on `stream`, local or peer source (cudaMemcpyAsync; no kernel is launched) */
int jblas_b200_ipc_export(const void* dptr, void* handle64, int64_t* offset);
int jblas_b200_ipc_open(const void* handle64, int64_t offset, void** peer_ptr);
int jblas_b200_ipc_close(void* peer_ptr);
int jblas_b200_copy_async(void* dst, const void* src, size_t bytes, void* stream);

/* ---- inputs: mrandn (src/randmat.jl:5-14): iid N(0,1) fill, here counter-based and seeded -----------
 * Writes elements [first, first+n) of the stream keyed by `seed` to dptr[0..n): element e depends only on
 * (seed, e), so a column shard of a matrix (first = col0*rows) holds the same values as the whole matrix. */
int jblas_b200_randn_fill(void* dptr, int64_t first, int64_t n, uint64_t seed, int dtype, void* stream);

/* ---- planner introspection --------------------------------------------------------------------------
 * The B200 counterpart of pick_kernel_size (src/kernel_structure.jl:76-99) and blocking_structure
 * (src/memory_management.jl:78-140): reports which kernel and tiling AUTO would use.
 * out[0]=kernel id, [1]=CTA tile M, [2]=CTA tile N, [3]=tile K, [4]=pipeline stages, [5]=threads per CTA,
 * out[6]=grid size (CTAs), [7]=raster group (tile rows), [8]=dynamic shared memory bytes, [9]=1 if 16-byte staging. */
int jblas_b200_plan(int dtype, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int selector,
                    int64_t out[10]);
/* Registry introspection: kernel i (0 <= i < num_kernels) can be forced with selector 100+i (tuning/tests). */
int jblas_b200_num_kernels(void);
const char* jblas_b200_kernel_name(int kidx);
/* Number of kernels this library launched since init (for bench.py's gpu_launches). */
int64_t jblas_b200_launch_count(void);
/* CUDA-event time in ms of the last host-pointer call's device work (kernel + copies). */
float jblas_b200_time_last_ms(void);

/* ---- pipe-rate probes (roofline denominators; registers only, no memory traffic) --------------------
 * kind: 0 = DFMA, 1 = DMMA m8n8k4, 2 = FFMA (independent chains, operands reused);
 *       3 = DMMA, 4 = DFMA, 5 = FFMA, 6 = FFMA2 in the GEMM micro-kernel operand pattern (8x8 outer product).
 *       (Development builds with -DJBLAS_B200_TUNING_PROBES add shared-memory ablation kinds 7+; not shipped.)
 * Returns achieved TFLOP/s (2 flop per FMA) in *tflops. */
int jblas_b200_probe_pipe(int kind, int iters, double* tflops, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* JBLAS_B200_H */

// common.cuh -- shared device helpers for the jblas_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jb {

// ---------------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) staging.  This is the replacement for the reference's only memory-hierarchy mechanism,
// the software prefetch statements of src/memory_management.jl:181-277 / src/gemm.jl:314-335: instead of
// hinting lines into L1 a few columns ahead, whole A/X tiles are staged asynchronously into a multi-stage
// shared-memory ring while the FMA/DMMA pipes work on earlier stages.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte copy, `src_bytes` (0..16) taken from global, the rest zero-filled.  .cg: bypass L1 (tiles are
// consumed from shared memory only; L2 carries the inter-CTA reuse).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// 8-byte / 4-byte element copies for operands whose leading dimension or base breaks 16-byte alignment
// (e.g. M = 1023 doubles per column).  .ca is the only variant that allows sizes below 16.
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Element copy of one T with zero fill when !valid.
template <typename T>
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const T* src, bool valid);
template <>
__device__ __forceinline__ void cp_async_elem<double>(uint32_t dst, const double* src, bool valid)
{
    cp_async8(dst, src, valid ? 8 : 0);
}
template <>
__device__ __forceinline__ void cp_async_elem<float>(uint32_t dst, const float* src, bool valid)
{
    cp_async4(dst, src, valid ? 4 : 0);
}

// fused multiply-add, one rounding: the reference's `fma` (src/gemm.jl:165) / `vmuladd` (src/kernels.jl:95)
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

// Tile rasterisation: the 1-D block index walks GROUP_M tile rows at a time so that the CTAs that are
// resident together share A row-panels and X column-panels in L2.  This is the B200 stand-in for the
// reference's (uncalled) cache-block planner, src/memory_management.jl:78-140.
__device__ __forceinline__ void raster(int bid, int tiles_m, int tiles_n, int group_m, int& tm, int& tn)
{
    int per_group = group_m * tiles_n;
    int g = bid / per_group;
    int first_m = g * group_m;
    int gm = min(tiles_m - first_m, group_m);
    int r = bid - g * per_group;
    tm = first_m + r % gm;
    tn = r / gm;
}

}  // namespace jb

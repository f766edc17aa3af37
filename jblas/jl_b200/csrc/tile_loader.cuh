// tile_loader.cuh -- global -> shared staging of one (A tile, X tile) pipeline stage with cp.async.
//
// Column-major operands, as the reference's MMatrix storage (src/gemm.jl:309-311):
//   A is M x K, element (m,k) at A[m + k*lda]   -> staged as sA[k][m]  (m contiguous, row pitch LDA)
//   X is K x N, element (k,n) at X[k + n*ldx]   -> staged as sB[n][k]  (k contiguous, row pitch LDB)
// Both are copied in their native orientation (no transpose): threads stream 16-byte chunks along the
// contiguous dimension, so global reads are fully coalesced 128-byte lines.
//
// ALIGNED=true : 16-byte cp.async; requires 16 B-aligned bases and lda, ldx multiples of 16/sizeof(T).
// ALIGNED=false: element-wise cp.async (8 B / 4 B) for ragged leading dimensions such as M = 1023.
// Out-of-range elements are zero-filled (src-size operand), addresses are clamped in range.
#pragma once
#include "common.cuh"

namespace jb {

template <typename T, int BM, int BN, int BK, int LDA, int LDB, int THREADS, bool ALIGNED>
__device__ __forceinline__ void load_stage(T* __restrict__ sA, T* __restrict__ sB, const T* __restrict__ A,
                                           const T* __restrict__ X, int64_t lda, int64_t ldx, int M, int N, int K,
                                           int m0, int n0, int k0, int tid)
{
    constexpr int VEC = ALIGNED ? (16 / (int)sizeof(T)) : 1;
    // ---- A tile: BK rows (k) of BM elements (m) ----
    {
        constexpr int CPR = BM / VEC;  // chunks per k-row
        constexpr int TOTAL = BK * CPR;
        static_assert(TOTAL % THREADS == 0, "A tile must split evenly over the CTA");
#pragma unroll
        for (int it = 0; it < TOTAL / THREADS; ++it) {
            int c = tid + it * THREADS;
            int kr = c / CPR, mc = (c % CPR) * VEC;
            int gm = m0 + mc, gk = k0 + kr;
            int valid = (gk < K) ? min(max(M - gm, 0), VEC) : 0;
            const T* src = A + (size_t)(valid ? gk : 0) * lda + (valid ? gm : 0);
            uint32_t dst = smem_u32(sA + kr * LDA + mc);
            if constexpr (ALIGNED)
                cp_async16(dst, src, valid * (int)sizeof(T));
            else
                cp_async_elem<T>(dst, src, valid != 0);
        }
    }
    // ---- X tile: BN columns (n) of BK elements (k) ----
    {
        constexpr int CPC = BK / VEC;  // chunks per column
        constexpr int TOTAL = BN * CPC;
        static_assert(TOTAL % THREADS == 0, "X tile must split evenly over the CTA");
#pragma unroll
        for (int it = 0; it < TOTAL / THREADS; ++it) {
            int c = tid + it * THREADS;
            int nc = c / CPC, kc = (c % CPC) * VEC;
            int gn = n0 + nc, gk = k0 + kc;
            int valid = (gn < N) ? min(max(K - gk, 0), VEC) : 0;
            const T* src = X + (size_t)(valid ? gn : 0) * ldx + (valid ? gk : 0);
            uint32_t dst = smem_u32(sB + nc * LDB + kc);
            if constexpr (ALIGNED)
                cp_async16(dst, src, valid * (int)sizeof(T));
            else
                cp_async_elem<T>(dst, src, valid != 0);
        }
    }
}

}  // namespace jb

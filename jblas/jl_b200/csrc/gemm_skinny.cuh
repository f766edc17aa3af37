// gemm_skinny.cuh -- tall-skinny FP64 product D(M x N) = A(M x K) * X(K x N) with N <= 64 and a short K:
// X stays RESIDENT on the SM, A is streamed from HBM exactly once, D is written exactly once.
//
// THREE kernels, in the order they were built (DESIGN.md s4 "Tall-skinny" has the measurements that led from one to the next):
//   1. gemm_skinny_f64_kernel       X in shared memory, warp-private TMA boxes, 16 x 64 items        (N <= 64, K <= 128, K % 8 == 0)
//   2. gemm_skinny_xreg_f64_kernel  X fragments in REGISTERS, warp-private boxes, 16 x 16 items     (K = 32 / 64; AUTO while A fits in L2)
//   3. gemm_skinny_team_f64_kernel  X fragments in registers, ONE box per row block shared by a team of four quarter-column
//                                   warps, 16 warps per SM                                          (K = 32 / 64; AUTO for N > 32)
// All three chain every element in ascending k from -0.0 (or C) with 4 k per DMMA: bit-identical to the reference chain on B200;
// all three are launched with programmatic stream serialisation (griddepcontrol.wait precedes their first global access).
//
// ---- kernel 1 ----
// Regime (BASELINE configs[3], 65536 x 64 x 64): 8 flop per byte -- the FP64 pipe (14.4 us) and HBM (10.3 us) bind almost
// together, so neither may idle.  The reference's analogue is fastmul!'s thin panel (src/kernels.jl:202-208): the whole X
// is small, A is a tall stack of row blocks.  A tile kernel is the wrong shape for it (K = 64 is two pipeline stages; every
// 64 x 32 tile pays a pipeline fill, an mbarrier round trip between warps and re-fetches its X block: 25-27 us,
// profiles/r1_ncu_dmma_tma_skinny_*).
//
// Design: ONE WARP owns one 16-row block of D for all N columns (2 row tiles x NI column tiles of mma.m8n8k4) and runs its
// own private pipeline -- there is no CTA-wide synchronisation after start-up and no producer warp:
//   * lane 0 of the warp issues ONE TMA box (16 rows x 64 k = 8 KiB) per (block, k chunk) into one of the warp's two private
//     shared-memory buffers, completion on the warp's own mbarrier; the box for the NEXT item is in flight while the current
//     one is multiplied (8 warps x 2 x 8 KiB = 128 KiB of loads outstanding per SM, no register cost).  TMA zero-fills
//     rows >= M.
//   * the two 8-row tiles are INTERLEAVED (tile mi holds rows 2g + mi): lane (g, t) reads rows (2g, 2g+1) of column 4s + t
//     with one LDS.128 and later writes rows (2g, 2g+1) of a D column with one 16-byte store -- 8 lanes cover a full
//     128-byte line, D is written straight from the accumulators.
//   * bank conflicts: an LDS.128 is served a quarter-warp at a time (g in {2q, 2q+1}, t in 0..3): its eight 16-byte chunks
//     must differ.  A is therefore described to TMA as a 4-D tensor (m, s_lo, t, s_hi) with k = 8 s_hi + 4 s_lo + t, box
//     {16, 2, 4, 8}: the 128-byte row of k-step s, lane column t lands at row 8 (s >> 1) + 2t + (s & 1), and the hardware
//     128B swizzle XORs the chunk index g with (row & 7) = 2t + (s & 1) -- g ^ 2t is distinct over a quarter-warp.
//   * X fragments come from shared memory, staged once per CTA with cp.async in FRAGMENT-MAJOR order (the 32 lane values of
//     one fragment are contiguous: conflict-free, and every load is one base register + an immediate); k >= K is -0.0 on the X
//     side (and 0 on the A side), so a padded product is -0.0 and leaves every accumulator untouched.
//   * fragments are double-buffered in registers one k-step ahead, which also pins the issue order: 16 independent DMMAs per
//     step (ptxas otherwise chains the steps of one accumulator back to back and stalls on the DMMA latency: measured 46 %
//     tensor pipe with register-resident A).
//   * accumulators start at -0.0 (or C): ascending k, 4 per DMMA -> bit-identical to the reference chain on B200.
// Work distribution.  Blocks are dealt round-robin over ALL warps of the grid with the CTA index fastest, so the blocks in
// flight at any moment are neighbours in memory.  65536 rows are 4096 blocks for 1184 warps = 3.46 each: dealing whole blocks
// costs a fourth round on every SM (measured: 26.6 us against 4.5 us of start-up + 3.46 x 4.8 us).  So only the
// floor(blocks / warps) full rounds are dealt as whole blocks; the REMAINDER is dealt as half blocks (16 rows x half of the
// column tiles: the A box is fetched twice, from L2 the second time), one or two per warp.  The second warp of every SM
// sub-partition (warps 4..7 of 8) takes its half items FIRST, the first warp takes them LAST: the two warps that share a DMMA
// pipe then reach their epilogues half a block apart instead of together (stores + accumulator re-initialisation + the wait
// for the next box of one warp hide behind the other warp's MMAs).
// Start-up: every warp requests only its FIRST box before X has landed and the second one after -- all first boxes are then
// served ahead of all second boxes by HBM (19 MB requested at t = 0 take 3 us to arrive, in no particular order).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_tma.cuh"

namespace jb {

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int NI_, int WARPS_, int KC_ = 64, int FLAGS_ = 3>
struct SkinnyCfg {
    static constexpr int FLAGS = FLAGS_;  // bit 0: stagger the half items (above); bit 1: second box requested after X has landed;
                                          // probe-only ablations (results are wrong): bit 2 = no stores of D, bit 3 = no A boxes (stale shared memory);
                                          // bit 4: anti-phase handshake between the warps of an SM sub-partition (below)
    static constexpr int NI = NI_, WARPS = WARPS_, THREADS = WARPS_ * 32, BN = NI_ * 8;
    static constexpr int KC = KC_;                    // k chunk of one TMA box (multiple of 8)
    static constexpr int BOX_BYTES = 16 * KC * 8;     // 16 rows x KC columns of doubles
    static constexpr int NBUF = 2;
    static_assert(NI_ % 2 == 0, "half items take NI / 2 column tiles");
    static_assert(KC_ % 8 == 0 && BOX_BYTES % 1024 == 0, "a box is a whole number of swizzle atoms");
    // X lives in shared memory FRAGMENT-MAJOR, sX[s][n][t]: the 32 doubles lane (g, t) = X[4s + t][8 ni + g] of k-step s, column
    // tile ni are contiguous in lane order, so every fragment load is base + lane*8 + immediate (conflict-free)
    static size_t x_bytes(int K) { return (size_t)(K / 4) * NI * 256; }
    static size_t smem(int K) { return (size_t)WARPS * NBUF * BOX_BYTES + x_bytes(K) + (WARPS * NBUF + 1 + (WARPS + 1) / 2) * sizeof(uint64_t) + 1024; }
};

template <int V>
struct IntC {
    static constexpr int value = V;
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_skinny_f64_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapX, double* __restrict__ D, int M, int N, int K,
                       int64_t ldd, const double* __restrict__ Cin, int64_t ldc, unsigned long long* trace = nullptr)
{
    // development aid (tools/skinny_probe): per-warp %globaltimer stamps -- [0] entry, [1] X staged, [2] first box landed, [3+i] item i stored
    auto stamp = [&](int slot) {
        if (trace && (threadIdx.x & 31) == 0 && slot < 12) {
            unsigned long long tns;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
            trace[((size_t)blockIdx.x * Cfg::WARPS + (threadIdx.x >> 5)) * 12 + slot] = tns;
        }
    };
    stamp(0);
    // programmatic dependent launch (see the team kernel below): no-ops unless the host launched with stream serialisation
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    constexpr int NI = Cfg::NI, KC = Cfg::KC, NBUF = Cfg::NBUF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));  // swizzle atoms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    unsigned char* myA = base + (size_t)warp * NBUF * Cfg::BOX_BYTES;                                 // this warp's two boxes
    double* sX = reinterpret_cast<double*>(base + (size_t)Cfg::WARPS * NBUF * Cfg::BOX_BYTES);
    const int ksteps = K >> 2;  // K % 8 == 0 (host-checked): whole pairs of k-steps
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + (size_t)ksteps * NI * 32);
    uint64_t* full = bars + warp * NBUF;

    // Anti-phase handshake (FLAGS bit 4).  The G warps that share an SM sub-partition (warp, warp + 4, ...) drift into lock-step
    // and then reach their item boundaries together, leaving the DMMA pipe without work (measured).  With the
    // handshake, the warp of rank r starts its item j only when the warp of rank r - 1 is 1/G of the way through ITS item j
    // (rank 0: through item j - 1 of rank G - 1): the boundaries of the G warps stay 1/G of a period apart.  Progress counters
    // (in G-ths of an item) live in shared memory; a warp that has run out of items publishes "infinitely far".
    constexpr int G = Cfg::WARPS / 4;
    volatile int* prog = reinterpret_cast<volatile int*>(bars + Cfg::WARPS * NBUF + 1);
    if (tid < Cfg::WARPS) prog[tid] = 0;
    const int rank = warp >> 2, prev_warp = ((rank + G - 1) % G) * 4 + (warp & 3);
    const int prog_scale = (G << 16) / (ksteps > 0 ? ksteps : 1);  // k-steps -> G-ths of an item, 16.16 fixed point

    const int nblocks = (M + 15) >> 4, nchunks = (K + KC - 1) / KC;
    const int W = gridDim.x * Cfg::WARPS;
    const int wid = warp * gridDim.x + blockIdx.x;  // CTA index fastest: every SM sub-partition gets the same number of items (+-1)
    // items of this warp: `rounds` whole blocks wid + i W, then the remainder of the block list as half blocks h = wid (+ W)
    const int rounds = nblocks / W, nhalves = 2 * (nblocks - rounds * W);
    const int nhalf = (wid < nhalves ? 1 : 0) + (wid + W < nhalves ? 1 : 0);
    const int nitems = rounds + nhalf;
    const bool halves_first = (Cfg::FLAGS & 1) && ((warp >> 2) & 1) && nhalf > 0;
    auto item = [&](int j, int& blk, int& half) {  // half: -1 = whole block, else which half of the column tiles
        const int jf = halves_first ? j - nhalf : j, jh = halves_first ? j : j - rounds;
        if (jf >= 0 && jf < rounds) {
            blk = wid + jf * W;
            half = -1;
        } else {
            const int h = wid + jh * W;
            blk = rounds * W + (h >> 1);
            half = h & 1;
        }
    };

    // ---- X first: ONE 3-D TMA box (t, n, s) -> sX[s][n][t], which is fragment-major: the 32 lane values X[4s + t][8 ni + g] of
    //      a fragment are contiguous in lane order (conflict-free, address = base + immediate).  Columns n >= N are zero-filled
    //      by TMA (their accumulators are never stored).  It is requested before any A box so that it is not queued behind
    //      the start-up burst of A requests. ----
    uint64_t* xbar = bars + Cfg::WARPS * NBUF;
    if (tid == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapA);
        mbar_init(xbar, 1);
        mbar_fence_init();
        mbar_expect_tx(xbar, (uint32_t)(ksteps * Cfg::BN * 32));
        tma_load_3d_hint(sX, &mapX, xbar, 0, 0, 0, kL2EvictLast);
    }
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        mbar_fence_init();
    }
    __syncwarp();
    int ij = 0, ic = 0;  // next box to REQUEST: item ij, chunk ic
    auto request = [&](int buf) {
        if (ij < nitems && lane == 0 && !(Cfg::FLAGS & 8)) {
            int blk, half;
            item(ij, blk, half);
            mbar_expect_tx(&full[buf], Cfg::BOX_BYTES);
            tma_load_4d(myA + buf * Cfg::BOX_BYTES, &mapA, &full[buf], blk * 16, 0, 0, ic * (KC / 8));
        }
        if (++ic == nchunks) { ic = 0; ++ij; }
    };
    request(0);
    if (!(Cfg::FLAGS & 2)) request(1);
    __syncthreads();  // xbar is initialised for everybody
    mbar_wait(xbar, 0);
    if (Cfg::FLAGS & 2) request(1);
    stamp(1);

    const uint32_t sB = smem_u32(sX + lane);                    // + ((k0/4 + s) * NI + ni) * 256
    // rows (2g, 2g+1) of column k = 4s + t of a box: row R = 8 (s >> 1) + 2t + (s & 1), chunk g ^ (R & 7)
    const uint32_t sA0 = smem_u32(myA);
    const uint32_t offA[2] = {(uint32_t)((2 * t) * 128 + ((g ^ (2 * t)) << 4)), (uint32_t)((2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4))};
    // Fast path (warp-uniform): a block whose 16 rows all exist and whose D (and C) columns are 16-byte aligned moves through
    // 16-byte accesses off one base pointer -- the element-wise path costs ~430 instructions per block.
    const bool vec_ok = (ldd & 1) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 &&
                        (!ACC || ((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0));

    int buf = 0, done = 0;
    uint32_t phase = 0;
    // One item: NIC column tiles starting at tile NI0 of block blk -- all k chunks, then the stores.
    auto process = [&](auto ni0_c, auto nic_c, int blk) {
        constexpr int NI0 = decltype(ni0_c)::value, NIC = decltype(nic_c)::value;
        double acc[2][NIC][2];
        if ((Cfg::FLAGS & 16) && G > 1) {
            const int need = G * (rank ? done : done - 1) + 1;
            if (need > 0)
                while (prog[prev_warp] < need) {}
        }
        const int m = blk * 16 + 2 * g;
        const int ncol0 = NI0 * 8 + 2 * t;  // this lane's first column; the others are ncol0 + 8 ni + c
        const bool fast = vec_ok && blk * 16 + 16 <= M;
        if constexpr (ACC) {
            if (fast) {
                const double* p = Cin + (int64_t)ncol0 * ldc + m;
#pragma unroll
                for (int ni = 0; ni < NIC; ++ni)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        double2 v = make_double2(0.0, 0.0);
                        if (ncol0 + ni * 8 + c < N) v = *reinterpret_cast<const double2*>(p + (int64_t)(ni * 8 + c) * ldc);
                        acc[0][ni][c] = v.x;
                        acc[1][ni][c] = v.y;
                    }
            } else {
#pragma unroll
                for (int ni = 0; ni < NIC; ++ni)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int n = ncol0 + ni * 8 + c;
                        double2 v = make_double2(0.0, 0.0);
                        if (n < N) {
                            const double* p = Cin + (size_t)n * ldc + m;
                            if (m < M) v.x = p[0];
                            if (m + 1 < M) v.y = p[1];
                        }
                        acc[0][ni][c] = v.x;
                        acc[1][ni][c] = v.y;
                    }
            }
        } else {
#pragma unroll
            for (int ni = 0; ni < NIC; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) acc[0][ni][c] = acc[1][ni][c] = -0.0;
        }
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            if (!(Cfg::FLAGS & 8)) mbar_wait(&full[buf], phase);
            if (done == 0 && chunk == 0) stamp(2);
            const int s0 = chunk * (KC / 4);
            const int steps = min(ksteps - s0, KC / 4);  // even: K % 8 == 0
            const uint32_t pa = sA0 + buf * Cfg::BOX_BYTES, pb = sB + (s0 * NI + NI0) * 256;
            double2 a[2];
            double b[2][NIC];
            auto load = [&](int s, int which) {
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a[which].x), "=d"(a[which].y) : "r"(pa + offA[which] + (s >> 1) * 1024));
#pragma unroll
                for (int ni = 0; ni < NIC; ++ni) b[which][ni] = lds_f64(pb + (s * NI + ni) * 256);
            };
            auto mma = [&](int which) {
#pragma unroll
                for (int ni = 0; ni < NIC; ++ni) {
                    dmma884(acc[0][ni][0], acc[0][ni][1], a[which].x, b[which][ni]);
                    dmma884(acc[1][ni][0], acc[1][ni][1], a[which].y, b[which][ni]);
                }
            };
            load(0, 0);
#pragma unroll 2
            for (int s = 0; s < steps; s += 2) {
                load(s + 1, 1);
                mma(0);
                if (s + 2 < steps) load(s + 2, 0);
                mma(1);
                if ((Cfg::FLAGS & 16) && lane == 0) {
                    int frac = ((s0 + s + 2) * prog_scale) >> 16;
                    prog[warp] = G * done + (frac < G ? frac : G - 1);
                }
            }
            __syncwarp();  // every lane has read the box: it may be overwritten
            request(buf);
            if (++buf == NBUF) { buf = 0; phase ^= 1; }
        }
        if (Cfg::FLAGS & 4) {  // keep the accumulators alive without storing them
            double sum = 0;
#pragma unroll
            for (int ni = 0; ni < NIC; ++ni) sum += acc[0][ni][0] * acc[1][ni][1];
            if (sum == 1.2345678) D[m] = sum;
        } else if (fast) {
            double* p = D + (int64_t)ncol0 * ldd + m;
#pragma unroll
            for (int ni = 0; ni < NIC; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    if (ncol0 + ni * 8 + c < N) *reinterpret_cast<double2*>(p + (int64_t)(ni * 8 + c) * ldd) = make_double2(acc[0][ni][c], acc[1][ni][c]);
        } else {
#pragma unroll
            for (int ni = 0; ni < NIC; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = ncol0 + ni * 8 + c;
                    if (n >= N) continue;
                    double* p = D + (size_t)n * ldd + m;
                    if (m < M) p[0] = acc[0][ni][c];
                    if (m + 1 < M) p[1] = acc[1][ni][c];
                }
        }
        stamp(3 + done);
        ++done;
        if ((Cfg::FLAGS & 16) && lane == 0) prog[warp] = G * done;
    };
    for (int j = 0; j < nitems; ++j) {
        int blk, half;
        item(j, blk, half);
        if (half < 0) process(IntC<0>{}, IntC<NI>{}, blk);
        else if (half == 0) process(IntC<0>{}, IntC<NI / 2>{}, blk);
        else process(IntC<NI / 2>{}, IntC<NI / 2>{}, blk);
    }
    if ((Cfg::FLAGS & 16) && lane == 0) prog[warp] = 0x7fffffff;
}

// =====================================================================================================================
// X IN REGISTERS (K = 32 or 64, N <= 64): the variant the per-instruction profile of the kernel above asked for.
//
// ncu, one warp per SM sub-partition (profiles/r2_ncu_skinny_lone_warp.txt): the warp is never throttled by the DMMA pipe; its
// time goes to the fixed issue cadence of its own instructions -- a DMMA.8x8x4 holds the warp for ~16-18 clk and every LDS
// (0.56 per DMMA above: eight X fragments + one A pair per k-step), its write-after-read scoreboard waits and the loop
// arithmetic ADD to that instead of overlapping with it: 72 % of the pipe for a lone warp, 80 % for two, 82 % for three.
// X is the same for every row block, so here it does not come from shared memory at all: each warp owns ONE half of the
// columns (32 = 4 column tiles) for the whole kernel and keeps its X fragments -- KSTEPS x 4 doubles per lane, 128
// registers for K = 64 -- in REGISTERS, loaded once from global memory.  The k loop is straight-line code: per k-step one
// LDS.128 (rows 2g, 2g+1 of A) and 8 DMMAs (0.125 loads per DMMA).  Everything else is the design above: warp-private
// TMA boxes (16 rows x K), 128B-swizzled 4-D layout, -0.0 start, 16-byte stores straight from the accumulators.  Items are
// (row block, column half); the two warps (2i, 2i+1) of a CTA walk the same blocks, so the second fetch of a box is an L2
// hit, and 4096 blocks over 592 warps per half are 6.92 rounds of 2048 pipe-clocks: 1 % of round quantisation instead of 13 %.
// =====================================================================================================================
template <int KSTEPS_, int WARPS_, int NBUF_ = 2, int NG_ = 4>
struct SkinnyRegCfg {
    static constexpr int KSTEPS = KSTEPS_, K = 4 * KSTEPS_, WARPS = WARPS_, THREADS = WARPS_ * 32, NBUF = NBUF_;
    static constexpr int NG = NG_, BN = 64;          // column tiles per warp item (4: two column halves; 2: four quarters), columns covered
    static constexpr int BOX_BYTES = 16 * K * 8;     // 16 rows x K columns of doubles
    static_assert(KSTEPS_ % 2 == 0 && BOX_BYTES % 1024 == 0, "a box is a whole number of swizzle atoms");
    static constexpr size_t SMEM = (size_t)WARPS * NBUF * BOX_BYTES + (size_t)WARPS * NBUF * sizeof(uint64_t) + 1024;
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_skinny_xreg_f64_kernel(const __grid_constant__ CUtensorMap mapA, const double* __restrict__ X, int64_t ldx, double* __restrict__ D, int M, int N,
                            int64_t ldd, const double* __restrict__ Cin, int64_t ldc)
{
    constexpr int KSTEPS = Cfg::KSTEPS, NG = Cfg::NG, NBUF = Cfg::NBUF;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");  // programmatic dependent launch, as in the team kernel below
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));  // swizzle atoms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    unsigned char* myA = base + (size_t)warp * NBUF * Cfg::BOX_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)Cfg::WARPS * NBUF * Cfg::BOX_BYTES) + warp * NBUF;

    // this warp's column half and its place among the warps of that half (CTA index fastest: neighbouring blocks are in flight together)
    constexpr int GW = 8 * NG;                       // columns per group
    const int halves = min((N + GW - 1) / GW, 64 / GW);  // column groups that exist ("halves" when NG = 4)
    const int half = warp % halves;
    const int hw = (warp / halves) * gridDim.x + blockIdx.x;
    const int HW = (Cfg::WARPS / halves) * gridDim.x;
    if (warp >= halves * (Cfg::WARPS / halves)) return;  // three column groups on eight warps: the two surplus warps sit the launch out
    const int nblocks = (M + 15) >> 4;

    if (lane == 0) {
        if (warp == 0) tma_prefetch_desc(&mapA);
#pragma unroll
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        mbar_fence_init();
    }
    __syncwarp();
    int rq = hw;  // next block to request
    auto request = [&](int buf) {
        if (rq < nblocks && lane == 0) {
            mbar_expect_tx(&full[buf], Cfg::BOX_BYTES);
            tma_load_4d(myA + buf * Cfg::BOX_BYTES, &mapA, &full[buf], rq * 16, 0, 0, 0);
        }
        rq += HW;
    };
    request(0);
    // X fragments of this warp's columns: lane (g, t) holds X[4s + t][32 half + 8 ni + g] for every k-step s and column tile ni.
    // Columns >= N read as 0 (their accumulators are never stored).
    double xr[KSTEPS][NG];
    {
        const double* xp = X + t;
#pragma unroll
        for (int ni = 0; ni < NG; ++ni) {
            const int n = GW * half + 8 * ni + g;
            const double* col = xp + (int64_t)n * ldx;
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) xr[s][ni] = n < N ? __ldg(col + 4 * s) : 0.0;
        }
    }
#pragma unroll
    for (int b = 1; b < NBUF; ++b) request(b);  // after the X loads: first boxes are served ahead of the later ones

    // rows (2g, 2g+1) of column k = 4s + t of a box: row R = 8 (s >> 1) + 2t + (s & 1), chunk g ^ (R & 7)
    const uint32_t sA0 = smem_u32(myA);
    const uint32_t offA[2] = {(uint32_t)((2 * t) * 128 + ((g ^ (2 * t)) << 4)), (uint32_t)((2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4))};
    const bool vec_ok = (ldd & 1) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 &&
                        (!ACC || ((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0));
    const int ncol0 = GW * half + 2 * t;  // this lane's first column; the others are ncol0 + 8 ni + c
    const bool cols_full = GW * half + GW <= N;

    // (Tried and measured to change nothing: software-pipelining the loop ACROSS items -- first fragment of the next item and the
    // wait for its box before the last k-step, stores of the previous item behind the second k-step, two accumulator sets.  The
    // per-item instructions cost the warp the same issue time wherever they stand; only the other warp of the sub-partition
    // hides them: tensor pipe 78 % with one warp per sub-partition, 90 % with two, profiles/r2_ncu_skinny_xreg_*.txt.)
    int buf = 0;
    uint32_t phase = 0;
    for (int blk = hw; blk < nblocks; blk += HW) {
        const int m = blk * 16 + 2 * g;
        const bool fast = vec_ok && blk * 16 + 16 <= M;
        double acc[2][NG][2];
        if constexpr (ACC) {
#pragma unroll
            for (int ni = 0; ni < NG; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = ncol0 + ni * 8 + c;
                    double2 v = make_double2(0.0, 0.0);
                    if (n < N) {
                        const double* p = Cin + (size_t)n * ldc + m;
                        if (fast) {
                            v = *reinterpret_cast<const double2*>(p);
                        } else {
                            if (m < M) v.x = p[0];
                            if (m + 1 < M) v.y = p[1];
                        }
                    }
                    acc[0][ni][c] = v.x;
                    acc[1][ni][c] = v.y;
                }
        }
        mbar_wait(&full[buf], phase);
        const uint32_t pa = sA0 + buf * Cfg::BOX_BYTES;
        // A fragments two k-steps ahead in FOUR register pairs: a load never overwrites a pair the DMMAs just issued still read
        double2 a[4];
        auto load = [&](int s) {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a[s & 3].x), "=d"(a[s & 3].y) : "r"(pa + offA[s & 1] + (s >> 1) * 1024));
        };
        load(0);
        load(1);
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s) {
            if (s + 2 < KSTEPS) load(s + 2);
            if (!ACC && s == 0) {  // the chain starts from -0.0: no accumulator initialisation pass
                const double nz = -0.0;
#pragma unroll
                for (int ni = 0; ni < NG; ++ni) {
                    dmma884_from(acc[0][ni][0], acc[0][ni][1], a[0].x, xr[0][ni], nz, nz);
                    dmma884_from(acc[1][ni][0], acc[1][ni][1], a[0].y, xr[0][ni], nz, nz);
                }
            } else {
#pragma unroll
                for (int ni = 0; ni < NG; ++ni) {
                    dmma884(acc[0][ni][0], acc[0][ni][1], a[s & 3].x, xr[s][ni]);
                    dmma884(acc[1][ni][0], acc[1][ni][1], a[s & 3].y, xr[s][ni]);
                }
            }
        }
        __syncwarp();  // every lane has read the box: it may be overwritten
        request(buf);
        if (++buf == NBUF) { buf = 0; phase ^= 1; }
        if (fast && cols_full) {  // 8 16-byte stores off one pointer, no predicates
            double* p = D + (int64_t)ncol0 * ldd + m;
            const int64_t ldd8 = 8 * ldd;
#pragma unroll
            for (int ni = 0; ni < NG; ++ni) {
                *reinterpret_cast<double2*>(p) = make_double2(acc[0][ni][0], acc[1][ni][0]);
                *reinterpret_cast<double2*>(p + ldd) = make_double2(acc[0][ni][1], acc[1][ni][1]);
                p += ldd8;
            }
        } else {
#pragma unroll
            for (int ni = 0; ni < NG; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = ncol0 + ni * 8 + c;
                    if (n >= N) continue;
                    double* p = D + (size_t)n * ldd + m;
                    if (fast) {
                        *reinterpret_cast<double2*>(p) = make_double2(acc[0][ni][c], acc[1][ni][c]);
                    } else {
                        if (m < M) p[0] = acc[0][ni][c];
                        if (m + 1 < M) p[1] = acc[1][ni][c];
                    }
                }
        }
    }
}

// =====================================================================================================================
// X IN REGISTERS, BOXES SHARED BY A TEAM OF WARPS.  The kernel above fetches every A box once per column group; the second fetch
// is an L2 hit only while the warps that share a row block stay close in time and A fits in L2: at 10^6 rows the two halves
// drift apart and DRAM traffic doubles (293 us against 258), and with four column QUARTERS per block -- 56 fewer registers,
// shorter start-up, 21.5 us at 65536 rows -- it quadruples (512 us).  Here the TEAM = 64 / (8 NG) warps that cover the column
// groups of a row block consume the SAME box: member 0 requests it (one TMA per block), every member waits on the box's `full`
// mbarrier, multiplies its own columns and arrives on the box's `empty` mbarrier; member 0 re-uses a buffer only when its
// `empty` phase has completed, and it asks one item later than it could, so that it practically never waits for a team mate.
// With quarters a warp needs ~120 registers: 16 warps per SM (four per sub-partition) hide each other's non-DMMA instructions.
// =====================================================================================================================
template <int KSTEPS_, int WARPS_, int NBUF_ = 3, int NG_ = 2, int BN_ = 64>
struct SkinnyTeamCfg {
    static constexpr int KSTEPS = KSTEPS_, K = 4 * KSTEPS_, WARPS = WARPS_, THREADS = WARPS_ * 32, NBUF = NBUF_, NG = NG_;
    static constexpr int TEAM = BN_ / (8 * NG_), TEAMS = WARPS_ / TEAM, BN = BN_;  // BN = 32: teams of two warps for N <= 32
    static constexpr int BOX_BYTES = 16 * K * 8;
    static_assert(WARPS_ % TEAM == 0 && NBUF_ >= 2, "whole teams, at least two boxes per team");
    static_assert(KSTEPS_ % 2 == 0 && BOX_BYTES % 1024 == 0, "a box is a whole number of swizzle atoms");
    static constexpr size_t SMEM = (size_t)TEAMS * NBUF * BOX_BYTES + (size_t)TEAMS * NBUF * 2 * sizeof(uint64_t) + 1024;
};

template <typename Cfg, bool ACC>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_skinny_team_f64_kernel(const __grid_constant__ CUtensorMap mapA, const double* __restrict__ X, int64_t ldx, double* __restrict__ D, int M, int N,
                            int64_t ldd, const double* __restrict__ Cin, int64_t ldc)
{
    constexpr int KSTEPS = Cfg::KSTEPS, NG = Cfg::NG, NBUF = Cfg::NBUF, TEAM = Cfg::TEAM, GW = 8 * NG;  // GW columns per member, TEAM GW = Cfg::BN >= N
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));  // swizzle atoms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int team = warp / TEAM, member = warp % TEAM;
    unsigned char* teamA = base + (size_t)team * NBUF * Cfg::BOX_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)Cfg::TEAMS * NBUF * Cfg::BOX_BYTES) + team * 2 * NBUF;
    uint64_t* empty = full + NBUF;

    const int nblocks = (M + 15) >> 4;
    const int tg = team * gridDim.x + blockIdx.x, TG = Cfg::TEAMS * gridDim.x;  // CTA index fastest: neighbouring blocks are in flight together
    const bool has_cols = GW * member < N;  // a member without columns only keeps the box protocol going

    // Programmatic dependent launch: when the host launches this kernel with programmatic stream serialisation, the NEXT kernel
    // of the stream may be scheduled while this one runs (its CTAs take the SMs as they become free) and does its prologue --
    // barrier init, tensor-map prefetch -- early; griddepcontrol.wait then holds it until every earlier kernel has completed
    // and flushed, BEFORE its first global-memory access.  Without the launch attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (member == 0 && lane == 0) {
        if (warp == 0) tma_prefetch_desc(&mapA);
#pragma unroll
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&full[b], 1);
            mbar_init(&empty[b], TEAM);
        }
        mbar_fence_init();
    }
    __syncthreads();  // the team's barriers exist before any member waits on them
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    // item i of the team = block tg + i TG, box buffer i % NBUF, parity (i / NBUF) & 1 on both barriers of that buffer
    auto request = [&](int i) {  // member 0 only
        const int blk = tg + i * TG;
        if (blk < nblocks && lane == 0) {
            const int b = i % NBUF;
            mbar_expect_tx(&full[b], Cfg::BOX_BYTES);
            tma_load_4d(teamA + b * Cfg::BOX_BYTES, &mapA, &full[b], blk * 16, 0, 0, 0);
        }
    };
    if (member == 0) request(0);
    // X fragments of this member's columns: lane (g, t) holds X[4s + t][GW member + 8 ni + g]; columns >= N read as 0
    double xr[KSTEPS][NG];
    {
        const double* xp = X + t;
#pragma unroll
        for (int ni = 0; ni < NG; ++ni) {
            const int n = GW * member + 8 * ni + g;
            const double* col = xp + (int64_t)n * ldx;
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) xr[s][ni] = n < N ? __ldg(col + 4 * s) : 0.0;
        }
    }
    if (member == 0) {
#pragma unroll
        for (int b = 1; b < NBUF; ++b) request(b);  // after the X loads: first boxes are served ahead of the later ones
    }

    const uint32_t sA0 = smem_u32(teamA);
    const uint32_t offA[2] = {(uint32_t)((2 * t) * 128 + ((g ^ (2 * t)) << 4)), (uint32_t)((2 * t + 1) * 128 + ((g ^ (2 * t + 1)) << 4))};
    const bool vec_ok = (ldd & 1) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 &&
                        (!ACC || ((ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(Cin) & 15) == 0));
    const int ncol0 = GW * member + 2 * t;
    const bool cols_full = GW * member + GW <= N;

    for (int i = 0, blk = tg; blk < nblocks; ++i, blk += TG) {
        const int b = i % NBUF;
        const uint32_t par = (uint32_t)(i / NBUF) & 1u;
        const int m = blk * 16 + 2 * g;
        const bool fast = vec_ok && blk * 16 + 16 <= M;
        double acc[2][NG][2];
        if constexpr (ACC) {
#pragma unroll
            for (int ni = 0; ni < NG; ++ni)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int n = ncol0 + ni * 8 + c;
                    double2 v = make_double2(0.0, 0.0);
                    if (n < N) {
                        const double* p = Cin + (size_t)n * ldc + m;
                        if (fast) {
                            v = *reinterpret_cast<const double2*>(p);
                        } else {
                            if (m < M) v.x = p[0];
                            if (m + 1 < M) v.y = p[1];
                        }
                    }
                    acc[0][ni][c] = v.x;
                    acc[1][ni][c] = v.y;
                }
        }
        mbar_wait(&full[b], par);
        if (has_cols) {
            const uint32_t pa = sA0 + b * Cfg::BOX_BYTES;
            double2 a[4];
            auto load = [&](int s) {
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a[s & 3].x), "=d"(a[s & 3].y) : "r"(pa + offA[s & 1] + (s >> 1) * 1024));
            };
            load(0);
            load(1);
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
                if (s + 2 < KSTEPS) load(s + 2);
                if (!ACC && s == 0) {  // the chain starts from -0.0: no accumulator initialisation pass
                    const double nz = -0.0;
#pragma unroll
                    for (int ni = 0; ni < NG; ++ni) {
                        dmma884_from(acc[0][ni][0], acc[0][ni][1], a[0].x, xr[0][ni], nz, nz);
                        dmma884_from(acc[1][ni][0], acc[1][ni][1], a[0].y, xr[0][ni], nz, nz);
                    }
                } else {
#pragma unroll
                    for (int ni = 0; ni < NG; ++ni) {
                        dmma884(acc[0][ni][0], acc[0][ni][1], a[s & 3].x, xr[s][ni]);
                        dmma884(acc[1][ni][0], acc[1][ni][1], a[s & 3].y, xr[s][ni]);
                    }
                }
            }
        }
        __syncwarp();  // every lane of this member has read the box
        if (lane == 0) mbar_arrive(&empty[b]);
        // member 0: the buffer of the PREVIOUS item gets item i - 1 + NBUF once the whole team has released it
        if (member == 0 && i >= 1) {
            mbar_wait(&empty[(i - 1) % NBUF], (uint32_t)((i - 1) / NBUF) & 1u);
            request(i - 1 + NBUF);
        }
        if (has_cols) {
            if (fast && cols_full) {  // 2 NG 16-byte stores off one pointer, no predicates
                double* p = D + (int64_t)ncol0 * ldd + m;
                const int64_t ldd8 = 8 * ldd;
#pragma unroll
                for (int ni = 0; ni < NG; ++ni) {
                    *reinterpret_cast<double2*>(p) = make_double2(acc[0][ni][0], acc[1][ni][0]);
                    *reinterpret_cast<double2*>(p + ldd) = make_double2(acc[0][ni][1], acc[1][ni][1]);
                    p += ldd8;
                }
            } else {
#pragma unroll
                for (int ni = 0; ni < NG; ++ni)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int n = ncol0 + ni * 8 + c;
                        if (n >= N) continue;
                        double* p = D + (size_t)n * ldd + m;
                        if (fast) {
                            *reinterpret_cast<double2*>(p) = make_double2(acc[0][ni][c], acc[1][ni][c]);
                        } else {
                            if (m < M) p[0] = acc[0][ni][c];
                            if (m + 1 < M) p[1] = acc[1][ni][c];
                        }
                    }
            }
        }
    }
}

}  // namespace jb

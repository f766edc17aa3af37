// capi.cu -- the C ABI of libjblas_b200.so (see include/jblas_b200.h) and the host-side planner.
//
// Host side of the replacement for jmul! (src/gemm.jl:244-348): where the reference's generator body picks a
// register tile (pick_kernel_size, src/gemm.jl:258) and emits the two tile loops, this file picks a CTA tile
// and kernel family per shape (plan()) and launches one grid whose 1-D block index is rasterised over the
// tile grid.  No CPU fallback exists anywhere in this file.
#include "../../../include/jblas_b200.h"

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "fastmul_batched.cuh"
#include "gemm_dmma.cuh"
#include "gemm_dmma_tma.cuh"
#include "gemm_f32_tma.cuh"
#include "gemm_simt.cuh"
#include "gemm_simt_f32x2.cuh"
#include "gemm_skinny.cuh"
#include "gemm_tf32x3.cuh"

using namespace jb;

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? JBLAS_B200_ENOMEM : JBLAS_B200_ECUDA,       \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// contexts.  jblas_b200_init binds the process to ONE GPU (one process per GPU: the torch.distributed mode); the
// single-process multi-GPU entries (jblas_b200_mgpu_*) create a context on every GPU they drive.  Every host routine
// below works on cur(): the context the calling thread is currently issuing to.
// ---------------------------------------------------------------------------------------------------------
static constexpr int kMaxDevices = 16;
// Work counters of the persistent TMA kernels: {next tile, CTAs done} pairs, all zero at rest (a kernel leaves its pair
// zeroed, gemm_dmma_tma.cuh).  Two LIVE kernels must never share a pair.  Kernels on one stream never overlap, so a pair
// belongs to a STREAM: the first kStreamSlots pairs are handed out per stream handle (cudaStreamPerThread: per thread);
// launches captured into a CUDA graph may replay anywhere later, so those take pairs from a separate ring instead
// (replays of ONE graph must not overlap each other).  When the stream table is full the launch falls back to the
// static tile stride (ctr == nullptr), which needs no counter.
static constexpr int kStreamSlots = 4096, kCaptureSlots = 4096, kTileCtrSlots = kStreamSlots + kCaptureSlots;
struct Context {
    int device = -1;
    int num_sms = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D for the host-pointer entry points
    cudaStream_t d2h_stream = nullptr;   // D2H (its own stream: PCIe is full duplex, one stream would serialise the two)
    cudaStream_t peer_stream = nullptr;  // multi-GPU: copy-engine pulls of A panels from the other GPUs over NVLink
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr;
    void* ws[3] = {nullptr, nullptr, nullptr};  // device staging for D, A, X (host-pointer entry points)
    size_t ws_bytes[3] = {0, 0, 0};
    float last_ms = 0.f;
    bool attrs_set = false;
    int* tile_ctr = nullptr;  // kTileCtrSlots pairs
    std::mutex slot_mu;
    std::map<cudaStream_t, int> slot_of;
    uint32_t capture_seq = 0;
    // rates of this box, measured once (calibrate()): they size the panel ramp of the host-pointer pipeline
    double h2d_bytes_per_s = 0, dmma_flops_per_s = 0, ffma2_flops_per_s = 0;
};
static Context g_ctxs[kMaxDevices];
static int g_primary = -1;                     // device bound by jblas_b200_init
static int g_mgpu = 0;                         // GPUs 0..g_mgpu-1 carry a context and peer access (jblas_b200_mgpu_init)
static thread_local Context* t_cur = nullptr;  // set while a multi-GPU routine issues work to another device
static Context& cur() { return t_cur ? *t_cur : g_ctxs[g_primary < 0 ? 0 : g_primary]; }
#define g_ctx (cur())
static std::mutex g_mu;  // host-pointer calls are serialised (SURVEY 8b "Threading")
static std::atomic<int64_t> g_launches{0};

static int require_init()
{
    if (g_primary < 0 && !t_cur) return fail(JBLAS_B200_ENOTINIT, "jblas_b200_init has not been called (no CPU fallback exists)");
    // entries may be called from any host thread, and a fresh thread's current device is 0, not the bound one
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d != cur().device) CUDA_TRY(cudaSetDevice(cur().device));
    return 0;
}
// The operands of a device entry must live on the bound GPU: a pointer from another device would be dereferenced by
// kernels launched here with this context's counters and streams (illegal address instead of a clean error).
static int check_on_device(const void* p, const char* what)
{
    if (!p) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;  // unknown to the runtime (e.g. a mapping the driver API made): not ours to reject
    }
    if (a.type == cudaMemoryTypeDevice && a.device != cur().device)
        return fail(JBLAS_B200_EINVAL, "%s lives on device %d but the library is bound to device %d", what, a.device, cur().device);
    return 0;
}

static int* tile_counter_for(cudaStream_t s)
{
    Context& c = cur();
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) {
        cudaGetLastError();  // legacy default stream while another stream captures: not capturing itself
        cap = cudaStreamCaptureStatusNone;
    }
    std::lock_guard<std::mutex> lk(c.slot_mu);
    if (cap == cudaStreamCaptureStatusActive) return c.tile_ctr + 2 * (kStreamSlots + (int)(c.capture_seq++ % kCaptureSlots));
    if (s == cudaStreamPerThread) {
        static thread_local int per_thread_slot[kMaxDevices];  // 0 = none yet, else slot + 1
        int& mine = per_thread_slot[c.device];
        if (!mine) {
            if ((int)c.slot_of.size() >= kStreamSlots) return nullptr;
            const int slot = (int)c.slot_of.size();
            c.slot_of[(cudaStream_t)(uintptr_t)(0x100000000ull + slot)] = slot;  // reserves the slot (key is no real handle)
            mine = slot + 1;
        }
        return c.tile_ctr + 2 * (mine - 1);
    }
    auto it = c.slot_of.find(s);
    if (it == c.slot_of.end()) {
        if ((int)c.slot_of.size() >= kStreamSlots) return nullptr;
        it = c.slot_of.emplace(s, (int)c.slot_of.size()).first;
    }
    return c.tile_ctr + 2 * it->second;
}

// Launch with programmatic stream serialisation (the kernel calls griddepcontrol.wait before its first global access): the next
// such launch on the stream is scheduled while this one runs.  JBLAS_B200_NO_PDL=1 launches plainly (A/B measurements).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t s, Args... args)
{
    static int no_pdl = -1;
    if (no_pdl < 0) {
        const char* e = getenv("JBLAS_B200_NO_PDL");
        no_pdl = (e && atoi(e)) ? 1 : 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------------
// kernel registry
// ---------------------------------------------------------------------------------------------------------
enum Family { FAM_SIMT = 0, FAM_DMMA = 1, FAM_TF32X3 = 2 };

// Cin/ldc: the matrix an accumulating launch starts from (== D for D += A*X; another matrix for D = A*X + C); unused otherwise
typedef int (*LaunchFn)(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                        int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc);

struct KernelInfo {
    const char* name;
    int dtype;   // JBLAS_B200_DT_*
    int family;  // Family
    int bm, bn, bk, stages, threads;
    size_t smem;
    float eff;           // measured throughput on many-wave shapes relative to the family's cp.async 128x128 kernel (sweep.json)
    bool needs_aligned;  // TMA kernels: 16-byte aligned bases and leading dimensions only
    bool persistent;     // grid = min(tiles, #SMs * ctas_per_sm), CTAs loop over the rasterised tile list
    int ctas_per_sm;     // resident CTAs per SM the kernel is built for (persistent kernels)
    LaunchFn launch[2][2];  // [aligned][acc]
    cudaError_t (*set_attr)();
    int (*query_occ)();  // non-persistent kernels: CTAs of the aligned variant the hardware keeps resident per SM
    bool has_ragged = false;  // needs_aligned kernels only: launch[0][*] is an element-wise staging variant, not an error
    int cluster = 1;          // CTAs (SMs) that work on ONE bm x bn tile together (2: tcgen05 cta_group::2 pair)
};
template <typename K>
static int occupancy_of(K kernel, int threads, size_t smem)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <typename T, typename Cfg, bool ALIGNED, bool ACC>
static int launch_simt(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                       int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    CUDA_TRY(launch_pdl(gemm_simt_kernel<T, Cfg, ALIGNED, ACC>, tiles_m * tiles_n, Cfg::THREADS, Cfg::SMEM, s, (T*)D, (const T*)A, (const T*)X, M, N, K, ldd, lda, ldx,
                        tiles_m, tiles_n, group_m, (const T*)Cin, ldc));
    return 0;
}
template <typename T, typename Cfg>
static cudaError_t attr_simt()
{
    cudaError_t e;
#define SET(AL, AC)                                                                                               \
    e = cudaFuncSetAttribute(gemm_simt_kernel<T, Cfg, AL, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                             (int)Cfg::SMEM);                                                                     \
    if (e != cudaSuccess) return e;
    SET(false, false) SET(false, true) SET(true, false) SET(true, true)
#undef SET
    return cudaSuccess;
}
template <typename Cfg, bool ALIGNED, bool ACC>
static int launch_simt_f32x2(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                             int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    CUDA_TRY(launch_pdl(gemm_simt_f32x2_kernel<Cfg, ALIGNED, ACC>, tiles_m * tiles_n, Cfg::THREADS, Cfg::SMEM, s, (float*)D, (const float*)A, (const float*)X, M, N, K,
                        ldd, lda, ldx, tiles_m, tiles_n, group_m, (const float*)Cin, ldc));
    return 0;
}
template <typename Cfg>
static cudaError_t attr_simt_f32x2()
{
    cudaError_t e;
#define SET(AL, AC)                                                                                               \
    e = cudaFuncSetAttribute(gemm_simt_f32x2_kernel<Cfg, AL, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)Cfg::SMEM);                                                                     \
    if (e != cudaSuccess) return e;
    SET(false, false) SET(false, true) SET(true, false) SET(true, true)
#undef SET
    return cudaSuccess;
}
template <typename Cfg, bool ALIGNED, bool ACC>
static int launch_dmma(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                       int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    CUDA_TRY(launch_pdl(gemm_dmma_kernel<Cfg, ALIGNED, ACC>, tiles_m * tiles_n, Cfg::THREADS, Cfg::SMEM, s, (double*)D, (const double*)A, (const double*)X, M, N, K,
                        ldd, lda, ldx, tiles_m, tiles_n, group_m, (const double*)Cin, ldc));
    return 0;
}
template <typename Cfg>
static cudaError_t attr_dmma()
{
    cudaError_t e;
#define SET(AL, AC)                                                                                               \
    e = cudaFuncSetAttribute(gemm_dmma_kernel<Cfg, AL, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                             (int)Cfg::SMEM);                                                                     \
    if (e != cudaSuccess) return e;
    SET(false, false) SET(false, true) SET(true, false) SET(true, true)
#undef SET
    return cudaSuccess;
}

// ---- TMA tensor maps (driver entry point resolved at run time: the library must load without libcuda.so.1) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

static int get_encode_tiled()
{
    if (g_encode_tiled) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return fail(JBLAS_B200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver (%s)", cudaGetErrorString(e));
    g_encode_tiled = (EncodeTiledFn)fn;
    return 0;
}

// 2-D column-major operand: dim0 = rows (contiguous), dim1 = cols (stride ld elements); box = box_r x box_c, 128B swizzle
static int make_tmap_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esize, uint64_t rows, uint64_t cols,
                        uint64_t ld, uint32_t box_r, uint32_t box_c, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    // cuTensorMapEncodeTiled costs 1-2 us, as much as the launch itself on small problems: callers that multiply the same
    // buffers again (every benchmark loop, every K panel of a pipeline) hit a small per-thread cache of encoded maps.
    // A tensor map is a pure function of these eight values, so a stale entry cannot exist.
    struct Entry { const void* base; uint64_t rows, cols, ld; uint32_t box_r, box_c; int dt, esize, swizzle; bool valid; CUtensorMap map; };
    static thread_local Entry cache[8] = {};
    static thread_local unsigned next_slot = 0;
    for (const Entry& e : cache)
        if (e.valid && e.base == base && e.rows == rows && e.cols == cols && e.ld == ld && e.box_r == box_r && e.box_c == box_c &&
            e.dt == (int)dt && e.esize == esize && e.swizzle == (int)swizzle) {
            *map = e.map;
            return 0;
        }
    if (int rc = get_encode_tiled()) return rc;
    cuuint64_t gdim[2] = {rows, cols};
    cuuint64_t gstride[1] = {ld * (uint64_t)esize};
    cuuint32_t box[2] = {box_r, box_c};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JBLAS_B200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    Entry& slot = cache[next_slot++ % 8];
    slot.base = base; slot.rows = rows; slot.cols = cols; slot.ld = ld; slot.box_r = box_r; slot.box_c = box_c;
    slot.dt = (int)dt; slot.esize = esize; slot.swizzle = (int)swizzle; slot.map = *map; slot.valid = true;
    return 0;
}

// 3-D: dim2 = product index of a batch (element stride `stride2`)
static int make_tmap_3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t batch, uint64_t ld, uint64_t stride2,
                        uint32_t box_r, uint32_t box_c)
{
    if (int rc = get_encode_tiled()) return rc;
    cuuint64_t gdim[3] = {rows, cols, batch};
    cuuint64_t gstride[2] = {ld * 8, stride2 * 8};
    cuuint32_t box[3] = {box_r, box_c, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JBLAS_B200_ECUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    return 0;
}

template <typename Cfg, bool ACC>
static int launch_dmma_tma(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                           int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    CUtensorMap mapA, mapX;
    if (int rc = make_tmap_2d(&mapA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 16, 16)) return rc;
    if (int rc = make_tmap_2d(&mapX, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, (uint64_t)K, (uint64_t)N, (uint64_t)ldx, 16, Cfg::BN)) return rc;
    int grid = tiles_m * tiles_n;
    if (grid > g_ctx.num_sms * Cfg::MIN_BLOCKS) grid = g_ctx.num_sms * Cfg::MIN_BLOCKS;
    // L2 eviction priority per operand.  Measured (ncu, 8192^3): marking X evict-first RAISES DRAM reads (9.97 vs 6.49 GB:
    // CTAs of a wave drift apart in k and the laggards then miss); JBLAS_B200_L2HINT selects a mode for experiments.
    static int hint_mode = -1;
    if (hint_mode < 0) {
        const char* e = getenv("JBLAS_B200_L2HINT");
        hint_mode = e ? atoi(e) : 0;
    }
    uint64_t pa = kL2EvictNormal, px = kL2EvictNormal;
    if (hint_mode == 1) { pa = kL2EvictLast; px = kL2EvictFirst; }
    else if (hint_mode == 2) { pa = kL2EvictLast; px = kL2EvictNormal; }
    else if (hint_mode == 3) { pa = kL2EvictLast; px = kL2EvictLast; }
    static int static_tiles = -1;
    if (static_tiles < 0) {
        const char* e = getenv("JBLAS_B200_STATIC_TILES");
        static_tiles = (e && atoi(e)) ? 1 : 0;
    }
    int* ctr = static_tiles ? nullptr : tile_counter_for(s);
    // programmatic dependent launch: the next product of the stream is scheduled while this one runs (1.7 us per call on small
    // shapes); the kernel waits for all earlier work before its first global access (counter, TMA, C)
    CUDA_TRY(launch_pdl(gemm_dmma_tma_kernel<Cfg, ACC, false, false>, grid, Cfg::THREADS, Cfg::SMEM, s, mapA, mapX, (double*)D, M, N, K, ldd, tiles_m, tiles_n,
                        group_m, pa, px, ctr, (const double*)Cin, ldc, 1, (int64_t)0, (const double*)nullptr, (const double*)nullptr, (int64_t)0, (int64_t)0));
    return 0;
}
template <typename Cfg, bool ACC>
static int launch_f32_tma(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                          int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    CUtensorMap mapA, mapX;
    if (int rc = make_tmap_2d(&mapA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)M, (uint64_t)K, (uint64_t)lda, Cfg::BM, Cfg::BK, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    if (int rc = make_tmap_2d(&mapX, X, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)K, (uint64_t)N, (uint64_t)ldx, Cfg::BK, Cfg::BN)) return rc;
    int grid = tiles_m * tiles_n;
    if (grid > g_ctx.num_sms) grid = g_ctx.num_sms;
    gemm_f32_tma_kernel<Cfg, ACC><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(mapA, mapX, (float*)D, M, N, K, ldd, tiles_m, tiles_n, group_m, tile_counter_for(s),
                                                                          (const float*)Cin, ldc);
    return 0;
}
template <typename Cfg>
static cudaError_t attr_f32_tma()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_f32_tma_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_f32_tma_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
}
// The same kernel with element-wise staging (RAGGED): operands whose base or leading dimension is not 16-byte aligned.
template <typename Cfg, bool ACC>
static int launch_dmma_tma_ragged(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                                  int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    static const CUtensorMap none = {};
    int grid = tiles_m * tiles_n;
    if (grid > g_ctx.num_sms * Cfg::MIN_BLOCKS) grid = g_ctx.num_sms * Cfg::MIN_BLOCKS;
    CUDA_TRY(launch_pdl(gemm_dmma_tma_kernel<Cfg, ACC, false, true>, grid, Cfg::THREADS, Cfg::SMEM, s, none, none, (double*)D, M, N, K, ldd, tiles_m, tiles_n, group_m,
                        kL2EvictNormal, kL2EvictNormal, (int*)nullptr, (const double*)Cin, ldc, 1, (int64_t)0, (const double*)A, (const double*)X, lda, ldx));
    return 0;
}
template <typename Cfg>
static cudaError_t attr_dmma_tma_ragged()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    return e;
}
// General tiled tensor map (rank 3 or 4, Float64) with a small per-thread cache, as make_tmap_2d.
static int make_tmap_nd(CUtensorMap* map, const void* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstride_bytes, const cuuint32_t* box,
                        CUtensorMapSwizzle swizzle)
{
    struct Entry { const void* base; cuuint64_t gdim[4], gstr[3]; cuuint32_t box[4]; int rank, swizzle; bool valid; CUtensorMap map; };
    static thread_local Entry cache[8] = {};
    static thread_local unsigned next_slot = 0;
    Entry key = {};
    key.base = base; key.rank = rank; key.swizzle = (int)swizzle;
    for (int i = 0; i < rank; ++i) { key.gdim[i] = gdim[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.gstr[i] = gstride_bytes[i];
    for (const Entry& e : cache)
        if (e.valid && e.base == key.base && e.rank == key.rank && e.swizzle == key.swizzle && !memcmp(e.gdim, key.gdim, sizeof(key.gdim)) &&
            !memcmp(e.gstr, key.gstr, sizeof(key.gstr)) && !memcmp(e.box, key.box, sizeof(key.box))) {
            *map = e.map;
            return 0;
        }
    if (int rc = get_encode_tiled()) return rc;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride_bytes, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JBLAS_B200_ECUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
    Entry& slot = cache[next_slot++ % 8];
    slot = key;
    slot.map = *map;
    slot.valid = true;
    return 0;
}

// Tall-skinny Float64 (gemm_skinny.cuh): N <= 64, K a multiple of 8 with X resident in shared memory, aligned operands.
using SK_n64 = SkinnyCfg<8, 12, 32, 3>;  // 12 warps, boxes of 16 rows x 32 k: 23.3 us at 65536 x 64 x 64 (profiles/r2_skinny_probe_v2.txt)
using SK_n32 = SkinnyCfg<4, 12, 32, 3>;
static constexpr int kSkinnyMaxK = 128, kSkinnyMaxN = 64;
static bool skinny_shape_ok(int64_t M, int64_t N, int64_t K) { return M >= 1 && N >= 1 && N <= kSkinnyMaxN && K >= 8 && K <= kSkinnyMaxK && K % 8 == 0; }
template <typename Cfg, bool ACC>
static int launch_skinny_cfg(double* D, const double* A, const double* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, cudaStream_t s,
                             const double* Cin, int64_t ldc)
{
    CUtensorMap mapA, mapX;
    {   // A as (m, s_lo, t, s_hi) with k = 8 s_hi + 4 s_lo + t; box = 16 rows x KC k
        const cuuint64_t gdim[4] = {(cuuint64_t)M, 2, 4, (cuuint64_t)(K / 8)};
        const cuuint64_t gstr[3] = {(cuuint64_t)(4 * lda * 8), (cuuint64_t)(lda * 8), (cuuint64_t)(8 * lda * 8)};
        const cuuint32_t box[4] = {16, 2, 4, (cuuint32_t)(Cfg::KC / 8)};
        if (int rc = make_tmap_nd(&mapA, A, 4, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {   // X as (t, n, s) with k = 4 s + t: ONE box = the whole matrix in fragment-major order
        const cuuint64_t gdim[3] = {4, (cuuint64_t)N, (cuuint64_t)(K / 4)};
        const cuuint64_t gstr[2] = {(cuuint64_t)(ldx * 8), 32};
        const cuuint32_t box[3] = {4, (cuuint32_t)Cfg::BN, (cuuint32_t)(K / 4)};
        if (int rc = make_tmap_nd(&mapX, X, 3, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
    }
    const int nblocks = (M + 15) / 16;
    int grid = (nblocks + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > g_ctx.num_sms) grid = g_ctx.num_sms;
    CUDA_TRY(launch_pdl(gemm_skinny_f64_kernel<Cfg, ACC>, grid, Cfg::THREADS, Cfg::smem(K), s, mapA, mapX, D, M, N, K, ldd, Cin, ldc, (unsigned long long*)nullptr));
    return 0;
}
template <bool ACC>
static int launch_skinny(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, int, int, int, cudaStream_t s,
                         const void* Cin, int64_t ldc)
{
    if (!skinny_shape_ok(M, N, K))
        return fail(JBLAS_B200_EUNSUPPORTED, "the tall-skinny kernel takes N <= %d and K <= %d, K a multiple of 8 (got N=%d K=%d)", kSkinnyMaxN, kSkinnyMaxK, N, K);
    if (N <= 32) return launch_skinny_cfg<SK_n32, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
    return launch_skinny_cfg<SK_n64, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
}
// X in registers (gemm_skinny.cuh, second kernel): K = 64 or 32 exactly
using SKR_k64 = SkinnyRegCfg<16, 8, 3, 2>;   // column quarters (16 columns per warp), private boxes, three per warp
using SKR_k32 = SkinnyRegCfg<8, 12, 2, 2>;
static bool skinny_xreg_shape_ok(int64_t M, int64_t N, int64_t K) { return M >= 1 && N >= 1 && N <= kSkinnyMaxN && (K == 64 || K == 32); }
template <typename Cfg, bool ACC>
static int launch_skinny_xreg_cfg(double* D, const double* A, const double* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, cudaStream_t s,
                                  const double* Cin, int64_t ldc)
{
    CUtensorMap mapA;
    const cuuint64_t gdim[4] = {(cuuint64_t)M, 2, 4, (cuuint64_t)(K / 8)};
    const cuuint64_t gstr[3] = {(cuuint64_t)(4 * lda * 8), (cuuint64_t)(lda * 8), (cuuint64_t)(8 * lda * 8)};
    const cuuint32_t box[4] = {16, 2, 4, (cuuint32_t)(K / 8)};
    if (int rc = make_tmap_nd(&mapA, A, 4, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int groups = (N + 8 * Cfg::NG - 1) / (8 * Cfg::NG);  // column groups in use; every warp owns one of them
    const int items = ((M + 15) / 16) * groups;
    int grid = (items + Cfg::WARPS - 1) / Cfg::WARPS;
    if (grid > g_ctx.num_sms) grid = g_ctx.num_sms;
    CUDA_TRY(launch_pdl(gemm_skinny_xreg_f64_kernel<Cfg, ACC>, grid, Cfg::THREADS, Cfg::SMEM, s, mapA, X, ldx, D, M, N, ldd, Cin, ldc));
    return 0;
}
template <bool ACC>
static int launch_skinny_xreg(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, int, int, int, cudaStream_t s,
                              const void* Cin, int64_t ldc)
{
    if (!skinny_xreg_shape_ok(M, N, K))
        return fail(JBLAS_B200_EUNSUPPORTED, "the X-in-registers tall-skinny kernel takes N <= %d and K = 32 or 64 (got N=%d K=%d)", kSkinnyMaxN, N, K);
    if (K == 32) return launch_skinny_xreg_cfg<SKR_k32, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
    return launch_skinny_xreg_cfg<SKR_k64, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
}
static cudaError_t attr_skinny_xreg()
{
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_xreg_f64_kernel<SKR_k64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKR_k64::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_xreg_f64_kernel<SKR_k64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKR_k64::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_xreg_f64_kernel<SKR_k32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKR_k32::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_xreg_f64_kernel<SKR_k32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKR_k32::SMEM);
    return e;
}
// X in registers, boxes shared by a team of four quarter-column warps (gemm_skinny.cuh, third kernel): K = 64 or 32 exactly.
// Launched with programmatic stream serialisation: a following launch of the same kind is scheduled while this one runs and
// waits (griddepcontrol.wait) before its first global access -- 1.7 us less per call in back-to-back streams of products.
//                                    KSTEPS WARPS NBUF NG  BN      (team = BN / (8 NG) warps share one box)
template <int KS> using SKT_n64 = SkinnyTeamCfg<KS, 16, 3, 2, 64>;  // 32 < N <= 64: four quarter-column warps
template <int KS> using SKT_n16 = SkinnyTeamCfg<KS, 16, 3, 1, 16>;  //      N <= 16: two warps of one column tile (262144 x 16 x 32: 18.0 vs 24.4 us,
                                                                     //                5.6 TB/s of algorithmic traffic)
template <typename Cfg, bool ACC>
static int launch_skinny_team_cfg(double* D, const double* A, const double* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, cudaStream_t s,
                                  const double* Cin, int64_t ldc)
{
    CUtensorMap mapA;
    const cuuint64_t gdim[4] = {(cuuint64_t)M, 2, 4, (cuuint64_t)(K / 8)};
    const cuuint64_t gstr[3] = {(cuuint64_t)(4 * lda * 8), (cuuint64_t)(lda * 8), (cuuint64_t)(8 * lda * 8)};
    const cuuint32_t box[4] = {16, 2, 4, (cuuint32_t)(K / 8)};
    if (int rc = make_tmap_nd(&mapA, A, 4, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    const int nblocks = (M + 15) / 16;
    int grid = (nblocks + Cfg::TEAMS - 1) / Cfg::TEAMS;
    if (grid > g_ctx.num_sms) grid = g_ctx.num_sms;
    CUDA_TRY(launch_pdl(gemm_skinny_team_f64_kernel<Cfg, ACC>, grid, Cfg::THREADS, Cfg::SMEM, s, mapA, X, ldx, D, M, N, ldd, Cin, ldc));
    return 0;
}
template <int KS, bool ACC>
static int launch_skinny_team_k(double* D, const double* A, const double* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, cudaStream_t s,
                                const double* Cin, int64_t ldc)
{
    if (N > 16) return launch_skinny_team_cfg<SKT_n64<KS>, ACC>(D, A, X, M, N, K, ldd, lda, ldx, s, Cin, ldc);  // (16 < N <= 32 when forced: two members idle)
    return launch_skinny_team_cfg<SKT_n16<KS>, ACC>(D, A, X, M, N, K, ldd, lda, ldx, s, Cin, ldc);
}
template <bool ACC>
static int launch_skinny_team(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx, int, int, int, cudaStream_t s,
                              const void* Cin, int64_t ldc)
{
    if (!skinny_xreg_shape_ok(M, N, K))
        return fail(JBLAS_B200_EUNSUPPORTED, "the team tall-skinny kernel takes N <= %d and K = 32 or 64 (got N=%d K=%d)", kSkinnyMaxN, N, K);
    if (K == 32) return launch_skinny_team_k<8, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
    return launch_skinny_team_k<16, ACC>((double*)D, (const double*)A, (const double*)X, M, N, K, ldd, lda, ldx, s, (const double*)Cin, ldc);
}
template <typename Cfg>
static cudaError_t attr_skinny_team_cfg()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_skinny_team_f64_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_team_f64_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    return e;
}
static cudaError_t attr_skinny_team()
{
    cudaError_t e = attr_skinny_team_cfg<SKT_n64<16>>();
    if (e == cudaSuccess) e = attr_skinny_team_cfg<SKT_n64<8>>();
    if (e == cudaSuccess) e = attr_skinny_team_cfg<SKT_n16<16>>();
    if (e == cudaSuccess) e = attr_skinny_team_cfg<SKT_n16<8>>();
    return e;
}
static cudaError_t attr_skinny()
{
    cudaError_t e = cudaSuccess;
    const int most = (int)SK_n64::smem(kSkinnyMaxK);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_f64_kernel<SK_n64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_f64_kernel<SK_n64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_f64_kernel<SK_n32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_skinny_f64_kernel<SK_n32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
    return e;
}
static int launch_needs_alignment(void*, const void*, const void*, int, int, int, int64_t, int64_t, int64_t, int, int, int,
                                  cudaStream_t, const void*, int64_t)
{
    return fail(JBLAS_B200_EUNSUPPORTED, "this kernel needs 16-byte aligned A/X bases and even leading dimensions (TMA)");
}
// 3xTF32: split A and X into (hi, lo) TF32 parts in stream-ordered scratch, then the tcgen05/TMEM kernel (PAIR: the
// cta_group::2 kernel, two CTAs per 256 x 256 tile).
template <typename Cfg, bool ACC, bool PAIR>
static int launch_tf32x3_any(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                             int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    const int64_t ldA2 = (K + 3) / 4 * 4, ldX2 = (K + 3) / 4 * 4;  // A is stored transposed: K contiguous, M columns
    float* parts = nullptr;  // [A^T_hi | A^T_lo | X_hi | X_lo]
    const size_t nA = (size_t)ldA2 * M, nX = (size_t)ldX2 * N;
    cudaError_t e = cudaMallocAsync((void**)&parts, (2 * nA + 2 * nX) * sizeof(float), s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(JBLAS_B200_ENOMEM, "3xTF32 split scratch of %zu bytes: %s", (2 * nA + 2 * nX) * sizeof(float), cudaGetErrorString(e));
    }
    float *Ahi = parts, *Alo = parts + nA, *Xhi = parts + 2 * nA, *Xlo = parts + 2 * nA + nX;
    {
        unsigned gy = (unsigned)((K + 31) / 32);
        split_tf32_transpose_kernel<<<dim3((unsigned)((M + 31) / 32), gy < 65535u ? gy : 65535u), dim3(32, 8), 0, s>>>((const float*)A, lda, M, K, Ahi, Alo, ldA2);
    }
    split_tf32_kernel<<<dim3((unsigned)((K + 255) / 256), (unsigned)(N < 65535 ? N : 65535)), 256, 0, s>>>((const float*)X, ldx, K, N, Xhi, Xlo, ldX2);
    g_launches += 2;
    constexpr uint32_t box_m = PAIR ? 128 : Cfg::BM, box_n = PAIR ? 128 : Cfg::BN;  // a pair's CTA stages half of the tile's rows and columns
    CUtensorMap mAh, mAl, mXh, mXl;
    int rc = make_tmap_2d(&mAh, Ahi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)K, (uint64_t)M, (uint64_t)ldA2, 32, box_m);
    if (!rc) rc = make_tmap_2d(&mAl, Alo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)K, (uint64_t)M, (uint64_t)ldA2, 32, box_m);
    if (!rc) rc = make_tmap_2d(&mXh, Xhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)K, (uint64_t)N, (uint64_t)ldX2, 32, box_n);
    if (!rc) rc = make_tmap_2d(&mXl, Xlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)K, (uint64_t)N, (uint64_t)ldX2, 32, box_n);
    if (!rc) {
        if constexpr (PAIR) {
            int grid = 2 * tiles_m * tiles_n;  // one cluster of two CTAs per tile, persistent over the tile list
            const int most = g_ctx.num_sms & ~1;
            if (grid > most) grid = most;
            gemm_tf32x3_pair_kernel<Cfg, ACC><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(mAh, mAl, mXh, mXl, (float*)D, M, N, K, ldd, tiles_m, tiles_n, group_m,
                                                                                     (const float*)Cin, ldc);
        } else {
            int grid = tiles_m * tiles_n;
            if (grid > g_ctx.num_sms) grid = g_ctx.num_sms;
            gemm_tf32x3_kernel<Cfg, ACC><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(mAh, mAl, mXh, mXl, (float*)D, M, N, K, ldd, tiles_m, tiles_n, group_m,
                                                                               (const float*)Cin, ldc);
        }
    }
    cudaFreeAsync(parts, s);
    return rc;
}
template <typename Cfg, bool ACC>
static int launch_tf32x3(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                         int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    return launch_tf32x3_any<Cfg, ACC, false>(D, A, X, M, N, K, ldd, lda, ldx, tiles_m, tiles_n, group_m, s, Cin, ldc);
}
template <typename Cfg, bool ACC>
static int launch_tf32x3_pair(void* D, const void* A, const void* X, int M, int N, int K, int64_t ldd, int64_t lda, int64_t ldx,
                              int tiles_m, int tiles_n, int group_m, cudaStream_t s, const void* Cin, int64_t ldc)
{
    return launch_tf32x3_any<Cfg, ACC, true>(D, A, X, M, N, K, ldd, lda, ldx, tiles_m, tiles_n, group_m, s, Cin, ldc);
}
template <typename Cfg>
static cudaError_t attr_tf32x3_pair()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tf32x3_pair_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
}
template <typename Cfg>
static cudaError_t attr_tf32x3()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tf32x3_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
}
template <typename Cfg>
static cudaError_t attr_dmma_tma()
{
    cudaError_t e = cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
}

#define SIMT_ENTRY(NAME, T, DT, CFG, EFF)                                                                          \
    {                                                                                                              \
        NAME, DT, FAM_SIMT, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF, false, false,    \
            CFG::MIN_BLOCKS,                                                                                       \
            {{launch_simt<T, CFG, false, false>, launch_simt<T, CFG, false, true>},                                \
             {launch_simt<T, CFG, true, false>, launch_simt<T, CFG, true, true>}},                                 \
            attr_simt<T, CFG>,                                                                                     \
            []() -> int { return occupancy_of(gemm_simt_kernel<T, CFG, true, false>, CFG::THREADS, CFG::SMEM); }   \
    }
#define SIMT_F32X2_ENTRY(NAME, CFG, EFF)                                                                           \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F32, FAM_SIMT, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF,   \
            false, false, CFG::MIN_BLOCKS,                                                                         \
            {{launch_simt_f32x2<CFG, false, false>, launch_simt_f32x2<CFG, false, true>},                          \
             {launch_simt_f32x2<CFG, true, false>, launch_simt_f32x2<CFG, true, true>}},                           \
            attr_simt_f32x2<CFG>,                                                                                  \
            []() -> int { return occupancy_of(gemm_simt_f32x2_kernel<CFG, true, false>, CFG::THREADS, CFG::SMEM); } \
    }
#define DMMA_ENTRY(NAME, CFG, EFF)                                                                                 \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F64, FAM_DMMA, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF,   \
            false, false, 1,                                                                                       \
            {{launch_dmma<CFG, false, false>, launch_dmma<CFG, false, true>},                                      \
             {launch_dmma<CFG, true, false>, launch_dmma<CFG, true, true>}},                                       \
            attr_dmma<CFG>,                                                                                        \
            []() -> int { return occupancy_of(gemm_dmma_kernel<CFG, true, false>, CFG::THREADS, CFG::SMEM); }      \
    }
#define TF32X3_ENTRY(NAME, CFG, EFF)                                                                               \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F32, FAM_TF32X3, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF, \
            false, true, 1,                                                                                        \
            {{launch_tf32x3<CFG, false>, launch_tf32x3<CFG, true>}, {launch_tf32x3<CFG, false>, launch_tf32x3<CFG, true>}}, \
            attr_tf32x3<CFG>, nullptr                                                                              \
    }
#define TF32X3_PAIR_ENTRY(NAME, CFG, EFF)                                                                          \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F32, FAM_TF32X3, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF, \
            false, true, 1,                                                                                        \
            {{launch_tf32x3_pair<CFG, false>, launch_tf32x3_pair<CFG, true>}, {launch_tf32x3_pair<CFG, false>, launch_tf32x3_pair<CFG, true>}}, \
            attr_tf32x3_pair<CFG>, nullptr, false, 2                                                               \
    }
#define DMMA_TMA_ENTRY(NAME, CFG, EFF)                                                                             \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F64, FAM_DMMA, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF,   \
            true, true, CFG::MIN_BLOCKS,                                                                           \
            {{launch_needs_alignment, launch_needs_alignment}, {launch_dmma_tma<CFG, false>, launch_dmma_tma<CFG, true>}}, \
            attr_dmma_tma<CFG>, nullptr                                                                            \
    }

// TMA kernels whose warp tile is small enough to also come with the element-wise (RAGGED) producers: unaligned operands take
// launch[0][*] = the ragged variant instead of an error
#define DMMA_TMA_RAGGED_ENTRY(NAME, CFG, EFF)                                                                      \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F64, FAM_DMMA, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF,   \
            true, true, CFG::MIN_BLOCKS,                                                                           \
            {{launch_dmma_tma_ragged<CFG, false>, launch_dmma_tma_ragged<CFG, true>}, {launch_dmma_tma<CFG, false>, launch_dmma_tma<CFG, true>}}, \
            attr_dmma_tma_ragged<CFG>, nullptr, true                                                               \
    }
#define F32_TMA_ENTRY(NAME, CFG, EFF)                                                                              \
    {                                                                                                              \
        NAME, JBLAS_B200_DT_F32, FAM_SIMT, CFG::BM, CFG::BN, CFG::BK, CFG::STAGES, CFG::THREADS, CFG::SMEM, EFF,   \
            true, true, 1,                                                                                         \
            {{launch_needs_alignment, launch_needs_alignment}, {launch_f32_tma<CFG, false>, launch_f32_tma<CFG, true>}}, \
            attr_f32_tma<CFG>, nullptr                                                                             \
    }

//                         T      WM WN BK ST MINB
using S64_128x128 = SimtCfg<double, 2, 4, 16, 4, 1>;
using S64_128x64 = SimtCfg<double, 2, 2, 16, 4, 1>;
using S64_64x64 = SimtCfg<double, 1, 2, 16, 4, 1>;
using S32_128x128 = SimtCfg<float, 2, 4, 16, 4, 2>;
using S32_128x64 = SimtCfg<float, 2, 2, 16, 4, 2>;
using S32_64x64 = SimtCfg<float, 1, 2, 16, 4, 4>;
using S32_128x128_k32 = SimtCfg<float, 2, 4, 32, 3, 2>;
using D64_128x128 = DmmaCfg<2, 4, 16, 4>;
using D64_128x64 = DmmaCfg<2, 2, 16, 4>;
using D64_64x64 = DmmaCfg<1, 2, 16, 4>;
//                             WARPS_M WARPS_N MI NI KSUB STAGES
using T64_k16s6 = DmmaTmaCfg<2, 4, 8, 4, 1, 6>;      // 128x128, 6 stages of 32 KiB
using T64_k32s3 = DmmaTmaCfg<2, 4, 8, 4, 2, 3>;      // 128x128, 3 stages of 64 KiB
using T64_128x64 = DmmaTmaCfg<2, 2, 8, 4, 2, 4>;     // 4 warps of 64x32, 4 stages of 48 KiB
using T64_96x64 = DmmaTmaCfg<2, 2, 6, 4, 2, 4>;      // 4 warps of 48x32: tile count just under #SMs on ragged shapes
using T64_64x64_k64 = DmmaTmaCfg<2, 2, 4, 4, 4, 3>;  // 4 warps of 32x32, BK = 64: tall-skinny (K = 64 is ONE stage)
using T64_96x64_w8 = DmmaTmaCfg<2, 4, 6, 2, 2, 4>;   // 8 warps of 48x16: two warps per sub-partition hide latency
using T64_64x64_x2 = DmmaTmaCfg<2, 2, 4, 4, 2, 3, 2>;  // 2 CTAs/SM (3 x 32 KiB each): epilogue of one overlaps the other
using T64_128x64_w8 = DmmaTmaCfg<2, 4, 8, 2, 2, 4>;  // 8 warps of 64x16
using T64_32x32_x2 = DmmaTmaCfg<2, 2, 2, 2, 4, 3, 2>;  // 4 warps of 16x16, 2 CTAs/SM: 256^3..512^3 fill the 148 SMs
using T64_64x32_x2 = DmmaTmaCfg<2, 2, 4, 2, 2, 4, 2>;  // 4 warps of 32x16, 2 CTAs/SM
//                            WM WN RI NJ BK ST MINB
using F2_64x64_w4 = F32x2Cfg<2, 2, 1, 8, 16, 4, 4>;  // 4 warps of 32x32 (thread 4x8)
using F2_64x32_w4 = F32x2Cfg<2, 2, 1, 4, 16, 4, 4>;  // 4 warps of 32x16 (thread 4x4)
using F2_32x32_w2 = F32x2Cfg<1, 2, 1, 4, 16, 4, 8>;  // 2 warps of 32x16
using F32T_s4 = F32TmaCfg<4>;          // 128 x 256 x 32, 4 stages of 48 KiB
using X3_128x256 = Tf32x3Cfg<256, 2>;  // 2 stages of 96 KiB, two 256-column TMEM accumulators
using X3_128x128 = Tf32x3Cfg<128, 3>;  // 3 stages of 64 KiB
using X3_PAIR = Tf32x3PairCfg<3>;      // cta_group::2: 256 x 256 per CTA pair, 3 stages of 64 KiB per CTA

// NOTE: indices are part of the tuning interface (selector 100+i); append, do not reorder.
static const KernelInfo g_kernels[] = {
    /* 0 */ SIMT_ENTRY("simt_f64_128x128x16", double, JBLAS_B200_DT_F64, S64_128x128, 1.00f),
    /* 1 */ SIMT_ENTRY("simt_f64_128x64x16", double, JBLAS_B200_DT_F64, S64_128x64, 1.05f),
    /* 2 */ SIMT_ENTRY("simt_f64_64x64x16", double, JBLAS_B200_DT_F64, S64_64x64, 0.97f),
    /* 3 */ DMMA_ENTRY("dmma_f64_128x128x16", D64_128x128, 1.00f),
    /* 4 */ DMMA_ENTRY("dmma_f64_128x64x16", D64_128x64, 1.08f),
    /* 5 */ DMMA_ENTRY("dmma_f64_64x64x16", D64_64x64, 0.96f),
    /* 6 */ SIMT_ENTRY("simt_f32_128x128x16", float, JBLAS_B200_DT_F32, S32_128x128, 1.00f),
    /* 7 */ SIMT_ENTRY("simt_f32_128x64x16", float, JBLAS_B200_DT_F32, S32_128x64, 1.08f),
    /* 8 */ SIMT_ENTRY("simt_f32_64x64x16", float, JBLAS_B200_DT_F32, S32_64x64, 1.05f),
    /* 9 */ DMMA_TMA_ENTRY("dmma_tma_f64_128x128x16_s6", T64_k16s6, 1.17f),   // 35.15-35.9 vs 30.2 TFLOP/s (8192^3, profiles/r1_sweep_*.json)
    /* 10 */ DMMA_TMA_ENTRY("dmma_tma_f64_128x128x32_s3", T64_k32s3, 1.20f),  // 36.1-36.3 TFLOP/s: the AUTO choice for big shapes
    /* 11 */ SIMT_F32X2_ENTRY("simt_f32x2_128x128x16", S32_128x128, 1.22f),
    /* 12 */ SIMT_F32X2_ENTRY("simt_f32x2_128x64x16", S32_128x64, 1.27f),
    /* 13 */ SIMT_F32X2_ENTRY("simt_f32x2_64x64x16", S32_64x64, 1.215f),
    /* 14 */ SIMT_F32X2_ENTRY("simt_f32x2_128x128x32", S32_128x128_k32, 1.28f),
    /* 15 */ DMMA_TMA_ENTRY("dmma_tma_f64_128x64x32_s4", T64_128x64, 1.02f),
    /* 16 */ DMMA_TMA_ENTRY("dmma_tma_f64_96x64x32_s4", T64_96x64, 1.01f),
    /* 17 */ DMMA_TMA_ENTRY("dmma_tma_f64_64x64x64_s3", T64_64x64_k64, 1.01f),
    /* 18 */ DMMA_TMA_RAGGED_ENTRY("dmma_tma_f64_96x64x32_s4_w8", T64_96x64_w8, 1.17f),
    /* 19 */ DMMA_TMA_RAGGED_ENTRY("dmma_tma_f64_64x64x32_s3_x2", T64_64x64_x2, 1.14f),
    /* 20 */ DMMA_TMA_RAGGED_ENTRY("dmma_tma_f64_128x64x32_s4_w8", T64_128x64_w8, 1.185f),
    /* 21 */ TF32X3_ENTRY("tf32x3_tcgen05_f32_128x256x32_s2", X3_128x256, 1.00f),
    /* 22 */ TF32X3_ENTRY("tf32x3_tcgen05_f32_128x128x32_s3", X3_128x128, 0.80f),
    /* 23 */ DMMA_TMA_RAGGED_ENTRY("dmma_tma_f64_32x32x64_s3_x2", T64_32x32_x2, 1.06f),
    /* 24 */ DMMA_TMA_RAGGED_ENTRY("dmma_tma_f64_64x32x32_s4_x2", T64_64x32_x2, 1.125f),
    /* 25 */ SIMT_F32X2_ENTRY("simt_f32x2_64x64x16_w4", F2_64x64_w4, 1.125f),
    /* 26 */ SIMT_F32X2_ENTRY("simt_f32x2_64x32x16_w4", F2_64x32_w4, 1.00f),
    /* 27 */ SIMT_F32X2_ENTRY("simt_f32x2_32x32x16_w2", F2_32x32_w2, 1.00f),
    /* 28 */ F32_TMA_ENTRY("simt_f32_tma_ffma2_128x256x32_s4", F32T_s4, 1.40f),
    /* 29 */ TF32X3_PAIR_ENTRY("tf32x3_tcgen05_2cta_f32_256x256x32_s3", X3_PAIR, 1.11f),  // 8192^3: 270.8 vs 243.7 TFLOP/s, 16384^3: 239 vs 214 (same box, incl. the split)
    /* 30 */ {"dmma_skinny_f64_16x64_xres_w12", JBLAS_B200_DT_F64, FAM_DMMA, 16, 64, 32, 2, SK_n64::THREADS, SK_n64::smem(64), 1.0f, true, true, 1,
              {{launch_needs_alignment, launch_needs_alignment}, {launch_skinny<false>, launch_skinny<true>}}, attr_skinny, nullptr},
    /* 31 */ {"dmma_skinny_f64_16x16_xreg_w8", JBLAS_B200_DT_F64, FAM_DMMA, 16, 16, 64, 3, SKR_k64::THREADS, SKR_k64::SMEM, 1.0f, true, true, 1,
              {{launch_needs_alignment, launch_needs_alignment}, {launch_skinny_xreg<false>, launch_skinny_xreg<true>}}, attr_skinny_xreg, nullptr},
    /* 32 */ {"dmma_skinny_f64_16x16_xreg_team_w16", JBLAS_B200_DT_F64, FAM_DMMA, 16, 64, 64, 3, SKT_n64<16>::THREADS, SKT_n64<16>::SMEM, 1.0f, true, true, 1,
              {{launch_needs_alignment, launch_needs_alignment}, {launch_skinny_team<false>, launch_skinny_team<true>}}, attr_skinny_team, nullptr},
};
static constexpr int NUM_KERNELS = (int)(sizeof(g_kernels) / sizeof(g_kernels[0]));
static constexpr int kSkinnyKernel = 30, kSkinnyXregKernel = 31, kSkinnyTeamKernel = 32;
static int g_occ[NUM_KERNELS] = {0};  // measured residency (filled at init); 0 = unknown, the planner uses ctas_per_sm
#define JBLAS_B200_EXPLICIT_BASE 100 /* selector 100+i forces g_kernels[i] (tuning / tests) */

// ---------------------------------------------------------------------------------------------------------
// planner: the B200 counterpart of pick_kernel_size + blocking_structure
// ---------------------------------------------------------------------------------------------------------
struct Plan {
    int kidx;
    int tiles_m, tiles_n, group_m;
    bool aligned;
};

// Fraction of an SM's issue rate that `warps` resident warps of these kernels sustain (measured on lone CTAs of 2, 4 and
// 8 warps, profiles/r1_size_sweep_f32_per_kernel.json; the persistent 4-warp DMMA CTAs show the same 0.8).
static double sm_rate(int warps)
{
    if (warps >= 8) return 1.0;
    if (warps >= 4) return 0.78 + (warps - 4) * 0.055;
    if (warps >= 2) return 0.49 + (warps - 2) * 0.145;
    return 0.25 * (warps > 0 ? warps : 1);
}

static bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int make_plan(int dtype, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldx, const void* A, const void* X,
                     int selector, Plan* out)
{
    const int num_sms = g_ctx.num_sms > 0 ? g_ctx.num_sms : 148;
    int family = -1, explicit_idx = -1;
    if (selector >= JBLAS_B200_EXPLICIT_BASE) {
        explicit_idx = selector - JBLAS_B200_EXPLICIT_BASE;
        if (explicit_idx >= NUM_KERNELS || g_kernels[explicit_idx].dtype != dtype)
            return fail(JBLAS_B200_EINVAL, "explicit kernel selector %d is not a %s kernel", selector,
                        dtype == JBLAS_B200_DT_F64 ? "Float64" : "Float32");
    } else if (dtype == JBLAS_B200_DT_F64) {
        if (selector == JBLAS_B200_F64_AUTO) family = FAM_DMMA;  // see DESIGN.md "kernel choice" (ncu evidence)
        else if (selector == JBLAS_B200_F64_DMMA) family = FAM_DMMA;
        else if (selector == JBLAS_B200_F64_SIMT) family = FAM_SIMT;
        else return fail(JBLAS_B200_EINVAL, "unknown Float64 kernel selector %d", selector);
    } else {
        if (selector == JBLAS_B200_F32_EXACT) family = FAM_SIMT;
        else if (selector == JBLAS_B200_F32_3XTF32) family = FAM_TF32X3;
        else return fail(JBLAS_B200_EINVAL, "unknown Float32 mode selector %d", selector);
    }
    const int vec = dtype == JBLAS_B200_DT_F64 ? 2 : 4;
    out->aligned = is_aligned16(A) && is_aligned16(X) && (lda % vec == 0) && (ldx % vec == 0);
    if (explicit_idx == 31 || explicit_idx == 32) out->aligned = is_aligned16(A) && (lda % vec == 0);  // the X-in-registers kernels read X element-wise: only A goes through TMA
    int best = -1;
    double best_t = 0;
    // Tall-skinny Float64 on the tensor pipe: X resident in shared memory, A streamed once by warp-private TMA pipelines
    // (gemm_skinny.cuh).  Measured against the best tile kernel on cold operands: 23.3 vs 24.9 us at 65536 x 64 x 64.
    // K = 64 or 32: the variant that keeps the X fragments in registers (22.6 us; tensor pipe 90 % in the steady state).
    if (explicit_idx < 0 && dtype == JBLAS_B200_DT_F64 && family == FAM_DMMA && out->aligned && M >= 16384) {
        // measured (profiles/r2_skinny_compare.txt): the team kernel (X fragments in registers, one box per row block shared by the
        // warps that cover its columns, 16 warps per SM) wins for 32 < N <= 64 (four quarter-column warps) and for N <= 16 (two
        // one-tile warps: 262144 x 16 x 32 in 18.0 us against 22.2), and it stays at the algorithmic DRAM traffic for any M (10^6
        // rows: 248 us against 293 for private boxes, 253 for the shared-memory variant); for 16 < N <= 32 the shared-memory variant's
        // 4-tile configuration is as fast or faster at every M (65536 x 32 x 64: 13.0 us against 12.7 / 14.9 for the team widths tried)
        // While A fits in L2 (up to 72 MB: 131072 x 64 x 64 still measures 35.8 against 36.6 us) the private-box variant is ~1 us faster: the four quarter-column warps of a row block
        // re-fetch its box from L2, not from DRAM, and need no hand-shake (65536 x 64 x 64: 19.9 against 20.8 us, 32768 rows: 11.7
        // against 12.8); beyond that the re-fetches go to DRAM (10^6 rows: 512 us) and the team kernel takes over.
        if (skinny_xreg_shape_ok(M, N, K) && N > 32 && K == 64 && (double)M * (double)K * 8.0 <= 72.0e6) best = kSkinnyXregKernel;  // (K = 32: team 23.8 vs 24.9 us)
        else if (skinny_xreg_shape_ok(M, N, K) && (N > 32 || N <= 16)) best = kSkinnyTeamKernel;
        else if (skinny_shape_ok(M, N, K)) best = kSkinnyKernel;
    }
    const bool by_shape_rule = best >= 0;
    for (int i = 0; i < NUM_KERNELS && !by_shape_rule; ++i) {
        const KernelInfo& k = g_kernels[i];
        if (explicit_idx < 0 && (i == kSkinnyKernel || i == kSkinnyXregKernel || i == kSkinnyTeamKernel)) continue;  // only through the shape rule above
        if (explicit_idx >= 0 ? (i != explicit_idx) : (k.dtype != dtype || k.family != family)) continue;
        if (explicit_idx < 0 && k.needs_aligned && !out->aligned && !k.has_ragged) continue;
        int64_t tiles = ((M + k.bm - 1) / k.bm) * ((N + k.bn - 1) / k.bn);
        double per_tile = (double)k.bm * k.bn / k.cluster;  // a pair works through its tile in half the time
        // rounds of resident CTAs; CTAs that share an SM (ctas_per_sm > 1) also share its pipes, so a round of c
        // co-resident CTAs costs c tile-times once more than one of them actually lands on an SM
        const int64_t units = num_sms / k.cluster;  // SMs, or SM pairs
        const int64_t resident = units * k.ctas_per_sm;
        const int64_t rounds = (tiles + resident - 1) / resident;
        const int64_t last = tiles - (rounds - 1) * resident;                          // CTAs in the last round
        const int64_t last_share = (last + units - 1) / units;                         // how many of them share an SM
        // a CTA built to share its SM but left alone on it runs at ~0.8 of the shared rate (nothing overlaps its
        // pipeline fill and epilogue): measured, profiles/r1_size_sweep_per_kernel.json
        double waves = (double)((rounds - 1) * k.ctas_per_sm) + (last_share < k.ctas_per_sm ? 1.25 * last_share : (double)last_share);
        int warps = k.threads / 32;
        double t;
        if (k.persistent) {
            t = waves * per_tile;
            double single = per_tile * 4.0 / (warps < 4 ? warps : 4);  // a lone CTA with <4 warps cannot fill an SM
            if (single > t) t = single;
        } else {
            // one CTA per tile: the busiest SM runs n tiles, c at a time; a batch of j co-resident CTAs advances at the
            // rate j*warps resident warps sustain (sm_rate: latency-bound below ~8 warps)
            const int c = g_occ[i] > 0 ? g_occ[i] : k.ctas_per_sm;
            const int64_t n = (tiles + num_sms - 1) / num_sms;
            const int64_t full = n / c, rem = n % c;
            t = (double)full * per_tile * c / sm_rate(c * warps);
            if (rem) t += per_tile * (double)rem / sm_rate((int)rem * warps);
        }
        t /= k.eff;
        if (k.needs_aligned && !out->aligned) t *= 1.05;  // element-wise staging: measured 5 % slower than the TMA boxes (1023 x 777 x 4097 vs 1024 x 776 x 4096)
        if (best < 0 || t < best_t) { best = i; best_t = t; }
    }
    if (best < 0) return fail(JBLAS_B200_EUNSUPPORTED, "no kernel for this dtype/selector");
    const KernelInfo& k = g_kernels[best];
    out->kidx = best;
    out->tiles_m = (int)((M + k.bm - 1) / k.bm);
    out->tiles_n = (int)((N + k.bn - 1) / k.bn);
    // raster group: keep ~sqrt(#SMs) tile rows together so a resident wave touches a near-square block
    static int group_override = -1;  // JBLAS_B200_GROUP_M: raster experiments (DRAM traffic vs group height, profiles/r2_group_m_sweep.txt)
    if (group_override < 0) {
        const char* e = getenv("JBLAS_B200_GROUP_M");
        group_override = e ? atoi(e) : 0;
    }
    const int want_group = group_override > 0 ? group_override : 12;
    out->group_m = out->tiles_m < want_group ? out->tiles_m : want_group;
    if (out->group_m < 1) out->group_m = 1;
    return 0;
}

static int validate(const void* D, const void* A, const void* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                    int64_t ldx)
{
    if (M < 0 || K < 0 || N < 0) return fail(JBLAS_B200_EINVAL, "negative dimension (M=%lld K=%lld N=%lld)", (long long)M, (long long)K, (long long)N);
    if (M > 0x7fffffffLL - 256 || N > 0x7fffffffLL - 256 || K > 0x7fffffffLL - 256)
        return fail(JBLAS_B200_EINVAL, "dimension exceeds 2^31-257");
    if (M == 0 || N == 0) return 0;
    if (!D) return fail(JBLAS_B200_EINVAL, "D is NULL");
    if (ldd < M) return fail(JBLAS_B200_EINVAL, "ldd=%lld < M=%lld", (long long)ldd, (long long)M);
    if (K > 0) {
        if (!A || !X) return fail(JBLAS_B200_EINVAL, "A or X is NULL");
        if (lda < M) return fail(JBLAS_B200_EINVAL, "lda=%lld < M=%lld", (long long)lda, (long long)M);
        if (ldx < K) return fail(JBLAS_B200_EINVAL, "ldx=%lld < K=%lld", (long long)ldx, (long long)K);
    }
    return 0;
}

template <typename T>
__global__ void zero_fill_kernel(T* D, int64_t M, int64_t N, int64_t ldd)
{
    int64_t total = M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        D[(i / M) * ldd + (i % M)] = T(0);
}

// 2-D copy into a buffer with an aligned leading dimension (threads along the contiguous rows: coalesced both ways)
template <typename T>
__global__ void realign_kernel(T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds, int rows, int cols)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int c = blockIdx.y; c < cols; c += gridDim.y) dst[(size_t)c * ldd + r] = src[(size_t)c * lds + r];
}
// One launch re-aligns up to two operands.  128 threads per 16-byte-row-group x 8-column patch: a thread reads R = 16/sizeof(T)
// consecutive rows of 8 columns with scalar loads (the source is only element-aligned; 8*R independent loads in flight) and
// writes one 16-byte vector per column (the destination has an aligned base and leading dimension; its padding rows get 0).
struct RealignJob {
    const void* src;
    void* dst;
    int64_t lds, ldd;
    int rows, cols, row_patches, patches;
};
template <typename T>
static RealignJob make_realign_job(const T* src, int64_t lds, T* dst, int64_t ldd, int64_t rows, int64_t cols)
{
    constexpr int R = 16 / (int)sizeof(T);
    RealignJob j;
    j.src = src; j.dst = dst; j.lds = lds; j.ldd = ldd; j.rows = (int)rows; j.cols = (int)cols;
    j.row_patches = (int)((rows + 128 * R - 1) / (128 * R));
    j.patches = j.row_patches * (int)((cols + 7) / 8);
    return j;
}
template <typename T>
__global__ void __launch_bounds__(128) realign2_kernel(const RealignJob j0, const RealignJob j1)
{
    constexpr int R = 16 / (int)sizeof(T);
    int b = blockIdx.x;
    const bool second = b >= j0.patches;
    const RealignJob& j = second ? j1 : j0;
    if (second) b -= j0.patches;
    const int rp = b % j.row_patches, cp = b / j.row_patches;
    const int r = (rp * 128 + threadIdx.x) * R;
    if (r >= j.rows) return;
    const T* __restrict__ src = static_cast<const T*>(j.src);
    T* __restrict__ dst = static_cast<T*>(j.dst);
    struct alignas(16) V { T v[R]; };
    V v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cc = cp * 8 + c;
#pragma unroll
        for (int i = 0; i < R; ++i) v[c].v[i] = (cc < j.cols && r + i < j.rows) ? src[(size_t)cc * j.lds + r + i] : T(0);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int cc = cp * 8 + c;
        if (cc < j.cols) *reinterpret_cast<V*>(dst + (size_t)cc * j.ldd + r) = v[c];
    }
}

// Prologue of D = A*(X + C) (src/memory_management.jl:72-76): dst = a + b, each element rounded once, written with an aligned
// leading dimension.  A separate HBM-speed pass on purpose: it moves 3*K*N elements against 2*M*N*K flops (0.8 % of the
// 8192^3 product), where adding inside the kernels would put a DADD per X fragment on the FP64 pipe of every tile row.
template <typename T>
__global__ void add_realign_kernel(T* __restrict__ dst, int64_t ldd, const T* __restrict__ a, int64_t lda_, const T* __restrict__ b,
                                   int64_t ldb_, int rows, int cols)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int c = blockIdx.y; c < cols; c += gridDim.y) dst[(size_t)c * ldd + r] = a[(size_t)c * lda_ + r] + b[(size_t)c * ldb_ + r];
}

static int set_all_attrs()
{
    if (g_ctx.attrs_set) return 0;
    for (int i = 0; i < NUM_KERNELS; ++i) {
        CUDA_TRY(g_kernels[i].set_attr());
        g_occ[i] = g_kernels[i].query_occ ? g_kernels[i].query_occ() : 0;
    }
    g_ctx.attrs_set = true;
    return 0;
}

// Misaligned operands (base or leading dimension off the 16-byte grid): when is a re-aligned scratch copy worth it?
//   * Float64 on the tensor pipe (AUTO / DMMA): products up to 3e10 flop run the RAGGED variants of the persistent kernels
//     (element-wise staging inside the kernel, no extra HBM pass and no scratch: 1023 x 777 x 4097 243 -> 235 us; the staging costs ~5 %, a copy pass less than that only beyond ~3e10 flop); beyond that the 128 x 128
//     TMA kernel wins by more than the copy costs (< 1 % of such a product);
//   * exact SIMT families and Float32: from 1e9 flop (their element-wise cp.async variants stage from every thread);
//   * an explicitly forced kernel: only if it cannot stage element-wise itself.
static bool wants_realign(int dtype, int selector, double flops)
{
    if (selector >= JBLAS_B200_EXPLICIT_BASE) {
        const int i = selector - JBLAS_B200_EXPLICIT_BASE;
        if (i < NUM_KERNELS && g_kernels[i].has_ragged) return false;
        return flops >= 1.0e9;
    }
    if (dtype == JBLAS_B200_DT_F64 && (selector == JBLAS_B200_F64_AUTO || selector == JBLAS_B200_F64_DMMA)) return flops >= 3.0e10;
    return flops >= 1.0e9;
}

// The general device-side product:  D = A*(X [+ Xadd]) [+ Cin].
//   Cin  == nullptr : overwrite (jmul!/initkernel!);  Cin == D : D += A*X (kernel!, src/kernels.jl:226);  any other Cin: the
//                     planned fused form D = A*X + C -- every element's chain starts from C[i,j] instead of -0.0;
//   Xadd != nullptr : the planned D = A*(X + C) -- X + Xadd is formed (one rounding per element) by add_realign_kernel.
template <typename T>
static int gemm_dev_ex(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                       int64_t ldx, const T* Cin, int64_t ldc, const T* Xadd, int64_t ldxa, int selector, cudaStream_t s)
{
    if (int rc = require_init()) return rc;
    if (int rc = validate(D, A, X, M, K, N, ldd, lda, ldx)) return rc;
    if (M == 0 || N == 0) return 0;
    if (int rc = check_on_device(D, "D")) return rc;
    if (Cin && ldc < M) return fail(JBLAS_B200_EINVAL, "ldc=%lld < M=%lld", (long long)ldc, (long long)M);
    if (Xadd && K > 0 && ldxa < K) return fail(JBLAS_B200_EINVAL, "ld of the matrix added to X = %lld < K=%lld", (long long)ldxa, (long long)K);
    const int accumulate = Cin != nullptr;
    // s == NULL is the CUDA default stream, exactly as for any CUDA API: the caller's stream ordering is kept
    if (K == 0) {  // empty contraction: jmul! would read X[1,j] out of bounds; defined here as D = 0 (or D = C / D unchanged)
        if (!accumulate) {
            zero_fill_kernel<T><<<g_ctx.num_sms * 4, 256, 0, s>>>(D, M, N, ldd);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
        } else if (Cin != D) {
            realign_kernel<T><<<dim3((unsigned)((M + 255) / 256), (unsigned)(N < 65535 ? N : 65535)), 256, 0, s>>>(D, ldd, Cin, ldc, (int)M, (int)N);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
        }
        return 0;
    }
    // Ragged leading dimensions (e.g. M = 1023 doubles per column) break the 16-byte alignment TMA and 16-byte cp.async
    // need.  For products big enough to matter, the misaligned operand is first copied into a stream-ordered scratch
    // buffer with an even leading dimension (HBM-speed pass, a few % of the GEMM), then the fast path runs.
    const int vec = 16 / (int)sizeof(T);
    struct Scratch {  // stream-ordered scratch, returned to the pool on EVERY exit path
        cudaStream_t s;
        void* p[2] = {nullptr, nullptr};
        ~Scratch()
        {
            for (void* q : p)
                if (q) cudaFreeAsync(q, s);
        }
    } scratch{s};
    T*& tmpA = reinterpret_cast<T*&>(scratch.p[0]);
    T*& tmpX = reinterpret_cast<T*&>(scratch.p[1]);
    if (Xadd) {
        const int64_t ld2 = (K + vec - 1) / vec * vec;
        CUDA_TRY(cudaMallocAsync((void**)&tmpX, (size_t)ld2 * N * sizeof(T), s));
        add_realign_kernel<T><<<dim3((unsigned)((K + 255) / 256), (unsigned)(N < 65535 ? N : 65535)), 256, 0, s>>>(tmpX, ld2, X, ldx, Xadd, ldxa,
                                                                                                                 (int)K, (int)N);
        g_launches++;
        X = tmpX;
        ldx = ld2;
    }
    // the 3xTF32 path re-splits both operands into its own K-major scratch anyway: re-aligning first would be a wasted HBM pass
    const bool splits_itself = dtype == JBLAS_B200_DT_F32 && (selector == JBLAS_B200_F32_3XTF32 ||
                                                               (selector >= JBLAS_B200_EXPLICIT_BASE && selector - JBLAS_B200_EXPLICIT_BASE < NUM_KERNELS &&
                                                                g_kernels[selector - JBLAS_B200_EXPLICIT_BASE].family == FAM_TF32X3));
    if (!splits_itself && wants_realign(dtype, selector, 2.0 * (double)M * (double)N * (double)K)) {
        RealignJob jobs[2];
        int njobs = 0;
        if (!is_aligned16(A) || lda % vec) {
            const int64_t ld2 = (M + vec - 1) / vec * vec;
            CUDA_TRY(cudaMallocAsync((void**)&tmpA, (size_t)ld2 * K * sizeof(T), s));
            jobs[njobs++] = make_realign_job<T>(A, lda, tmpA, ld2, M, K);
            A = tmpA;
            lda = ld2;
        }
        if (!tmpX && (!is_aligned16(X) || ldx % vec)) {  // (the X + C scratch above is born aligned)
            const int64_t ld2 = (K + vec - 1) / vec * vec;
            CUDA_TRY(cudaMallocAsync((void**)&tmpX, (size_t)ld2 * N * sizeof(T), s));
            jobs[njobs++] = make_realign_job<T>(X, ldx, tmpX, ld2, K, N);
            X = tmpX;
            ldx = ld2;
        }
        if (njobs) {  // both operands in ONE launch
            if (njobs == 1) { jobs[1] = jobs[0]; jobs[1].patches = 0; }
            realign2_kernel<T><<<(unsigned)(jobs[0].patches + jobs[1].patches), 128, 0, s>>>(jobs[0], jobs[1]);
            g_launches++;
        }
    }
    Plan p;
    if (int rc = make_plan(dtype, M, K, N, lda, ldx, A, X, selector, &p)) return rc;
    if (int rc = set_all_attrs()) return rc;
    const KernelInfo& k = g_kernels[p.kidx];
    if (int rc = k.launch[p.aligned ? 1 : 0][accumulate ? 1 : 0](D, A, X, (int)M, (int)N, (int)K, ldd, lda, ldx, p.tiles_m, p.tiles_n,
                                                                  p.group_m, s, Cin, ldc))
        return rc;
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <typename T>
static int gemm_dev(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                    int64_t ldx, int accumulate, int selector, cudaStream_t s)
{
    return gemm_dev_ex<T>(dtype, D, A, X, M, K, N, ldd, lda, ldx, accumulate ? D : nullptr, ldd, nullptr, 0, selector, s);
}

// ---------------------------------------------------------------------------------------------------------
// host-pointer path
// ---------------------------------------------------------------------------------------------------------
static int ensure_ws(int i, size_t bytes)
{
    if (g_ctx.ws_bytes[i] >= bytes) return 0;
    if (g_ctx.ws[i]) { cudaFree(g_ctx.ws[i]); g_ctx.ws[i] = nullptr; g_ctx.ws_bytes[i] = 0; }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&g_ctx.ws[i], want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&g_ctx.ws[i], want);
    }
    if (e != cudaSuccess) return fail(JBLAS_B200_ENOMEM, "device workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
    g_ctx.ws_bytes[i] = want;
    return 0;
}

// Optional timeline of the host-pointer pipeline (JBLAS_B200_TRACE=1): one CUDA event per stage, printed to stderr
// (times are relative to the start of the call on the SAME GPU: events of different devices cannot be subtracted).
struct HostTrace {
    bool on = false;
    struct Mark { int g; cudaEvent_t ev; std::string tag; };
    std::vector<Mark> marks;
    void mark(int g, const char* what, int64_t idx, cudaStream_t st)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.push_back({g, e, std::string(what) + " " + std::to_string(idx)});
    }
    void dump(int G, Context** ctx)
    {
        if (!on) return;
        for (int g = 0; g < G; ++g)
            for (const Mark& m : marks) {
                if (m.g != g) continue;
                float ms = 0;
                cudaSetDevice(ctx[g]->device);
                cudaEventElapsedTime(&ms, ctx[g]->ev0, m.ev);
                fprintf(stderr, "[jblas_b200 trace] gpu %d %8.3f ms  %s\n", ctx[g]->device, ms, m.tag.c_str());
                cudaEventDestroy(m.ev);
            }
        marks.clear();
    }
};

// Rates of THIS box that size the panel ramp below, measured once per context (a slower PCIe slot or a power-capped GPU
// must not get a ramp tuned on another machine): pinned H2D bandwidth (32 MiB copy) and the register-only DMMA / FFMA2
// warp-tile probes.  JBLAS_B200_H2D_GBS / JBLAS_B200_F64_TFLOPS / JBLAS_B200_F32_TFLOPS override the measurement.
static int calibrate()
{
    Context& c = cur();
    if (c.h2d_bytes_per_s > 0) return 0;
    auto env = [](const char* name) -> double { const char* e = getenv(name); return e ? atof(e) : 0.0; };
    double h2d = env("JBLAS_B200_H2D_GBS") * 1e9, f64 = env("JBLAS_B200_F64_TFLOPS") * 1e12, f32 = env("JBLAS_B200_F32_TFLOPS") * 1e12;
    cudaStream_t st = c.stream;
    float ms = 0.f;
    if (h2d <= 0) {
        const size_t bytes = (size_t)32 << 20;
        void *h = nullptr, *d = nullptr;
        CUDA_TRY(cudaMallocHost(&h, bytes));
        if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaFreeHost(h); cudaGetLastError(); return fail(JBLAS_B200_ENOMEM, "calibration buffer"); }
        memset(h, 0, bytes);
        for (int rep = 0; rep < 3; ++rep) {  // the last repetition counts
            cudaEventRecord(c.ev0, st);
            cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
            cudaEventRecord(c.ev1, st);
            cudaStreamSynchronize(st);
            cudaEventElapsedTime(&ms, c.ev0, c.ev1);
        }
        cudaFree(d);
        cudaFreeHost(h);
        CUDA_TRY(cudaGetLastError());
        h2d = ms > 0 ? (double)bytes / (ms * 1e-3) : 50e9;
    }
    void* out = nullptr;
    if (f64 <= 0 || f32 <= 0) CUDA_TRY(cudaMalloc(&out, 256));
    if (f64 <= 0) {
        const int iters = 1500;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(c.ev0, st);
            probe_dmma_tile_kernel<<<c.num_sms, 256, 0, st>>>((double*)out, iters, 1.0000001, 1e-9);
            cudaEventRecord(c.ev1, st);
            cudaStreamSynchronize(st);
            cudaEventElapsedTime(&ms, c.ev0, c.ev1);
        }
        g_launches += 4;
        f64 = ms > 0 ? 2.0 * 256 * 32 * iters * (double)c.num_sms * 8 / (ms * 1e-3) : 36e12;
    }
    if (f32 <= 0) {
        const int iters = 4000;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(c.ev0, st);
            probe_ffma2_tile_kernel<<<c.num_sms * 2, 256, 0, st>>>((float*)out, iters, 1.0000001f, 1e-9f);
            cudaEventRecord(c.ev1, st);
            cudaStreamSynchronize(st);
            cudaEventElapsedTime(&ms, c.ev0, c.ev1);
        }
        g_launches += 4;
        f32 = ms > 0 ? 2.0 * 64 * iters * (double)c.num_sms * 2 * 256 / (ms * 1e-3) : 60e12;
    }
    if (out) cudaFree(out);
    CUDA_TRY(cudaGetLastError());
    c.h2d_bytes_per_s = h2d;
    c.dmma_flops_per_s = f64;
    c.ffma2_flops_per_s = f32;
    if (getenv("JBLAS_B200_TRACE"))
        fprintf(stderr, "[jblas_b200 trace] calibration on device %d: H2D %.1f GB/s, DMMA tile %.2f TFLOP/s, FFMA2 tile %.2f TFLOP/s\n", c.device,
                h2d / 1e9, f64 / 1e12, f32 / 1e12);
    return 0;
}

// Event pool of the host pipeline (timing disabled; reused across calls)
struct EventPool {
    std::vector<cudaEvent_t> ev;
    size_t next = 0;
    int get(cudaEvent_t* out)
    {
        if (next == ev.size()) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev.push_back(e);
        }
        *out = ev[next++];
        return 0;
    }
};
static EventPool g_pool[kMaxDevices];

// [c0, c1) of the N columns owned by shard g of G; remainder columns go to the LAST shards (multigpu.column_shard).
static void column_shard(int64_t N, int G, int g, int64_t* c0, int64_t* c1)
{
    const int64_t base = N / G, rem = N % G;
    auto count = [&](int r) { return base + (r >= G - rem ? 1 : 0); };
    int64_t at = 0;
    for (int r = 0; r < g; ++r) at += count(r);
    *c0 = at;
    *c1 = at + count(g);
}

// Synchronous host-pointer GEMM: the literal drop-in for jmul!(D, A, X), on `G` GPUs of this process (G = 1: the bound one).
//
// Partition: GPU g owns a column block of X and D -- the outer `cc` loop of jmul! (src/gemm.jl:313) -- and needs
// all of A.  PCIe (one link per GPU, full duplex), NVLink and the tensor pipes are overlapped with a two-phase schedule over
// the (K panel, column block) grid of every GPU:
//   phase 1 -- K-PANEL major over the FIRST half of the GPU's columns.  Panel p of A is cut into G column slices; GPU u
//              uploads slice u over ITS OWN PCIe link (so A crosses the host links once, in parallel), every other GPU pulls
//              that slice from u's memory with its copy engines over NVLink (no kernel, no SM); with the panel complete the
//              compute stream does  D[:, :N1] (+)= A[:, kp] * X[kp, :N1]  (accumulate = p > 0: ascending k per element,
//              i.e. the chain -- and for every kernel here every bit -- of a single launch; kernel! semantics,
//              src/kernels.jl:226).  The matching rows of X travel with the panel.
//   phase 2 -- A is now resident: the remaining columns go COLUMN-BLOCK major, one full-K launch per block, while the
//              finished blocks (first the whole phase-1 half) travel back on a separate D2H stream.
// Only the last block's D2H is exposed.  Small problems degenerate to one panel / one block.  One host thread issues
// everything (all calls are asynchronous); an event is always recorded before any other stream is told to wait on it.
template <typename T>
static int gemm_host_issue(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                           int64_t ldx, int accumulate, int selector, int G, Context** ctx, const int64_t* shard_c0, const int64_t* shard_ns,
                           HostTrace& trace)
{
    const size_t es = sizeof(T);
    const int vec = 16 / (int)es;
    // device copies are dense with even leading dimensions so the 16-byte / TMA staging paths apply
    const int64_t dM = (M + vec - 1) / vec * vec, dK = (K + vec - 1) / vec * vec;
    struct Dev {
        int64_t c0 = 0, ns = 0, N1 = 0;
        T *dD = nullptr, *dA = nullptr, *dX = nullptr;
    } dv[kMaxDevices];
    auto on = [&](int g) -> int {  // make GPU g the one this thread issues to
        t_cur = ctx[g];
        CUDA_TRY(cudaSetDevice(ctx[g]->device));
        return 0;
    };
    int64_t ns_max = 0;
    (void)N;
    for (int g = 0; g < G; ++g) {  // every shard here is non-empty (the caller dropped GPUs without columns)
        dv[g].c0 = shard_c0[g];
        dv[g].ns = shard_ns[g];
        if (dv[g].ns > ns_max) ns_max = dv[g].ns;
        if (int rc = on(g)) return rc;
        g_pool[ctx[g]->device].next = 0;
        if (int rc = ensure_ws(0, (size_t)dM * dv[g].ns * es)) return rc;
        if (K > 0) {
            if (int rc = ensure_ws(1, (size_t)dM * K * es)) return rc;
            if (int rc = ensure_ws(2, (size_t)dK * dv[g].ns * es)) return rc;
        }
        dv[g].dD = (T*)ctx[g]->ws[0];
        dv[g].dA = (T*)ctx[g]->ws[1];
        dv[g].dX = (T*)ctx[g]->ws[2];
        CUDA_TRY(cudaEventRecord(ctx[g]->ev0, ctx[g]->copy_stream));
        if (accumulate)
            CUDA_TRY(cudaMemcpy2DAsync(dv[g].dD, dM * es, D + dv[g].c0 * ldd, ldd * es, M * es, dv[g].ns, cudaMemcpyHostToDevice, ctx[g]->copy_stream));
    }
    auto d2h = [&](int g, int64_t n0, int64_t nc) -> int {  // compute stream -> D2H stream hand-over for local columns [n0, n0+nc)
        Context& c = *ctx[g];
        CUDA_TRY(cudaEventRecord(c.ev_copy, c.stream));
        CUDA_TRY(cudaStreamWaitEvent(c.d2h_stream, c.ev_copy, 0));
        CUDA_TRY(cudaMemcpy2DAsync(D + (dv[g].c0 + n0) * ldd, ldd * es, dv[g].dD + n0 * dM, dM * es, M * es, nc, cudaMemcpyDeviceToHost, c.d2h_stream));
        trace.mark(g, "d2h done, n0 =", n0, c.d2h_stream);
        return 0;
    };
    if (K == 0) {
        for (int g = 0; g < G; ++g) {
            if (int rc = on(g)) return rc;
            Context& c = *ctx[g];
            CUDA_TRY(cudaEventRecord(c.ev_copy, c.copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_copy, 0));
            if (int rc = gemm_dev<T>(dtype, dv[g].dD, dv[g].dA, dv[g].dX, M, 0, dv[g].ns, dM, dM, 1, accumulate, selector, c.stream)) return rc;
            if (int rc = d2h(g, 0, dv[g].ns)) return rc;
        }
        return 0;
    }
    // ---- schedule (shared by all GPUs: the K panels are a collective object) ----
    const size_t panel_bytes = (size_t)64 << 20;
    const bool big = (size_t)(M + ns_max) * K * es > 2 * panel_bytes;
    // K panels of ~64 MiB of A (at least 512 columns: a shorter accumulate pass re-reads its D tiles too often); the first
    // one is a quarter panel so the multiply starts early
    int64_t kp = K;
    if (big) {
        kp = (int64_t)(panel_bytes / ((size_t)M * es));
        kp = kp / 64 * 64;
        if (kp < 512) kp = 512;
        if (kp > K) kp = K;
    }
    const int64_t kfirst = (kp < K && kp >= 1024) ? kp / 4 : (kp < K && kp >= 512 ? 256 : kp);
    // phase-1 columns: everything for small problems, otherwise the first half (multiple of 128) on one GPU.  Phase 1 exists
    // to keep the tensor pipe busy while A is on its way, and with G GPUs every link carries only 1/G of A: the phase-1 share
    // shrinks with G (half / G, at least 256 columns), so that A is complete early, the column-block phase with its
    // overlapped D2H starts early, and D -- the transfer that saturates the host side first when several GPUs write to one
    // host (measured: 8 GPUs, profiles/r2_pcie_probe_8gpu.txt) -- starts to flow back as soon as possible.
    for (int g = 0; g < G; ++g) {
        dv[g].N1 = dv[g].ns;
        if (big && (size_t)M * dv[g].ns * es > panel_bytes && dv[g].ns >= 512) {
            dv[g].N1 = ((dv[g].ns / 2 / G) + 127) / 128 * 128;
            if (dv[g].N1 < 256) dv[g].N1 = 256;
        }
    }
    // phase-2 column blocks: ~64 MiB of D each
    int64_t nb = (int64_t)(panel_bytes / ((size_t)M * es));
    nb = nb / 128 * 128;
    if (nb < 128) nb = 128;
    // Panel ramp.  Panel i+1 is uploaded while panel i is multiplied, so without a bubble it can exceed panel i only by the
    // ratio of the two rates in contraction columns per second: a GPU's link moves (M/G + N1) elements per column, the
    // multiply spends 2*M*N1 flops on it.  8192^3 f64 on one GPU: 573 vs 515 columns per ms -> x1.11 per step (a fixed
    // 256 -> 1024 jump left the tensor pipe idle for 1.3 ms at the start, JBLAS_B200_TRACE timeline).
    double growth = 0.0;
    if (big && kfirst < kp) {
        if (int rc = on(0)) return rc;
        if (int rc = calibrate()) return rc;
        const Context& c = *ctx[0];
        // sustained rates of the ACCUMULATE passes relative to the register-only probes (a pass re-reads its D tiles):
        // 34.5 of 36.8 TFLOP/s in f64, 52 of 61 exact f32; 3xTF32 runs at ~6.2x the DMMA rate
        const double flops_per_s = dtype == JBLAS_B200_DT_F64 ? 0.9375 * c.dmma_flops_per_s
                                                              : (selector == JBLAS_B200_F32_3XTF32 ? 6.2 * c.dmma_flops_per_s : 0.85 * c.ffma2_flops_per_s);
        const int64_t n1 = dv[0].N1 > 0 ? dv[0].N1 : 1;
        growth = (c.h2d_bytes_per_s / (((double)M / G + (double)n1) * es)) / (flops_per_s / (2.0 * (double)M * (double)n1));
        if (growth > 2.0) growth = 2.0;
        if (growth < 1.02) growth = 0.0;  // copy-bound: nothing to gain, keep the two-size scheme
    }
    // ---- phase 1 ----
    double ramp = (double)kfirst;
    int panel = 0;
    std::vector<cudaEvent_t> a_up(G), x_up(G);
    for (int64_t k0 = 0; k0 < K; ++panel) {
        int64_t kstep = (k0 == 0) ? kfirst : kp;
        if (growth > 0.0 && k0 > 0) {
            ramp *= growth;
            kstep = ((int64_t)(ramp / 64.0 + 0.5)) * 64;
            if (kstep < kfirst) kstep = kfirst;
            if (kstep > kp) kstep = kp;
        }
        const int64_t kc = (K - k0 < kstep) ? (K - k0) : kstep;
        const int acc = (accumulate || k0 > 0) ? 1 : 0;
        // slice u of the panel: columns [k0 + kc*u/G, k0 + kc*(u+1)/G), rounded to 16 columns
        auto slice = [&](int u) -> int64_t { return u >= G ? k0 + kc : k0 + (kc * u / G) / 16 * 16; };
        for (int u = 0; u < G; ++u) {  // uploads, each GPU over its own link
            if (int rc = on(u)) return rc;
            Context& c = *ctx[u];
            const int64_t ka = slice(u), kb = slice(u + 1);
            if (kb > ka)
                CUDA_TRY(cudaMemcpy2DAsync(dv[u].dA + ka * dM, dM * es, A + ka * lda, lda * es, M * es, kb - ka, cudaMemcpyHostToDevice, c.copy_stream));
            if (G > 1) {
                if (int rc = g_pool[c.device].get(&a_up[u])) return rc;
                CUDA_TRY(cudaEventRecord(a_up[u], c.copy_stream));
            }
            CUDA_TRY(cudaMemcpy2DAsync(dv[u].dX + k0, dK * es, X + dv[u].c0 * ldx + k0, ldx * es, kc * es, dv[u].N1, cudaMemcpyHostToDevice, c.copy_stream));
            trace.mark(u, "h2d A slice + X rows (phase-1 columns) done, k0 =", k0, c.copy_stream);
            if (int rc = g_pool[c.device].get(&x_up[u])) return rc;
            CUDA_TRY(cudaEventRecord(x_up[u], c.copy_stream));
        }
        for (int g = 0; g < G; ++g) {  // pulls over NVLink, then the panel product
            if (int rc = on(g)) return rc;
            Context& c = *ctx[g];
            if (G > 1) {
                for (int i = 1; i < G; ++i) {
                    const int u = (g + i) % G;  // every GPU starts with a different peer: no slice owner serves G-1 pulls at once
                    const int64_t ka = slice(u), kb = slice(u + 1);
                    if (kb == ka) continue;
                    CUDA_TRY(cudaStreamWaitEvent(c.peer_stream, a_up[u], 0));
                    CUDA_TRY(cudaMemcpyPeerAsync(dv[g].dA + ka * dM, c.device, dv[u].dA + ka * dM, ctx[u]->device, (size_t)(kb - ka) * dM * es, c.peer_stream));
                }
                cudaEvent_t have;
                if (int rc = g_pool[c.device].get(&have)) return rc;
                CUDA_TRY(cudaEventRecord(have, c.peer_stream));
                trace.mark(g, "A panel complete (peer slices pulled), k0 =", k0, c.peer_stream);
                CUDA_TRY(cudaStreamWaitEvent(c.stream, have, 0));
            }
            CUDA_TRY(cudaStreamWaitEvent(c.stream, x_up[g], 0));
            if (int rc = gemm_dev<T>(dtype, dv[g].dD, dv[g].dA + k0 * dM, dv[g].dX + k0, M, kc, dv[g].N1, dM, dM, dK, acc, selector, c.stream)) return rc;
            trace.mark(g, "gemm phase-1 panel done, k0 =", k0, c.stream);
        }
        k0 += kc;
    }
    // ---- phase 2 ----
    const int64_t tail = 256;  // the last block's D2H is the one transfer nothing hides: split a short tail block off the final one
    for (int g = 0; g < G; ++g) {
        if (int rc = on(g)) return rc;
        Context& c = *ctx[g];
        if (int rc = d2h(g, 0, dv[g].N1)) return rc;
        for (int64_t n0 = dv[g].N1, nc = 0; n0 < dv[g].ns; n0 += nc) {
            const int64_t left = dv[g].ns - n0;
            nc = left < nb ? left : nb;
            if (big && left <= nb && left >= 3 * tail) nc = left - tail;
            CUDA_TRY(cudaMemcpy2DAsync(dv[g].dX + n0 * dK, dK * es, X + (dv[g].c0 + n0) * ldx, ldx * es, K * es, nc, cudaMemcpyHostToDevice, c.copy_stream));
            trace.mark(g, "h2d X column block done, n0 =", n0, c.copy_stream);
            CUDA_TRY(cudaEventRecord(c.ev_copy, c.copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_copy, 0));
            if (int rc = gemm_dev<T>(dtype, dv[g].dD + n0 * dM, dv[g].dA, dv[g].dX + n0 * dK, M, K, nc, dM, dM, dK, accumulate ? 1 : 0, selector, c.stream))
                return rc;
            trace.mark(g, "gemm phase-2 block done, n0 =", n0, c.stream);
            if (int rc = d2h(g, n0, nc)) return rc;
        }
    }
    return 0;
}

template <typename T>
static int gemm_host_multi(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                           int64_t ldx, int accumulate, int selector, int G, Context** ctx)
{
    if (int rc = validate(D, A, X, M, K, N, ldd, lda, ldx)) return rc;
    if (M == 0 || N == 0) return 0;
    Context* const saved = t_cur;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    HostTrace trace;
    trace.on = getenv("JBLAS_B200_TRACE") != nullptr;
    // column shards (multigpu.column_shard: remainder columns go to the last GPUs); GPUs left without a column sit the call out
    Context* act[kMaxDevices];
    int64_t c0s[kMaxDevices], nss[kMaxDevices];
    const int G_all = G;
    G = 0;
    for (int g = 0; g < G_all; ++g) {
        int64_t c0, c1;
        column_shard(N, G_all, g, &c0, &c1);
        if (c1 == c0) continue;
        act[G] = ctx[g];
        c0s[G] = c0;
        nss[G] = c1 - c0;
        ++G;
    }
    ctx = act;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = gemm_host_issue<T>(dtype, D, A, X, M, K, N, ldd, lda, ldx, accumulate, selector, G, ctx, c0s, nss, trace);
    // Success or not, nothing may stay in flight: the copies reference the caller's host buffers and the shared workspaces.
    cudaError_t first = cudaSuccess;
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(ctx[g]->device);
        if (!rc) cudaEventRecord(ctx[g]->ev1, ctx[g]->d2h_stream);
        for (cudaStream_t st : {ctx[g]->d2h_stream, ctx[g]->copy_stream, ctx[g]->peer_stream, ctx[g]->stream}) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess && first == cudaSuccess) first = e;
        }
    }
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!rc && first != cudaSuccess) rc = fail(JBLAS_B200_ECUDA, "host-pointer pipeline: %s", cudaGetErrorString(first));
    float ms = 0.f;
    Context& report = g_ctxs[g_primary >= 0 ? g_primary : 0];
    if (!rc && G == 1 && cudaEventElapsedTime(&ms, ctx[0]->ev0, ctx[0]->ev1) == cudaSuccess) report.last_ms = ms;  // device time of the pipeline
    else report.last_ms = rc ? 0.f : (float)wall_ms;  // several GPUs: events of different devices cannot be subtracted
    cudaGetLastError();
    if (!rc) trace.dump(G, ctx);
    t_cur = saved;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

template <typename T>
static int gemm_host(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda,
                     int64_t ldx, int accumulate, int selector)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = require_init()) return rc;
    Context* one[1] = {&cur()};
    return gemm_host_multi<T>(dtype, D, A, X, M, K, N, ldd, lda, ldx, accumulate, selector, 1, one);
}

// Host-pointer form of the fused products: plain staging (everything up, one product, D down) on the compute stream.
// These are convenience entries for host callers; the pipelined path above is the one the benchmark times.
template <typename T>
static int fused_host(int dtype, T* D, const T* A, const T* X, const T* C, int64_t M, int64_t K, int64_t N, int64_t ldd,
                      int64_t lda, int64_t ldx, int64_t ldc, bool x_plus_c, int selector)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = require_init()) return rc;
    if (int rc = validate(D, A, X, M, K, N, ldd, lda, ldx)) return rc;
    if (M == 0 || N == 0) return 0;
    const int64_t crows = x_plus_c ? K : M;
    if (crows > 0 && (!C || ldc < crows)) return fail(JBLAS_B200_EINVAL, "C is NULL or ldc=%lld < %lld rows", (long long)ldc, (long long)crows);
    const size_t es = sizeof(T);
    const int vec = 16 / (int)es;
    const int64_t dM = (M + vec - 1) / vec * vec, dK = (K + vec - 1) / vec * vec, dC = x_plus_c ? dK : dM;
    cudaStream_t s = g_ctx.stream;
    T *dD = nullptr, *dA = nullptr, *dX = nullptr, *dCm = nullptr;
    auto release = [&]() {
        if (dD) cudaFreeAsync(dD, s);
        if (dA) cudaFreeAsync(dA, s);
        if (dX) cudaFreeAsync(dX, s);
        if (dCm) cudaFreeAsync(dCm, s);
    };
    cudaError_t e = cudaMallocAsync((void**)&dD, (size_t)dM * N * es, s);
    if (e == cudaSuccess && K > 0) e = cudaMallocAsync((void**)&dA, (size_t)dM * K * es, s);
    if (e == cudaSuccess && K > 0) e = cudaMallocAsync((void**)&dX, (size_t)dK * N * es, s);
    if (e == cudaSuccess && crows > 0) e = cudaMallocAsync((void**)&dCm, (size_t)dC * N * es, s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        release();
        return fail(JBLAS_B200_ENOMEM, "device staging for the fused product: %s", cudaGetErrorString(e));
    }
    int rc = 0;
    auto up = [&](T* dst, int64_t dld, const T* src, int64_t sld, int64_t rows, int64_t cols) {
        if (rc || rows == 0 || cols == 0) return;
        cudaError_t ce = cudaMemcpy2DAsync(dst, dld * es, src, sld * es, rows * es, cols, cudaMemcpyHostToDevice, s);
        if (ce != cudaSuccess) rc = fail(JBLAS_B200_ECUDA, "H2D: %s", cudaGetErrorString(ce));
    };
    cudaEventRecord(g_ctx.ev0, s);
    up(dA, dM, A, lda, M, K);
    up(dX, dK, X, ldx, K, N);
    up(dCm, dC, C, ldc, crows, N);
    if (!rc)
        rc = x_plus_c ? gemm_dev_ex<T>(dtype, dD, dA, dX, M, K, N, dM, dM, dK, nullptr, 0, K > 0 ? dCm : nullptr, dC, selector, s)
                      : gemm_dev_ex<T>(dtype, dD, dA, dX, M, K, N, dM, dM, dK, dCm, dC, nullptr, 0, selector, s);
    if (!rc) {
        cudaError_t ce = cudaMemcpy2DAsync(D, ldd * es, dD, dM * es, M * es, N, cudaMemcpyDeviceToHost, s);
        if (ce != cudaSuccess) rc = fail(JBLAS_B200_ECUDA, "D2H: %s", cudaGetErrorString(ce));
    }
    cudaEventRecord(g_ctx.ev1, s);
    release();
    cudaError_t ce = cudaStreamSynchronize(s);
    if (!rc && ce != cudaSuccess) rc = fail(JBLAS_B200_ECUDA, "fused product: %s", cudaGetErrorString(ce));
    if (!rc) cudaEventElapsedTime(&g_ctx.last_ms, g_ctx.ev0, g_ctx.ev1);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// context lifetime
// ---------------------------------------------------------------------------------------------------------
static int create_context(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(JBLAS_B200_ECUDA, "no CUDA device available (%s); jblas_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n || device >= kMaxDevices) return fail(JBLAS_B200_EINVAL, "device %d out of range [0,%d)", device, n < kMaxDevices ? n : kMaxDevices);
    Context& c = g_ctxs[device];
    if (c.device == device) return 0;
    int prev = -1;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(JBLAS_B200_EUNSUPPORTED, "device %d is sm_%d%d; this library contains sm_100a code only", device,
                    prop.major, prop.minor);
    c.num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.d2h_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.peer_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c.ev0));
    CUDA_TRY(cudaEventCreate(&c.ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc((void**)&c.tile_ctr, (size_t)kTileCtrSlots * 2 * sizeof(int)));
    CUDA_TRY(cudaMemset(c.tile_ctr, 0, (size_t)kTileCtrSlots * 2 * sizeof(int)));
    CUDA_TRY(cudaDeviceSynchronize());  // zero before any launch on the (non-blocking) library or caller streams
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;  // re-align scratch (cudaMallocAsync) stays cached across synchronisations
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    c.device = device;
    c.attrs_set = false;
    Context* saved = t_cur;
    t_cur = &c;
    int rc = set_all_attrs();
    t_cur = saved;
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    return rc;
}

static void destroy_context(Context& c)
{
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 3; ++i) {
        if (c.ws[i]) cudaFree(c.ws[i]);
        c.ws[i] = nullptr;
        c.ws_bytes[i] = 0;
    }
    if (c.tile_ctr) cudaFree(c.tile_ctr);
    if (c.ev0) cudaEventDestroy(c.ev0);
    if (c.ev1) cudaEventDestroy(c.ev1);
    if (c.ev_copy) cudaEventDestroy(c.ev_copy);
    for (cudaStream_t st : {c.stream, c.copy_stream, c.d2h_stream, c.peer_stream})
        if (st) cudaStreamDestroy(st);
    c.tile_ctr = nullptr;
    c.ev0 = c.ev1 = c.ev_copy = nullptr;
    c.stream = c.copy_stream = c.d2h_stream = c.peer_stream = nullptr;
    c.slot_of.clear();
    c.capture_seq = 0;
    c.num_sms = 0;
    c.attrs_set = false;
    c.last_ms = 0.f;
    c.h2d_bytes_per_s = c.dmma_flops_per_s = c.ffma2_flops_per_s = 0;
    c.device = -1;
}

// ---------------------------------------------------------------------------------------------------------
// exported C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int jblas_b200_version(void) { return JBLAS_B200_VERSION; }
const char* jblas_b200_last_error(void) { return g_err; }

int jblas_b200_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(JBLAS_B200_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return n;
}

int jblas_b200_init(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_primary == device) return 0;
    if (g_primary >= 0) return fail(JBLAS_B200_EINVAL, "already initialised on device %d (one process per GPU; the multi-GPU entries drive the others)", g_primary);
    if (int rc = create_context(device)) return rc;
    g_primary = device;
    CUDA_TRY(cudaSetDevice(device));
    return 0;
}

int jblas_b200_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d = 0; d < kMaxDevices; ++d) {
        if (g_ctxs[d].device >= 0) cudaSetDevice(d);
        for (cudaEvent_t e : g_pool[d].ev) cudaEventDestroy(e);
        g_pool[d].ev.clear();
        g_pool[d].next = 0;
        destroy_context(g_ctxs[d]);
    }
    if (g_primary >= 0) cudaSetDevice(g_primary);
    g_primary = -1;
    g_mgpu = 0;
    return 0;
}

// ---- fused forms (src/memory_management.jl:72-76) ----
#define FUSED_DEV(SUFFIX, T, DT)                                                                                                   \
    int jblas_b200_gemm_plus_c_##SUFFIX##_dev(T* D, const T* A, const T* X, const T* C, int64_t M, int64_t K, int64_t N,           \
                                              int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel, void* stream)        \
    {                                                                                                                              \
        if (!C && M > 0 && N > 0) return fail(JBLAS_B200_EINVAL, "C is NULL");                                                     \
        return gemm_dev_ex<T>(DT, D, A, X, M, K, N, ldd, lda, ldx, C, ldc, nullptr, 0, kernel, (cudaStream_t)stream);              \
    }                                                                                                                              \
    int jblas_b200_gemm_x_plus_c_##SUFFIX##_dev(T* D, const T* A, const T* X, const T* C, int64_t M, int64_t K, int64_t N,         \
                                                int64_t ldd, int64_t lda, int64_t ldx, int64_t ldc, int kernel, void* stream)      \
    {                                                                                                                              \
        if (!C && K > 0 && N > 0 && M > 0) return fail(JBLAS_B200_EINVAL, "C is NULL");                                            \
        return gemm_dev_ex<T>(DT, D, A, X, M, K, N, ldd, lda, ldx, nullptr, 0, C, ldc, kernel, (cudaStream_t)stream);              \
    }                                                                                                                              \
    int jblas_b200_gemm_plus_c_##SUFFIX(T* D, const T* A, const T* X, const T* C, int64_t M, int64_t K, int64_t N, int64_t ldd,    \
                                        int64_t lda, int64_t ldx, int64_t ldc, int kernel)                                         \
    {                                                                                                                              \
        return fused_host<T>(DT, D, A, X, C, M, K, N, ldd, lda, ldx, ldc, false, kernel);                                          \
    }                                                                                                                              \
    int jblas_b200_gemm_x_plus_c_##SUFFIX(T* D, const T* A, const T* X, const T* C, int64_t M, int64_t K, int64_t N, int64_t ldd,  \
                                          int64_t lda, int64_t ldx, int64_t ldc, int kernel)                                       \
    {                                                                                                                              \
        return fused_host<T>(DT, D, A, X, C, M, K, N, ldd, lda, ldx, ldc, true, kernel);                                           \
    }
FUSED_DEV(f64, double, JBLAS_B200_DT_F64)
FUSED_DEV(f32, float, JBLAS_B200_DT_F32)
#undef FUSED_DEV

int jblas_b200_gemm_f64_dev(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                            int64_t lda, int64_t ldx, int accumulate, int kernel, void* stream)
{
    return gemm_dev<double>(JBLAS_B200_DT_F64, D, A, X, M, K, N, ldd, lda, ldx, accumulate, kernel, (cudaStream_t)stream);
}
int jblas_b200_gemm_f32_dev(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                            int64_t lda, int64_t ldx, int accumulate, int mode, void* stream)
{
    return gemm_dev<float>(JBLAS_B200_DT_F32, D, A, X, M, K, N, ldd, lda, ldx, accumulate, mode, (cudaStream_t)stream);
}
int jblas_b200_gemm_f64(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                        int64_t lda, int64_t ldx, int accumulate, int kernel)
{
    return gemm_host<double>(JBLAS_B200_DT_F64, D, A, X, M, K, N, ldd, lda, ldx, accumulate, kernel);
}
int jblas_b200_gemm_f32(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                        int64_t lda, int64_t ldx, int accumulate, int mode)
{
    return gemm_host<float>(JBLAS_B200_DT_F32, D, A, X, M, K, N, ldd, lda, ldx, accumulate, mode);
}

// ---- single-process multi-GPU mode: the same host-pointer contract on `ngpus` GPUs ----

int jblas_b200_mgpu_init(int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(JBLAS_B200_ECUDA, "no CUDA device available (%s); jblas_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (ngpus <= 0) ngpus = n < kMaxDevices ? n : kMaxDevices;
    if (ngpus > n || ngpus > kMaxDevices) return fail(JBLAS_B200_EINVAL, "%d GPUs requested, %d visible (at most %d supported)", ngpus, n, kMaxDevices);
    if (g_primary >= ngpus) return fail(JBLAS_B200_EINVAL, "the process is bound to device %d, outside the requested set 0..%d", g_primary, ngpus - 1);
    if (ngpus <= g_mgpu) return ngpus;
    int prev = -1;
    cudaGetDevice(&prev);
    for (int d = 0; d < ngpus; ++d)
        if (int rc = create_context(d)) return rc;
    for (int d = 0; d < ngpus; ++d) {
        CUDA_TRY(cudaSetDevice(d));
        for (int q = 0; q < ngpus; ++q) {
            if (q == d) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, d, q);
            if (!can) continue;  // cudaMemcpyPeerAsync still works (staged through the host), only slower
            cudaError_t pe = cudaDeviceEnablePeerAccess(q, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                return fail(JBLAS_B200_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", d, q, cudaGetErrorString(pe));
            cudaGetLastError();
        }
    }
    if (g_primary < 0) g_primary = 0;
    cudaSetDevice(prev >= 0 ? prev : g_primary);
    g_mgpu = ngpus;
    return ngpus;
}

}  // extern "C" (templates cannot have C linkage)
template <typename T>
static int mgpu_gemm(int dtype, T* D, const T* A, const T* X, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda, int64_t ldx,
                     int accumulate, int selector, int ngpus)
{
    if (ngpus <= 0) return fail(JBLAS_B200_EINVAL, "ngpus must be positive");
    if (ngpus > g_mgpu) {
        int rc = jblas_b200_mgpu_init(ngpus);
        if (rc < 0) return rc;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    Context* ctx[kMaxDevices];
    for (int g = 0; g < ngpus; ++g) ctx[g] = &g_ctxs[g];
    return gemm_host_multi<T>(dtype, D, A, X, M, K, N, ldd, lda, ldx, accumulate, selector, ngpus, ctx);
}
extern "C" {
int jblas_b200_mgpu_gemm_f64(double* D, const double* A, const double* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                             int64_t lda, int64_t ldx, int accumulate, int kernel, int ngpus)
{
    return mgpu_gemm<double>(JBLAS_B200_DT_F64, D, A, X, M, K, N, ldd, lda, ldx, accumulate, kernel, ngpus);
}
int jblas_b200_mgpu_gemm_f32(float* D, const float* A, const float* X, int64_t M, int64_t K, int64_t N, int64_t ldd,
                             int64_t lda, int64_t ldx, int accumulate, int mode, int ngpus)
{
    return mgpu_gemm<float>(JBLAS_B200_DT_F32, D, A, X, M, K, N, ldd, lda, ldx, accumulate, mode, ngpus);
}

// jBLAS naming: D is MxP, A is MxN, X is NxP (src/gemm.jl:244); dense MMatrix storage.
int jblas_b200_jmul_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P)
{
    return gemm_host<double>(JBLAS_B200_DT_F64, D, A, X, M, N, P, M > 0 ? M : 1, M > 0 ? M : 1, N > 0 ? N : 1, 0,
                             JBLAS_B200_F64_AUTO);
}
int jblas_b200_jmul_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P)
{
    return gemm_host<float>(JBLAS_B200_DT_F32, D, A, X, M, N, P, M > 0 ? M : 1, M > 0 ? M : 1, N > 0 ? N : 1, 0,
                            JBLAS_B200_F32_EXACT);
}
// fastmul! (src/kernels.jl:202-208): small static matrices, any M (row remainder masked in the reference).
// Small products are issue/latency bound either way; the exact SIMT kernels with predicated edges serve them.
int jblas_b200_fastmul_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P)
{
    return gemm_host<double>(JBLAS_B200_DT_F64, D, A, X, M, N, P, M > 0 ? M : 1, M > 0 ? M : 1, N > 0 ? N : 1, 0,
                             JBLAS_B200_F64_SIMT);
}
int jblas_b200_fastmul_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P)
{
    return gemm_host<float>(JBLAS_B200_DT_F32, D, A, X, M, N, P, M > 0 ? M : 1, M > 0 ? M : 1, N > 0 ? N : 1, 0,
                            JBLAS_B200_F32_EXACT);
}
// kernel!/initkernel! (src/kernels.jl:239,273): stride_AD is shared by A and D (src/kernels.jl:213-215).
// The reference throws unless Mk is a multiple of the vector width (src/kernels.jl:219,249); here any Mk works.
int jblas_b200_kernel_f64(double* pD, const double* pA, const double* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                          int64_t stride_X, int64_t N)
{
    return gemm_host<double>(JBLAS_B200_DT_F64, pD, pA, pX, Mk, N, Pk, stride_AD, stride_AD, stride_X, 1, JBLAS_B200_F64_SIMT);
}
int jblas_b200_initkernel_f64(double* pD, const double* pA, const double* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                              int64_t stride_X, int64_t N)
{
    return gemm_host<double>(JBLAS_B200_DT_F64, pD, pA, pX, Mk, N, Pk, stride_AD, stride_AD, stride_X, 0, JBLAS_B200_F64_SIMT);
}
int jblas_b200_kernel_f32(float* pD, const float* pA, const float* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                          int64_t stride_X, int64_t N)
{
    return gemm_host<float>(JBLAS_B200_DT_F32, pD, pA, pX, Mk, N, Pk, stride_AD, stride_AD, stride_X, 1, JBLAS_B200_F32_EXACT);
}
int jblas_b200_initkernel_f32(float* pD, const float* pA, const float* pX, int64_t Mk, int64_t Pk, int64_t stride_AD,
                              int64_t stride_X, int64_t N)
{
    return gemm_host<float>(JBLAS_B200_DT_F32, pD, pA, pX, Mk, N, Pk, stride_AD, stride_AD, stride_X, 0, JBLAS_B200_F32_EXACT);
}

}  // extern "C" (templates cannot have C linkage)

// Batched products too big for one warp (M or P > 32), Float64, 16-byte aligned: the persistent TMA/DMMA GEMM kernel with 3-D
// tensor maps -- every tile of every product is one entry of the same dynamically scheduled tile list, so the TMA producer
// streams from product to product without draining the pipeline.  BLAS names here: D(MxN) = A(MxK) * X(KxN) per product.
template <typename Cfg>
static int launch_dmma_tma_batched(double* D, const double* A, const double* X, int M, int K, int N, int64_t batch, int64_t strideD,
                                   int64_t strideA, int64_t strideX, cudaStream_t s)
{
    static bool attr_done[kMaxDevices] = {};  // function attributes are per device
    if (!attr_done[g_ctx.device]) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_dmma_tma_kernel<Cfg, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_done[g_ctx.device] = true;
    }
    CUtensorMap mapA, mapX;
    if (int rc = make_tmap_3d(&mapA, A, (uint64_t)M, (uint64_t)K, (uint64_t)batch, (uint64_t)M, (uint64_t)strideA, 16, 16)) return rc;
    if (int rc = make_tmap_3d(&mapX, X, (uint64_t)K, (uint64_t)N, (uint64_t)batch, (uint64_t)K, (uint64_t)strideX, 16, Cfg::BN)) return rc;
    const int tiles_m = (M + Cfg::BM - 1) / Cfg::BM, tiles_n = (N + Cfg::BN - 1) / Cfg::BN;
    const int64_t tiles = (int64_t)tiles_m * tiles_n * batch;
    if (tiles > 0x7fffffffLL) return fail(JBLAS_B200_EINVAL, "batch too large: %lld tiles", (long long)tiles);
    int64_t grid = (int64_t)g_ctx.num_sms * Cfg::MIN_BLOCKS;
    if (grid > tiles) grid = tiles;
    int* ctr = tile_counter_for(s);
    gemm_dmma_tma_kernel<Cfg, false, true><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, s>>>(
        mapA, mapX, D, M, N, K, (int64_t)M, tiles_m, tiles_n, tiles_m, kL2EvictNormal, kL2EvictNormal, ctr, nullptr, 0, (int)batch, strideD);
    return 0;
}

// fastmul!-class batched small products, device pointers, jBLAS dimension names (D MxP, A MxN, X NxP).
template <typename T>
static int fastmul_batched_dev(T* D, const T* A, const T* X, int64_t M, int64_t N, int64_t P, int64_t batch, int64_t strideD,
                               int64_t strideA, int64_t strideX, cudaStream_t s)
{
    if (int rc = require_init()) return rc;
    if (M < 0 || N < 0 || P < 0 || batch < 0) return fail(JBLAS_B200_EINVAL, "negative dimension");
    if (batch == 0 || M == 0 || P == 0) return 0;
    if (N == 0) return fail(JBLAS_B200_EINVAL, "empty contraction (N = 0) is not defined for fastmul_batched");
    if (!D || !A || !X) return fail(JBLAS_B200_EINVAL, "NULL matrix pointer");
    if (strideD < M * P || strideA < M * N || strideX < N * P) return fail(JBLAS_B200_EINVAL, "batch stride smaller than one matrix");
    if (M > 4096 || N > 4096 || P > 4096) return fail(JBLAS_B200_EUNSUPPORTED, "fastmul_batched is for small matrices; use gemm");
    if constexpr (sizeof(T) == 8) {
        // Float64 up to 32 x N x 32: one warp per product on the tensor pipe, fragments loaded straight from HBM
        static int force_simt = -1;
        if (force_simt < 0) {
            const char* e = getenv("JBLAS_B200_BATCHED_SIMT");  // A/B switch for measurements
            force_simt = (e && atoi(e)) ? 1 : 0;
        }
        const bool tma_ok = (M > 32 || P > 32) && M % 2 == 0 && N % 2 == 0 && strideA % 2 == 0 && strideX % 2 == 0 && is_aligned16(A) &&
                            is_aligned16(X) && batch <= 0x7fffffffLL;
        if (!force_simt && tma_ok) {
            int rc;
            if (M <= 64 && P <= 64)
                rc = launch_dmma_tma_batched<T64_64x64_x2>((double*)D, (const double*)A, (const double*)X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, s);
            else if (P <= 64)
                rc = launch_dmma_tma_batched<T64_128x64_w8>((double*)D, (const double*)A, (const double*)X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, s);
            else
                rc = launch_dmma_tma_batched<T64_k32s3>((double*)D, (const double*)A, (const double*)X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, s);
            if (rc) return rc;
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return 0;
        }
        if (!force_simt && M <= 128 && P <= 128) {  // beyond that a product is a GEMM of its own: the shared-memory limit below applies
            const bool one_block = M <= 32 && P <= 32;
            const int mblocks = (int)((M + 31) / 32), pblocks = (int)((P + 31) / 32);
            const int mi = one_block ? (int)((M + 7) / 8) : 4, ni = one_block ? (int)((P + 7) / 8) : 4;
            const bool tiny = mi == 1 && ni == 1 && N <= 16;
            const int64_t warps_needed = tiny ? (batch + 3) / 4 : batch * mblocks * pblocks;
            int64_t grid = (int64_t)g_ctx.num_sms * 16;  // 128-thread CTAs; the hardware keeps as many resident as registers allow
            if (grid * 4 > warps_needed) grid = (warps_needed + 3) / 4;
#define BATCHED_DMMA_LAUNCH(MI_, NI_, KC_, U_, ONE_)                                                                                 \
    fastmul_batched_dmma_kernel<MI_, NI_, KC_, U_, ONE_><<<(unsigned)grid, 128, 0, s>>>(                                               \
        (double*)D, (const double*)A, (const double*)X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, mblocks, pblocks)
#define BATCHED_DMMA(MI_, NI_, KC_, U_) BATCHED_DMMA_LAUNCH(MI_, NI_, KC_, U_, true)
#define BATCHED_DMMA_ROW(MI_)                                                                                                        \
    switch (ni) {                                                                                                                    \
        case 1: BATCHED_DMMA(MI_, 1, ((MI_) + 1 <= 4 ? 8 : 4), 1); break;                                                            \
        case 2: BATCHED_DMMA(MI_, 2, ((MI_) + 2 <= 4 ? 8 : 4), 1); break;                                                            \
        case 3: BATCHED_DMMA(MI_, 3, 4, 1); break;                                                                                   \
        default: BATCHED_DMMA(MI_, 4, 4, 1); break;                                                                                  \
    }
            if (tiny) {
                BATCHED_DMMA(1, 1, 2, 4);
            } else if (!one_block) {
                BATCHED_DMMA_LAUNCH(4, 4, 4, 1, false);
            } else {
                switch (mi) {
                    case 1: BATCHED_DMMA_ROW(1) break;
                    case 2: BATCHED_DMMA_ROW(2) break;
                    case 3: BATCHED_DMMA_ROW(3) break;
                    default: BATCHED_DMMA_ROW(4) break;
                }
            }
#undef BATCHED_DMMA_ROW
#undef BATCHED_DMMA
#undef BATCHED_DMMA_LAUNCH
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            return 0;
        }
    }
    if constexpr (sizeof(T) == 4) {
        // Float32 up to 32 x N x 16 (even M, N % 4 == 0, 16-byte aligned): warp-private staging, FFMA2, no CTA barriers
        static int force_generic = -1;
        if (force_generic < 0) {
            const char* e = getenv("JBLAS_B200_BATCHED_SIMT");
            force_generic = (e && atoi(e)) ? 1 : 0;
        }
        const bool ok = !force_generic && M % 2 == 0 && M <= 32 && P <= 16 && N % 4 == 0 && strideA % 4 == 0 && strideX % 4 == 0 &&
                        strideD % 2 == 0 && is_aligned16(A) && is_aligned16(X) && ((reinterpret_cast<uintptr_t>(D) & 7) == 0);
        if (ok) {
            const int lpp = M <= 16 ? 8 : 16, ppw = 32 / lpp, pc = (int)((P + 1) / 2 * 2);
            const int slot_floats = (int)(((M * N + N * pc) + 31) / 32 * 32 + 16);  // +16: the two products of a half-warp in different banks
            const int warps_per_cta = 2;
            const size_t smem = (size_t)warps_per_cta * 2 * ppw * slot_floats * sizeof(float) + (size_t)warps_per_cta * 2 * sizeof(uint64_t);
            if (smem <= (size_t)200 * 1024) {
                int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
                if (per_sm < 1) per_sm = 1;
                if (per_sm > 8) per_sm = 8;
                const int64_t items = (batch + ppw - 1) / ppw;
                int64_t grid = (int64_t)g_ctx.num_sms * per_sm;
                if (grid * warps_per_cta > items) grid = (items + warps_per_cta - 1) / warps_per_cta;
#define F32_WARP(LPP_, PC_)                                                                                                           \
    {                                                                                                                                 \
        static bool attr_done[kMaxDevices] = {};                                                                                     \
        if (!attr_done[g_ctx.device]) {                                                                                              \
            CUDA_TRY(cudaFuncSetAttribute(fastmul_batched_f32_warp_kernel<LPP_, PC_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          (int)(200 * 1024)));                                                                      \
            attr_done[g_ctx.device] = true;                                                                                          \
        }                                                                                                                            \
        fastmul_batched_f32_warp_kernel<LPP_, PC_><<<(unsigned)grid, 32 * warps_per_cta, smem, s>>>(                                  \
            (float*)D, (const float*)A, (const float*)X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, slot_floats);     \
    }
#define F32_WARP_PC(LPP_)                                                                                                             \
    switch (pc) {                                                                                                                    \
        case 2: F32_WARP(LPP_, 2) break;                                                                                             \
        case 4: F32_WARP(LPP_, 4) break;                                                                                             \
        case 6: F32_WARP(LPP_, 6) break;                                                                                             \
        case 8: F32_WARP(LPP_, 8) break;                                                                                             \
        case 10: F32_WARP(LPP_, 10) break;                                                                                           \
        case 12: F32_WARP(LPP_, 12) break;                                                                                           \
        case 14: F32_WARP(LPP_, 14) break;                                                                                           \
        default: F32_WARP(LPP_, 16) break;                                                                                           \
    }
                if (lpp == 8) { F32_WARP_PC(8) } else { F32_WARP_PC(16) }
#undef F32_WARP_PC
#undef F32_WARP
                g_launches++;
                CUDA_TRY(cudaGetLastError());
                return 0;
            }
        }
    }
    const int xpitch = (int)(N | 1);  // odd column pitch: the column lanes of a warp hit distinct banks
    const size_t slot = ((((size_t)M * N + (size_t)xpitch * P) + 1) & ~(size_t)1) * sizeof(T);  // even element count, as in the kernel
    if (2 * slot > (size_t)200 * 1024)
        return fail(JBLAS_B200_EUNSUPPORTED, "one product needs %zu bytes of shared memory; fastmul_batched is for small matrices, use gemm", slot);
    const int bpp = (int)(((M + 1) / 2) * ((P + 1) / 2));
    int G = bpp >= 256 ? 1 : 256 / bpp;                     // products multiplied side by side by one CTA
    const size_t budget = (size_t)36 * 1024;                // per stage: keeps >= 3 CTAs per SM resident
    while (G > 1 && G * slot > budget) --G;
    if ((int64_t)G > batch) G = (int)batch;
    const size_t smem = 2 * (size_t)G * slot;
    static size_t attr_set[kMaxDevices][2] = {};
    const int ti = sizeof(T) == 8 ? 0 : 1;
    if (smem > 48 * 1024 && smem > attr_set[g_ctx.device][ti]) {
        CUDA_TRY(cudaFuncSetAttribute(fastmul_batched_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
        attr_set[g_ctx.device][ti] = 220 * 1024;
    }
    int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    const int64_t ngroups = (batch + G - 1) / G;
    int64_t grid = (int64_t)g_ctx.num_sms * per_sm;
    if (grid > ngroups) grid = ngroups;
    fastmul_batched_kernel<T><<<(unsigned)grid, 256, smem, s>>>(D, A, X, (int)M, (int)N, (int)P, batch, strideD, strideA, strideX, G, xpitch);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Host-pointer form: the batch is cut into chunks of ~32 MiB that travel H2D, are multiplied and travel back D2H on three
// streams, three device slots deep, so PCIe (full duplex) and the kernel overlap.  PCIe-bound by construction (1.5 flop/B).
template <typename T>
static int fastmul_batched_host(T* D, const T* A, const T* X, int64_t M, int64_t N, int64_t P, int64_t batch, int64_t strideD,
                                int64_t strideA, int64_t strideX)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = require_init()) return rc;
    if (M < 0 || N < 0 || P < 0 || batch < 0) return fail(JBLAS_B200_EINVAL, "negative dimension");
    if (batch == 0 || M == 0 || P == 0) return 0;
    if (N == 0) return fail(JBLAS_B200_EINVAL, "empty contraction (N = 0) is not defined for fastmul_batched");
    if (!D || !A || !X) return fail(JBLAS_B200_EINVAL, "NULL matrix pointer");
    if (strideD < M * P || strideA < M * N || strideX < N * P) return fail(JBLAS_B200_EINVAL, "batch stride smaller than one matrix");
    const size_t es = sizeof(T);
    const int64_t eA = M * N, eX = N * P, eD = M * P;  // dense on the device
    const size_t per_product = (size_t)(eA + eX + eD) * es;
    int64_t nb = (int64_t)(((size_t)32 << 20) / per_product);
    if (nb < 1) nb = 1;
    if (nb > batch) nb = batch;
    constexpr int SLOTS = 3;
    if (int rc = ensure_ws(0, (size_t)SLOTS * nb * eD * es)) return rc;
    if (int rc = ensure_ws(1, (size_t)SLOTS * nb * eA * es)) return rc;
    if (int rc = ensure_ws(2, (size_t)SLOTS * nb * eX * es)) return rc;
    static cudaEvent_t ev_in[SLOTS] = {nullptr}, ev_k[SLOTS] = {nullptr}, ev_out[SLOTS] = {nullptr};
    for (int i = 0; i < SLOTS; ++i) {
        if (!ev_in[i]) {
            CUDA_TRY(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t cs = g_ctx.copy_stream, ks = g_ctx.stream, os = g_ctx.d2h_stream;
    auto copy = [&](void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int64_t rows, cudaMemcpyKind kind,
                    cudaStream_t st) -> cudaError_t {
        if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * (size_t)rows, kind, st);  // dense: one run
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, (size_t)rows, kind, st);  // strided batch: gaps untouched
    };
    // Success or not, nothing may stay in flight when the call returns: the copies reference the caller's host buffers and the
    // shared workspaces (every early return below runs this), and a failed call must not leave a stale time behind.
    struct Drain {
        cudaStream_t a, b, c;
        float* last_ms;
        bool ok = false;
        ~Drain()
        {
            if (ok) return;
            cudaStreamSynchronize(a);
            cudaStreamSynchronize(b);
            cudaStreamSynchronize(c);
            cudaGetLastError();
            *last_ms = 0.f;
        }
    } drain{os, cs, ks, &g_ctx.last_ms};
    CUDA_TRY(cudaEventRecord(g_ctx.ev0, cs));
    int64_t c = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += nb, ++c) {
        const int slot = (int)(c % SLOTS);
        const int64_t n = batch - b0 < nb ? batch - b0 : nb;
        T* dD = (T*)g_ctx.ws[0] + (size_t)slot * nb * eD;
        T* dA = (T*)g_ctx.ws[1] + (size_t)slot * nb * eA;
        T* dX = (T*)g_ctx.ws[2] + (size_t)slot * nb * eX;
        if (c >= SLOTS) CUDA_TRY(cudaStreamWaitEvent(cs, ev_out[slot], 0));  // the slot's previous result has left the device
        CUDA_TRY(copy(dA, eA * es, A + b0 * strideA, strideA * es, eA * es, n, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(copy(dX, eX * es, X + b0 * strideX, strideX * es, eX * es, n, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(cudaEventRecord(ev_in[slot], cs));
        CUDA_TRY(cudaStreamWaitEvent(ks, ev_in[slot], 0));
        if (int rc = fastmul_batched_dev<T>(dD, dA, dX, M, N, P, n, eD, eA, eX, ks)) return rc;
        CUDA_TRY(cudaEventRecord(ev_k[slot], ks));
        CUDA_TRY(cudaStreamWaitEvent(os, ev_k[slot], 0));
        CUDA_TRY(copy(D + b0 * strideD, strideD * es, dD, eD * es, eD * es, n, cudaMemcpyDeviceToHost, os));
        CUDA_TRY(cudaEventRecord(ev_out[slot], os));
    }
    CUDA_TRY(cudaEventRecord(g_ctx.ev1, os));
    CUDA_TRY(cudaStreamSynchronize(os));
    CUDA_TRY(cudaStreamSynchronize(cs));
    CUDA_TRY(cudaStreamSynchronize(ks));
    drain.ok = true;
    cudaEventElapsedTime(&g_ctx.last_ms, g_ctx.ev0, g_ctx.ev1);
    return 0;
}

extern "C" {

int jblas_b200_fastmul_batched_f64(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                   int64_t strideD, int64_t strideA, int64_t strideX)
{
    return fastmul_batched_host<double>(D, A, X, M, N, P, batch, strideD, strideA, strideX);
}
int jblas_b200_fastmul_batched_f32(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                   int64_t strideD, int64_t strideA, int64_t strideX)
{
    return fastmul_batched_host<float>(D, A, X, M, N, P, batch, strideD, strideA, strideX);
}
int jblas_b200_fastmul_batched_f64_dev(double* D, const double* A, const double* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                       int64_t strideD, int64_t strideA, int64_t strideX, void* stream)
{
    return fastmul_batched_dev<double>(D, A, X, M, N, P, batch, strideD, strideA, strideX, (cudaStream_t)stream);
}
int jblas_b200_fastmul_batched_f32_dev(float* D, const float* A, const float* X, int64_t M, int64_t N, int64_t P, int64_t batch,
                                       int64_t strideD, int64_t strideA, int64_t strideX, void* stream)
{
    return fastmul_batched_dev<float>(D, A, X, M, N, P, batch, strideD, strideA, strideX, (cudaStream_t)stream);
}

int jblas_b200_alloc(void** dptr, size_t bytes)
{
    if (int rc = require_init()) return rc;
    if (!dptr) return fail(JBLAS_B200_EINVAL, "dptr is NULL");
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(JBLAS_B200_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}
int jblas_b200_free(void* dptr)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaFree(dptr));
    return 0;
}
int jblas_b200_h2d(void* dst, const void* src, size_t bytes)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int jblas_b200_d2h(void* dst, const void* src, size_t bytes)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int jblas_b200_host_register(void* host, size_t bytes)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaHostRegister(host, bytes, cudaHostRegisterPortable));  // pinned for every GPU of the process (multi-GPU entries)
    return 0;
}
int jblas_b200_host_unregister(void* host)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaHostUnregister(host));
    return 0;
}
int jblas_b200_stream_sync(void* stream)
{
    if (int rc = require_init()) return rc;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

// ---- peer memory (CUDA IPC) ----
static std::map<void*, void*> g_ipc_maps;  // pointer handed to the caller -> base of the mapping

int jblas_b200_ipc_export(const void* dptr, void* handle64, int64_t* offset)
{
    if (int rc = require_init()) return rc;
    if (!dptr || !handle64 || !offset) return fail(JBLAS_B200_EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    // cudaIpcGetMemHandle names the whole ALLOCATION dptr lies in (a caching allocator may have carved dptr out of a larger
    // block), so the offset of dptr inside it travels with the handle
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRangeFn get_range = nullptr;
    if (!get_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(JBLAS_B200_ECUDA, "cuMemGetAddressRange is not available from this driver");
        get_range = (GetRangeFn)fn;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult r = get_range(&base, &size, (CUdeviceptr)dptr);
    if (r != CUDA_SUCCESS) return fail(JBLAS_B200_ECUDA, "cuMemGetAddressRange failed with CUresult %d", (int)r);
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, (void*)base));
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((CUdeviceptr)dptr - base);
    return 0;
}
int jblas_b200_ipc_open(const void* handle64, int64_t offset, void** peer_ptr)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = require_init()) return rc;
    if (!handle64 || !peer_ptr || offset < 0) return fail(JBLAS_B200_EINVAL, "bad ipc_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = (char*)base + offset;
    g_ipc_maps[*peer_ptr] = base;
    return 0;
}
int jblas_b200_ipc_close(void* peer_ptr)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = require_init()) return rc;
    auto it = g_ipc_maps.find(peer_ptr);
    if (it == g_ipc_maps.end()) return fail(JBLAS_B200_EINVAL, "pointer was not returned by jblas_b200_ipc_open");
    void* base = it->second;
    g_ipc_maps.erase(it);
    CUDA_TRY(cudaIpcCloseMemHandle(base));
    return 0;
}
int jblas_b200_copy_async(void* dst, const void* src, size_t bytes, void* stream)
{
    if (int rc = require_init()) return rc;
    if (bytes == 0) return 0;
    if (!dst || !src) return fail(JBLAS_B200_EINVAL, "NULL pointer");
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int jblas_b200_randn_fill(void* dptr, int64_t first, int64_t n, uint64_t seed, int dtype, void* stream)
{
    if (int rc = require_init()) return rc;
    if (n < 0 || first < 0 || (n > 0 && !dptr)) return fail(JBLAS_B200_EINVAL, "bad randn_fill arguments");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;  // NULL = the CUDA default stream
    int blocks = g_ctx.num_sms * 8;
    if (dtype == JBLAS_B200_DT_F64)
        randn_fill_kernel<double><<<blocks, 256, 0, s>>>((double*)dptr, first, n, seed);
    else if (dtype == JBLAS_B200_DT_F32)
        randn_fill_kernel<float><<<blocks, 256, 0, s>>>((float*)dptr, first, n, seed);
    else
        return fail(JBLAS_B200_EINVAL, "unknown dtype %d", dtype);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int jblas_b200_plan(int dtype, int64_t M, int64_t K, int64_t N, int64_t ldd, int64_t lda, int64_t ldx, int selector,
                    int64_t out[10])
{
    if (!out) return fail(JBLAS_B200_EINVAL, "out is NULL");
    if (dtype != JBLAS_B200_DT_F64 && dtype != JBLAS_B200_DT_F32) return fail(JBLAS_B200_EINVAL, "unknown dtype %d", dtype);
    if (M <= 0 || N <= 0 || K <= 0 || ldd < M || lda < M || ldx < K) return fail(JBLAS_B200_EINVAL, "bad dimensions");
    Plan p;
    // alignment of device bases is assumed (cudaMalloc gives 256 B); leading dimensions decide the staging path
    const int vec = dtype == JBLAS_B200_DT_F64 ? 2 : 4;
    bool realigned = false;
    if (wants_realign(dtype, selector, 2.0 * (double)M * (double)N * (double)K) && (lda % vec || ldx % vec)) {  // same rule as gemm_dev
        lda = (lda + vec - 1) / vec * vec;
        ldx = (ldx + vec - 1) / vec * vec;
        realigned = true;
    }
    if (int rc = make_plan(dtype, M, K, N, lda, ldx, nullptr, nullptr, selector, &p)) return rc;
    const KernelInfo& k = g_kernels[p.kidx];
    out[0] = p.kidx; out[1] = k.bm; out[2] = k.bn; out[3] = k.bk; out[4] = k.stages; out[5] = k.threads;
    out[6] = (int64_t)p.tiles_m * p.tiles_n * k.cluster;  // CTAs: a cta_group::2 tile takes a pair
    const int64_t resident = (int64_t)(g_ctx.num_sms > 0 ? g_ctx.num_sms : 148) / k.cluster * k.cluster * k.ctas_per_sm;
    if (k.persistent && out[6] > resident) out[6] = resident;
    out[7] = p.group_m; out[8] = (int64_t)k.smem; out[9] = realigned ? 2 : (p.aligned ? 1 : 0);
    return 0;
}
const char* jblas_b200_kernel_name(int kidx) { return (kidx >= 0 && kidx < NUM_KERNELS) ? g_kernels[kidx].name : ""; }
int jblas_b200_num_kernels(void) { return NUM_KERNELS; }

int64_t jblas_b200_launch_count(void) { return g_launches.load(); }
float jblas_b200_time_last_ms(void) { return g_primary >= 0 ? g_ctxs[g_primary].last_ms : 0.f; }

int jblas_b200_probe_pipe(int kind, int iters, double* tflops, float* ms_out)
{
    if (int rc = require_init()) return rc;
    if (iters <= 0 || !tflops) return fail(JBLAS_B200_EINVAL, "bad probe arguments");
    cudaStream_t s = g_ctx.stream;
    void* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, 256));
    const int blocks = g_ctx.num_sms * 4, threads = 256;
    double flops = 0;
    // 3 warm-up launches (clocks ramp from idle over tens of ms), then the best of 5 timed launches
    float best_ms = 0.f;
    for (int rep = 0; rep < 8; ++rep) {
        CUDA_TRY(cudaEventRecord(g_ctx.ev0, s));
        if (kind == 0) {
            probe_dfma_kernel<16><<<blocks, threads, 0, s>>>((double*)out, iters, 1.0000001, 1e-9);
            flops = 2.0 * 16 * iters * (double)blocks * threads;
        } else if (kind == 1) {
            probe_dmma_kernel<16><<<blocks, threads, 0, s>>>((double*)out, iters, 1.0000001, 1e-9);
            flops = 2.0 * 256 * 16 * iters * (double)blocks * (threads / 32);
        } else if (kind == 2) {
            probe_ffma_kernel<32><<<blocks, threads, 0, s>>>((float*)out, iters, 1.0000001f, 1e-9f);
            flops = 2.0 * 32 * iters * (double)blocks * threads;
        } else if (kind == 3) {  // DMMA, the GEMM warp-tile pattern (8x4 tiles, 2 warps per sub-partition)
            probe_dmma_tile_kernel<<<g_ctx.num_sms, threads, 0, s>>>((double*)out, iters, 1.0000001, 1e-9);
            flops = 2.0 * 256 * 32 * iters * (double)g_ctx.num_sms * (threads / 32);
        } else if (kind == 4) {  // DFMA, 8x8 outer-product pattern
            probe_fma_tile_kernel<double><<<g_ctx.num_sms, threads, 0, s>>>((double*)out, iters, 1.0000001, 1e-9);
            flops = 2.0 * 64 * iters * (double)g_ctx.num_sms * threads;
        } else if (kind == 5) {  // FFMA, 8x8 outer-product pattern, 2 CTAs per SM
            probe_fma_tile_kernel<float><<<g_ctx.num_sms * 2, threads, 0, s>>>((float*)out, iters, 1.0000001f, 1e-9f);
            flops = 2.0 * 64 * iters * (double)g_ctx.num_sms * 2 * threads;
        } else if (kind == 6) {  // FFMA2 (fma.rn.f32x2), 8x8 outer-product pattern, 2 CTAs per SM
            probe_ffma2_tile_kernel<<<g_ctx.num_sms * 2, threads, 0, s>>>((float*)out, iters, 1.0000001f, 1e-9f);
            flops = 2.0 * 64 * iters * (double)g_ctx.num_sms * 2 * threads;
#ifdef JBLAS_B200_TUNING_PROBES
        } else if (kind >= 7 && kind < 7 + 64) {  // FFMA2 inner loop of the exact-FP32 kernel with its shared-memory operand loads (mode = kind - 7)
            const int grid = g_ctx.num_sms * 2;
            switch (kind - 7) {
#define LDS_PROBE(MODE_) case MODE_: probe_ffma2_lds_kernel<MODE_><<<grid, threads, 0, s>>>((float*)out, iters); break;
                LDS_PROBE(0) LDS_PROBE(1) LDS_PROBE(2) LDS_PROBE(3) LDS_PROBE(4) LDS_PROBE(5) LDS_PROBE(6) LDS_PROBE(7)
                LDS_PROBE(8) LDS_PROBE(9) LDS_PROBE(10) LDS_PROBE(11) LDS_PROBE(12) LDS_PROBE(13) LDS_PROBE(14) LDS_PROBE(15)
                LDS_PROBE(16) LDS_PROBE(18) LDS_PROBE(20) LDS_PROBE(24) LDS_PROBE(22)
                LDS_PROBE(32) LDS_PROBE(33) LDS_PROBE(36) LDS_PROBE(40)
                default: cudaFree(out); return fail(JBLAS_B200_EINVAL, "probe mode %d is not built", kind - 7);
#undef LDS_PROBE
            }
            flops = 2.0 * 64 * 32 * iters * (double)grid * threads;  // 32 k per iteration, 64 FMAs per thread per k
        } else if (kind >= 100 && kind < 104) {  // the same loop with an 8 x 16 thread tile (one CTA per SM), 16 k per iteration
            const int grid = g_ctx.num_sms;
            switch (kind - 100) {
                case 0: probe_ffma2_lds_kernel<0, 16, 1><<<grid, threads, 0, s>>>((float*)out, iters); break;
                case 1: probe_ffma2_lds_kernel<1, 16, 1><<<grid, threads, 0, s>>>((float*)out, iters); break;
                case 2: probe_ffma2_lds_kernel<12, 16, 1><<<grid, threads, 0, s>>>((float*)out, iters); break;
                default: probe_ffma2_lds_kernel<13, 16, 1><<<grid, threads, 0, s>>>((float*)out, iters); break;
            }
            flops = 2.0 * 128 * 16 * iters * (double)grid * threads;
#endif
        } else {
            cudaFree(out);
            return fail(JBLAS_B200_EINVAL, "unknown probe kind %d (kinds 7+ exist only in a library built with JBLAS_B200_TUNING_PROBES)", kind);
        }
        g_launches++;
        CUDA_TRY(cudaEventRecord(g_ctx.ev1, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        float rep_ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&rep_ms, g_ctx.ev0, g_ctx.ev1));
        if (rep >= 3 && (best_ms == 0.f || rep_ms < best_ms)) best_ms = rep_ms;
    }
    float ms = best_ms;
    CUDA_TRY(cudaFree(out));
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return 0;
}

}  // extern "C"

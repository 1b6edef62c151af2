// Host-side plumbing shared by the C-ABI entry points: status codes, last-error text,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <utility>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op unless a profiler is attached

namespace dyt {

// Status convention of include/dyt_b200.h: 0 ok, <0 argument error, >0 cudaError_t.
constexpr int DYT_OK = 0;
constexpr int DYT_EINVAL = -1;
constexpr int DYT_EUNSUPPORTED = -2;
constexpr int DYT_EDRIVER = -3;

inline char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DYT_OK;
  snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

#define DYT_CHECK_ARG(cond, ...)                               \
  do {                                                         \
    if (!(cond)) return ::dyt::fail(::dyt::DYT_EINVAL, __VA_ARGS__); \
  } while (0)

#define DYT_CUDA(call)                                      \
  do {                                                      \
    int _s = ::dyt::cuda_status((call), #call);             \
    if (_s != 0) return _s;                                 \
  } while (0)

// NVTX range around a host-side launch sequence (shows up as a named span in nsys / ncu --nvtx;
// costs one pointer check per call when no tool is attached).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  }
  return fn;
}

// 2-D fp16 row-major tensor [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols]
// (64 halves = 128 B = the swizzle span), SWIZZLE_128B, OOB elements read as zero.
inline int make_tmap_f16_sw128(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols,
                               uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return fail(DYT_EDRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(DYT_EINVAL, "TMA operand must be 16-byte aligned with a 16-byte multiple row stride");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DYT_EDRIVER, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return DYT_OK;
}

// Host-side caches are per device and lock-free: one process may drive several GPUs from several
// threads (the C ABI is a plain stream-based API with no such restriction).
constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// Library option (dyt_configure): size every persistent grid for at most this many SMs (0 = all).  Two
// forwards on two streams, each limited to half of the SMs, then really run side by side (a full-size
// persistent kernel leaves no room for another one's CTAs).  For experiments with concurrent streams.
inline std::atomic<int>& sm_limit_option() {
  static std::atomic<int> v{0};
  return v;
}

inline int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  const int lim = sm_limit_option().load(std::memory_order_relaxed);
  if (lim >= 2 && lim < v) v = lim & ~1;
  return v;
}

// Library option (dyt_configure): launch the forward-path kernels with programmatic stream
// serialization so that each kernel's prologue overlaps its predecessor's tail.  On by default
// since the end of round 2: with the shorter kernels of that round the interleaved same-box A/B
// (scripts/ab_step.py, six rounds of 20 replays) gives 9.16-9.22 ms without, 9.09-9.17 ms with
// (-0.08 ms, bit-identical logits); the first measurement (9.43-9.62 vs 9.50-10.01 ms) was noise.
inline std::atomic<int>& pdl_option() {
  static std::atomic<int> v{1};
  return v;
}

// Library option (dyt_configure): cut the tiles of a GEMM's last, partial round into column
// sub-tiles (gemm_tn.cuh GemmItems).  On by default.
inline std::atomic<int>& tail_split_option() {
  static std::atomic<int> v{1};
  return v;
}

// Library option (dyt_configure): the block forward computes the adapter's up projection inside the
// scatter-merge kernel (merge_up.cu).  On by default.
inline std::atomic<int>& fuse_up_option() {
  static std::atomic<int> v{1};
  return v;
}

// Library option (dyt_configure): the block forward also computes the adapter's down projection
// inside the merge kernel (no fp16 copy of x1, no `down` buffer, no side stream).  Off by default:
// 167 MB less HBM traffic per layer, but the pre-pass over the tile's x1 rows costs the merge kernel
// 36 us against 18 us for the separate down GEMM, and the step does not move (9.16 vs 9.31 ms).
inline std::atomic<int>& fuse_down_option() {
  static std::atomic<int> v{0};
  return v;
}

// Library option (dyt_configure), bit mask: GEMMs of dyt_block_fwd that walk their row tiles from the
// last to the first (1 = qkv, 2 = proj, 4 = fc2, 8 = fc1).  Every producer of the block writes its rows
// in ascending order, so the rows it wrote last are the ones still in the 126 MB L2 when the consumer
// starts; a consumer that begins there finds them (and leaves the rows IT writes last -- the low ones --
// for the ascending kernel behind it: attention after qkv, the dispatcher after proj).  Results are
// bit-identical; only the tile order changes.  Default 7 (qkv, proj, fc2 descending; the chain is then
// merge asc -> qkv desc -> attention asc -> proj desc -> dispatcher asc -> fc1 asc -> fc2 desc -> merge
// asc): same-box interleaved A/B 8.86 ms with 0, 8.78 ms with 7 (3 / 5 / 15: 8.79 / 8.79 / 8.80 ms).
inline std::atomic<int>& tile_order_option() {
  static std::atomic<int> v{7};
  return v;
}

// launch flags of the internal gemm_tn() entry (internal.h)
constexpr int GEMM_FLAG_REVERSE = 1;     // row tiles descending (L2-aware order)
constexpr int GEMM_FLAG_HALF_GRID = 4;   // at most half of the clusters (a GEMM that shares the machine
                                         // with a kernel on another stream)

// Library option (dyt_configure), bit mask: how the adapter's down GEMM (side stream) shares the machine
// with the dispatcher (caller's stream); both become runnable when the proj GEMM ends.  A persistent GEMM
// CTA (384 threads x 96 registers, 185 KB of shared memory) leaves no room for a dispatcher CTA (256 x 128
// registers) on its SM, so the two kernels mostly run one after the other whatever the streams say.
// bit 0: the down GEMM takes at most half of the clusters (37 on a B200), leaving the other SMs to the
// dispatcher; bit 1: no side stream (the GEMM stays on the caller's stream, before the dispatcher);
// bit 2: the side-stream branch is launched after the dispatcher instead of before it.
// Measured (same box, twelve interleaved rounds of 30 replays): 9.023 / 9.013 / 9.007 ms for 0 / 1 / 2 --
// no difference, so the default stays 0.  (A first measurement that showed -0.07 ms for bit 0 was taken with
// a flag mix-up that also put the fc2 GEMM on half of the SMs in every arm; a dispatcher held to 112
// registers so that it fits beside a GEMM CTA spills and was slower: removed.)
inline std::atomic<int>& side_plan_option() {
  static std::atomic<int> v{0};
  return v;
}

// Library option (dyt_configure): uniform sequences of 161..256 tokens run the four-stream
// attention kernel (attn_split.cu) instead of the two-stream one (attn_varlen.cu).
inline std::atomic<int>& attn_split_option() {
  // On by default: 96 -> 88 us alone, 112.6 -> 103.8 us per launch inside the step (-0.1 ms).  The key
  // halves of a row share one maximum, so the fp16 rounding points are those of the two-stream kernel.
  static std::atomic<int> v{1};
  return v;
}

// <<<grid, block, smem, stream>>> with the PDL attribute (the kernel must call pdl_wait() before
// its first dependent global access)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_option().load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): set it once per device to
// the kernel's fixed maximum (racing threads write the same value, so the race is benign).
struct SmemAttrCache {
  std::atomic<int> done[kMaxDevices];
};
template <typename Kernel>
inline int ensure_dyn_smem(Kernel kern, int max_bytes, SmemAttrCache& cache) {
  const int dev = current_device();
  if (cache.done[dev].load(std::memory_order_acquire) == max_bytes) return DYT_OK;
  DYT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
  cache.done[dev].store(max_bytes, std::memory_order_release);
  return DYT_OK;
}

}  // namespace dyt

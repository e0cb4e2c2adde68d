// Host-side TMA descriptor (CUtensorMap) construction + cache.
// The driver entry point is fetched through the runtime (cudaGetDriverEntryPoint) so the shared
// library has no link-time dependency on libcuda and loads on a GPU-less build box.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <unordered_map>

namespace i360 {

// error codes returned through the C ABI (negative = failure)
enum : int {
  I360_OK = 0,
  I360_ERR_ARG = -1,       // bad shape / alignment / null pointer
  I360_ERR_CUDA = -2,      // CUDA runtime error (launch, attribute, ...)
  I360_ERR_TMAP = -3,      // cuTensorMapEncodeTiled failed or unavailable
  I360_ERR_UNSUPPORTED = -4
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t dim[4];
  uint64_t stride[3];  // bytes, dims 1..3
  uint32_t box[4];
  uint32_t estride[4];  // traversal ("element") strides: box[i] elements are traversed, every estride[i]-th is loaded
  uint32_t rank;
  uint32_t swizzle;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
  }
};

// bf16 tensor map, rank 2..4. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// swizzle: 0 none, 1 = 32B, 2 = 64B, 3 = 128B (CUtensorMapSwizzle numbering)
inline int get_tmap_bf16(CUtensorMap* out, const void* ptr, uint32_t rank, const uint64_t* dim,
                         const uint64_t* stride_bytes, const uint32_t* box, uint32_t swizzle,
                         const uint32_t* elem_stride = nullptr) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank; key.swizzle = swizzle;
  for (uint32_t i = 0; i < rank; ++i) { key.dim[i] = dim[i]; key.box[i] = box[i]; key.estride[i] = elem_stride ? elem_stride[i] : 1; }
  for (uint32_t i = 0; i + 1 < rank; ++i) key.stride[i] = stride_bytes[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return I360_OK; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return I360_ERR_TMAP;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return I360_ERR_ARG;
  cuuint64_t gdim[4]; cuuint64_t gstr[3]; cuuint32_t bx[4]; cuuint32_t es[4];
  for (uint32_t i = 0; i < rank; ++i) { gdim[i] = dim[i]; bx[i] = box[i]; es[i] = elem_stride ? elem_stride[i] : 1; }
  for (uint32_t i = 0; i + 1 < rank; ++i) {
    if (stride_bytes[i] % 16 != 0) return I360_ERR_ARG;
    gstr[i] = stride_bytes[i];
  }
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr,
                   bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, static_cast<CUtensorMapSwizzle>(swizzle),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return I360_ERR_TMAP;
  {
    std::lock_guard<std::mutex> g(mu);
    // bounded: a denoising run re-uses a few thousand (pointer, shape) pairs; long-lived processes that keep allocating new
    // buffers (eager mode, many resolutions) restart the cache instead of growing without limit (encode costs ~1 us)
    if (cache.size() > 32768) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return I360_OK;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// EXPERIMENT, off by default: I360_PDL=1 launches every kernel with the programmatic-dependent-launch attribute (without it
// the griddepcontrol instructions in the kernels are no-ops).  Measured on one B200, CUDA-graphed 16x512x1024 step, same box,
// two alternating repetitions: 319.3 / 319.3 ms with the early trigger, 313.8 / 313.9 ms with the attribute but no explicit
// trigger (-DI360_PDL_EARLY_TRIGGER=0), 312.5 / 312.6 ms without PDL; the 24x768x1536 step 1380-1389 vs 1360 ms, the
// 16x256x512 single-branch step 23.6-23.8 vs 23.5-23.9 ms.  Inside a graph the kernel-to-kernel gap is already small, and
// the early-launched CTAs (resident next to the primary grid whenever their shared memory fits, e.g. the normalisation
// kernels next to a GEMM) take more from the running grid than their hidden set-up gives back.
inline bool pdl_enabled() {
  static const bool on = getenv("I360_PDL") != nullptr && atoi(getenv("I360_PDL")) != 0;
  return on;
}

// <<<grid, block, smem, stream>>> with the PDL attribute (see common.cuh)
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through cudaGetLastError()
}

// launch errors (and sticky errors of an earlier kernel, e.g. a device assertion) are named on stderr: the reference's
// script swallows exceptions around the pipeline call (inference_dual_p2e.py:596-597)
#define I360_CUDA_CHECK_LAUNCH()                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                                 \
    if (e__ != cudaSuccess) {                                                                             \
      fprintf(stderr, "imagine360_b200: CUDA error '%s' after a launch in %s (a trap = a kernel invariant failed: "    \
              "smem alignment or an mbarrier wait past the deadlock guard)\n", cudaGetErrorString(e__), __func__);       \
      return i360::I360_ERR_CUDA;                                                                         \
    }                                                                                                     \
  } while (0)

}  // namespace i360

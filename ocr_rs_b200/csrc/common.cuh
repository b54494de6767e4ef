// Shared host/device plumbing for libocrb (context, error reporting, workspace buffers).
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ocrb.h"

namespace ocrb {

void set_error(const char *fmt, ...);

#define OCRB_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ocrb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return OCRB_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define OCRB_TRY(call)            \
  do {                            \
    int rc__ = (call);            \
    if (rc__ != OCRB_OK) return rc__; \
  } while (0)

#define OCRB_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ocrb::set_error(__VA_ARGS__);    \
      return OCRB_ERR_INVALID;         \
    }                                  \
  } while (0)

// grow-only device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return OCRB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
      return OCRB_ERR_CUDA;
    }
    cap = want;
    return OCRB_OK;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// grow-only pinned host buffer
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return OCRB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMallocHost(%zu) -> %s", want, cudaGetErrorString(e));
      return OCRB_ERR_CUDA;
    }
    cap = want;
    return OCRB_OK;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct PostprocWorkspace;  // postproc.cu
struct PipelineWorkspace;  // pipeline.cu

// per-launch CUDA-event timeline of one ctx (the B200 counterpart of the reference's
// measure_time! macro, macros.rs:46-71): when enabled, check_launch()/prof_mark() record an
// event after every launch or copy; the span between consecutive events is attributed to
// the later one's name.  Off by default (zero cost: one branch per launch).
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  size_t used = 0;
};

}  // namespace ocrb

struct ocrb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  // staging for host-pointer arguments (indexed slots so one call can stage several)
  ocrb::DevBuf stage[6];
  ocrb::PinBuf pin[3];
  ocrb::DevBuf decode_rgba;     // decoded RGBA arena of ocrb_preprocess_files (decode.cu)
  ocrb::DevBuf ccl_tile_empty;  // one byte per CCL tile of the last labelling: 1 = no foreground pixel (ccl.cu)
  ocrb::DevBuf ccl_seam_list;   // [0] = count, [1..] = tiles whose seams need a whole warp (ccl.cu)
  ocrb::PostprocWorkspace *pp = nullptr;
  ocrb::PipelineWorkspace *pipe = nullptr;  // streams / events / buffers of ocrb_detect_and_recognize, created on first use
  std::vector<const void *> smem_attr_done;  // kernels whose dynamic shared-memory limit this ctx has raised
  int sm_limit = 0;                          // > 0: persistent convolution kernels use at most this many SMs (pipeline.cu)
  ocrb::Profiler prof;
  // SMs a persistent kernel of this context may occupy
  int sm_budget() const { return sm_limit > 0 && sm_limit < sm_count ? sm_limit : sm_count; }
};

namespace ocrb {

inline bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Returns a device view of `p` (nbytes).  Host memory is copied into ctx->stage[slot].
inline int to_device(ocrb_ctx *ctx, int slot, const void *p, size_t nbytes, const void **out) {
  if (is_device_ptr(p)) {
    *out = p;
    return OCRB_OK;
  }
  OCRB_TRY(ctx->stage[slot].reserve(nbytes));
  OCRB_CUDA(cudaMemcpyAsync(ctx->stage[slot].p, p, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  *out = ctx->stage[slot].p;
  return OCRB_OK;
}

// Returns a device buffer to write results into; if `p` is host memory the buffer is
// ctx->stage[slot] and finish_output() copies it back.
inline int out_device(ocrb_ctx *ctx, int slot, void *p, size_t nbytes, void **out) {
  if (is_device_ptr(p)) {
    *out = p;
    return OCRB_OK;
  }
  OCRB_TRY(ctx->stage[slot].reserve(nbytes));
  *out = ctx->stage[slot].p;
  return OCRB_OK;
}

inline int finish_output(ocrb_ctx *ctx, void *user, const void *dev, size_t nbytes) {
  if (user != dev) OCRB_CUDA(cudaMemcpyAsync(user, dev, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  return OCRB_OK;
}

inline int sync(ocrb_ctx *ctx) {
  OCRB_CUDA(cudaStreamSynchronize(ctx->stream));
  return OCRB_OK;
}

inline void prof_mark(ocrb_ctx *ctx, const char *what) {
  Profiler &pr = ctx->prof;
  if (!pr.on) return;
  if (pr.used == pr.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { pr.on = false; return; }
    pr.ev.push_back(e);
    pr.names.emplace_back();
  }
  pr.names[pr.used] = what;
  cudaEventRecord(pr.ev[pr.used], ctx->stream);
  pr.used += 1;
}

inline int check_launch(ocrb_ctx *ctx, const char *what) {
  ctx->launches += 1;
  prof_mark(ctx, what);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch %s -> %s", what, cudaGetErrorString(e));
    return OCRB_ERR_CUDA;
  }
  return OCRB_OK;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (device, function); remembered per ctx, so that two contexts
// (devices, host threads) never share state
template <class K>
inline int ensure_dyn_smem(ocrb_ctx *ctx, K kern, int bytes) {
  const void *key = reinterpret_cast<const void *>(kern);
  for (const void *k : ctx->smem_attr_done)
    if (k == key) return OCRB_OK;
  OCRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  ctx->smem_attr_done.push_back(key);
  return OCRB_OK;
}

// One-item-per-thread kernels whose items follow wildly different control flow (border tracing,
// Douglas-Peucker, polygon offsetting) run only SPARSE_LANES lanes per warp: a warp executes
// its divergent lanes one after another, so fewer lanes per warp = shorter critical path, and
// the extra warps spread over all SMs.  Index of this thread's item, or -1.
constexpr int SPARSE_LANES = 4;
#ifdef __CUDACC__
__device__ __forceinline__ int64_t sparse_item_index() {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  return lane < SPARSE_LANES ? gw * SPARSE_LANES + lane : -1;
}
#endif
inline unsigned sparse_grid(int64_t n_items, int block_threads) {
  return (unsigned)cdiv(cdiv(n_items, SPARSE_LANES) * 32, block_threads);
}

}  // namespace ocrb

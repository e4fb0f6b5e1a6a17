// m3tsz_ctx.h -- internal: the context object and host helpers shared by the C-ABI
// translation units (m3tsz_capi.cu, m3tsz_stream.cu).  Not installed.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct m3tsz_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // used by the *_host entry points
  cudaStream_t stream2 = nullptr;  // second lane of the chunked H2D / kernel / D2H pipeline
  cudaEvent_t ev = nullptr;
  uint64_t launches = 0;
  char last_error[256] = {0};
  Scratch s[48];
  int32_t *d_flag = nullptr;  // [2]
  void *h_stage[2] = {nullptr, nullptr};  // pinned staging for per-chunk offsets
  size_t h_stage_bytes[2] = {0, 0};
  std::mutex handle_mu;  // serialises the streaming handles' use of the scratch below
};

namespace m3tsz {
namespace host {

inline int set_cuda_error(m3tsz_ctx *ctx, cudaError_t e, const char *where) {
  if (ctx) snprintf(ctx->last_error, sizeof(ctx->last_error), "%s: %s", where, cudaGetErrorString(e));
  return M3TSZ_ERR_CUDA;
}

#define CK(call)                                                 \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return set_cuda_error(ctx, _e, #call); \
  } while (0)

inline int ensure(m3tsz_ctx *ctx, int slot, size_t bytes, void **out) {
  Scratch &sc = ctx->s[slot];
  if (bytes == 0) bytes = 16;
  if (sc.bytes < bytes) {
    if (sc.ptr) CK(cudaFree(sc.ptr));
    sc.ptr = nullptr;
    sc.bytes = 0;
    size_t want = bytes + bytes / 8;  // grow-only with slack
    cudaError_t e = cudaMalloc(&sc.ptr, want);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      want = bytes;
      CK(cudaMalloc(&sc.ptr, want));
    }
    sc.bytes = want;
  }
  *out = sc.ptr;
  return M3TSZ_OK;
}

inline int ensure_stage(m3tsz_ctx *ctx, int i, size_t bytes) {
  if (ctx->h_stage_bytes[i] >= bytes) return M3TSZ_OK;
  if (ctx->h_stage[i]) CK(cudaFreeHost(ctx->h_stage[i]));
  ctx->h_stage[i] = nullptr;
  ctx->h_stage_bytes[i] = 0;
  CK(cudaMallocHost(&ctx->h_stage[i], bytes));
  ctx->h_stage_bytes[i] = bytes;
  return M3TSZ_OK;
}

// Series per pipeline chunk: ~8 chunks per call, but never tiny ones.
inline uint64_t pick_chunk(uint64_t n_series, uint64_t bytes_per_series) {
  uint64_t ch = (n_series + 7) / 8;
  uint64_t min_ch = (16ull << 20) / (bytes_per_series ? bytes_per_series : 1);
  if (min_ch < 2048) min_ch = 2048;
  if (ch < min_ch) ch = min_ch;
  if (ch > n_series) ch = n_series;
  ch = (ch + 127) & ~127ull;  // whole thread blocks
  return ch;
}

// Entry points run on the context's device and restore the caller's current device on
// return (a process that drives several GPUs must not be retargeted by a library call).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) {
      err = cudaSetDevice(device);
      ok = (err == cudaSuccess);
    } else {
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline bool valid_opts(const m3tsz_options *o) {
  return o && (o->int_optimized == 0 || o->int_optimized == 1) && o->default_time_unit >= 0 &&
         o->default_time_unit <= 8;
}

}  // namespace host
}  // namespace m3tsz


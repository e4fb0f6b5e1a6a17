// m3tsz_query.cu -- the two consumers directly after the decode path:
//
//  * Prometheus conversion epilogue: iteratorToPromResult,
//    /root/reference/src/query/storage/prom_converter.go:42-120 (ns -> ms timestamps,
//    valueDecreaseTolerance, counter-reset cumulativeSum per resolution window), over the
//    decoded / merged (ts, value) arrays that m3tsz_decode_batch / m3tsz_merge_series_batch
//    leave in HBM -- the last per-datapoint host loop of a fetch.
//  * Tile aggregation glue (storage.TileAggregator, src/dbnode/storage/types.go:1444-1472): the
//    fused decode+downsample kernel leaves window-major Gauge aggregates, the encoder's IN = 2 input
//    stage (m3tsz_encode.cu) turns them into (window end, Gauge.ValueOf(type)) datapoints on the fly;
//    only small helper kernels live here.
#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

namespace m3tsz {

// ---------------------------------------------------------------------------
// Prometheus epilogue
// ---------------------------------------------------------------------------
// Plain conversion (no tolerance, no counter normalisation): every datapoint is
// independent -- one warp per series, lanes stride over the datapoints (coalesced).
__global__ void prom_simple_kernel(const PromParams p) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= p.n_series) return;
  const uint32_t n = p.n_points[w];
  const uint32_t m = n < p.out_cap ? n : (uint32_t)p.out_cap;
  const int64_t *ts = p.ts + w * p.cap;
  const double *val = p.val + w * p.cap;
  int64_t *to = p.ts_out + w * p.out_cap;
  double *vo = p.val_out + w * p.out_cap;
  for (uint32_t i = lane; i < m; i += 32) {
    to[i] = ts[i] / 1000000;  // TimeToPromTimestamp, converter.go:388-391 (truncating)
    vo[i] = val[i];
  }
  if (lane == 0) {
    p.n_out[w] = n;
    if (p.status) p.status[w] = n > p.out_cap ? M3TSZ_ERR_CAPACITY : M3TSZ_OK;
  }
}

// General form: the tolerance clamp and the cumulative sum are sequential recurrences in
// floating point (order matters bit for bit), so one thread walks one series.
__global__ void prom_general_kernel(const PromParams p) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_series) return;
  const uint32_t n = p.n_points[s];
  const int64_t *ts = p.ts + s * p.cap;
  const double *val = p.val + s * p.cap;
  int64_t *to = p.ts_out + s * p.out_cap;
  double *vo = p.val_out + s * p.out_cap;
  const bool handle = p.handle_resets && p.handle_resets[s] != 0;
  const double tol = p.tolerance;
  bool first = true;
  double cumulative = 0.0, prev_v = 0.0;
  int64_t prev_t = 0;
  uint64_t k = 0;
  for (uint32_t i = 0; i < n; i++) {
    const int64_t t = ts[i];
    double v = val[i];
    if (tol > 0 && t < p.tolerance_until) {  // prom_converter.go:68-72
      if (!first && v < prev_v && v > __dmul_rn(prev_v, __dsub_rn(1.0, tol))) v = prev_v;
    }
    if (handle) {  // :84-98
      if (p.resolution != 0 && t / p.resolution != prev_t / p.resolution && !first) {
        if (k < p.out_cap) {
          to[k] = prev_t / 1000000;
          vo[k] = cumulative;
        }
        k++;
      }
      if (v < prev_v)
        cumulative = __dadd_rn(cumulative, v);  // counter reset
      else
        cumulative = __dadd_rn(cumulative, __dsub_rn(v, prev_v));
    } else {
      if (k < p.out_cap) {
        to[k] = t / 1000000;
        vo[k] = v;
      }
      k++;
    }
    prev_t = t;
    prev_v = v;
    first = false;
  }
  if (handle && !first) {  // :113-118
    if (k < p.out_cap) {
      to[k] = prev_t / 1000000;
      vo[k] = cumulative;
    }
    k++;
  }
  p.n_out[s] = (uint32_t)k;
  if (p.status) p.status[s] = k > p.out_cap ? M3TSZ_ERR_CAPACITY : M3TSZ_OK;
}

cudaError_t launch_prom(const PromParams &p, cudaStream_t stream) {
  if (p.n_series == 0) return cudaSuccess;
  const unsigned tb = 256;
  if (!(p.tolerance > 0) && !p.handle_resets) {
    const uint64_t threads = p.n_series * 32ull;
    const uint64_t blocks = (threads + tb - 1) / tb;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    prom_simple_kernel<<<(unsigned)blocks, tb, 0, stream>>>(p);
  } else {
    const uint64_t blocks = (p.n_series + tb - 1) / tb;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    prom_general_kernel<<<(unsigned)blocks, tb, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Tile aggregation glue (the re-encode itself is the encoder's IN = 2 input stage, m3tsz_encode.cu)
// ---------------------------------------------------------------------------
__global__ void fill_i64_kernel(int64_t *dst, int64_t v, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
cudaError_t launch_fill_i64(int64_t *dst, int64_t v, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned tb = 256;
  fill_i64_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, stream>>>(dst, v, n);
  return cudaGetLastError();
}

// status merge: a series whose source stream failed to decode reports that error
__global__ void tiles_status_kernel(const int32_t *src_status, int32_t *status, uint64_t n) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n && src_status[s] != 0) status[s] = src_status[s];
}
cudaError_t launch_tiles_status(const int32_t *src_status, int32_t *status, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned tb = 256;
  tiles_status_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, stream>>>(src_status, status, n);
  return cudaGetLastError();
}

}  // namespace m3tsz

// host launcher of the warp-per-series prototype (see decode_warp_per_series.cu)
#include <cuda_runtime.h>
#include <stdint.h>
extern "C" __global__ void decode_warp_per_series(const uint8_t *, const uint64_t *, uint64_t, uint32_t, int64_t *,
                                                  double *);
extern "C" int launch_decode_warp_per_series(const uint8_t *streams, const uint64_t *offsets, uint64_t n_series,
                                             uint32_t n_points, int64_t *ts, double *val, void *stream) {
  const uint64_t blocks = (n_series + 3) / 4;
  decode_warp_per_series<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(streams, offsets, n_series, n_points, ts, val);
  return (int)cudaGetLastError();
}

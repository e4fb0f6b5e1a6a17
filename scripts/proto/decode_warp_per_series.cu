// PROTOTYPE (profiling artefact, not part of the product library): the mapping the north star
// words literally -- ONE WARP PER SERIES -- for the float-mode grammar, to put a measured number
// next to the lane-per-series kernel.
//
// The warp cooperates on memory exactly as the north star says: the stream is staged through
// shared memory with coalesced 128-bit loads (32 lanes x 16 B per step), decoded datapoints are
// staged in shared memory and written back as coalesced 256-byte rows.  The parse itself is a
// serial dependency chain (every field width depends on the previous XOR / delta), so one lane
// walks it while 31 wait at the tile boundaries -- that idle issue width is what the capture shows.
//
// Grammar subset: intOptimized = false, unit = second, no annotations / unit changes, n_points
// known (the bench streams).  Build: see scripts/r2_warp_per_series.py.
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int TILE_WORDS = 256;   // 1 KB of stream per refill (64 datapoints of <= 79 bits + the first one)
constexpr int OUT_TILE = 64;      // datapoints staged per write-back

__device__ __forceinline__ uint64_t peek64(const uint32_t *w, uint32_t pos) {  // MSB-first, words big-endian swapped
  const uint32_t i = pos >> 5;
  const uint32_t w0 = __byte_perm(w[i], 0, 0x0123), w1 = __byte_perm(w[i + 1], 0, 0x0123),
                 w2 = __byte_perm(w[i + 2], 0, 0x0123);
  const uint32_t hi = __funnelshift_l(w1, w0, pos), lo = __funnelshift_l(w2, w1, pos);
  return ((uint64_t)hi << 32) | lo;
}

extern "C" __global__ void __launch_bounds__(128) decode_warp_per_series(
    const uint8_t *streams, const uint64_t *offsets, uint64_t n_series, uint32_t n_points, int64_t *ts_out,
    double *val_out) {
  __shared__ __align__(16) uint32_t s_in[4][TILE_WORDS + 4];
  __shared__ int64_t s_t[4][OUT_TILE];
  __shared__ uint64_t s_v[4][OUT_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t sidx = (uint64_t)blockIdx.x * 4 + warp;
  if (sidx >= n_series) return;
  const uint64_t o0 = offsets[sidx], o1 = offsets[sidx + 1];
  const uint8_t *base = streams + (o0 & ~15ull);  // 16-byte aligned staging base
  uint32_t *in = s_in[warp];
  uint64_t tile_byte0 = 0;  // byte offset (from base) of in[0]
  auto refill = [&](uint64_t from_byte) {  // coalesced: 32 lanes x 16 B per iteration
    tile_byte0 = from_byte & ~15ull;
    const uint4 *src = reinterpret_cast<const uint4 *>(base + tile_byte0);
    const uint64_t avail = (o1 - (o0 & ~15ull)) - tile_byte0 + 16;
    for (int q = lane; q < TILE_WORDS / 4 + 1; q += 32) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if ((uint64_t)q * 16 < avail) v = __ldg(src + q);
      reinterpret_cast<uint4 *>(in)[q] = v;
    }
    __syncwarp();
  };
  refill(0);
  // lane 0's serial state
  uint64_t bitpos = (o0 & 15ull) * 8;  // from base
  int64_t prev_time = 0, prev_delta = 0;
  uint64_t prev_bits = 0, prev_xor = 0;
  int64_t *to = ts_out + sidx * n_points;
  uint64_t *vo = reinterpret_cast<uint64_t *>(val_out) + sidx * n_points;
  for (uint32_t d0 = 0; d0 < n_points; d0 += OUT_TILE) {
    const uint32_t nd = min((uint32_t)OUT_TILE, n_points - d0);
    if (lane == 0) {
      for (uint32_t k = 0; k < nd; k++) {
        uint32_t p = (uint32_t)(bitpos - tile_byte0 * 8);
        const bool first = (d0 + k == 0);
        if (first) {
          prev_time = (int64_t)peek64(in, p);
          p += 64;
        }
        // delta-of-delta, second scheme
        uint64_t w = peek64(in, p);
        int64_t dod = 0;
        if (w >> 63) {
          int nb, hb;
          if (!((w >> 62) & 1)) { hb = 2; nb = 7; }
          else if (!((w >> 61) & 1)) { hb = 3; nb = 9; }
          else if (!((w >> 60) & 1)) { hb = 4; nb = 12; }
          else { hb = 4; nb = 32; }
          dod = ((int64_t)(w << hb)) >> (64 - nb);
          p += hb + nb;
        } else {
          p += 1;
        }
        prev_delta += dod * 1000000000ll;
        prev_time += prev_delta;
        // value
        if (first) {
          prev_bits = peek64(in, p);
          prev_xor = prev_bits;
          p += 64;
        } else {
          w = peek64(in, p);
          if (!(w >> 63)) {
            prev_xor = 0;
            p += 1;
          } else if (!((w >> 62) & 1)) {
            const int lz = prev_xor ? __clzll((long long)prev_xor) : 64;
            const int tz = prev_xor ? (__ffsll((long long)prev_xor) - 1) : 0;
            const int n = 64 - lz - tz;
            p += 2;
            const uint64_t m = n ? (peek64(in, p) >> (64 - n)) : 0ull;
            p += n;
            prev_xor = m << tz;
            prev_bits ^= prev_xor;
          } else {
            const int lz = (int)((w >> 56) & 63), n = (int)((w >> 50) & 63) + 1;
            p += 14;
            const uint64_t m = peek64(in, p) >> (64 - n);
            p += n;
            const int tz = 64 - lz - n;
            prev_xor = tz < 0 ? 0ull : (m << tz);
            prev_bits ^= prev_xor;
          }
        }
        bitpos = tile_byte0 * 8 + p;
        s_t[warp][k] = prev_time;
        s_v[warp][k] = prev_bits;
      }
    }
    __syncwarp();
    // every output tile is followed by a refill at lane 0's current position
    const uint64_t bp = __shfl_sync(0xffffffffu, bitpos, 0);
    for (uint32_t k = lane; k < nd; k += 32) {  // coalesced write-back
      to[d0 + k] = s_t[warp][k];
      vo[d0 + k] = s_v[warp][k];
    }
    __syncwarp();
    refill(bp >> 3);
  }
}

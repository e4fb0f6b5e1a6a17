// m3tsz_checksum.cu -- batch segment checksums on device:
// the Adler-32 the fileset read path verifies for every stream it reads
//   ts.Segment.CalculateChecksum   src/dbnode/ts/segment.go:60-76
//   digest.Checksum                src/dbnode/digest/digest.go:36-38
//   check sites                    src/dbnode/persist/fs/read.go:395-397, seek.go:370-373
// over the same CSR layout the decoder takes (= the data file + index entries).
//
// Mapping: one WARP per stream.  With a = 1 + sum d_i and b = n + sum (n - i) d_i
// (both mod 65521) the checksum is a function of S1 = sum d_i and W = sum i d_i,
// which are plain sums: lanes take 16-byte chunks round-robin (one coalesced
// 512-byte segment per warp instruction), add up 4 bytes at a time with dp4a,
// and a shuffle reduction combines them.  64-bit accumulators cannot overflow for
// streams below 2^28 bytes (W < 255 * 2^55).  HBM-bound: reads every byte once.
#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

namespace m3tsz {

constexpr uint32_t kAdlerMod = 65521u;
constexpr int CK_WARPS = 8;

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
  return v;
}

__global__ void __launch_bounds__(CK_WARPS * 32) checksum_kernel(const ChecksumParams p) {
  const int lane = threadIdx.x & 31;
  const uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= p.n_series) return;
  const uint64_t o0 = p.offsets[s];
  // index-entry addressing (Offset, Size): the streams may sit anywhere in the buffer, in any
  // order (the packed encoder places them in completion order); CSR otherwise
  const uint64_t o1 = p.lengths ? o0 + p.lengths[s] : p.offsets[s + 1];
  if (o1 < o0 || o1 > p.streams_bytes || o1 - o0 >= (1ull << 28)) {
    if (lane == 0) {
      if (p.out) p.out[s] = 0;
      if (p.status) p.status[s] = (o1 < o0 || o1 > p.streams_bytes) ? M3TSZ_ERR_INVALID_ARG : M3TSZ_ERR_STREAM_TOO_LARGE;
    }
    return;
  }
  const uint64_t n = o1 - o0;
  const uint8_t *base = p.streams + o0;
  // [0, head): bytes before the first 16-byte boundary; then n_chunks aligned chunks; then the tail
  uint64_t head = (16u - (uint32_t)((uintptr_t)base & 15u)) & 15u;
  if (head > n) head = n;
  const uint64_t n_chunks = (n - head) >> 4;
  const uint64_t tail0 = head + (n_chunks << 4);
  uint64_t S1 = 0, W = 0;
  if ((uint64_t)lane < head) {
    const uint64_t d = base[lane];
    S1 += d;
    W += (uint64_t)lane * d;
  }
  if (lane >= 16 && tail0 + (uint64_t)(lane - 16) < n) {
    const uint64_t i = tail0 + (uint64_t)(lane - 16);
    const uint64_t d = base[i];
    S1 += d;
    W += i * d;
  }
  const uint4 *body = reinterpret_cast<const uint4 *>(base + head);
  for (uint64_t c = lane; c < n_chunks; c += 32) {
    const uint4 v = __ldg(body + c);
    // per 32-bit word: s_k = sum of its 4 bytes, t_k = sum j * byte_j (j = 0..3, little endian)
    const uint32_t s0 = __dp4a(v.x, 0x01010101u, 0u), s1 = __dp4a(v.y, 0x01010101u, 0u),
                   s2 = __dp4a(v.z, 0x01010101u, 0u), s3 = __dp4a(v.w, 0x01010101u, 0u);
    uint32_t t = __dp4a(v.x, 0x03020100u, 0u);
    t = __dp4a(v.y, 0x03020100u, t);
    t = __dp4a(v.z, 0x03020100u, t);
    t = __dp4a(v.w, 0x03020100u, t);
    t += 4u * s1 + 8u * s2 + 12u * s3;  // byte index inside the chunk = 4k + j
    const uint32_t sc = s0 + s1 + s2 + s3;
    const uint64_t i0 = head + (c << 4);
    S1 += sc;
    W += i0 * (uint64_t)sc + t;
  }
  S1 = warp_sum_u64(S1);
  W = warp_sum_u64(W);
  if (lane == 0) {
    const uint64_t nm = n % kAdlerMod, s1m = S1 % kAdlerMod, wm = W % kAdlerMod;
    const uint32_t a = (uint32_t)((1ull + s1m) % kAdlerMod);
    const uint32_t b = (uint32_t)((nm + (nm * s1m) % kAdlerMod + kAdlerMod - wm) % kAdlerMod);
    const uint32_t ck = (b << 16) | a;
    if (p.out) p.out[s] = ck;
    if (p.status) p.status[s] = (p.expected && p.expected[s] != ck) ? M3TSZ_ERR_CHECKSUM_MISMATCH : M3TSZ_OK;
  }
}

cudaError_t launch_checksum(const ChecksumParams &p, cudaStream_t stream) {
  if (p.n_series == 0) return cudaSuccess;
  const uint64_t blocks = (p.n_series + CK_WARPS - 1) / CK_WARPS;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  checksum_kernel<<<(unsigned)blocks, CK_WARPS * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace m3tsz

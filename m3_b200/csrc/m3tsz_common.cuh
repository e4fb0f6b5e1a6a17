// m3tsz_common.cuh -- shared device helpers for the sm_100a M3TSZ kernels.
//
// Format constants follow the reference (paths under
// /root/reference/src/dbnode/encoding): scheme.go:30-62 (marker + time buckets),
// m3tsz/m3tsz.go:28-62 (value opcodes), SURVEY.md Appendix A (normative summary).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/m3tsz_b200.h"

namespace m3tsz {

constexpr unsigned FULL_MASK = 0xffffffffu;

// marker scheme (scheme.go:30-38): 9-bit opcode 0x100 + 2-bit value
constexpr uint32_t kMarkerOpcode = 0x100;
constexpr int kMarkerBits = 11;
constexpr int kMarkerEOS = 0, kMarkerAnnotation = 1, kMarkerTimeUnit = 2;

constexpr int kMaxMult = 6;  // m3tsz.go:60

// Time-encoding scheme kinds (scheme.go:40-53,109-120,160-165)
enum SchemeKind : int {
  kSchemeNone = 0,  // nil scheme (errNoTimeSchemaForUnit)
  kScheme32 = 1,    // s, ms: buckets 7/9/12, default 32 bits
  kScheme64 = 2,    // us, ns: buckets 7/9/12, default 64 bits
  kSchemeZero = 3   // m, h, d, y: the all-zero scheme left in the table
};

__host__ __device__ __forceinline__ bool unit_is_valid(int u) { return u > 0 && u < 9; }

__host__ __device__ __forceinline__ int64_t unit_nanos(int u) {  // unit.go:185-195
  switch (u) {
    case 1: return 1000000000LL;
    case 2: return 1000000LL;
    case 3: return 1000LL;
    case 4: return 1LL;
    case 5: return 60LL * 1000000000LL;
    case 6: return 3600LL * 1000000000LL;
    case 7: return 86400LL * 1000000000LL;
    case 8: return 365LL * 86400LL * 1000000000LL;
    default: return 0;
  }
}

__host__ __device__ __forceinline__ int scheme_kind_for_unit(int u) {
  if (!unit_is_valid(u)) return kSchemeNone;
  if (u <= 2) return kScheme32;
  if (u <= 4) return kScheme64;
  return kSchemeZero;
}

// initialTimeUnit, m3tsz/timestamp_encoder.go:248-259
__host__ __device__ __forceinline__ int initial_time_unit(int64_t start, int unit) {
  if (!unit_is_valid(unit)) return 0;
  return (start % unit_nanos(unit) == 0) ? unit : 0;
}

// Go shift semantics: x >> n == 0 and x << n == 0 for n >= 64.
__device__ __forceinline__ uint64_t shr64(uint64_t x, int n) { return n >= 64 ? 0ull : x >> n; }
__device__ __forceinline__ uint64_t shl64(uint64_t x, int n) { return n >= 64 ? 0ull : x << n; }
// PTX shifts: amounts above 63 (incl. "negative" ones read as unsigned) are clamped and shift everything
// out -- defined behaviour, so no guard compare + select around the hot path's shifts
__device__ __forceinline__ uint64_t shr64_clamp(uint64_t x, uint32_t n) {
  uint64_t r;
  asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n));
  return r;
}
__device__ __forceinline__ uint64_t shl64_clamp(uint64_t x, uint32_t n) {
  uint64_t r;
  asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n));
  return r;
}

// encoding.SignExtend (encoding.go:45-49) for 1 <= n <= 64
__device__ __forceinline__ int64_t sign_extend(uint64_t v, int n) {
  int sh = 64 - n;
  return ((int64_t)(v << sh)) >> sh;
}

// encoding.LeadingAndTrailingZeros (encoding.go:33-43): (64, 0) for v == 0
__device__ __forceinline__ void lz_tz(uint64_t v, int &lz, int &tz) {
  uint32_t hi, lo;  // unpacked in PTX: "hi != 0" must stay a 32-bit test (not v >= 2^32: two compares)
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  lz = hi ? __clz((int)hi) : 32 + __clz((int)lo);  // 64 for v == 0
  const int tlo = __clz((int)__brev(lo)), thi = 32 + __clz((int)__brev(hi));
  tz = lo ? tlo : (hi ? thi : 0);
}

__device__ __forceinline__ int num_sig(uint64_t v) { return 64 - __clzll((long long)v); }  // encoding.go:29-31

// multipliers[], m3tsz/m3tsz.go:131-140 (exact powers of ten)
__device__ __forceinline__ double mult_pow10(int m) {
  switch (m) {
    case 0: return 1.0;
    case 1: return 10.0;
    case 2: return 100.0;
    case 3: return 1000.0;
    case 4: return 10000.0;
    case 5: return 100000.0;
    default: return 1000000.0;
  }
}

// Big-endian 32-bit word `widx` of a byte buffer whose base is 4-byte aligned.
// Words (partly) past `nbytes` are zero-filled, never read out of bounds.
__device__ __forceinline__ uint32_t load_be32(const uint8_t *base, uint64_t nbytes, uint64_t widx) {
  uint64_t b = widx << 2;
  if (b + 4 <= nbytes) {
    uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(base + b));
    return __byte_perm(v, 0, 0x0123);
  }
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t x = (b + i < nbytes) ? (uint32_t)base[b + i] : 0u;
    v = (v << 8) | x;
  }
  return v;
}

}  // namespace m3tsz

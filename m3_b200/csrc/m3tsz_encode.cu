// m3tsz_encode.cu -- batch M3TSZ encode for sm_100a (+ stream compaction).
//
// Mapping: one LANE per series, 32 series per warp.  The
// int-optimised value grammar is a sequential state machine per series
// (maxMult / isFloat / IntSigBitsTracker, m3tsz/encoder.go:148-250), so each
// lane runs its series' encoder while the WARP cooperates on memory:
// (ts, value) inputs are staged through a transposed shared-memory tile from
// coalesced 128-byte loads, each datapoint's code (<= 32-bit header + <= 64-bit
// payload) is merged into a per-lane bit accumulator with funnel shifts, full
// 32-bit words go to a transposed output tile, and the tile is written back
// with coalesced 128-byte stores per series.
//
// Format: SURVEY.md Appendix A; reference encode path
//   m3tsz/encoder.go:90-250, m3tsz/timestamp_encoder.go:72-259,
//   m3tsz/float_encoder_iterator.go:69-103, m3tsz/int_sig_bits_tracker.go:35-91,
//   m3tsz/m3tsz.go:78-119, ostream.go:133-221, scheme.go:198-220.
#include <cstdlib>
#include <cstring>
#include <cub/device/device_scan.cuh>
#include <cuda.h>  // CUtensorMap (types only: the encoder is resolved through cudaGetDriverEntryPoint)

#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

namespace m3tsz {

constexpr int ENC_WARPS = 4;
constexpr int ENC_STRIDE = 33;
constexpr int ENC_IN_T = 8;     // datapoints per input tile (double buffered, cp.async)
#ifndef M3_ENC_OUT_W
#define M3_ENC_OUT_W 24  // best of the 16/20/24/40 sweep at 1M x 1440 (smaller tile = more L1)
#endif
#ifndef M3_ENC_MIN_BLOCKS
#define M3_ENC_MIN_BLOCKS 4
#endif
#ifndef M3_ENC_UNROLL_PM
#define M3_ENC_UNROLL_PM 0  // 1: unroll the datapoint loop over one input tile (measured: 8.74 -> 9.54 ms, 9.6 K instructions)
#endif
#ifndef M3_ENC_MIN_BLOCKS_PM
#define M3_ENC_MIN_BLOCKS_PM 4  // point-major input stage (smaller tiles: more blocks fit)
#endif
constexpr int ENC_OUT_W = M3_ENC_OUT_W;   // output tile words per lane
constexpr int ENC_GUARD = 10;   // words a single datapoint (no annotation) may add
constexpr int ENC_OUT_TILE_WORDS = ENC_OUT_W * ENC_STRIDE;
#ifndef M3_ENC_IN_T_PM
// rows per input tile of the point-major input stage (IN = 1).  With the cp.async fills: 9.60 / 9.10 / 9.58 ms for
// 8 / 4 / 2 rows (a smaller tile leaves more L1 to the fills).  With tensor copies (M3_ENC_BULK_PM = 2, which do not
// allocate L1 lines) one copy per array costs the same whatever the box: 8.65 ms with 8 rows, 9.06 with 4 (the
// cp.async build of the same day: 8.73).
#define M3_ENC_IN_T_PM 8
#endif
// input stage IN: 0 series-major tiles (8 rows: one 64-byte segment per series and array), 1 point-major
// tiles, 2 Gauge aggregates read directly (no input tiles)
#ifndef M3_ENC_BULK_PM
// 1: the point-major input stage (IN = 1) fills its tiles with TMA 1-D bulk copies (cp.async.bulk, one 256-byte
// row of 32 series per copy, issued by one lane, completion on a per-warp mbarrier) whenever a batch is a full
// warp of equally long series and the arrays / row pitch are 16-byte aligned; other batches keep the lane-local
// 8-byte cp.async fills.  A tile row is then 32 datapoints exactly (no padding column: bulk destinations must be
// 16-byte aligned and the lane-local reads are conflict-free either way).
// 2: the same tiles by TMA TENSOR copies: the arrays are described as 2-D tensors [points][series] of 8-byte
// elements (cuTensorMapEncodeTiled, box = IN_T rows x 32 series), so one cp.async.bulk.tensor.2d per array moves a
// whole tile; rows past the end and series past n_series are zero-filled by the hardware, which makes ragged
// batches and the last partial tile the same code.  Needs 16-byte aligned arrays and an even n_series < 2^31,
// else the cp.async fills.  Measured at 1 M x 1440 (profiles/r02_decode_history.md): 0 -> 8.73 ms (4-row tiles),
// 2 -> 8.65 ms with 8-row tiles / 9.06 with 4-row tiles; the full GPU suite is green on the build with 2.
#define M3_ENC_BULK_PM 2
#endif
template <int IN>
__host__ __device__ constexpr int enc_in_t() {
  return IN == 1 ? M3_ENC_IN_T_PM : ENC_IN_T;
}
template <int IN>
__host__ __device__ constexpr int enc_bulk() {
  return IN == 1 ? M3_ENC_BULK_PM : 0;
}
// the two tensor maps of the point-major input arrays (M3_ENC_BULK_PM == 2), a __grid_constant__ kernel parameter
struct EncTensorMaps {
  alignas(64) CUtensorMap ts;
  alignas(64) CUtensorMap val;
  int ok;  // 0: not built (unaligned arrays, odd n_series, no driver entry point): cp.async fills
};
// datapoints (8-byte cells) per input tile row
template <int IN>
__host__ __device__ constexpr int enc_in_stride() {
  return enc_bulk<IN>() ? 32 : ENC_STRIDE;
}
template <int IN>
__host__ __device__ constexpr size_t enc_warp_smem() {
  const size_t raw = ((IN == 0 || IN == 1) ? 4 * (size_t)(enc_in_t<IN>() * enc_in_stride<IN>()) * 8 : 0) +
                     (size_t)ENC_OUT_TILE_WORDS * 4 + (enc_bulk<IN>() ? 16 : 0);  // (+ two mbarriers per warp)
  return enc_bulk<IN>() == 2 ? ((raw + 127) & ~(size_t)127) : raw;  // tensor copies land on 128-byte boundaries
}

struct EncLane {
  uint32_t carry, sh, k, words_out;
  int64_t prev_time, prev_delta;
  uint64_t prev_bits;
  int plz, ptz;  // leading / trailing zeros of the previous XOR ((64, 0) for a zero XOR)
  double int_val;
  int unit;
  int num_sig, hi_lower_sig, n_lower_sig, max_mult;
  int err;
  uint32_t n_enc;
  bool is_float;
};

// ---- bit output -----------------------------------------------------------
// appends the low n (1..32) bits of v
__device__ __forceinline__ void put32(EncLane &s, uint32_t *tile, int lane, uint32_t v, int n) {
  const uint32_t c0 = v << (32 - n);
  const uint32_t m0 = s.carry | __funnelshift_rc(c0, 0u, s.sh);
  const uint32_t m1 = __funnelshift_rc(0u, c0, s.sh);
  const uint32_t tot = s.sh + (uint32_t)n;
  if (tot >= 32) {
    tile[s.k * ENC_STRIDE + lane] = m0;
    s.k++;
    s.carry = m1;
    s.sh = tot - 32;
  } else {
    s.carry = m0;
    s.sh = tot;
  }
}
// appends the low n (0..64) bits of v
__device__ __forceinline__ void put64(EncLane &s, uint32_t *tile, int lane, uint64_t v, int n) {
  if (n > 32) {
    put32(s, tile, lane, (uint32_t)(v >> 32), n - 32);
    put32(s, tile, lane, (uint32_t)v, 32);
  } else if (n > 0) {
    put32(s, tile, lane, (uint32_t)v, n);
  }
}

// appends header (low hb bits of hdr, hb 0..32) followed by a payload given
// LEFT-ALIGNED in P (its top plen bits, plen 0..64; bits below must be zero): one
// merge of up to 96 bits into the accumulator.
__device__ __forceinline__ void emit_code_p(EncLane &s, uint32_t *tile, int lane, uint32_t hdr, int hb,
                                            uint64_t P, int plen) {
  const uint32_t H = hb ? (hdr << (32 - hb)) : 0u;
  const uint32_t Ph = (uint32_t)(P >> 32), Pl = (uint32_t)P;
  const uint32_t c0 = H | __funnelshift_rc(Ph, 0u, (uint32_t)hb);
  const uint32_t c1 = __funnelshift_rc(Pl, Ph, (uint32_t)hb);
  const uint32_t c2 = __funnelshift_rc(0u, Pl, (uint32_t)hb);
  const uint32_t m0 = s.carry | __funnelshift_rc(c0, 0u, s.sh);
  const uint32_t m1 = __funnelshift_rc(c1, c0, s.sh);
  const uint32_t m2 = __funnelshift_rc(c2, c1, s.sh);
  const uint32_t m3 = __funnelshift_rc(0u, c2, s.sh);
  const uint32_t tot = s.sh + (uint32_t)(hb + plen);
  const uint32_t full = tot >> 5;
  // predicated stores (kept out of the compiler's hands: it turns the chain into a
  // divergent branch region otherwise)
  const uint32_t ta = (uint32_t)__cvta_generic_to_shared(tile + s.k * ENC_STRIDE + lane);
  asm volatile(
      "{\n\t.reg .pred p0, p1, p2;\n\t"
      "setp.gt.u32 p0, %0, 0;\n\tsetp.gt.u32 p1, %0, 1;\n\tsetp.gt.u32 p2, %0, 2;\n\t"
      "@p0 st.shared.u32 [%1], %2;\n\t@p1 st.shared.u32 [%1+132], %3;\n\t@p2 st.shared.u32 [%1+264], %4;\n\t}"
      ::"r"(full), "r"(ta), "r"(m0), "r"(m1), "r"(m2)
      : "memory");
  static_assert(ENC_STRIDE * 4 == 132, "store offsets above");
  const uint32_t c01 = full == 0 ? m0 : m1, c23 = full == 2 ? m2 : m3;
  s.carry = full < 2 ? c01 : c23;
  s.k += full;
  s.sh = tot & 31u;
}
// same with the payload right-aligned (its low plen bits)
__device__ __forceinline__ void emit_code(EncLane &s, uint32_t *tile, int lane, uint32_t hdr, int hb,
                                          uint64_t payload, int plen) {
  emit_code_p(s, tile, lane, hdr, hb, plen ? (payload << (64 - plen)) : 0ull, plen);
}

// ---- convertToIntFloat (m3tsz.go:78-119) ----------------------------------
// Cheap NECESSARY condition for convertToIntFloat(v, cur) to return an int for
// any cur in [0,6]: if some v*10^m lies within one ulp of an integer, then
// fl(|v|*1e6) lies within 4 ulps of an integer (DESIGN.md §4.2 has the proof);
// tested with margin 8.  Returns true ("maybe int") for everything it cannot
// rule out, including NaN/Inf/huge/tiny values.
__device__ __forceinline__ bool maybe_int(double v) {
  const double a = fabs(v);
  const double p = __dmul_rn(a, 1000000.0);
  // on the (otherwise idle) FP64 pipe: p - rint(p) is exact, and p * 2^-49 lies in
  // [8, 16) ulps of p, so "far" implies more than 8 ulps from the nearest integer
  const double r = __dsub_rn(p, rint(p));
  const bool far = fabs(r) > __dmul_rn(p, 0x1p-49);
  const bool in_rng = (p >= 1.0) && (p < 0x1p48);   // NaN fails both: "maybe int"
  const bool tiny_v = (p < 0.5) && (a >= 1e-300);   // only N = 0 is near, i.e. v ~ 0
  return !((in_rng && far) || tiny_v);
}

// Go int64(float64) on amd64 (CVTTSD2SQ): out of range / NaN -> 0x8000000000000000
__device__ __forceinline__ uint64_t go_f64_to_u64_via_i64(double a) {
  if (!(a >= -9223372036854775808.0 && a < 9223372036854775808.0)) return 0x8000000000000000ull;
  return (uint64_t)__double2ll_rz(a);
}

__device__ __noinline__ void convert_to_int_float(double v, int cur, double &val, int &mult,
                                                  bool &is_float) {
  if (cur == 0 && v < 9223372036854775807.0) {
    const double i = trunc(v);
    if (__dsub_rn(v, i) == 0.0) {
      val = i;
      mult = 0;
      is_float = false;
      return;
    }
  }
  const double sign = (v < 0.0) ? -1.0 : 1.0;
  for (int m = cur; m <= kMaxMult; m++) {
    const double x = __dmul_rn(__dmul_rn(v, mult_pow10(m)), sign);
    if (x >= 1e13) break;
    const double i = trunc(x);
    const double r = __dsub_rn(x, i);
    if (r == 0.0) {
      val = __dmul_rn(sign, i);
      mult = m;
      is_float = false;
      return;
    } else if (r < 0.1) {
      const double below = __longlong_as_double(__double_as_longlong(x) - 1);  // Nextafter(x, 0)
      if (below <= i) {
        val = __dmul_rn(sign, i);
        mult = m;
        is_float = false;
        return;
      }
    } else if (r > 0.9) {
      const double next = __dadd_rn(i, 1.0);
      const double above = __longlong_as_double(__double_as_longlong(x) + 1);  // Nextafter(x, next)
      if (above >= next) {
        val = __dmul_rn(sign, next);
        mult = m;
        is_float = false;
        return;
      }
    }
  }
  val = v;
  mult = 0;
  is_float = true;
}

// IntSigBitsTracker.TrackNewSig, int_sig_bits_tracker.go:68-91
__device__ __forceinline__ int track_new_sig(EncLane &s, int ns) {
  int new_sig = s.num_sig;
  if (ns > s.num_sig) {
    new_sig = ns;
  } else if (s.num_sig - ns >= 3) {
    if (s.n_lower_sig == 0)
      s.hi_lower_sig = ns;
    else if (ns > s.hi_lower_sig)
      s.hi_lower_sig = ns;
    s.n_lower_sig++;
    if (s.n_lower_sig >= 5) {
      new_sig = s.hi_lower_sig;
      s.n_lower_sig = 0;
    }
  } else {
    s.n_lower_sig = 0;
  }
  return new_sig;
}

// writeIntSigMult (encoder.go:235-250) + WriteIntSig (int_sig_bits_tracker.go:48-62):
// appends to the header accumulator
__device__ __forceinline__ void sig_mult_hdr(EncLane &s, int sig, int mult, bool float_changed,
                                             uint32_t &hdr, int &hb) {
  if (s.num_sig != sig) {
    if (sig == 0) {
      hdr = (hdr << 2) | 2u;  // '1' '0'
      hb += 2;
    } else {
      hdr = (hdr << 8) | (3u << 6) | (uint32_t)(sig - 1);  // '1' '1' + 6 bits
      hb += 8;
    }
  } else {
    hdr <<= 1;  // '0'
    hb += 1;
  }
  s.num_sig = sig;
  if (mult > s.max_mult) {
    hdr = (hdr << 4) | 8u | (uint32_t)mult;
    hb += 4;
    s.max_mult = mult;
  } else if (s.max_mult == mult && float_changed) {
    hdr = (hdr << 4) | 8u | (uint32_t)s.max_mult;
    hb += 4;
  } else {
    hdr <<= 1;
    hb += 1;
  }
}

// XOR code (float_encoder_iterator.go:75-103) appended after `hb` prefix bits
__device__ __forceinline__ void xor_code(EncLane &s, uint64_t fb, uint32_t &hdr, int &hb,
                                         uint64_t &payload, int &plen) {
  const uint64_t x = s.prev_bits ^ fb;
  int cl, ct;
  lz_tz(x, cl, ct);  // (64, 0) for x == 0
  if (x == 0) {
    hdr <<= 1;
    hb += 1;
    plen = 0;
  } else if (cl >= s.plz && ct >= s.ptz) {
    hdr = (hdr << 2) | 2u;
    hb += 2;
    payload = x >> s.ptz;
    plen = 64 - s.plz - s.ptz;
  } else {
    const int nm = 64 - cl - ct;
    hdr = (hdr << 14) | (3u << 12) | ((uint32_t)cl << 6) | (uint32_t)(nm - 1);
    hb += 14;
    payload = x >> ct;
    plen = nm;
  }
  s.plz = cl;  // PrevXOR := x
  s.ptz = ct;
  s.prev_bits = fb;
}

// value grammar: encoder.go:112-231.  Produces header bits appended to
// (hdr, hb) and the payload.
template <bool INT_OPT>
__device__ __forceinline__ void encode_value(EncLane &s, double v, uint32_t &hdr, int &hb,
                                             uint64_t &payload, int &plen) {
  const bool first = (s.n_enc == 0);
  const uint64_t vbits = (uint64_t)__double_as_longlong(v);
  payload = 0;
  plen = 0;
  if (!INT_OPT) {
    if (first) {  // writeFullFloat
      s.prev_bits = vbits;
      lz_tz(vbits, s.plz, s.ptz);  // PrevXOR := bits
      payload = vbits;
      plen = 64;
    } else {
      xor_code(s, vbits, hdr, hb, payload, plen);
    }
    return;
  }
  double val = v;
  int mult = 0;
  bool isf = true;
  if (maybe_int(v)) {
    // out-parameters of the (noinline) exact classifier live on the stack: keep
    // them out of the common path
    double v2;
    int m2;
    bool f2;
    convert_to_int_float(v, first ? 0 : s.max_mult, v2, m2, f2);
    val = v2;
    mult = m2;
    isf = f2;
  }
  if (first) {  // writeFirstValue :112-146
    if (isf) {
      hdr = (hdr << 1) | 1u;
      hb += 1;
      s.prev_bits = vbits;
      lz_tz(vbits, s.plz, s.ptz);  // PrevXOR := bits
      payload = vbits;
      plen = 64;
      s.is_float = true;
      s.max_mult = mult;
    } else {
      hdr <<= 1;  // opcodeIntMode
      hb += 1;
      s.int_val = val;
      uint32_t neg_diff = 1;
      double a = val;
      if (val < 0.0) {
        neg_diff = 0;
        a = -val;
      }
      const uint64_t vb = go_f64_to_u64_via_i64(a);
      const int ns = num_sig(vb);
      sig_mult_hdr(s, ns, mult, false, hdr, hb);
      hdr = (hdr << 1) | neg_diff;
      hb += 1;
      payload = vb;
      plen = s.num_sig;
    }
    return;
  }
  // writeNextValue :148-172
  double diff = 0.0;
  if (!isf) diff = __dsub_rn(s.int_val, val);
  if (isf || diff >= 9223372036854775807.0 || diff <= -9223372036854775808.0) {
    const uint64_t fb = (uint64_t)__double_as_longlong(val);  // writeFloatVal :176-198
    if (!s.is_float) {
      hdr = (hdr << 3) | 1u;  // update, no-repeat, float-mode
      hb += 3;
      s.prev_bits = fb;
      lz_tz(fb, s.plz, s.ptz);  // PrevXOR := bits
      payload = fb;
      plen = 64;
      s.is_float = true;
      s.max_mult = mult;
    } else if (fb == s.prev_bits) {
      hdr = (hdr << 2) | 1u;  // update, repeat
      hb += 2;
    } else {
      hdr = (hdr << 1) | 1u;  // no-update
      hb += 1;
      xor_code(s, fb, hdr, hb, payload, plen);
    }
    return;
  }
  // writeIntVal :201-231
  if (diff == 0.0 && !s.is_float && mult == s.max_mult) {
    hdr = (hdr << 2) | 1u;
    hb += 2;
    return;
  }
  uint32_t neg = 0;
  if (diff < 0.0) {
    neg = 1;
    diff = -diff;
  }
  const uint64_t db = go_f64_to_u64_via_i64(diff);
  const int ns = num_sig(db);
  const int new_sig = track_new_sig(s, ns);
  const bool float_changed = s.is_float;
  if (mult > s.max_mult || s.num_sig != new_sig || float_changed) {
    hdr <<= 3;  // update, no-repeat, int-mode
    hb += 3;
    sig_mult_hdr(s, new_sig, mult, float_changed, hdr, hb);
    hdr = (hdr << 1) | neg;
    hb += 1;
    payload = db;
    plen = s.num_sig;
    s.is_float = false;
  } else {
    hdr = (hdr << 2) | 2u | neg;  // no-update + sign
    hb += 2;
    payload = db;
    plen = s.num_sig;
  }
  s.int_val = val;
}

// timestamp grammar: timestamp_encoder.go:104-246 (annotation handled by caller).
// Small codes are returned in (hdr, hb); 32/64-bit fields are written directly.
__device__ __forceinline__ void encode_time(EncLane &s, uint32_t *tile, int lane, int64_t t, int u,
                                            uint32_t &hdr, int &hb) {
  hdr = 0;
  hb = 0;
  const int64_t delta = (int64_t)((uint64_t)t - (uint64_t)s.prev_time);
  const int64_t dd = (int64_t)((uint64_t)delta - (uint64_t)s.prev_delta);
  s.prev_time = t;
  if (u != s.unit && unit_is_valid(u)) {  // maybeWriteTimeUnitChange :141-162
    put32(s, tile, lane, (kMarkerOpcode << 2) | (uint32_t)kMarkerTimeUnit, kMarkerBits);
    put32(s, tile, lane, (uint32_t)u, 8);
    s.unit = u;
    put64(s, tile, lane, (uint64_t)dd, 64);  // writeDeltaOfDeltaTimeUnitChanged :197-203
    s.prev_delta = 0;
    return;
  }
  if (!unit_is_valid(u)) {
    s.err = M3TSZ_ERR_UNRECOGNIZED_UNIT;
    s.prev_delta = delta;
    return;
  }
  s.prev_delta = delta;
  const int kind = scheme_kind_for_unit(u);
  if (dd == 0) {
    hb = (kind == kSchemeZero) ? 0 : 1;
    return;
  }
  const int64_t dod = dd / unit_nanos(u);  // ToNormalizedDuration, x/time/time.go:55-57
  if (u <= 2 && dod != (int64_t)(int32_t)dod) {
    s.err = M3TSZ_ERR_DOD_OVERFLOW;
    return;
  }
  if (kind == kSchemeZero) return;
  if (dod == 0) {
    hb = 1;
  } else if (dod >= -64 && dod <= 63) {
    hdr = (2u << 7) | ((uint32_t)dod & 0x7fu);
    hb = 9;
  } else if (dod >= -256 && dod <= 255) {
    hdr = (6u << 9) | ((uint32_t)dod & 0x1ffu);
    hb = 12;
  } else if (dod >= -2048 && dod <= 2047) {
    hdr = (14u << 12) | ((uint32_t)dod & 0xfffu);
    hb = 16;
  } else {
    put32(s, tile, lane, 0xfu, 4);
    put64(s, tile, lane, (uint64_t)dod, kind == kScheme32 ? 32 : 64);
  }
}

// ---- second-tier candidate (see the kernel's datapoint step) ---------------
struct Tier2 {
  uint32_t thdr;  // delta-of-delta code
  int thb;
  int kind;       // 0 float XOR, 1 float repeat, 2 int diff, 3 int repeat
  uint32_t neg;
  uint64_t db;    // |diff| as integer (kind 2)
  double val;     // converted value (kind 2)
  int n_lower, hi_lower;  // tracker state after this datapoint (kind 2)
  int new_sig;            // significant bits in force after this datapoint (kind 2)
};

// ToNormalizedDuration (x/time/time.go:55-57) with the unit's constant divisor
__device__ __forceinline__ int64_t div_unit(int64_t dd, int u) {
  switch (u) {
    case 1: return dd / 1000000000LL;
    case 2: return dd / 1000000LL;
    case 3: return dd / 1000LL;
    default: return dd;
  }
}

// True when this datapoint can be coded without touching the mode / header
// state: same s/ms/us/ns unit, not the first datapoint, delta-of-delta within the
// 12-bit bucket, and the value is (a) float mode, certainly not int-like, or
// (b) int mode, int-like at the current multiplier, diff in int64 range (a
// significant-bits update is fine, a multiplier or mode change is not).
// Nothing in `s` is modified.
template <bool INT_OPT>
__device__ __forceinline__ bool tier2_candidate(const EncLane &s, bool base, int64_t delta, uint64_t fb,
                                                double v, Tier2 &c) {
  bool ok = base && (s.unit >= 1 && s.unit <= 4) && s.n_enc > 0;
  const int64_t dd = (int64_t)((uint64_t)delta - (uint64_t)s.prev_delta);
  const int64_t dod = div_unit(dd, s.unit);
  ok = ok && (dod >= -2048 && dod <= 2047);
  const uint32_t ud = (uint32_t)dod;
  if (dod == 0) {
    c.thdr = 0;
    c.thb = 1;
  } else if (dod >= -64 && dod <= 63) {
    c.thdr = (2u << 7) | (ud & 0x7fu);
    c.thb = 9;
  } else if (dod >= -256 && dod <= 255) {
    c.thdr = (6u << 9) | (ud & 0x1ffu);
    c.thb = 12;
  } else {
    c.thdr = (14u << 12) | (ud & 0xfffu);
    c.thb = 16;
  }
  c.kind = 0;
  c.neg = 0;
  c.db = 0;
  c.val = 0.0;
  c.n_lower = s.n_lower_sig;
  c.hi_lower = s.hi_lower_sig;
  c.new_sig = s.num_sig;
  if (!INT_OPT) return ok;
  if (s.is_float) {
    c.kind = (fb == s.prev_bits) ? 1 : 0;
    return ok && !maybe_int(v);
  }
  // convertToIntFloat at m = max_mult only (m3tsz.go:78-119); anything that would
  // move on to a larger multiplier is left to the general path
  const double sign = (v < 0.0) ? -1.0 : 1.0;
  const double x = __dmul_rn(__dmul_rn(v, mult_pow10(s.max_mult)), sign);
  const double i = trunc(x);
  const double r = __dsub_rn(x, i);
  const bool exact = (r == 0.0);
  const bool lo = (r < 0.1) && (__longlong_as_double(__double_as_longlong(x) - 1) <= i);
  const double next = __dadd_rn(i, 1.0);
  const bool hi = (r > 0.9) && (__longlong_as_double(__double_as_longlong(x) + 1) >= next);
  ok = ok && (x < 1e13) && (exact || lo || hi);
  c.val = __dmul_rn(sign, (hi && !exact) ? next : i);
  double diff = __dsub_rn(s.int_val, c.val);
  ok = ok && (diff < 9223372036854775807.0) && (diff > -9223372036854775808.0);
  if (diff == 0.0) {
    c.kind = 3;
    return ok;
  }
  c.kind = 2;
  if (diff < 0.0) {
    c.neg = 1;
    diff = -diff;
  }
  c.db = ok ? (uint64_t)__double2ll_rz(diff) : 1ull;
  const int ns = num_sig(c.db);
  // IntSigBitsTracker.TrackNewSig without committing (int_sig_bits_tracker.go:68-91)
  c.new_sig = s.num_sig;
  if (ns > s.num_sig) {
    c.new_sig = ns;
  } else if (s.num_sig - ns >= 3) {
    c.hi_lower = (c.n_lower == 0 || ns > c.hi_lower) ? ns : c.hi_lower;
    c.n_lower++;
    if (c.n_lower >= 5) {
      c.new_sig = c.hi_lower;
      c.n_lower = 0;
    }
  } else {
    c.n_lower = 0;
  }
  return ok;
}

template <bool INT_OPT>
__device__ __forceinline__ void tier2_commit(EncLane &s, uint32_t *tile, int lane, uint64_t fb, const Tier2 &c) {
  if (!INT_OPT || c.kind == 0) {
    const uint32_t pre = INT_OPT ? ((c.thdr << 1) | 1u) : c.thdr;  // + '1' no-update
    const int pb = c.thb + (INT_OPT ? 1 : 0);
    const uint64_t x = s.prev_bits ^ fb;
    int cl, ct;
    lz_tz(x, cl, ct);
    const bool zero = (x == 0);
    const bool cont = !zero && cl >= s.plz && ct >= s.ptz;
    const int nm = 64 - cl - ct;
    const uint64_t P = shl64(x, cont ? s.plz : cl);
    const int plen = zero ? 0 : (cont ? 64 - s.plz - s.ptz : nm);
    const uint32_t hdr = zero ? (pre << 1)
                              : (cont ? ((pre << 2) | 2u)
                                      : ((pre << 14) | (3u << 12) | ((uint32_t)cl << 6) | (uint32_t)(nm - 1)));
    const int hb = pb + (zero ? 1 : (cont ? 2 : 14));
    s.plz = cl;
    s.ptz = ct;
    s.prev_bits = fb;
    emit_code_p(s, tile, lane, hdr, hb, P, plen);
  } else if (c.kind == 2) {  // '1' no-update + sign + num_sig bits (encoder.go:224-230)
    s.n_lower_sig = c.n_lower;
    s.hi_lower_sig = c.hi_lower;
    s.int_val = c.val;
    uint32_t hdr = (c.thdr << 1) | 1u;  // '1' no-update
    int hb = c.thb + 1;
    if (c.new_sig != s.num_sig) {
      // '0' update '0' no-repeat '0' int mode, new significant bits, '0' same multiplier
      // (writeIntSigMult, encoder.go:235-250)
      if (c.new_sig == 0) {
        hdr = (c.thdr << 6) | (2u << 1);
        hb = c.thb + 6;
      } else {
        hdr = (c.thdr << 12) | (3u << 7) | ((uint32_t)(c.new_sig - 1) << 1);
        hb = c.thb + 12;
      }
      s.num_sig = c.new_sig;
    }
    const uint64_t P = s.num_sig ? (c.db << (64 - s.num_sig)) : 0ull;
    emit_code_p(s, tile, lane, (hdr << 1) | c.neg, hb + 1, P, s.num_sig);
  } else {  // '0' update + '1' repeat (float: encoder.go:186-189, int: :202-206)
    emit_code_p(s, tile, lane, (c.thdr << 2) | 1u, c.thb + 2, 0ull, 0);
  }
}

__device__ __forceinline__ uint32_t enc_smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void enc_cp_async8(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
// ---- TMA 1-D bulk copies + mbarrier (M3_ENC_BULK_PM) ----
__device__ __forceinline__ void enc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void enc_mbar_init_fence() {
  // make the initialised barriers visible to the async proxy (the bulk copies' complete_tx)
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void enc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completion counted
// in bytes on the mbarrier
__device__ __forceinline__ void enc_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// one box of a 2-D tensor (tensor map in param space) -> shared memory; c0 = innermost coordinate
__device__ __forceinline__ void enc_tma_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void enc_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "ENC_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra ENC_MBAR_DONE;\n\t"
      "bra ENC_MBAR_WAIT;\n"
      "ENC_MBAR_DONE:\n\t}"
      ::"r"(bar), "r"(parity)
      : "memory");
}

// PACKED = false: one warp per 32-series batch, every series writes its own slot
//   p.out + s * p.out_stride (fixed stride).
// PACKED = true : persistent warps fetch batches from a counter, encode into a per-warp
//   scratch slot set (reused batch after batch), then allocate exact space in the packed
//   buffer with ONE atomicAdd per batch (sum of the 32 aligned lengths) and copy their
//   streams there with coalesced 16-byte loads/stores -- the separate compaction pass
//   (scan + gather over all bytes) and the [S][stride] slot buffer disappear; the copies
//   of finished warps overlap the bit packing of the others.  Streams land in completion
//   order: p.packed_off[s] / p.out_len[s] are the index entry (Offset, Size) of stream s
//   (persist/schema/types.go:70-78).
// IN = 0: series-major inputs ts/val[S][stride], staged through transposed shared-memory tiles;
// IN = 1: point-major inputs ts/val[stride][S]: the same tiles, filled lane-locally (a tile row is 32
//         consecutive elements of the array: coalesced sources, no shuffles).  Direct LDG.64 with two rows
//         in flight was measured first: 13.0 ms vs 10.3 -- the loads' latency is exposed
//         (profiles/r02_encode_pm_1Mx1440.ncu_summary.txt: long_scoreboard 4.3);
// IN = 2: the Gauge aggregates of the fused decode+downsample kernel (window-major): datapoint i of a
//         series = (tile_start + (i+1)*tile_step, Gauge.ValueOf(agg)) unless window i is empty
//         (count == 0: skipped) -- the tile aggregation's re-encode without a gather pass.
template <bool INT_OPT, bool PACKED, int IN>
__global__ void __launch_bounds__(ENC_WARPS * 32, (IN == 1) ? M3_ENC_MIN_BLOCKS_PM : M3_ENC_MIN_BLOCKS)
    encode_kernel(const EncodeParams p
#if M3_ENC_BULK_PM == 2
                  ,
                  const __grid_constant__ EncTensorMaps tmaps
#endif
    ) {
  extern __shared__ __align__(128) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // the direct input stages (IN != 0) need no input tiles: only the output tile lives in shared memory
  constexpr bool STAGED = (IN == 0 || IN == 1);  // inputs go through the shared-memory tiles
  constexpr int IN_T = enc_in_t<IN>();  // datapoints per input tile
  constexpr int IN_STRIDE = enc_in_stride<IN>();  // cells per tile row
  constexpr int IN_TILE_DW = IN_T * IN_STRIDE;  // one array (ts or val), one buffer
  constexpr int BULK = enc_bulk<IN>();  // 0 cp.async fills, 1 1-D bulk copies per row, 2 one 2-D tensor copy per array
  constexpr size_t WARP_SMEM = enc_warp_smem<IN>();
  uint8_t *wbase = reinterpret_cast<uint8_t *>(smem) + warp * WARP_SMEM;
  // in_tiles: [buffer][array (0 ts, 1 val)][row][lane]
  uint64_t *in_tiles = reinterpret_cast<uint64_t *>(wbase);
  uint32_t *out_tile = STAGED ? reinterpret_cast<uint32_t *>(in_tiles + 4 * IN_TILE_DW)
                              : reinterpret_cast<uint32_t *>(wbase);

  // BULK: two mbarriers per warp (one per tile buffer) behind the output tile; `bulk_seq` counts the tiles this
  // warp has requested through them (buffer = seq & 1, phase parity = (seq >> 1) & 1), across batches
  const uint32_t mbar0 = BULK ? enc_smem_addr(out_tile + ENC_OUT_TILE_WORDS) : 0u;
  uint32_t bulk_seq = 0;
#if M3_ENC_BULK_PM == 2
  const bool bulk_args = BULK && tmaps.ok != 0;
#else
  const bool bulk_args = BULK && ((((uintptr_t)p.ts | (uintptr_t)p.val) & 15u) == 0) && ((p.n_series & 1ull) == 0);
#endif
  if (BULK) {
    if (lane == 0) {
      enc_mbar_init(mbar0, 1);
      enc_mbar_init(mbar0 + 8, 1);
      enc_mbar_init_fence();
    }
    __syncwarp();
  }

  const uint64_t n_batches = (p.n_series + 31) >> 5;
  const uint64_t warp_slot = (uint64_t)blockIdx.x * ENC_WARPS + warp;  // PACKED: scratch slot set
  if (PACKED && p.stagger_ns) {
    // Batches of equal length keep the persistent warps in lockstep: everyone packs bits (ALU
    // bound, DRAM half idle), then everyone copies (DRAM bound, ALU idle).  Spreading the warps'
    // phases over one batch period lets the copies of some overlap the bit packing of the others.
    const uint32_t h = (uint32_t)warp_slot * 2654435761u;  // golden-ratio hash: phases mix within an SM
    const uint64_t delay = ((uint64_t)(h >> 8) * p.stagger_ns) >> 24;
    uint64_t t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 >= delay) break;
      const uint64_t rem = delay - (t1 - t0);
      __nanosleep((unsigned)(rem > 20000 ? 20000 : rem));
    }
  }
  for (uint64_t batch_iter = 0;; batch_iter++) {
  uint64_t batch;
  if (PACKED) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(p.batch_counter, 1ull);
    batch = __shfl_sync(FULL_MASK, b, 0);
  } else {
    if (batch_iter) break;
    batch = warp_slot;
  }
  if (batch >= n_batches) break;
  const uint64_t warp_s0 = batch * 32ull;
  const uint64_t sidx = warp_s0 + lane;
  const bool valid = sidx < p.n_series;

  EncLane s;
  s.carry = 0;
  s.sh = 0;
  s.k = 0;
  s.words_out = 0;
  s.prev_time = 0;
  s.prev_delta = 0;
  s.prev_bits = 0;
  s.plz = 64;
  s.ptz = 0;
  s.int_val = 0.0;
  s.unit = 0;
  s.num_sig = 0;
  s.hi_lower_sig = 0;
  s.n_lower_sig = 0;
  s.max_mult = 0;
  s.err = 0;
  s.n_enc = 0;
  s.is_float = false;

  uint32_t n_pts = 0;
  uint64_t ann_i = 0, ann_end = 0;
  uint64_t last_ann_off = 0;
  uint32_t last_ann_len = 0;  // 0 = no annotation written yet
  const uint32_t slot_words = (uint32_t)(p.out_stride >> 2);
  if (valid) {
    n_pts = (IN != 2 && p.n_points) ? p.n_points[sidx] : (uint32_t)p.points_stride;
    if (IN == 2 && p.tile_src_status && p.tile_src_status[sidx] != 0) n_pts = 0;  // failed source stream: no tiles
    if (n_pts > p.points_stride) {
      s.err = M3TSZ_ERR_INVALID_ARG;
      n_pts = 0;
    }
    const int64_t start = p.start[sidx];
    s.prev_time = start;
    s.unit = initial_time_unit(start, p.default_unit);  // encoder.reset :268-272
    if (p.ann_series_off) {
      ann_i = p.ann_series_off[sidx];
      ann_end = p.ann_series_off[sidx + 1];
    }
    if (n_pts > 0 && s.err == 0) {
      if (slot_words < (uint32_t)ENC_GUARD + 4u) {
        s.err = M3TSZ_ERR_CAPACITY;
      } else {
        put64(s, out_tile, lane, (uint64_t)start, 64);  // WriteFirstTime :89-102
      }
    }
  }
  const uint32_t max_pts = __reduce_max_sync(FULL_MASK, n_pts);

  // Lane-local flush: every lane streams the complete 16-byte groups of ITS OWN
  // column to its slot (big-endian byte order) and keeps the <= 3 leftover words.
  uint8_t *my_out = PACKED ? p.out + (warp_slot * 32ull + (uint64_t)lane) * p.out_stride
                           : p.out + sidx * p.out_stride;
  auto flush_out = [&]() {
    __syncwarp();
    const uint32_t kmax = __reduce_max_sync(FULL_MASK, s.k);
    const uint32_t full = s.k >> 2;
    uint4 *dst = reinterpret_cast<uint4 *>(my_out + (uint64_t)s.words_out * 4ull);
    for (uint32_t q = 0; q < (kmax >> 2); q++) {
      if (q < full) {
        const uint32_t *tp = out_tile + (4 * q) * ENC_STRIDE + lane;
        uint4 v;
        v.x = __byte_perm(tp[0], 0, 0x0123);
        v.y = __byte_perm(tp[ENC_STRIDE], 0, 0x0123);
        v.z = __byte_perm(tp[2 * ENC_STRIDE], 0, 0x0123);
        v.w = __byte_perm(tp[3 * ENC_STRIDE], 0, 0x0123);
        dst[q] = v;
      }
    }
    const uint32_t rem = s.k & 3u;
    uint32_t r0 = 0, r1 = 0, r2 = 0;
    {
      const uint32_t *tp = out_tile + (4 * full) * ENC_STRIDE + lane;
      if (rem > 0) r0 = tp[0];
      if (rem > 1) r1 = tp[ENC_STRIDE];
      if (rem > 2) r2 = tp[2 * ENC_STRIDE];
    }
    __syncwarp();
    if (rem > 0) out_tile[lane] = r0;
    if (rem > 1) out_tile[ENC_STRIDE + lane] = r1;
    if (rem > 2) out_tile[2 * ENC_STRIDE + lane] = r2;
    s.words_out += 4 * full;
    s.k = rem;
    __syncwarp();
  };

  // asynchronous staging of input tile `tile` (rows tile*IN_T ..) into buffer tile&1:
  // lanes 0-7 / 8-15: ts / value rows of series 2i, lanes 16-23 / 24-31: of series 2i+1
  const int st_r = lane & (IN_T - 1);
  const int st_arr = (lane >> 3) & 1;
  const int st_jo = lane >> 4;
  const bool uniform_n = __all_sync(FULL_MASK, !valid || n_pts == max_pts) && __all_sync(FULL_MASK, valid);
  // this batch's tiles come by bulk copy: a full warp of series of one length (rows of 32 x 8 bytes, all in range)
  // (tensor copies zero-fill what is out of range, so they serve every batch)
  const bool bulk = BULK && bulk_args && (BULK == 2 || uniform_n);
  const uint32_t tile_base = bulk ? bulk_seq : 0u;  // buffer of tile t: (tile_base + t) & 1
  auto stage = [&](uint32_t tile) {
    const uint32_t row0 = tile * IN_T;
    if (row0 >= max_pts) return;
#if M3_ENC_BULK_PM == 2
    if (BULK == 2 && bulk) {
      if (lane == 0) {
        const uint32_t g = tile_base + tile;
        const uint32_t bar = mbar0 + (g & 1u) * 8u;
        enc_mbar_expect_tx(bar, 2u * (uint32_t)IN_TILE_DW * 8u);  // the whole box counts, in range or not
        const uint32_t d0 = enc_smem_addr(in_tiles + ((g & 1u) * 2u) * IN_TILE_DW);
        enc_tma_2d(d0, &tmaps.ts, (int)warp_s0, (int)row0, bar);
        enc_tma_2d(d0 + (uint32_t)IN_TILE_DW * 8u, &tmaps.val, (int)warp_s0, (int)row0, bar);
      }
      return;
    }
#endif
    if (BULK == 1 && bulk) {
      if (lane == 0) {
        const uint32_t g = tile_base + tile;
        const uint32_t bar = mbar0 + (g & 1u) * 8u;
        const uint32_t nrow = (max_pts - row0 < (uint32_t)IN_T) ? max_pts - row0 : (uint32_t)IN_T;
        enc_mbar_expect_tx(bar, nrow * 512u);
        const uint64_t *st = reinterpret_cast<const uint64_t *>(p.ts) + (uint64_t)row0 * p.n_series + warp_s0;
        const uint64_t *sv = reinterpret_cast<const uint64_t *>(p.val) + (uint64_t)row0 * p.n_series + warp_s0;
        const uint32_t d0 = enc_smem_addr(in_tiles + ((g & 1u) * 2u) * IN_TILE_DW);
#pragma unroll
        for (int r = 0; r < IN_T; r++) {
          if ((uint32_t)r < nrow) {
            enc_bulk_g2s(d0 + (uint32_t)(r * IN_STRIDE) * 8u, st, 256u, bar);
            enc_bulk_g2s(d0 + (uint32_t)(IN_TILE_DW + r * IN_STRIDE) * 8u, sv, 256u, bar);
          }
          st += p.n_series;
          sv += p.n_series;
        }
      }
      return;
    }
    if (IN == 1) {
      // point-major inputs: a row of the tile is 32 consecutive elements of the array -- every lane
      // copies its own column (8 rows x 2 arrays), sources coalesced across the warp, no shuffles
      const uint64_t *st = reinterpret_cast<const uint64_t *>(p.ts) + (uint64_t)row0 * p.n_series + sidx;
      const uint64_t *sv = reinterpret_cast<const uint64_t *>(p.val) + (uint64_t)row0 * p.n_series + sidx;
      const uint32_t d0 = enc_smem_addr(in_tiles + ((tile & 1u) * 2u) * IN_TILE_DW + lane);
      const uint32_t lim = s.err == 0 ? n_pts : 0u;  // (0 for a lane without a series)
#pragma unroll
      for (int r = 0; r < IN_T; r++) {
        if (row0 + (uint32_t)r < lim) {
          enc_cp_async8(d0 + (uint32_t)(r * IN_STRIDE) * 8u, st);
          enc_cp_async8(d0 + (uint32_t)(IN_TILE_DW + r * IN_STRIDE) * 8u, sv);
        }
        st += p.n_series;
        sv += p.n_series;
      }
      return;
    }
    const uint64_t *src = (st_arr ? reinterpret_cast<const uint64_t *>(p.val)
                                  : reinterpret_cast<const uint64_t *>(p.ts)) +
                          (warp_s0 + st_jo) * p.points_stride + row0 + st_r;
    const uint32_t dst0 = enc_smem_addr(in_tiles + ((tile & 1u) * 2u + (uint32_t)st_arr) * IN_TILE_DW +
                                        st_r * IN_STRIDE + st_jo);
    const uint64_t step = 2ull * p.points_stride;
    if (uniform_n && row0 + IN_T <= max_pts) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        enc_cp_async8(dst0 + (uint32_t)i * 16u, src);
        src += step;
      }
    } else {
      for (int i = 0; i < 16; i++) {
        const uint32_t nj = __shfl_sync(FULL_MASK, (valid && s.err == 0) ? n_pts : 0u, 2 * i + st_jo);
        if (row0 + (uint32_t)st_r < nj) enc_cp_async8(dst0 + (uint32_t)i * 16u, src);
        src += step;
      }
    }
  };

  if (STAGED) {
    stage(0);
    if (!(BULK && bulk)) asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  // direct input stages: element `row` of this lane's series
  auto load_row = [&](uint32_t row, int64_t &t_o, uint64_t &v_o, bool &skip_o) {
    skip_o = false;
    t_o = 0;
    v_o = 0;
    if (!valid || row >= n_pts) return;
    const uint64_t o = (uint64_t)row * p.n_series + sidx;
    if (IN == 2) {
      const int64_t c = __ldg(p.tile_count + o);
      skip_o = (c == 0);
      t_o = p.tile_start + (int64_t)(row + 1) * p.tile_step;  // the window's end boundary
      double v = 0.0;
      switch (p.agg_type) {  // Gauge.ValueOf, gauge.go:144-165 (uniform switch)
        case M3TSZ_AGG_LAST: v = __ldg(p.tile_last + o); break;
        case M3TSZ_AGG_MIN: v = __ldg(p.tile_min + o); break;
        case M3TSZ_AGG_MAX: v = __ldg(p.tile_max + o); break;
        case M3TSZ_AGG_MEAN: v = c == 0 ? 0.0 : __ddiv_rn(__ldg(p.tile_sum + o), __ll2double_rn(c)); break;
        case M3TSZ_AGG_COUNT: v = __ll2double_rn(c); break;
        case M3TSZ_AGG_SUM: v = __ldg(p.tile_sum + o); break;
        default: break;
      }
      v_o = (uint64_t)__double_as_longlong(v);
    }
  };
  int64_t t_pf1 = 0;
  uint64_t fb_pf1 = 0;
  bool sk_pf = false, sk_pf1 = false;

  uint32_t iter = 0;
  int64_t t_pf = 0;
  uint64_t fb_pf = 0;
  if (!STAGED) {  // two rows in flight
    load_row(0, t_pf, fb_pf, sk_pf);
    load_row(1, t_pf1, fb_pf1, sk_pf1);
  }
  bool steady = false;  // same s/ms/us/ns unit as the batch's and not the first datapoint
  bool hot_ok = false;  // steady, and float mode when int-optimised: both only change on the slow path
  const uint64_t *in_next = in_tiles + lane;  // this lane's ts cell of the row to fetch next (value: + one tile)
  // Optional (-DM3_ENC_UNROLL_PM=1): the point-major stage walks a tile with a statically unrolled inner loop
  // (the phase of `iter` inside the tile is the unroll index, so the tile tests below are compile-time).
  // Measured slower (9.54 vs 8.74 ms: four copies of the three tiers do not pay for seven instructions).
  constexpr int UNR = (IN == 1 && M3_ENC_UNROLL_PM) ? IN_T : 1;
  bool more = true;
  while (more) {
#pragma unroll
  for (int ph = 0; ph < UNR; ph++) {
    if (iter >= max_pts) {  // warp-uniform (n_pts is 0 for lanes without a series)
      more = false;
      break;
    }
    const bool tile_first = UNR > 1 ? (ph == 0) : ((iter & (IN_T - 1)) == 0);
    const bool tile_last = UNR > 1 ? (ph == IN_T - 1) : ((iter & (IN_T - 1)) == IN_T - 1);
    const bool active = s.err == 0 && iter < n_pts && !(IN == 2 && sk_pf);  // (n_pts is 0 without a series)

    // ---- input pipeline: request tile t+1, wait for tile t ----
    if (STAGED && tile_first) {
      __syncwarp();  // everyone is done reading the buffer about to be overwritten
      stage(iter / IN_T + 1);
      const uint32_t g = tile_base + iter / IN_T;  // (tile_base is 0 unless this batch is bulk-filled)
      if (BULK && bulk) {
        enc_mbar_wait(mbar0 + (g & 1u) * 8u, (g >> 1) & 1u);  // every lane: the tile's bytes have landed
      } else {
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
      }
      __syncwarp();
      // row 0 of the tile that just landed; rows 1..7 are fetched one datapoint ahead below
      const uint64_t *tile = in_tiles + ((g & 1u) * 2u) * IN_TILE_DW + lane;
      t_pf = (int64_t)tile[0];
      fb_pf = tile[IN_TILE_DW];
      in_next = tile + IN_STRIDE;
    }

    // ---- make room in the output tile ----
    const bool tight = active && (s.k > (uint32_t)(ENC_OUT_W - ENC_GUARD));
    if (__any_sync(FULL_MASK, tight)) flush_out();

    // ---- annotations (warp-cooperative; timestamp_encoder.go:166-195) ----
    bool want_ann = false;
    uint64_t a_off = 0;
    uint32_t a_len = 0;
    if (p.ann_series_off) {
      if (active && ann_i < ann_end) {
        const m3tsz_annotation_entry e = p.ann_entries[ann_i];
        if (e.dp_index == iter) {
          ann_i++;
          a_off = e.byte_offset;
          a_len = e.length;
          if (a_len > 0) {
            // rewritten only when it differs from the last annotation written
            // (xxhash64 inequality in the reference == byte inequality here)
            bool same = (a_len == last_ann_len);
            for (uint32_t i = 0; same && i < a_len; i++)
              same = p.ann_bytes[a_off + i] == p.ann_bytes[last_ann_off + i];
            want_ann = !same;
          }
        }
      }
      if (__any_sync(FULL_MASK, want_ann)) {
        uint32_t rem = 0, done_bytes = 0;
        if (want_ann) {
          put32(s, out_tile, lane, (kMarkerOpcode << 2) | (uint32_t)kMarkerAnnotation, kMarkerBits);
          uint64_t ux = ((uint64_t)(a_len - 1)) << 1;  // binary.PutVarint(len-1), len >= 1
          while (ux >= 0x80) {
            put32(s, out_tile, lane, (uint32_t)(ux & 0x7f) | 0x80u, 8);
            ux >>= 7;
          }
          put32(s, out_tile, lane, (uint32_t)ux, 8);
          rem = a_len;
          last_ann_off = a_off;
          last_ann_len = a_len;
        }
        for (;;) {
          if (rem > 0) {
            // capacity: stop this series rather than overrun its slot
            const uint64_t need_words = (uint64_t)s.words_out + s.k + (rem + 3) / 4 + ENC_GUARD + 4;
            if (need_words > slot_words) {
              s.err = M3TSZ_ERR_CAPACITY;
              rem = 0;
            }
          }
          while (rem > 0 && s.k < (uint32_t)(ENC_OUT_W - ENC_GUARD)) {
            put32(s, out_tile, lane, (uint32_t)p.ann_bytes[a_off + done_bytes], 8);
            done_bytes++;
            rem--;
          }
          if (!__any_sync(FULL_MASK, rem > 0)) break;
          flush_out();
        }
      }
    }

    // ---- encode one datapoint ----
    {
      // this datapoint was fetched from shared memory one datapoint ago (the load latency
      // hides behind the previous bit packing); fetch the next row of the tile now
      const int64_t t = t_pf;
      const uint64_t fb = fb_pf;
      if (STAGED) {
        if (!tile_last) {
          t_pf = (int64_t)in_next[0];
          fb_pf = in_next[IN_TILE_DW];
          in_next += IN_STRIDE;
        }
      } else {  // rotate the two-deep prefetch; request row iter + 2
        t_pf = t_pf1;
        fb_pf = fb_pf1;
        sk_pf = sk_pf1;
        load_row(iter + 2, t_pf1, fb_pf1, sk_pf1);
      }
      const double v = __longlong_as_double((long long)fb);
      const bool room = s.words_out + s.k + (uint32_t)(ENC_GUARD + 4) <= slot_words;  // < 2^31 words
      // hot candidate: same (valid s/ms/us/ns) unit, zero delta-of-delta, not the
      // first datapoint, a float-mode XOR code (value is certainly not int-like)
      const int64_t delta = (int64_t)((uint64_t)t - (uint64_t)s.prev_time);
      // (the int-likeness filter runs on the FP64 pipe for every lane: no divergent region around it)
      const bool not_int = !INT_OPT || !maybe_int(v);
      bool hot = (active & room & hot_ok & not_int) && delta == s.prev_delta;
      if (INT_OPT) hot = hot && fb != s.prev_bits;
      Tier2 c2;
      if (__all_sync(FULL_MASK, hot || !active)) {
        // every live lane: '0' (zero DoD) [+ '1' no-update] + XOR code, one merge.  The state update is
        // unconditional where an inactive lane is a FINISHED lane (IN 0 / 1); with IN = 2 a lane may sit
        // out one row (an empty window) and must keep its state.
        const bool upd = (IN != 2) || active;
        if (upd) s.prev_time = t;
        const uint32_t pre = INT_OPT ? 1u : 0u;  // '0' zero DoD [+ '1' no-update]
        const int pb = INT_OPT ? 2 : 1;
        const uint64_t x = s.prev_bits ^ fb;
        int cl, ct;
        lz_tz(x, cl, ct);  // (64, 0) for x == 0
        const bool zero = !INT_OPT && (x == 0);  // int-optimised: a repeat is not a hot datapoint (fb != prev_bits)
        const bool cont = !zero && cl >= s.plz && ct >= s.ptz;
        const int nm = 64 - cl - ct;
        // payload left-aligned: (x >> ptz) << (64 - plen) == x << plz, (x >> ct) << (64 - nm) == x << cl
        const uint64_t P = shl64_clamp(x, (uint32_t)(cont ? s.plz : cl));  // (x == 0: cl = 64 shifts everything out)
        const int plen = zero ? 0 : (cont ? 64 - s.plz - s.ptz : nm);
        const uint32_t hdr = zero ? (pre << 1)
                                  : (cont ? ((pre << 2) | 2u)
                                          : ((pre << 14) | (3u << 12) | ((uint32_t)cl << 6) | (uint32_t)(nm - 1)));
        const int hb = pb + (zero ? 1 : (cont ? 2 : 14));
        if (upd) {
          s.plz = cl;  // PrevXOR := x
          s.ptz = ct;
          s.prev_bits = fb;
        }
        if (active) {
          emit_code_p(s, out_tile, lane, hdr, hb, P, plen);
          s.n_enc++;
        }
      } else if (__all_sync(FULL_MASK, tier2_candidate<INT_OPT>(s, active && room && steady, delta, fb, v, c2) ||
                                           !active)) {
        // second tier: small delta-of-delta + (float XOR | float repeat | int diff | int
        // repeat) without a mode / header update, still one merge per lane
        if (active) {
          s.prev_time = t;
          s.prev_delta = delta;
          tier2_commit<INT_OPT>(s, out_tile, lane, fb, c2);
          s.n_enc++;
        }
      } else if (active && s.err == 0) {
        if (!room) {
          s.err = M3TSZ_ERR_CAPACITY;
        } else {
          const int u = p.units ? (int)p.units[sidx * p.points_stride + iter] : p.unit;
          uint32_t hdr;
          int hb;
          encode_time(s, out_tile, lane, t, u, hdr, hb);
          if (s.err == 0) {
            uint64_t payload;
            int plen;
            encode_value<INT_OPT>(s, v, hdr, hb, payload, plen);
            emit_code(s, out_tile, lane, hdr, hb, payload, plen);
            s.n_enc++;
          }
          // the fast tiers' slow-changing preconditions: the unit only changes here
          steady = !p.units && p.unit == s.unit && (s.unit >= 1 && s.unit <= 4) && s.n_enc > 0;
          hot_ok = steady && (!INT_OPT || s.is_float);
        }
      }
    }
    iter++;
  }  // ph
  }  // while (more)
  if (STAGED) asm volatile("cp.async.wait_all;\n" ::: "memory");
  if (BULK && bulk) bulk_seq += (max_pts + (uint32_t)IN_T - 1u) / (uint32_t)IN_T;  // every requested tile was waited for

  // ---- tail: end-of-stream marker + zero padding (scheme.go:198-211) ----
  uint64_t total_bits = 0;
  if (valid && s.n_enc > 0) {
    put32(s, out_tile, lane, (kMarkerOpcode << 2) | (uint32_t)kMarkerEOS, kMarkerBits);
    total_bits = ((uint64_t)s.words_out + s.k) * 32ull + s.sh;
    if (s.sh > 0) {  // flush the partial word (zero padded)
      out_tile[s.k * ENC_STRIDE + lane] = s.carry;
      s.k++;
      s.carry = 0;
      s.sh = 0;
    }
    // pad to a whole 16-byte group (zeros; bytes past out_len are unspecified)
    while (s.k & 3u) {
      out_tile[s.k * ENC_STRIDE + lane] = 0u;
      s.k++;
    }
  } else if (valid) {
    s.k = 0;  // nothing encoded: empty stream (encoder.go:285-289)
    s.words_out = 0;
  } else {
    s.k = 0;
  }
  flush_out();
  const uint64_t my_len = (total_bits + 7) >> 3;
  if (valid && IN == 2 && p.n_tiles_out) p.n_tiles_out[sidx] = s.n_enc;
  if (valid && p.out_bits) p.out_bits[sidx] = total_bits;  // incl. the end-of-stream marker, before padding
  if (valid && p.last_value) {
    // Encoder.LastEncoded().Value (encoder.go:305-319): the float when the encoder is in float
    // mode, else its intVal -- the SCALED integer in int mode, and 0 without the int optimisation
    // (isFloat is never set there)
    p.last_value[sidx] = (INT_OPT && s.is_float) ? __longlong_as_double((long long)s.prev_bits)
                                                 : (INT_OPT ? s.int_val : 0.0);
  }
  if (!PACKED) {
    if (valid) {
      if (p.out_len) p.out_len[sidx] = my_len;
      if (p.status) p.status[sidx] = s.err;
    }
  } else {
    // ---- allocate exact space and move the batch's streams into the packed buffer ----
    const uint64_t amask = (uint64_t)p.align - 1;
    const uint64_t alen = valid ? ((my_len + amask) & ~amask) : 0ull;
    uint64_t incl = alen;  // inclusive warp scan
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t o = __shfl_up_sync(FULL_MASK, incl, d);
      if (lane >= d) incl += o;
    }
    const uint64_t total = __shfl_sync(FULL_MASK, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(p.packed_cursor, (unsigned long long)total);
    base = __shfl_sync(FULL_MASK, base, 0);
    const bool fits = base + total <= p.packed_capacity;
    const uint64_t my_off = base + incl - alen;
    if (valid) {
      p.packed_off[sidx] = my_off;
      if (p.out_len) p.out_len[sidx] = fits ? my_len : 0ull;
      if (p.status) p.status[sidx] = (!fits && s.err == 0) ? M3TSZ_ERR_CAPACITY : s.err;
    }
    __syncwarp();  // the slots were written lane-locally; the copy below reads other lanes' slots
    if (fits && total) {
      const uint8_t *slot0 = p.out + warp_slot * 32ull * p.out_stride;
      if (p.align >= 16) {
        // series after series, 32 lanes x 16 bytes, sixteen loads in flight per lane; the flush
        // above zero-padded every stream to a whole 16-byte group inside its slot
        for (int j = 0; j < 32; j++) {
          const uint64_t lj = __shfl_sync(FULL_MASK, my_len, j);
          const uint64_t oj = __shfl_sync(FULL_MASK, my_off, j);
          if (lj == 0) continue;
          const uint4 *src = reinterpret_cast<const uint4 *>(slot0 + (uint64_t)j * p.out_stride);
          uint4 *dst = reinterpret_cast<uint4 *>(p.packed + oj);
          const uint32_t nv = (uint32_t)((lj + 15) >> 4);
          uint32_t i = lane;
          for (; i + 480 < nv; i += 512) {  // 16 loads in flight per lane: 8 KB per warp
            uint4 r[16];
#pragma unroll
            for (int u = 0; u < 16; u++) r[u] = __ldcg(src + i + 32 * u);
#pragma unroll
            for (int u = 0; u < 16; u++) __stcs(dst + i + 32 * u, r[u]);
          }
          for (; i + 96 < nv; i += 128) {
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; u++) r[u] = __ldcg(src + i + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; u++) __stcs(dst + i + 32 * u, r[u]);
          }
          for (; i < nv; i += 32) __stcs(dst + i, __ldcg(src + i));
        }
      } else {
        for (int j = 0; j < 32; j++) {
          const uint64_t lj = __shfl_sync(FULL_MASK, my_len, j);
          const uint64_t oj = __shfl_sync(FULL_MASK, my_off, j);
          const uint8_t *src = slot0 + (uint64_t)j * p.out_stride;
          uint8_t *dst = p.packed + oj;
          for (uint64_t i = lane; i < lj; i += 32) dst[i] = __ldcg(src + i);
        }
      }
    }
    __syncwarp();  // the next batch overwrites the slots
  }
  }  // batch loop
}

template <int IN>
constexpr size_t enc_block_smem() {
  return enc_warp_smem<IN>() * ENC_WARPS;
}

// tuning knob: M3TSZ_ENC_CARVEOUT_KB=K asks for a K KB shared-memory carveout (the rest of 228 KB is L1)
static int enc_carveout_kb() {
  static const int kb = [] {
    const char *c = getenv("M3TSZ_ENC_CARVEOUT_KB");
    return c ? atoi(c) : 0;
  }();
  return kb;
}

#if M3_ENC_BULK_PM == 2
// Tensor maps of the point-major input arrays: ts / val as [points_stride][n_series] tensors of 8-byte elements,
// box = enc_in_t<1>() rows x 32 series, no swizzle, zero fill outside.  ok = 0 when the arrays cannot be described
// (then the kernel keeps its cp.async fills).
static void enc_build_tensor_maps(const EncodeParams &p, EncTensorMaps &tm) {
  memset(&tm, 0, sizeof(tm));
  typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
  static encode_tiled_fn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<encode_tiled_fn>(f);
  }();
  if (!fn || getenv("M3TSZ_ENC_NO_TMA")) return;
  if (p.in_mode != 1 || !p.ts || !p.val) return;
  if ((((uintptr_t)p.ts | (uintptr_t)p.val) & 15u) != 0) return;
  if ((p.n_series & 1ull) != 0 || p.n_series >= (1ull << 31) || p.points_stride == 0 ||
      p.points_stride >= (1ull << 31))
    return;
  const cuuint64_t dims[2] = {(cuuint64_t)p.n_series, (cuuint64_t)p.points_stride};
  const cuuint64_t strides[1] = {(cuuint64_t)p.n_series * 8ull};
  const cuuint32_t box[2] = {32u, (cuuint32_t)enc_in_t<1>()};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult a = fn(&tm.ts, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<int64_t *>(p.ts), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const CUresult b = fn(&tm.val, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<double *>(p.val), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  tm.ok = (a == CUDA_SUCCESS && b == CUDA_SUCCESS) ? 1 : 0;
}
#endif

template <bool INT_OPT, bool PACKED, int IN>
static cudaError_t launch_encode_one(const EncodeParams &p, cudaStream_t stream) {
  constexpr size_t smem = enc_block_smem<IN>();
  const uint64_t per_block = (uint64_t)ENC_WARPS * 32ull;
  uint64_t blocks = (p.n_series + per_block - 1) / per_block;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(encode_kernel<INT_OPT, PACKED, IN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (enc_carveout_kb() > 0) {
    e = cudaFuncSetAttribute(encode_kernel<INT_OPT, PACKED, IN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             enc_carveout_kb() * 100 / 228);
    if (e != cudaSuccess) return e;
  }
  if (PACKED) {
    const uint64_t resident = encode_packed_resident_blocks();
    if (resident == 0) return cudaErrorInvalidValue;
    if (blocks > resident) blocks = resident;
  }
#if M3_ENC_BULK_PM == 2
  EncTensorMaps tm;
  enc_build_tensor_maps(p, tm);  // (ok = 0 for the other input stages)
  encode_kernel<INT_OPT, PACKED, IN><<<(unsigned)blocks, ENC_WARPS * 32, smem, stream>>>(p, tm);
#else
  encode_kernel<INT_OPT, PACKED, IN><<<(unsigned)blocks, ENC_WARPS * 32, smem, stream>>>(p);
#endif
  return cudaGetLastError();
}

// Persistent grid of the packed mode: resident blocks on the current device (also the number
// of scratch slot sets: resident blocks x ENC_WARPS x 32 slots of out_stride bytes).
uint64_t encode_packed_resident_blocks() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  int best = 0;
  auto probe = [&](auto kernel, size_t sm) {
    int n = 0;
    if (enc_carveout_kb() > 0)
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, enc_carveout_kb() * 100 / 228);
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, ENC_WARPS * 32, sm) == cudaSuccess && n > best)
      best = n;
  };
  probe(encode_kernel<true, true, 0>, enc_block_smem<0>());
  probe(encode_kernel<false, true, 0>, enc_block_smem<0>());
  probe(encode_kernel<true, true, 1>, enc_block_smem<1>());
  probe(encode_kernel<false, true, 1>, enc_block_smem<1>());
  probe(encode_kernel<true, true, 2>, enc_block_smem<2>());
  probe(encode_kernel<false, true, 2>, enc_block_smem<2>());
  return (uint64_t)sms * (uint64_t)best;
}
uint64_t encode_packed_scratch_slots(uint64_t n_series) {
  const uint64_t per_block = (uint64_t)ENC_WARPS * 32ull;
  uint64_t blocks = (n_series + per_block - 1) / per_block;
  const uint64_t resident = encode_packed_resident_blocks();
  if (blocks > resident) blocks = resident;
  return blocks * per_block;
}

// in_mode 0: series-major inputs; 1: point-major inputs; 2: Gauge aggregates
// (packed output only)
cudaError_t launch_encode(const EncodeParams &p, bool int_optimized, cudaStream_t stream) {
  if (p.in_mode == 2) {
    if (!p.packed) return cudaErrorInvalidValue;
    return int_optimized ? launch_encode_one<true, true, 2>(p, stream) : launch_encode_one<false, true, 2>(p, stream);
  }
  if (p.in_mode == 1) {
    if (p.packed)
      return int_optimized ? launch_encode_one<true, true, 1>(p, stream) : launch_encode_one<false, true, 1>(p, stream);
    return int_optimized ? launch_encode_one<true, false, 1>(p, stream)
                         : launch_encode_one<false, false, 1>(p, stream);
  }
  if (p.packed) {
    return int_optimized ? launch_encode_one<true, true, 0>(p, stream) : launch_encode_one<false, true, 0>(p, stream);
  }
  return int_optimized ? launch_encode_one<true, false, 0>(p, stream) : launch_encode_one<false, false, 0>(p, stream);
}

// ---------------------------------------------------------------------------
// compaction: slots -> packed buffer with CSR offsets
// ---------------------------------------------------------------------------
__global__ void compact_lens_kernel(const uint64_t *len, uint64_t n, uint64_t mask, uint64_t *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (len[i] + mask) & ~mask;
  if (i == n) out[i] = 0;
}

__global__ void compact_gather_kernel(const uint8_t *slots, uint64_t slot_stride, const uint64_t *len,
                                      uint64_t n, uint8_t *packed, uint64_t cap,
                                      const uint64_t *offsets, int32_t *overflow) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint64_t o = offsets[w], l = len[w];
  if (o + l > cap) {
    if (lane == 0) atomicExch(overflow, 1);
    return;
  }
  const uint8_t *src = slots + w * slot_stride;
  uint8_t *dst = packed + o;
  if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
    // 16-byte vector copies, four in flight per lane
    const uint64_t nv = l >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    uint64_t i = lane;
    for (; i + 96 < nv; i += 128) {
      const uint4 a = __ldg(s4 + i), b = __ldg(s4 + i + 32), c = __ldg(s4 + i + 64), d = __ldg(s4 + i + 96);
      d4[i] = a;
      d4[i + 32] = b;
      d4[i + 64] = c;
      d4[i + 96] = d;
    }
    for (; i < nv; i += 32) d4[i] = __ldg(s4 + i);
    for (uint64_t j = (nv << 4) + lane; j < l; j += 32) dst[j] = src[j];
  } else if ((((uintptr_t)dst | (uintptr_t)src) & 3u) == 0) {
    const uint64_t nw = l >> 2;
    for (uint64_t i = lane; i < nw; i += 32)
      reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(src)[i];
    for (uint64_t i = (nw << 2) + lane; i < l; i += 32) dst[i] = src[i];
  } else {
    for (uint64_t i = lane; i < l; i += 32) dst[i] = src[i];
  }
}

size_t compact_scan_tmp_bytes(uint64_t n_series) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                (int)(n_series + 1));
  return tmp + (n_series + 1) * sizeof(uint64_t);
}

cudaError_t launch_compact(const uint8_t *slots, uint64_t slot_stride, const uint64_t *len,
                           uint64_t n_series, uint32_t align, uint8_t *packed,
                           uint64_t packed_capacity, uint64_t *offsets, void *scan_tmp,
                           size_t scan_tmp_bytes, int32_t *overflow_flag, cudaStream_t stream) {
  if (n_series + 1 > 0x7fffffffull) return cudaErrorInvalidValue;
  uint64_t *alen = reinterpret_cast<uint64_t *>(scan_tmp);
  void *cub_tmp = reinterpret_cast<uint8_t *>(scan_tmp) + (n_series + 1) * sizeof(uint64_t);
  size_t cub_bytes = scan_tmp_bytes - (n_series + 1) * sizeof(uint64_t);
  const uint64_t mask = (uint64_t)align - 1;
  const unsigned tb = 256;
  compact_lens_kernel<<<(unsigned)((n_series + 1 + tb - 1) / tb), tb, 0, stream>>>(len, n_series, mask, alen);
  cudaError_t e = cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, alen, offsets, (int)(n_series + 1), stream);
  if (e != cudaSuccess) return e;
  if (n_series == 0) return cudaGetLastError();
  const uint64_t threads = n_series * 32ull;
  compact_gather_kernel<<<(unsigned)((threads + tb - 1) / tb), tb, 0, stream>>>(
      slots, slot_stride, len, n_series, packed, packed_capacity, offsets, overflow_flag);
  return cudaGetLastError();
}

}  // namespace m3tsz

// m3tsz_decode.cu -- batch M3TSZ decode for sm_100a (+ fused downsample).
//
// Mapping: one LANE per series, 32 series per warp.  A stream is
// a serial bit-dependency chain (every field width depends on decoded state),
// so the per-series state machine runs on one lane while the WARP cooperates
// on memory: each lane's compressed words are staged into shared memory with
// coalesced 128-byte global loads (transposed [word][lane] tile, stride 33),
// every datapoint is parsed from four shared-memory words with funnel shifts,
// and decoded (ts, value) pairs go through a second transposed tile so global
// stores are 128-byte coalesced per series.
//
// Format: SURVEY.md Appendix A; reference decode path
//   m3tsz/iterator.go:81-219, m3tsz/timestamp_iterator.go:80-326,
//   m3tsz/float_encoder_iterator.go:105-165, istream.go:73-115.
#include "m3tsz_common.cuh"
#include "m3tsz_kernels.h"

namespace m3tsz {

constexpr int DEC_WARPS = 4;     // warps per block
constexpr int DEC_IN_W = 64;     // staged words per lane
constexpr int DEC_STRIDE = 33;   // tile row stride (words / dwords): conflict-free transposes
constexpr int DEC_OUT_T = 16;    // output tile rows (datapoints per flush)
constexpr int DEC_FAST_WORDS = 4;  // the fast path reads 4 consecutive words

constexpr int DEC_IN_TILE_WORDS = DEC_IN_W * DEC_STRIDE;                 // u32
constexpr int DEC_OUT_TILE_DWORDS = DEC_OUT_T * DEC_STRIDE;              // u64
constexpr size_t DEC_WARP_SMEM_PLAIN =
    (size_t)DEC_IN_TILE_WORDS * 4 + 2 * (size_t)DEC_OUT_TILE_DWORDS * 8;
constexpr size_t DEC_WARP_SMEM_DS = (size_t)DEC_IN_TILE_WORDS * 4;
static_assert((DEC_IN_TILE_WORDS * 4) % 8 == 0, "u64 tiles must stay 8-byte aligned");

constexpr uint64_t kGoNaNBits = 0x7FF8000000000001ull;  // math.NaN()

struct DecState {
  uint64_t wbase;  // global word index of this stream's 4-byte-aligned base
  uint32_t pos;    // current bit position relative to wbase
  uint32_t end;    // end-of-stream bit position (exclusive)
  int64_t prev_time, prev_delta;
  uint64_t prev_bits, prev_xor;
  double int_val;
  int64_t unit_ns;
  int sig, mult, unit, scheme;
  int err;
  uint32_t n;
  bool is_float, done;
  uint32_t ann_count, ann_len;
  uint32_t ann_bit;
};

// ---------------------------------------------------------------------------
// generic bit reader over global memory (slow path: first datapoint, markers,
// annotations, unit changes, 32/64-bit delta-of-delta, int-mode headers)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t gpeek64(const uint8_t *base, uint64_t nbytes, uint64_t wbase,
                                            uint32_t pos) {
  uint64_t w = wbase + (pos >> 5);
  uint32_t w0 = load_be32(base, nbytes, w);
  uint32_t w1 = load_be32(base, nbytes, w + 1);
  uint32_t w2 = load_be32(base, nbytes, w + 2);
  uint32_t hi = __funnelshift_l(w1, w0, pos);
  uint32_t lo = __funnelshift_l(w2, w1, pos);
  return ((uint64_t)hi << 32) | lo;
}

#define M3_RD(nbits, var)                                                        \
  uint64_t var;                                                                  \
  {                                                                              \
    const int _n = (nbits);                                                      \
    if (s.pos + (uint32_t)_n > s.end) {                                          \
      s.err = M3TSZ_ERR_EOF;                                                     \
      return false;                                                              \
    }                                                                            \
    var = _n ? (gpeek64(base, nbytes, s.wbase, s.pos) >> (64 - _n)) : 0ull;       \
    s.pos += (uint32_t)_n;                                                       \
  }

// Decodes exactly one datapoint with the complete grammar.  Returns true when a
// datapoint was produced (iterator.Next() == true *or* the value was read
// without error); false on end-of-stream (s.done) or error (s.err).
template <bool INT_OPT>
__device__ __noinline__ bool decode_dp_slow(DecState &s, const uint8_t *base, uint64_t nbytes,
                                            int default_unit, int64_t &out_t, uint64_t &out_v) {
  // ---- ReadTimestamp, timestamp_iterator.go:80-113 ----
  const bool first = (s.prev_time == 0);
  int64_t nt = 0;
  if (first) {  // readFirstTimestamp :136-162
    M3_RD(64, x);
    nt = (int64_t)x;
    if (s.unit == 0) {
      s.unit = initial_time_unit(nt, default_unit);
      s.unit_ns = unit_nanos(s.unit);
    }
    int k = scheme_kind_for_unit(s.unit);
    if (k != kSchemeNone) s.scheme = k;
  }
  bool unit_changed = false;
  int64_t dod = 0;
  // readMarkerOrDeltaOfDelta :233-244 / tryReadMarker :175-231 (recursion unrolled)
  for (;;) {
    if (s.pos + kMarkerBits <= s.end) {
      uint32_t p = (uint32_t)(gpeek64(base, nbytes, s.wbase, s.pos) >> 53);
      if ((p >> 2) == kMarkerOpcode) {
        int m = (int)(p & 3);
        if (m == kMarkerEOS) {
          s.pos += kMarkerBits;
          s.done = true;
          return false;
        }
        if (m == kMarkerAnnotation) {  // readAnnotation :328-356
          s.pos += kMarkerBits;
          uint64_t ux = 0;
          int shift = 0;
          bool fin = false;
          for (int i = 0; i < 10; i++) {  // binary.ReadUvarint
            if (s.pos + 8 > s.end) {
              s.err = (i > 0) ? M3TSZ_ERR_UNEXPECTED_EOF : M3TSZ_ERR_EOF;
              return false;
            }
            uint32_t b = (uint32_t)(gpeek64(base, nbytes, s.wbase, s.pos) >> 56);
            s.pos += 8;
            if (b < 0x80) {
              if (i == 9 && b > 1) {
                s.err = M3TSZ_ERR_VARINT_OVERFLOW;
                return false;
              }
              ux |= shl64((uint64_t)b, shift);
              fin = true;
              break;
            }
            ux |= shl64((uint64_t)(b & 0x7f), shift);
            shift += 7;
          }
          if (!fin) {
            s.err = M3TSZ_ERR_VARINT_OVERFLOW;
            return false;
          }
          int64_t alen = (int64_t)(ux >> 1);
          if (ux & 1) alen = ~alen;
          alen += 1;
          if (alen <= 0) {
            s.err = M3TSZ_ERR_ANNOTATION_LEN;
            return false;
          }
          if ((uint64_t)alen * 8ull > (uint64_t)(s.end - s.pos)) {
            s.err = M3TSZ_ERR_EOF;
            return false;
          }
          if (s.ann_count == 0) {
            s.ann_bit = s.pos;
            s.ann_len = (uint32_t)alen;
          }
          s.ann_count++;
          s.pos += (uint32_t)alen * 8u;
          continue;
        }
        if (m == kMarkerTimeUnit) {  // ReadTimeUnit :118-134
          s.pos += kMarkerBits;
          M3_RD(8, tub);
          int tu = (int)tub;
          if (unit_is_valid(tu) && tu != s.unit) {
            unit_changed = true;
            s.scheme = scheme_kind_for_unit(tu);
          }
          s.unit = tu;
          s.unit_ns = unit_nanos(tu);
          continue;
        }
        // marker value 3: not a marker, parse as delta-of-delta
      }
    }
    break;
  }
  // readDeltaOfDelta :246-305
  if (unit_changed) {  // readFullTimestamp :307-326
    int k = scheme_kind_for_unit(s.unit);
    if (k == kSchemeNone) {
      s.err = M3TSZ_ERR_NO_TIME_SCHEME;
      return false;
    }
    s.scheme = k;
    M3_RD(64, x);
    dod = (int64_t)x;
  } else if (s.scheme == kSchemeNone) {
    s.err = M3TSZ_ERR_NO_TIME_SCHEME;
    return false;
  } else {
    M3_RD(1, cb);
    if (cb != 0 && s.scheme != kSchemeZero) {
      int nb = (s.scheme == kScheme32) ? 32 : 64;
      bool swallowed = false;
      for (int i = 0; i < 3; i++) {
        if (s.pos + 1 > s.end) {  // error swallowed by the reference (:271-274)
          swallowed = true;
          break;
        }
        uint64_t b = gpeek64(base, nbytes, s.wbase, s.pos) >> 63;
        s.pos += 1;
        if (b == 0) {
          nb = (i == 0) ? 7 : (i == 1 ? 9 : 12);
          break;
        }
      }
      if (!swallowed) {
        M3_RD(nb, bits);
        int64_t d = sign_extend(bits, nb);
        dod = unit_is_valid(s.unit) ? (int64_t)((uint64_t)d * (uint64_t)s.unit_ns) : 0;
      }
    }
    // kSchemeZero: no buckets, default bucket has 0 value bits => dod = 0
  }
  s.prev_delta = (int64_t)((uint64_t)s.prev_delta + (uint64_t)dod);
  if (first)
    s.prev_time = (int64_t)((uint64_t)nt + (uint64_t)s.prev_delta);
  else
    s.prev_time = (int64_t)((uint64_t)s.prev_time + (uint64_t)s.prev_delta);
  if (unit_changed) s.prev_delta = 0;

  // ---- value: iterator.go:108-219 ----
  bool full_float = false, next_float = false, int_hdr = false, int_diff = false;
  if (!INT_OPT) {
    full_float = first;
    next_float = !first;
  } else if (first) {
    M3_RD(1, b);
    if (b == 1) {
      full_float = true;
      s.is_float = true;
    } else {
      int_hdr = true;
      int_diff = true;
    }
  } else {
    M3_RD(1, b);
    if (b == 0) {  // opcodeUpdate
      M3_RD(1, r);
      if (r == 1) {
        // repeat: value unchanged
      } else {
        M3_RD(1, f);
        if (f == 1) {
          full_float = true;
          s.is_float = true;
        } else {
          int_hdr = true;
          int_diff = true;
          // isFloat = false is applied after the diff is read (iterator.go:148)
        }
      }
    } else if (s.is_float) {
      next_float = true;
    } else {
      int_diff = true;
    }
  }
  if (full_float) {  // readFullFloat float_encoder_iterator.go:105-115
    M3_RD(64, vb);
    s.prev_bits = vb;
    s.prev_xor = vb;
  }
  if (next_float) {  // readNextFloat :117-165
    M3_RD(1, cb);
    if (cb == 0) {
      s.prev_xor = 0;
    } else {
      M3_RD(1, c2);
      if (c2 == 0) {
        int pl, pt;
        lz_tz(s.prev_xor, pl, pt);
        int nm = 64 - pl - pt;
        M3_RD(nm, mb);
        s.prev_xor = shl64(mb, pt);
        s.prev_bits ^= s.prev_xor;
      } else {
        M3_RD(12, hdr);
        int nlz = (int)((hdr >> 6) & 63);
        int nm = (int)(hdr & 63) + 1;
        M3_RD(nm, mb);
        int ntz = 64 - nlz - nm;
        s.prev_xor = (ntz < 0) ? 0ull : shl64(mb, ntz);
        s.prev_bits ^= s.prev_xor;
      }
    }
  }
  if (int_hdr) {  // readIntSigMult iterator.go:178-193
    M3_RD(1, us);
    if (us == 1) {
      M3_RD(1, nz);
      if (nz == 0) {
        s.sig = 0;
      } else {
        M3_RD(6, sb);
        s.sig = (int)sb + 1;
      }
    }
    M3_RD(1, um);
    if (um == 1) {
      M3_RD(3, mb);
      s.mult = (int)mb;
      if (s.mult > kMaxMult) {
        s.err = M3TSZ_ERR_INVALID_MULT;
        return false;
      }
    }
  }
  if (int_diff) {  // readIntValDiff(+Slow) iterator.go:195-219
    uint64_t neg, mag;
    if (s.sig == 64) {
      M3_RD(1, sg);
      M3_RD(64, mg);
      neg = sg;
      mag = mg;
    } else {
      M3_RD(s.sig + 1, bits);
      neg = bits >> s.sig;
      mag = bits & ((1ull << s.sig) - 1ull);
    }
    double m = __ull2double_rn(mag);
    s.int_val = neg ? __dadd_rn(s.int_val, m) : __dsub_rn(s.int_val, m);
    if (int_hdr) s.is_float = false;
  }
  out_t = s.prev_time;
  if (!INT_OPT || s.is_float) {
    out_v = s.prev_bits;
  } else {
    double v = s.mult ? __ddiv_rn(s.int_val, mult_pow10(s.mult)) : s.int_val;  // m3tsz.go:121-127
    out_v = (uint64_t)__double_as_longlong(v);
  }
  return true;
}
#undef M3_RD

// 64 bits starting at bit q (0 <= q < 64) of the 128-bit window w0:w1:w2:w3
__device__ __forceinline__ uint64_t extract64(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                              uint32_t q) {
  const bool up = q >= 32;
  const uint32_t a = up ? w1 : w0, b = up ? w2 : w1, c = up ? w3 : w2;
  const uint32_t hi = __funnelshift_l(b, a, q);
  const uint32_t lo = __funnelshift_l(c, b, q);
  return ((uint64_t)hi << 32) | lo;
}

struct DsAcc {  // fused-downsample per-lane accumulator (aggregation.Gauge, gauge.go:31-106)
  int64_t cur_w, hi_w, w_start;
  double sum, mn, mx;
  int64_t cnt;
};

template <bool INT_OPT, int MODE>
__global__ void __launch_bounds__(DEC_WARPS * 32)
    decode_kernel(const DecodeParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr size_t warp_smem = (MODE == 0) ? DEC_WARP_SMEM_PLAIN : DEC_WARP_SMEM_DS;
  uint32_t *in_tile = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(smem) + warp * warp_smem);
  uint64_t *ts_tile = reinterpret_cast<uint64_t *>(in_tile + DEC_IN_TILE_WORDS);
  uint64_t *val_tile = ts_tile + DEC_OUT_TILE_DWORDS;

  const uint64_t warp_s0 = ((uint64_t)blockIdx.x * DEC_WARPS + warp) * 32ull;
  if (warp_s0 >= p.n_series) return;
  const uint64_t sidx = warp_s0 + lane;
  const bool valid = sidx < p.n_series;

  DecState s;
  s.wbase = 0;
  s.pos = 0;
  s.end = 0;
  s.prev_time = 0;
  s.prev_delta = 0;
  s.prev_bits = 0;
  s.prev_xor = 0;
  s.int_val = 0.0;
  s.unit_ns = 0;
  s.sig = 0;
  s.mult = 0;
  s.unit = 0;
  s.scheme = kSchemeNone;
  s.err = 0;
  s.n = 0;
  s.is_float = false;
  s.done = !valid;
  s.ann_count = 0;
  s.ann_len = 0;
  s.ann_bit = 0;
  uint32_t pos0 = 0;
  if (valid) {
    const uint64_t o0 = p.offsets[sidx], o1 = p.offsets[sidx + 1];
    if (o1 < o0 || o1 > p.streams_bytes) {
      s.err = M3TSZ_ERR_INVALID_ARG;
    } else if (o1 - o0 >= (1ull << 28)) {
      s.err = M3TSZ_ERR_STREAM_TOO_LARGE;
    } else {
      s.wbase = o0 >> 2;
      s.pos = (uint32_t)(o0 & 3) * 8u;
      s.end = s.pos + (uint32_t)(o1 - o0) * 8u;
      pos0 = s.pos;
    }
  }

  DsAcc acc;
  acc.cur_w = -1;
  acc.hi_w = -1;
  acc.w_start = 0;
  acc.sum = 0.0;
  acc.mn = __longlong_as_double((long long)kGoNaNBits);
  acc.mx = acc.mn;
  acc.cnt = 0;
  const int64_t range_end = p.range_start + (int64_t)p.n_windows * p.window;
  (void)range_end;

  uint32_t tile_w0 = 0;
  bool tile_valid = false;
  uint32_t iter = 0;       // warp-uniform datapoint index
  uint32_t tile_row0 = 0;  // datapoint index of output tile row 0

  for (;;) {
    const bool active = !s.done && s.err == 0;
    if (!__any_sync(FULL_MASK, active)) break;

    // ---- (re)stage the compressed words of all 32 series ----
    uint32_t k = (s.pos >> 5) - tile_w0;
    const bool need = active && (!tile_valid || k > (uint32_t)(DEC_IN_W - DEC_FAST_WORDS));
    if (__any_sync(FULL_MASK, need)) {
      const uint32_t amask = __ballot_sync(FULL_MASK, active);
      tile_w0 = s.pos >> 5;
      const uint64_t my_src = s.wbase + tile_w0;
      __syncwarp();
#pragma unroll 4
      for (int j = 0; j < 32; j++) {
        const uint64_t src = __shfl_sync(FULL_MASK, my_src, j);
        if ((amask >> j) & 1u) {
#pragma unroll
          for (int hh = 0; hh < DEC_IN_W / 32; hh++) {
            const int wi = lane + 32 * hh;
            in_tile[wi * DEC_STRIDE + j] = load_be32(p.streams, p.streams_bytes, src + wi);
          }
        }
      }
      __syncwarp();
      tile_valid = true;
      k = 0;
    }

    int64_t t = 0;
    uint64_t v = 0;
    bool emitted = false;
    if (active) {
      // ---------------- fast path: parse from 4 staged words ----------------
      const uint32_t *tp = in_tile + k * DEC_STRIDE + lane;
      const uint32_t w0 = tp[0], w1 = tp[DEC_STRIDE], w2 = tp[2 * DEC_STRIDE],
                     w3 = tp[3 * DEC_STRIDE];
      const uint32_t sh = s.pos & 31u;
      const uint32_t h = __funnelshift_l(w1, w0, sh);
      bool ok = (s.prev_time != 0) && (s.scheme == kScheme32 || s.scheme == kScheme64) &&
                (s.unit >= 1 && s.unit <= 4);
      uint32_t c = 1;  // bits consumed
      int64_t dod = 0;
      if (h >> 31) {  // non-zero delta-of-delta bucket, marker, or default bucket
        if ((h >> 23) == kMarkerOpcode) {
          ok = false;
        } else if ((h >> 30) == 2u) {
          c = 9;
          dod = (int64_t)(((int32_t)(h << 2)) >> 25);
        } else if ((h >> 29) == 6u) {
          c = 12;
          dod = (int64_t)(((int32_t)(h << 3)) >> 23);
        } else if ((h >> 28) == 14u) {
          c = 16;
          dod = (int64_t)(((int32_t)(h << 4)) >> 20);
        } else {
          ok = false;
        }
        dod = (int64_t)((uint64_t)dod * (uint64_t)s.unit_ns);
      }
      uint32_t x = h << c;
      // value grammar (iterator.go:128-176)
      int kind = 0;  // 0 float-next, 1 int-diff, 2 repeat
      if (INT_OPT) {
        if (x >> 31) {
          x <<= 1;
          c += 1;
          kind = s.is_float ? 0 : 1;
          if (kind == 1 && s.sig == 64) ok = false;
        } else if ((x >> 30) == 1u) {
          c += 2;
          kind = 2;
        } else {
          ok = false;
        }
      }
      if (ok) {
        int n = 0, tz = 0;
        bool zero_xor = false;
        if (kind == 0) {
          if (!(x >> 31)) {
            zero_xor = true;
            c += 1;
          } else if (!(x & 0x40000000u)) {
            int pl, pt;
            lz_tz(s.prev_xor, pl, pt);
            n = 64 - pl - pt;
            tz = pt;
            c += 2;
          } else {
            const int lz = (int)((x >> 24) & 63u);
            n = (int)((x >> 18) & 63u) + 1;
            tz = 64 - lz - n;
            c += 14;
          }
        } else if (kind == 1) {
          n = s.sig + 1;
        }
        uint64_t payload = 0;
        if (n > 0) payload = extract64(w0, w1, w2, w3, sh + c) >> (64 - n);
        c += (uint32_t)n;
        if (s.pos + c > s.end) {
          s.err = M3TSZ_ERR_EOF;  // truncated stream: datapoint is not produced
        } else {
          s.pos += c;
          s.prev_delta = (int64_t)((uint64_t)s.prev_delta + (uint64_t)dod);
          s.prev_time = (int64_t)((uint64_t)s.prev_time + (uint64_t)s.prev_delta);
          if (kind == 0) {
            const uint64_t xr = zero_xor ? 0ull : ((tz < 0) ? 0ull : (payload << tz));
            s.prev_xor = xr;
            s.prev_bits ^= xr;
          } else if (kind == 1) {
            const uint64_t neg = payload >> s.sig;
            const uint64_t mag = payload & ((1ull << s.sig) - 1ull);
            const double m = __ull2double_rn(mag);
            s.int_val = neg ? __dadd_rn(s.int_val, m) : __dsub_rn(s.int_val, m);
          }
          t = s.prev_time;
          if (!INT_OPT || s.is_float) {
            v = s.prev_bits;
          } else {
            const double dv = s.mult ? __ddiv_rn(s.int_val, mult_pow10(s.mult)) : s.int_val;
            v = (uint64_t)__double_as_longlong(dv);
          }
          emitted = true;
        }
      } else {
        // copy-in / copy-out keeps the lane state in registers on the fast path
        DecState tmp = s;
        int64_t st = 0;
        uint64_t sv = 0;
        emitted = decode_dp_slow<INT_OPT>(tmp, p.streams, p.streams_bytes, p.default_unit, st, sv);
        s = tmp;
        t = st;
        v = sv;
      }
    }

    // ---------------- sink ----------------
    if (MODE == 0) {
      const int row = (int)(iter - tile_row0);
      if (emitted) {
        ts_tile[row * DEC_STRIDE + lane] = (uint64_t)t;
        val_tile[row * DEC_STRIDE + lane] = v;
        s.n++;
      }
    } else {
      if (emitted) {
        s.n++;
        if (t >= p.range_start && t < range_end) {
          if (!(acc.cur_w >= 0 && t >= acc.w_start && t - acc.w_start < p.window)) {
            // commit the window we are leaving
            if (acc.cur_w >= 0) {
              const uint64_t o = (uint64_t)acc.cur_w * p.n_series + sidx;
              p.ds_sum[o] = acc.sum;
              p.ds_count[o] = acc.cnt;
              p.ds_min[o] = acc.mn;
              p.ds_max[o] = acc.mx;
            }
            int64_t nw;
            if (acc.cur_w >= 0 && t >= acc.w_start + p.window && t - acc.w_start < 2 * p.window)
              nw = acc.cur_w + 1;
            else
              nw = (int64_t)((uint64_t)(t - p.range_start) / (uint64_t)p.window);
            if (nw > acc.hi_w) {
              for (int64_t w = acc.hi_w + 1; w < nw; w++) {
                const uint64_t o = (uint64_t)w * p.n_series + sidx;
                p.ds_sum[o] = 0.0;
                p.ds_count[o] = 0;
                p.ds_min[o] = __longlong_as_double((long long)kGoNaNBits);
                p.ds_max[o] = __longlong_as_double((long long)kGoNaNBits);
              }
              acc.hi_w = nw;
              acc.sum = 0.0;
              acc.cnt = 0;
              acc.mn = __longlong_as_double((long long)kGoNaNBits);
              acc.mx = acc.mn;
            } else {  // out-of-order timestamp: reopen a committed window
              const uint64_t o = (uint64_t)nw * p.n_series + sidx;
              acc.sum = p.ds_sum[o];
              acc.cnt = p.ds_count[o];
              acc.mn = p.ds_min[o];
              acc.mx = p.ds_max[o];
            }
            acc.cur_w = nw;
            acc.w_start = p.range_start + nw * p.window;
          }
          const double dv = __longlong_as_double((long long)v);
          acc.cnt++;
          if (dv == dv) {  // gauge.go:88-101
            acc.sum = __dadd_rn(acc.sum, dv);
            if (acc.mx != acc.mx || acc.mx < dv) acc.mx = dv;
            if (acc.mn != acc.mn || acc.mn > dv) acc.mn = dv;
          }
        }
      }
    }
    iter++;

    // ---------------- flush a full output tile ----------------
    if (MODE == 0 && iter - tile_row0 == (uint32_t)DEC_OUT_T) {
      __syncwarp();
      const uint32_t lim = s.n < (uint32_t)p.cap ? s.n : (uint32_t)p.cap;
      const uint32_t my_rows = lim > tile_row0 ? min(lim - tile_row0, (uint32_t)DEC_OUT_T) : 0u;
      const int r = lane & (DEC_OUT_T - 1);
      const bool isval = lane >= DEC_OUT_T;
      const uint64_t *tile = isval ? val_tile : ts_tile;
      uint64_t *dst0 = (isval ? reinterpret_cast<uint64_t *>(p.val) : reinterpret_cast<uint64_t *>(p.ts)) +
                       warp_s0 * p.cap + tile_row0 + r;
#pragma unroll 4
      for (int j = 0; j < 32; j++) {
        const uint32_t rows = __shfl_sync(FULL_MASK, my_rows, j);
        if ((uint32_t)r < rows) dst0[(uint64_t)j * p.cap] = tile[r * DEC_STRIDE + j];
      }
      __syncwarp();
      tile_row0 = iter;
    }
  }

  // ---------------- epilogue ----------------
  if (MODE == 0) {
    if (iter > tile_row0) {
      __syncwarp();
      const uint32_t lim = s.n < (uint32_t)p.cap ? s.n : (uint32_t)p.cap;
      const uint32_t my_rows = lim > tile_row0 ? min(lim - tile_row0, (uint32_t)DEC_OUT_T) : 0u;
      const int r = lane & (DEC_OUT_T - 1);
      const bool isval = lane >= DEC_OUT_T;
      const uint64_t *tile = isval ? val_tile : ts_tile;
      uint64_t *dst0 = (isval ? reinterpret_cast<uint64_t *>(p.val) : reinterpret_cast<uint64_t *>(p.ts)) +
                       warp_s0 * p.cap + tile_row0 + r;
      for (int j = 0; j < 32; j++) {
        const uint32_t rows = __shfl_sync(FULL_MASK, my_rows, j);
        if ((uint32_t)r < rows) dst0[(uint64_t)j * p.cap] = tile[r * DEC_STRIDE + j];
      }
    }
  } else if (valid) {
    if (acc.cur_w >= 0) {
      const uint64_t o = (uint64_t)acc.cur_w * p.n_series + sidx;
      p.ds_sum[o] = acc.sum;
      p.ds_count[o] = acc.cnt;
      p.ds_min[o] = acc.mn;
      p.ds_max[o] = acc.mx;
    }
    for (int64_t w = acc.hi_w + 1; w < (int64_t)p.n_windows; w++) {
      const uint64_t o = (uint64_t)w * p.n_series + sidx;
      p.ds_sum[o] = 0.0;
      p.ds_count[o] = 0;
      p.ds_min[o] = __longlong_as_double((long long)kGoNaNBits);
      p.ds_max[o] = __longlong_as_double((long long)kGoNaNBits);
    }
  }
  if (valid) {
    if (p.n_points) p.n_points[sidx] = s.n;
    int st = s.err;
    if (MODE == 0 && st == 0 && s.n > p.cap) st = M3TSZ_ERR_CAPACITY;
    if (p.status) p.status[sidx] = st;
    if (p.unit_out) p.unit_out[sidx] = (uint8_t)s.unit;
    if (p.ann_out) {
      m3tsz_annotation_ref a;
      a.bit_offset = s.ann_count ? (uint64_t)(s.ann_bit - pos0) : 0ull;
      a.length = s.ann_len;
      a.count = s.ann_count;
      p.ann_out[sidx] = a;
    }
  }
}

template <bool INT_OPT, int MODE>
static cudaError_t launch_one(const DecodeParams &p, cudaStream_t stream) {
  constexpr size_t warp_smem = (MODE == 0) ? DEC_WARP_SMEM_PLAIN : DEC_WARP_SMEM_DS;
  constexpr size_t smem = warp_smem * DEC_WARPS;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(decode_kernel<INT_OPT, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const uint64_t per_block = (uint64_t)DEC_WARPS * 32ull;
  const uint64_t blocks = (p.n_series + per_block - 1) / per_block;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
  decode_kernel<INT_OPT, MODE><<<(unsigned)blocks, DEC_WARPS * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeParams &p, bool int_optimized, bool downsample,
                          cudaStream_t stream) {
  if (!downsample)
    return int_optimized ? launch_one<true, 0>(p, stream) : launch_one<false, 0>(p, stream);
  return int_optimized ? launch_one<true, 1>(p, stream) : launch_one<false, 1>(p, stream);
}

}  // namespace m3tsz

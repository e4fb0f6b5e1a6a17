/*
 * m3tsz_oracle.c -- CPU restatement of the reference M3TSZ codec (TEST
 * INFRASTRUCTURE ONLY; see m3tsz_oracle.h for the rules on who may call it).
 *
 * Structured like the reference on purpose (byte-append bit writer, 64-bit
 * buffered bit reader, per-datapoint state machines) so that, timed on the host
 * cores, it is a fair stand-in for the reference's Go CPU path ("port" baseline).
 *
 * Reference paths are relative to /root/reference/src/dbnode/encoding.
 */
#include "m3tsz_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Go semantics helpers                                                        */
/* ------------------------------------------------------------------------- */

static inline uint64_t f64_bits(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return b;
}
static inline double f64_from_bits(uint64_t b) {
  double v;
  memcpy(&v, &b, 8);
  return v;
}
static inline int is_nan(double v) { return v != v; }

/* Go: x >> n and x << n are 0 for n >= 64 (Appendix B.7 of SURVEY.md). */
static inline uint64_t shr64(uint64_t x, unsigned n) { return n >= 64 ? 0 : x >> n; }
static inline uint64_t shl64(uint64_t x, unsigned n) { return n >= 64 ? 0 : x << n; }

/* Go on amd64: int64(float64) is CVTTSD2SQ; out-of-range and NaN give
 * 0x8000000000000000 (reachable via encoder.go:143,214). */
static inline int64_t go_f64_to_i64(double v) {
  if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)v;
}

/* math.Modf (Go 1.22 pure-Go path used on amd64, math/modf.go). */
static void go_modf(double f, double *ip, double *frac) {
  if (f < 1) {
    if (f < 0) {
      double i2, f2;
      go_modf(-f, &i2, &f2);
      *ip = -i2;
      *frac = -f2;
      return;
    }
    if (f == 0) {
      *ip = f;
      *frac = f;
      return;
    }
    *ip = 0;
    *frac = f;
    return;
  }
  uint64_t x = f64_bits(f);
  unsigned e = (unsigned)((x >> 52) & 0x7ff) - 1023u;
  if (e < 64 - 12) x &= ~((((uint64_t)1) << (64 - 12 - e)) - 1);
  *ip = f64_from_bits(x);
  *frac = f - *ip;
}

/* math.Nextafter (math/nextafter.go). */
static double go_nextafter(double x, double y) {
  if (is_nan(x) || is_nan(y)) return f64_from_bits(0x7ff8000000000001ULL);
  if (x == y) return x;
  if (x == 0) {
    uint64_t b = 1;
    if (f64_bits(y) >> 63) b |= 0x8000000000000000ULL;
    return f64_from_bits(b);
  }
  if ((y > x) == (x > 0)) return f64_from_bits(f64_bits(x) + 1);
  return f64_from_bits(f64_bits(x) - 1);
}

/* src/x/time/unit.go:185-195 unitsToDuration */
static const int64_t k_unit_ns[M3O_UNIT_COUNT] = {
    0,
    1000000000LL,
    1000000LL,
    1000LL,
    1LL,
    60LL * 1000000000LL,
    3600LL * 1000000000LL,
    24LL * 3600LL * 1000000000LL,
    365LL * 24LL * 3600LL * 1000000000LL,
};

static inline int unit_is_valid(int u) { return u > 0 && u < M3O_UNIT_COUNT; } /* unit.go:91-93 */

/* Unit.Value(), unit.go:56-61 */
static int unit_value(int u, int64_t *out) {
  if (u < 1 || u >= M3O_UNIT_COUNT) return M3O_ERR_UNRECOGNIZED_UNIT;
  *out = k_unit_ns[u];
  return M3O_OK;
}

/* ------------------------------------------------------------------------- */
/* encoding.go:29-49                                                           */
/* ------------------------------------------------------------------------- */

int m3o_num_sig(uint64_t v) { return v == 0 ? 0 : 64 - __builtin_clzll(v); }

void m3o_leading_trailing_zeros(uint64_t v, int *lz, int *tz) {
  if (v == 0) {
    *lz = 64;
    *tz = 0;
    return;
  }
  *lz = __builtin_clzll(v);
  *tz = __builtin_ctzll(v);
}

int64_t m3o_sign_extend(uint64_t v, int nbits) {
  unsigned shift = 64u - (unsigned)nbits;
  if (shift >= 64) return 0; /* Go: (int64(v) << 64) >> 64 == 0 */
  return ((int64_t)(v << shift)) >> shift;
}

/* ------------------------------------------------------------------------- */
/* ostream.go:86-221                                                           */
/* ------------------------------------------------------------------------- */

struct m3o_ostream {
  uint8_t *buf;
  size_t len, cap;
  int pos; /* bits used in the last byte: 0 (empty), 1..8 */
};

static void os_init(m3o_ostream *os) {
  os->buf = NULL;
  os->len = os->cap = 0;
  os->pos = 0;
}
static void os_release(m3o_ostream *os) {
  free(os->buf);
  os_init(os);
}
m3o_ostream *m3o_ostream_new(void) {
  m3o_ostream *os = (m3o_ostream *)malloc(sizeof(*os));
  os_init(os);
  return os;
}
void m3o_ostream_free(m3o_ostream *os) {
  if (!os) return;
  free(os->buf);
  free(os);
}
void m3o_ostream_reset(m3o_ostream *os) {
  os->len = 0;
  os->pos = 0;
}

/* ensureCapacityFor, ostream.go:98-131 (doubling growth; initial 1024, ostream.go:29) */
static void os_ensure(m3o_ostream *os, size_t n) {
  if (os->cap - os->len >= n) return;
  size_t ncap = os->cap * 2;
  if (ncap < os->len + n) ncap = os->len + n;
  if (ncap < 1024) ncap = 1024;
  os->buf = (uint8_t *)realloc(os->buf, ncap);
  os->cap = ncap;
}
static inline int os_has_unused(const m3o_ostream *os) { return os->pos > 0 && os->pos < 8; } /* :86-88 */
static inline void os_grow(m3o_ostream *os, uint8_t v, int np) { /* :91-96 */
  os_ensure(os, 1);
  os->buf[os->len++] = v;
  os->pos = np;
}
static inline void os_fill_unused(m3o_ostream *os, uint8_t v) { /* :133-135 */
  os->buf[os->len - 1] |= (uint8_t)(v >> os->pos);
}
void m3o_ostream_write_bit(m3o_ostream *os, int bit) { /* :137-145 */
  uint8_t v = (uint8_t)((bit & 1) << 7);
  if (!os_has_unused(os)) {
    os_grow(os, v, 1);
    return;
  }
  os_fill_unused(os, v);
  os->pos++;
}
void m3o_ostream_write_byte(m3o_ostream *os, uint8_t v) { /* :147-154 */
  if (!os_has_unused(os)) {
    os_grow(os, v, 8);
    return;
  }
  os_fill_unused(os, v);
  os_grow(os, (uint8_t)(v << (8 - os->pos)), os->pos);
}
void m3o_ostream_write_bytes(m3o_ostream *os, const uint8_t *p, size_t n) { /* :156-178 */
  os_ensure(os, n);
  if (!os_has_unused(os)) {
    if (n) memcpy(os->buf + os->len, p, n);
    os->len += n;
    os->pos = 8;
    return;
  }
  for (size_t i = 0; i < n; i++) m3o_ostream_write_byte(os, p[i]);
}
void m3o_ostream_write_bits(m3o_ostream *os, uint64_t v, int nbits) { /* :185-221 */
  if (nbits == 0) return;
  if (nbits > 64) nbits = 64;
  v <<= (unsigned)(64 - nbits);
  while (nbits >= 32) {
    m3o_ostream_write_byte(os, (uint8_t)(v >> 56));
    m3o_ostream_write_byte(os, (uint8_t)(v >> 48));
    m3o_ostream_write_byte(os, (uint8_t)(v >> 40));
    m3o_ostream_write_byte(os, (uint8_t)(v >> 32));
    v <<= 32;
    nbits -= 32;
  }
  while (nbits >= 8) {
    m3o_ostream_write_byte(os, (uint8_t)(v >> 56));
    v <<= 8;
    nbits -= 8;
  }
  uint8_t rem = (uint8_t)(v >> 56);
  while (nbits > 0) {
    uint8_t val = rem & 0x80;
    if (os_has_unused(os)) {
      os_fill_unused(os, val);
      os->pos++;
    } else {
      os_grow(os, val, 1);
    }
    rem = (uint8_t)(rem << 1);
    nbits--;
  }
}
size_t m3o_ostream_raw(const m3o_ostream *os, const uint8_t **data, int *pos) { /* :250-252 */
  if (data) *data = os->buf;
  if (pos) *pos = os->pos;
  return os->len;
}

/* ------------------------------------------------------------------------- */
/* x/xio/reader64.go:40-81 + istream.go:48-125                                 */
/* ------------------------------------------------------------------------- */

struct m3o_istream {
  const uint8_t *data;
  size_t len, index;
  uint64_t current;
  unsigned remaining;
};

static void is_reset(m3o_istream *is, const uint8_t *data, size_t len) { /* istream.go:127-133 */
  is->data = data;
  is->len = len;
  is->index = 0;
  is->current = 0;
  is->remaining = 0;
}
m3o_istream *m3o_istream_new(const uint8_t *data, size_t len) {
  m3o_istream *is = (m3o_istream *)malloc(sizeof(*is));
  is_reset(is, data, len);
  return is;
}
void m3o_istream_free(m3o_istream *is) { free(is); }

/* BytesReader64.Read64 / Peek64, reader64.go:40-81 */
static int r64(m3o_istream *is, int advance, uint64_t *word, unsigned *nbytes) {
  if (is->index + 8 <= is->len) {
    const uint8_t *p = is->data + is->index;
    *word = ((uint64_t)p[0] << 56) | ((uint64_t)p[1] << 48) | ((uint64_t)p[2] << 40) |
            ((uint64_t)p[3] << 32) | ((uint64_t)p[4] << 24) | ((uint64_t)p[5] << 16) |
            ((uint64_t)p[6] << 8) | (uint64_t)p[7];
    *nbytes = 8;
    if (advance) is->index += 8;
    return M3O_OK;
  }
  if (is->index >= is->len) {
    *word = 0;
    *nbytes = 0;
    return M3O_ERR_EOF;
  }
  uint64_t res = 0;
  unsigned bytes = 0;
  size_t i = is->index;
  for (; i < is->len; i++) {
    res = (res << 8) | is->data[i];
    bytes++;
  }
  if (advance) is->index = i;
  *word = res << (64 - 8 * bytes);
  *nbytes = bytes;
  return M3O_OK;
}

int m3o_istream_read_bits(m3o_istream *is, int nbits, uint64_t *out) { /* istream.go:73-98 */
  unsigned n = (unsigned)nbits;
  uint64_t res = shr64(is->current, 64 - n);
  if (n <= is->remaining) {
    is->current = shl64(is->current, n);
    is->remaining -= n;
    *out = res;
    return M3O_OK;
  }
  unsigned needed = n - is->remaining;
  uint64_t cur;
  unsigned nb;
  int err = r64(is, 1, &cur, &nb);
  if (err) {
    *out = 0;
    return err;
  }
  nb *= 8;
  if (nb < needed) {
    *out = 0;
    return M3O_ERR_EOF;
  }
  is->current = shl64(cur, needed);
  is->remaining = nb - needed;
  *out = res | shr64(cur, 64 - needed);
  return M3O_OK;
}

int m3o_istream_peek_bits(m3o_istream *is, int nbits, uint64_t *out) { /* istream.go:100-115 */
  unsigned n = (unsigned)nbits;
  if (n <= is->remaining) {
    *out = shr64(is->current, 64 - n);
    return M3O_OK;
  }
  uint64_t res = shr64(is->current, 64 - n);
  unsigned needed = n - is->remaining;
  uint64_t next;
  unsigned nb;
  int err = r64(is, 0, &next, &nb);
  if (err) {
    *out = 0;
    return err;
  }
  if (8 * nb < needed) {
    *out = 0;
    return M3O_ERR_EOF;
  }
  *out = res | shr64(next, 64 - needed);
  return M3O_OK;
}

int m3o_istream_remaining_bits_in_current_byte(const m3o_istream *is) { /* :117-120 */
  return (int)(is->remaining % 8);
}

/* ------------------------------------------------------------------------- */
/* scheme.go:40-144, 171-242                                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
  int64_t min, max;
  uint64_t opcode;
  int n_opcode_bits, n_value_bits;
} time_bucket;

typedef struct {
  time_bucket zero;
  time_bucket buckets[3];
  int n_buckets;
  time_bucket dflt;
} time_scheme;

static time_bucket new_time_bucket(uint64_t opcode, int nop, int nval) { /* scheme.go:75-84 */
  time_bucket b;
  b.opcode = opcode;
  b.n_opcode_bits = nop;
  b.n_value_bits = nval;
  b.min = -(((int64_t)1) << (nval - 1));
  b.max = (((int64_t)1) << (nval - 1)) - 1;
  return b;
}

static time_scheme g_schemes[M3O_UNIT_COUNT];
static pthread_once_t g_schemes_once = PTHREAD_ONCE_INIT;

/* NewTimeEncodingScheme, scheme.go:124-144 with defaults :40-53.  Units without an
 * entry in defaultTimeEncodingSchemes (m, h, d, y) get the all-zero scheme that
 * NewTimeEncodingSchemes leaves in the slice (scheme.go:109-120). */
static void init_schemes(void) {
  memset(g_schemes, 0, sizeof(g_schemes));
  static const int nvb[3] = {7, 9, 12};
  for (int u = M3O_UNIT_SECOND; u <= M3O_UNIT_NANOSECOND; u++) {
    time_scheme *s = &g_schemes[u];
    /* defaultZeroBucket = NewTimeBucket(0x0, 1, 0), scheme.go:42; its min/max come
     * from 1<<uint(-1) == 0 in Go, i.e. (0,-1), and are never consulted. */
    s->zero.opcode = 0;
    s->zero.n_opcode_bits = 1;
    s->zero.n_value_bits = 0;
    s->zero.min = 0;
    s->zero.max = -1;
    int nop = 1;
    uint64_t opcode = 0;
    for (int i = 0; i < 3; i++) {
      opcode = (((uint64_t)1) << (i + 1)) | opcode;
      s->buckets[i] = new_time_bucket(opcode, nop + 1, nvb[i]);
      nop++;
    }
    s->n_buckets = 3;
    int dbits = (u == M3O_UNIT_SECOND || u == M3O_UNIT_MILLISECOND) ? 32 : 64;
    s->dflt = new_time_bucket(opcode | 0x1, nop, dbits);
  }
}

/* TimeEncodingSchemes.SchemeForUnit, scheme.go:160-165 */
static const time_scheme *scheme_for_unit(int u) {
  pthread_once(&g_schemes_once, init_schemes);
  if (!unit_is_valid(u)) return NULL;
  return &g_schemes[u];
}

/* default marker scheme, scheme.go:30-38 */
#define MARKER_OPCODE 0x100u
#define MARKER_OPCODE_BITS 9
#define MARKER_VALUE_BITS 2
#define MARKER_EOS 0
#define MARKER_ANNOTATION 1
#define MARKER_TIMEUNIT 2

static void write_special_marker(m3o_ostream *os, int marker) { /* scheme.go:217-220 */
  m3o_ostream_write_bits(os, MARKER_OPCODE, MARKER_OPCODE_BITS);
  m3o_ostream_write_bits(os, (uint64_t)marker, MARKER_VALUE_BITS);
}

/* MarkerEncodingScheme.Tail, scheme.go:198-211,242: bytes for (lastByte,pos) + EOS */
static size_t make_tail(uint8_t last, int pos, uint8_t out[4]) {
  m3o_ostream tmp;
  os_init(&tmp);
  m3o_ostream_write_bits(&tmp, (uint64_t)(last >> (8 - pos)), pos);
  write_special_marker(&tmp, MARKER_EOS);
  size_t n = tmp.len;
  memcpy(out, tmp.buf, n);
  os_release(&tmp);
  return n;
}

/* ------------------------------------------------------------------------- */
/* XXH64 seed 0 (= cespare/xxhash/v2 Sum64, go.mod:10), public algorithm       */
/* ------------------------------------------------------------------------- */
#define XP1 11400714785074694791ULL
#define XP2 14029467366897019727ULL
#define XP3 1609587929392839161ULL
#define XP4 9650029242287828579ULL
#define XP5 2870177450012600261ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64le(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v; /* host is little-endian x86-64 */
}
static inline uint32_t rd32le(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
static inline uint64_t xround(uint64_t acc, uint64_t in) {
  acc += in * XP2;
  acc = rotl64(acc, 31);
  return acc * XP1;
}
static inline uint64_t xmerge(uint64_t acc, uint64_t val) {
  val = xround(0, val);
  acc ^= val;
  return acc * XP1 + XP4;
}
uint64_t m3o_xxh64(const uint8_t *p, size_t n) {
  const uint8_t *end = p + n;
  uint64_t h;
  if (n >= 32) {
    uint64_t v1 = XP1 + XP2, v2 = XP2, v3 = 0, v4 = 0 - XP1;
    const uint8_t *lim = end - 32;
    do {
      v1 = xround(v1, rd64le(p));
      v2 = xround(v2, rd64le(p + 8));
      v3 = xround(v3, rd64le(p + 16));
      v4 = xround(v4, rd64le(p + 24));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xmerge(h, v1);
    h = xmerge(h, v2);
    h = xmerge(h, v3);
    h = xmerge(h, v4);
  } else {
    h = XP5;
  }
  h += (uint64_t)n;
  while (p + 8 <= end) {
    h ^= xround(0, rd64le(p));
    h = rotl64(h, 27) * XP1 + XP4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= (uint64_t)rd32le(p) * XP1;
    h = rotl64(h, 23) * XP2 + XP3;
    p += 4;
  }
  while (p < end) {
    h ^= (uint64_t)(*p) * XP5;
    h = rotl64(h, 11) * XP1;
    p++;
  }
  h ^= h >> 33;
  h *= XP2;
  h ^= h >> 29;
  h *= XP3;
  h ^= h >> 32;
  return h;
}

/* ------------------------------------------------------------------------- */
/* m3tsz/m3tsz.go:28-140                                                       */
/* ------------------------------------------------------------------------- */
#define OPC_ZERO_SIG 0x0
#define OPC_NONZERO_SIG 0x1
#define NUM_SIG_BITS 6
#define OPC_ZERO_XOR 0x0
#define OPC_CONTAINED_XOR 0x2
#define OPC_UNCONTAINED_XOR 0x3
#define OPC_NO_UPDATE_SIG 0x0
#define OPC_UPDATE_SIG 0x1
#define OPC_UPDATE 0x0
#define OPC_NO_UPDATE 0x1
#define OPC_UPDATE_MULT 0x1
#define OPC_NO_UPDATE_MULT 0x0
#define OPC_POSITIVE 0x0
#define OPC_NEGATIVE 0x1
#define OPC_REPEAT 0x1
#define OPC_NO_REPEAT 0x0
#define OPC_FLOAT_MODE 0x1
#define OPC_INT_MODE 0x0
#define SIG_DIFF_THRESHOLD 3
#define SIG_REPEAT_THRESHOLD 5
#define MAX_MULT 6
#define NUM_MULT_BITS 3

static const double k_max_int = 9223372036854775807.0;  /* float64(math.MaxInt64) == 2^63 */
static const double k_min_int = -9223372036854775808.0; /* float64(math.MinInt64) */
static const double k_max_opt_int = 1e13;               /* math.Pow(10, 13) */
/* createMultipliers, m3tsz.go:131-140: repeated base*10 is exact up to 1e6 */
static const double k_multipliers[MAX_MULT + 1] = {1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0};

int m3o_convert_to_int_float(double v, int cur_max_mult, double *val, int *mult,
                             int *is_float) { /* m3tsz.go:78-119 */
  if (cur_max_mult == 0 && v < k_max_int) {
    double i, r;
    go_modf(v, &i, &r);
    if (r == 0) {
      *val = i;
      *mult = 0;
      *is_float = 0;
      return M3O_OK;
    }
  }
  if (cur_max_mult > MAX_MULT) {
    *val = 0.0;
    *mult = 0;
    *is_float = 0;
    return M3O_ERR_INVALID_MULT;
  }
  double sign = 1.0;
  if (v < 0) sign = -1.0;
  for (int m = cur_max_mult; m <= MAX_MULT; m++) {
    /* two separate roundings, as Go/amd64 never fuses */
    volatile double t = v * k_multipliers[m];
    double x = t * sign;
    if (x >= k_max_opt_int) break;
    double i, r;
    go_modf(x, &i, &r);
    if (r == 0) {
      *val = sign * i;
      *mult = m;
      *is_float = 0;
      return M3O_OK;
    } else if (r < 0.1) {
      if (go_nextafter(x, 0) <= i) {
        *val = sign * i;
        *mult = m;
        *is_float = 0;
        return M3O_OK;
      }
    } else if (r > 0.9) {
      double next = i + 1;
      if (go_nextafter(x, next) >= next) {
        *val = sign * next;
        *mult = m;
        *is_float = 0;
        return M3O_OK;
      }
    }
  }
  *val = v;
  *mult = 0;
  *is_float = 1;
  return M3O_OK;
}

double m3o_convert_from_int_float(double val, int mult) { /* m3tsz.go:121-127 */
  if (mult == 0) return val;
  return val / k_multipliers[mult];
}

int m3o_initial_time_unit(int64_t start_ns, int unit) { /* timestamp_encoder.go:248-259 */
  int64_t tv;
  if (unit_value(unit, &tv)) return M3O_UNIT_NONE;
  if (start_ns % tv == 0) return unit;
  return M3O_UNIT_NONE;
}

/* ------------------------------------------------------------------------- */
/* m3tsz/float_encoder_iterator.go                                             */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint64_t prev_xor, prev_bits;
} float_state;

void m3o_write_xor(m3o_ostream *os, uint64_t prev_xor, uint64_t cur) { /* :82-103 */
  if (cur == 0) {
    m3o_ostream_write_bits(os, OPC_ZERO_XOR, 1);
    return;
  }
  int pl, pt, cl, ct;
  m3o_leading_trailing_zeros(prev_xor, &pl, &pt);
  m3o_leading_trailing_zeros(cur, &cl, &ct);
  if (cl >= pl && ct >= pt) {
    m3o_ostream_write_bits(os, OPC_CONTAINED_XOR, 2);
    m3o_ostream_write_bits(os, shr64(cur, (unsigned)pt), 64 - pl - pt);
    return;
  }
  m3o_ostream_write_bits(os, OPC_UNCONTAINED_XOR, 2);
  m3o_ostream_write_bits(os, (uint64_t)cl, 6);
  int nm = 64 - cl - ct;
  m3o_ostream_write_bits(os, (uint64_t)(nm - 1), 6);
  m3o_ostream_write_bits(os, shr64(cur, (unsigned)ct), nm);
}
static void write_full_float(float_state *f, m3o_ostream *os, uint64_t val) { /* :69-73 */
  f->prev_bits = val;
  f->prev_xor = val;
  m3o_ostream_write_bits(os, val, 64);
}
static void write_next_float(float_state *f, m3o_ostream *os, uint64_t val) { /* :75-80 */
  uint64_t x = f->prev_bits ^ val;
  m3o_write_xor(os, f->prev_xor, x);
  f->prev_xor = x;
  f->prev_bits = val;
}
static int read_full_float(float_state *f, m3o_istream *is) { /* :105-115 */
  uint64_t vb;
  int err = m3o_istream_read_bits(is, 64, &vb);
  if (err) return err;
  f->prev_bits = vb;
  f->prev_xor = vb;
  return M3O_OK;
}
static int read_next_float(float_state *f, m3o_istream *is) { /* :117-165 */
  uint64_t cb, ncb;
  int err = m3o_istream_read_bits(is, 1, &cb);
  if (err) return err;
  if (cb == OPC_ZERO_XOR) {
    f->prev_xor = 0;
    return M3O_OK;
  }
  err = m3o_istream_read_bits(is, 1, &ncb);
  if (err) return err;
  cb = (cb << 1) | ncb;
  if (cb == OPC_CONTAINED_XOR) {
    int pl, pt;
    m3o_leading_trailing_zeros(f->prev_xor, &pl, &pt);
    int nm = (64 - pl - pt) & 0xff;
    uint64_t mb;
    err = m3o_istream_read_bits(is, nm, &mb);
    if (err) return err;
    f->prev_xor = shl64(mb, (unsigned)pt);
    f->prev_bits ^= f->prev_xor;
    return M3O_OK;
  }
  uint64_t hdr;
  err = m3o_istream_read_bits(is, 12, &hdr);
  if (err) return err;
  uint64_t nlz = (hdr & 4032) >> 6;
  uint64_t nm = (hdr & 63) + 1;
  uint64_t mb;
  err = m3o_istream_read_bits(is, (int)nm, &mb);
  if (err) return err;
  uint64_t ntz = 64 - nlz - nm; /* wraps like Go uint64 arithmetic */
  f->prev_xor = shl64(mb, (unsigned)(ntz > 0xffffffffULL ? 64u : (unsigned)ntz));
  f->prev_bits ^= f->prev_xor;
  return M3O_OK;
}

/* ------------------------------------------------------------------------- */
/* m3tsz/int_sig_bits_tracker.go                                               */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint8_t num_sig, cur_highest_lower_sig, num_lower_sig;
} sig_tracker;

static void write_int_val_diff(sig_tracker *t, m3o_ostream *os, uint64_t bits, int neg) { /* :35-44 */
  m3o_ostream_write_bit(os, neg ? OPC_NEGATIVE : OPC_POSITIVE);
  m3o_ostream_write_bits(os, bits, (int)t->num_sig);
}
static void write_int_sig(sig_tracker *t, m3o_ostream *os, uint8_t sig) { /* :48-62 */
  if (t->num_sig != sig) {
    m3o_ostream_write_bit(os, OPC_UPDATE_SIG);
    if (sig == 0) {
      m3o_ostream_write_bit(os, OPC_ZERO_SIG);
    } else {
      m3o_ostream_write_bit(os, OPC_NONZERO_SIG);
      m3o_ostream_write_bits(os, (uint64_t)(sig - 1), NUM_SIG_BITS);
    }
  } else {
    m3o_ostream_write_bit(os, OPC_NO_UPDATE_SIG);
  }
  t->num_sig = sig;
}
static uint8_t track_new_sig(sig_tracker *t, uint8_t ns) { /* :68-91 */
  uint8_t new_sig = t->num_sig;
  if (ns > t->num_sig) {
    new_sig = ns;
  } else if ((uint8_t)(t->num_sig - ns) >= SIG_DIFF_THRESHOLD) {
    if (t->num_lower_sig == 0) {
      t->cur_highest_lower_sig = ns;
    } else if (ns > t->cur_highest_lower_sig) {
      t->cur_highest_lower_sig = ns;
    }
    t->num_lower_sig++;
    if (t->num_lower_sig >= SIG_REPEAT_THRESHOLD) {
      new_sig = t->cur_highest_lower_sig;
      t->num_lower_sig = 0;
    }
  } else {
    t->num_lower_sig = 0;
  }
  return new_sig;
}

/* ------------------------------------------------------------------------- */
/* m3tsz/timestamp_encoder.go                                                  */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t prev_time, prev_time_delta;
  uint64_t prev_ann_checksum;
  int time_unit;
  int unit_encoded_manually;
  int has_written_first;
} ts_encoder;

static uint64_t empty_ann_checksum(void) { return m3o_xxh64(NULL, 0); } /* :56 */

static ts_encoder new_ts_encoder(int64_t start, int unit) { /* :59-69 */
  ts_encoder e;
  memset(&e, 0, sizeof(e));
  e.prev_time = start;
  e.time_unit = m3o_initial_time_unit(start, unit);
  e.prev_ann_checksum = empty_ann_checksum();
  return e;
}

void m3o_write_dod_unit_changed(m3o_ostream *os, int64_t prev_delta, int64_t cur_delta) { /* :197-203 */
  int64_t dod = (int64_t)((uint64_t)cur_delta - (uint64_t)prev_delta);
  m3o_ostream_write_bits(os, (uint64_t)dod, 64);
}

int m3o_write_dod_unit_unchanged(m3o_ostream *os, int64_t prev_delta, int64_t cur_delta,
                                 int unit) { /* :205-246 */
  int64_t u;
  int err = unit_value(unit, &u);
  if (err) return err;
  int64_t dod = (int64_t)((uint64_t)cur_delta - (uint64_t)prev_delta) / u; /* x/time/time.go:55-57 */
  if (unit == M3O_UNIT_MILLISECOND || unit == M3O_UNIT_SECOND) {
    int32_t d32 = (int32_t)dod;
    if ((int64_t)d32 != dod) return M3O_ERR_DOD_OVERFLOW;
  }
  const time_scheme *tes = scheme_for_unit(unit);
  if (!tes) return M3O_ERR_NO_TIME_SCHEME;
  if (dod == 0) {
    m3o_ostream_write_bits(os, tes->zero.opcode, tes->zero.n_opcode_bits);
    return M3O_OK;
  }
  for (int i = 0; i < tes->n_buckets; i++) {
    if (dod >= tes->buckets[i].min && dod <= tes->buckets[i].max) {
      m3o_ostream_write_bits(os, tes->buckets[i].opcode, tes->buckets[i].n_opcode_bits);
      m3o_ostream_write_bits(os, (uint64_t)dod, tes->buckets[i].n_value_bits);
      return M3O_OK;
    }
  }
  m3o_ostream_write_bits(os, tes->dflt.opcode, tes->dflt.n_opcode_bits);
  m3o_ostream_write_bits(os, (uint64_t)dod, tes->dflt.n_value_bits);
  return M3O_OK;
}

/* binary.PutVarint (Go stdlib): zig-zag then LEB128 */
static size_t put_varint(uint8_t *buf, int64_t x) {
  uint64_t ux = ((uint64_t)x) << 1;
  if (x < 0) ux = ~ux;
  size_t i = 0;
  while (ux >= 0x80) {
    buf[i++] = (uint8_t)(ux | 0x80);
    ux >>= 7;
  }
  buf[i++] = (uint8_t)ux;
  return i;
}

static void ts_write_annotation(ts_encoder *e, m3o_ostream *os, const uint8_t *ann, size_t n) { /* :166-195 */
  if (n == 0) return; /* shouldWriteAnnotation :156-164 */
  uint64_t checksum = m3o_xxh64(ann, n);
  if (checksum == e->prev_ann_checksum) return;
  write_special_marker(os, MARKER_ANNOTATION);
  uint8_t buf[10];
  size_t vl = put_varint(buf, (int64_t)n - 1);
  m3o_ostream_write_bytes(os, buf, vl);
  m3o_ostream_write_bytes(os, ann, n);
  e->prev_ann_checksum = checksum;
}

static int ts_maybe_write_unit_change(ts_encoder *e, m3o_ostream *os, int unit) { /* :141-162 */
  if (!unit_is_valid(unit) || unit == e->time_unit) return 0;
  write_special_marker(os, MARKER_TIMEUNIT);
  m3o_ostream_write_byte(os, (uint8_t)unit); /* WriteTimeUnit :133-137 */
  e->time_unit = unit;
  e->unit_encoded_manually = 1;
  return 1;
}

static int ts_write_next_time(ts_encoder *e, m3o_ostream *os, int64_t t, const uint8_t *ann,
                              size_t ann_len, int unit) { /* :104-129 */
  ts_write_annotation(e, os, ann, ann_len);
  int changed = ts_maybe_write_unit_change(e, os, unit);
  int64_t delta = (int64_t)((uint64_t)t - (uint64_t)e->prev_time);
  e->prev_time = t;
  if (changed || e->unit_encoded_manually) {
    m3o_write_dod_unit_changed(os, e->prev_time_delta, delta);
    e->prev_time_delta = 0;
    e->unit_encoded_manually = 0;
    return M3O_OK;
  }
  int err = m3o_write_dod_unit_unchanged(os, e->prev_time_delta, delta, unit);
  e->prev_time_delta = delta;
  return err;
}

static int ts_write_time(ts_encoder *e, m3o_ostream *os, int64_t t, const uint8_t *ann,
                         size_t ann_len, int unit) { /* :72-102 */
  if (!e->has_written_first) {
    m3o_ostream_write_bits(os, (uint64_t)e->prev_time, 64); /* WriteFirstTime :89-102 */
    int err = ts_write_next_time(e, os, t, ann, ann_len, unit);
    if (err) return err;
    e->has_written_first = 1;
    return M3O_OK;
  }
  return ts_write_next_time(e, os, t, ann, ann_len, unit);
}

/* ------------------------------------------------------------------------- */
/* m3tsz/encoder.go                                                            */
/* ------------------------------------------------------------------------- */
struct m3o_encoder {
  m3o_ostream os;
  ts_encoder ts;
  float_state fl;
  sig_tracker sig;
  double int_val;
  uint32_t num_encoded;
  uint8_t max_mult;
  int int_optimized, is_float, closed;
  int default_unit;
};

m3o_encoder *m3o_encoder_new(int64_t start_ns, int int_optimized, int default_unit) { /* :64-85 */
  m3o_encoder *e = (m3o_encoder *)calloc(1, sizeof(*e));
  os_init(&e->os);
  e->default_unit = default_unit;
  e->ts = new_ts_encoder(start_ns, default_unit);
  e->int_optimized = int_optimized;
  return e;
}
void m3o_encoder_free(m3o_encoder *e) {
  if (!e) return;
  os_release(&e->os);
  free(e);
}
void m3o_encoder_reset(m3o_encoder *e, int64_t start_ns) { /* :262-279 */
  m3o_ostream_reset(&e->os);
  int tu = m3o_initial_time_unit(start_ns, e->default_unit);
  e->ts = new_ts_encoder(start_ns, tu);
  memset(&e->fl, 0, sizeof(e->fl));
  e->int_val = 0;
  e->is_float = 0;
  e->max_mult = 0;
  memset(&e->sig, 0, sizeof(e->sig));
  e->num_encoded = 0;
  e->closed = 0;
}

static void enc_write_int_sig_mult(m3o_encoder *e, uint8_t sig, uint8_t mult, int float_changed) { /* :235-250 */
  write_int_sig(&e->sig, &e->os, sig);
  if (mult > e->max_mult) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE_MULT);
    m3o_ostream_write_bits(&e->os, (uint64_t)mult, NUM_MULT_BITS);
    e->max_mult = mult;
  } else if (e->sig.num_sig == sig && e->max_mult == mult && float_changed) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE_MULT);
    m3o_ostream_write_bits(&e->os, (uint64_t)e->max_mult, NUM_MULT_BITS);
  } else {
    m3o_ostream_write_bit(&e->os, OPC_NO_UPDATE_MULT);
  }
}

static int enc_write_first_value(m3o_encoder *e, double v) { /* :112-146 */
  if (!e->int_optimized) {
    write_full_float(&e->fl, &e->os, f64_bits(v));
    return M3O_OK;
  }
  double val;
  int mult, is_float;
  int err = m3o_convert_to_int_float(v, 0, &val, &mult, &is_float);
  if (err) return err;
  if (is_float) {
    m3o_ostream_write_bit(&e->os, OPC_FLOAT_MODE);
    write_full_float(&e->fl, &e->os, f64_bits(v));
    e->is_float = 1;
    e->max_mult = (uint8_t)mult;
    return M3O_OK;
  }
  m3o_ostream_write_bit(&e->os, OPC_INT_MODE);
  e->int_val = val;
  int neg_diff = 1;
  if (val < 0) {
    neg_diff = 0;
    val = -1 * val;
  }
  uint64_t val_bits = (uint64_t)go_f64_to_i64(val);
  uint8_t ns = (uint8_t)m3o_num_sig(val_bits);
  enc_write_int_sig_mult(e, ns, (uint8_t)mult, 0);
  write_int_val_diff(&e->sig, &e->os, val_bits, neg_diff);
  return M3O_OK;
}

static void enc_write_float_val(m3o_encoder *e, uint64_t val, uint8_t mult) { /* :176-198 */
  if (!e->is_float) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE);
    m3o_ostream_write_bit(&e->os, OPC_NO_REPEAT);
    m3o_ostream_write_bit(&e->os, OPC_FLOAT_MODE);
    write_full_float(&e->fl, &e->os, val);
    e->is_float = 1;
    e->max_mult = mult;
    return;
  }
  if (val == e->fl.prev_bits) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE);
    m3o_ostream_write_bit(&e->os, OPC_REPEAT);
    return;
  }
  m3o_ostream_write_bit(&e->os, OPC_NO_UPDATE);
  write_next_float(&e->fl, &e->os, val);
}

static void enc_write_int_val(m3o_encoder *e, double val, uint8_t mult, int is_float,
                              double val_diff) { /* :201-231 */
  if (val_diff == 0 && is_float == e->is_float && mult == e->max_mult) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE);
    m3o_ostream_write_bit(&e->os, OPC_REPEAT);
    return;
  }
  int neg = 0;
  if (val_diff < 0) {
    neg = 1;
    val_diff = -1 * val_diff;
  }
  uint64_t diff_bits = (uint64_t)go_f64_to_i64(val_diff);
  uint8_t ns = (uint8_t)m3o_num_sig(diff_bits);
  uint8_t new_sig = track_new_sig(&e->sig, ns);
  int float_changed = is_float != e->is_float;
  if (mult > e->max_mult || e->sig.num_sig != new_sig || float_changed) {
    m3o_ostream_write_bit(&e->os, OPC_UPDATE);
    m3o_ostream_write_bit(&e->os, OPC_NO_REPEAT);
    m3o_ostream_write_bit(&e->os, OPC_INT_MODE);
    enc_write_int_sig_mult(e, new_sig, mult, float_changed);
    write_int_val_diff(&e->sig, &e->os, diff_bits, neg);
    e->is_float = 0;
  } else {
    m3o_ostream_write_bit(&e->os, OPC_NO_UPDATE);
    write_int_val_diff(&e->sig, &e->os, diff_bits, neg);
  }
  e->int_val = val;
}

static int enc_write_next_value(m3o_encoder *e, double v) { /* :148-172 */
  if (!e->int_optimized) {
    write_next_float(&e->fl, &e->os, f64_bits(v));
    return M3O_OK;
  }
  double val;
  int mult, is_float;
  int err = m3o_convert_to_int_float(v, e->max_mult, &val, &mult, &is_float);
  if (err) return err;
  double val_diff = 0;
  if (!is_float) val_diff = e->int_val - val;
  if (is_float || val_diff >= k_max_int || val_diff <= k_min_int) {
    enc_write_float_val(e, f64_bits(val), (uint8_t)mult);
    return M3O_OK;
  }
  enc_write_int_val(e, val, (uint8_t)mult, is_float, val_diff);
  return M3O_OK;
}

int m3o_encoder_encode(m3o_encoder *e, int64_t ts_ns, double value, int unit, const uint8_t *ann,
                       size_t ann_len) { /* :90-110 */
  if (e->closed) return M3O_ERR_ENCODER_CLOSED;
  int err = ts_write_time(&e->ts, &e->os, ts_ns, ann, ann_len, unit);
  if (err) return err;
  if (e->num_encoded == 0)
    err = enc_write_first_value(e, value);
  else
    err = enc_write_next_value(e, value);
  if (!err) e->num_encoded++;
  return err;
}

int m3o_encoder_num_encoded(const m3o_encoder *e) { return (int)e->num_encoded; } /* :299-302 */

int m3o_encoder_last_encoded(const m3o_encoder *e, int64_t *ts_ns, double *value) { /* :305-319 */
  if (e->num_encoded == 0) return M3O_ERR_NO_DATAPOINTS;
  *ts_ns = e->ts.prev_time;
  if (e->is_float)
    *value = f64_from_bits(e->fl.prev_bits);
  else
    *value = e->int_val;
  return M3O_OK;
}
int m3o_encoder_last_annotation_checksum(const m3o_encoder *e, uint64_t *sum) { /* :321-327 */
  if (e->num_encoded == 0) return M3O_ERR_NO_DATAPOINTS;
  *sum = e->ts.prev_ann_checksum;
  return M3O_OK;
}
int m3o_encoder_empty(const m3o_encoder *e) { return e->os.len == 0; } /* :330-332 */

size_t m3o_encoder_len(const m3o_encoder *e) { /* :336-354 */
  if (e->os.len == 0) return 0;
  uint8_t tail[4];
  size_t tl = make_tail(e->os.buf[e->os.len - 1], e->os.pos, tail);
  return e->os.len - 1 + tl;
}
size_t m3o_encoder_stream(const m3o_encoder *e, uint8_t *out, size_t cap) { /* :282-297,394-429 */
  if (e->os.len == 0) return 0;
  uint8_t tail[4];
  size_t tl = make_tail(e->os.buf[e->os.len - 1], e->os.pos, tail);
  size_t total = e->os.len - 1 + tl;
  if (total > cap) return total;
  memcpy(out, e->os.buf, e->os.len - 1);
  memcpy(out + e->os.len - 1, tail, tl);
  return total;
}
size_t m3o_encoder_raw(const m3o_encoder *e, const uint8_t **data, int *pos) {
  return m3o_ostream_raw(&e->os, data, pos);
}
void m3o_encoder_close(m3o_encoder *e) { /* :357-370 */
  if (e->closed) return;
  e->closed = 1;
  m3o_ostream_reset(&e->os);
}

/* ------------------------------------------------------------------------- */
/* m3tsz/timestamp_iterator.go                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t prev_time, prev_time_delta;
  uint8_t *prev_ant;
  size_t prev_ant_len, prev_ant_cap;
  int has_ant;
  int time_unit, default_unit;
  const time_scheme *scheme;
  int unit_changed, done, skip_markers;
} ts_iter;

static void ts_iter_init(ts_iter *t, int default_unit, int skip_markers) { /* :66-77 */
  uint8_t *buf = t->prev_ant;
  size_t cap = t->prev_ant_cap;
  memset(t, 0, sizeof(*t));
  t->prev_ant = buf;
  t->prev_ant_cap = cap;
  t->default_unit = default_unit;
  t->skip_markers = skip_markers;
}

static int ts_read_delta_of_delta(ts_iter *t, m3o_istream *is, int64_t *dod);
static int ts_read_marker_or_dod(ts_iter *t, m3o_istream *is, int64_t *dod);

/* Go binary.ReadUvarint/ReadVarint over IStream.ReadByte (istream.go:61-64) */
static int ts_read_varint(m3o_istream *is, int64_t *out) {
  uint64_t x = 0;
  unsigned s = 0;
  int err = M3O_ERR_VARINT_OVERFLOW; /* loop exhausted: errOverflow */
  for (int i = 0; i < 10; i++) {
    uint64_t b;
    int e2 = m3o_istream_read_bits(is, 8, &b);
    if (e2) {
      err = (i > 0 && e2 == M3O_ERR_EOF) ? M3O_ERR_UNEXPECTED_EOF : e2;
      break;
    }
    b &= 0xff;
    if (b < 0x80) {
      if (i == 9 && b > 1) break; /* errOverflow */
      x |= shl64(b, s);
      err = M3O_OK;
      break;
    }
    x |= shl64(b & 0x7f, s);
    s += 7;
  }
  int64_t r = (int64_t)(x >> 1);
  if (x & 1) r = ~r;
  *out = r;
  return err;
}

static int ts_read_annotation(ts_iter *t, m3o_istream *is) { /* :328-356 */
  int64_t ant_len;
  int err = ts_read_varint(is, &ant_len);
  if (err) return err;
  ant_len = ant_len + 1;
  if (ant_len <= 0) return M3O_ERR_ANNOTATION_LEN;
  /* the reference allocates ant_len bytes then reads; bound the allocation by
   * what the stream can possibly still hold (same observable result: EOF). */
  size_t avail = (is->len - is->index) + 8;
  if ((uint64_t)ant_len > (uint64_t)avail) return M3O_ERR_EOF;
  if ((size_t)ant_len > t->prev_ant_cap) {
    t->prev_ant = (uint8_t *)realloc(t->prev_ant, (size_t)ant_len);
    t->prev_ant_cap = (size_t)ant_len;
  }
  for (int64_t i = 0; i < ant_len; i++) { /* IStream.Read, istream.go:48-58 */
    uint64_t b;
    err = m3o_istream_read_bits(is, 8, &b);
    if (err) return err;
    t->prev_ant[i] = (uint8_t)b;
  }
  t->prev_ant_len = (size_t)ant_len;
  t->has_ant = 1;
  return M3O_OK;
}

static int ts_read_time_unit(ts_iter *t, m3o_istream *is) { /* :118-134 */
  uint64_t tu_bits;
  int err = m3o_istream_read_bits(is, 8, &tu_bits);
  if (err) return err;
  int tu = (int)tu_bits;
  if (unit_is_valid(tu) && tu != t->time_unit) {
    t->unit_changed = 1;
    const time_scheme *s = scheme_for_unit(tu);
    if (s) t->scheme = s;
  }
  t->time_unit = tu;
  return M3O_OK;
}

/* tryReadMarker :175-231; returns err, *success */
static int ts_try_read_marker(ts_iter *t, m3o_istream *is, int64_t *dod, int *success) {
  const int num_bits = MARKER_OPCODE_BITS + MARKER_VALUE_BITS;
  uint64_t ov, tmp;
  *dod = 0;
  *success = 0;
  if (m3o_istream_peek_bits(is, num_bits, &ov)) return M3O_OK;
  uint64_t opcode = ov >> MARKER_VALUE_BITS;
  if (opcode != MARKER_OPCODE) return M3O_OK;
  int marker = (int)(ov & ((1u << MARKER_VALUE_BITS) - 1));
  int err;
  switch (marker) {
    case MARKER_EOS:
      err = m3o_istream_read_bits(is, num_bits, &tmp);
      if (err) return err;
      t->done = 1;
      *success = 1;
      return M3O_OK;
    case MARKER_ANNOTATION:
      err = m3o_istream_read_bits(is, num_bits, &tmp);
      if (err) return err;
      err = ts_read_annotation(t, is);
      if (err) return err;
      err = ts_read_marker_or_dod(t, is, dod);
      if (err) {
        *dod = 0;
        return err;
      }
      *success = 1;
      return M3O_OK;
    case MARKER_TIMEUNIT:
      err = m3o_istream_read_bits(is, num_bits, &tmp);
      if (err) return err;
      err = ts_read_time_unit(t, is);
      if (err) return err;
      err = ts_read_marker_or_dod(t, is, dod);
      if (err) {
        *dod = 0;
        return err;
      }
      *success = 1;
      return M3O_OK;
    default:
      return M3O_OK;
  }
}

static int ts_read_marker_or_dod(ts_iter *t, m3o_istream *is, int64_t *dod) { /* :233-244 */
  if (!t->skip_markers) {
    int success;
    int err = ts_try_read_marker(t, is, dod, &success);
    if (success || err || t->done) return err;
  }
  return ts_read_delta_of_delta(t, is, dod);
}

static int ts_read_full_timestamp(ts_iter *t, m3o_istream *is, int64_t *dod) { /* :307-326 */
  const time_scheme *s = scheme_for_unit(t->time_unit);
  *dod = 0;
  if (!s) return M3O_ERR_NO_TIME_SCHEME;
  t->scheme = s;
  uint64_t bits;
  int err = m3o_istream_read_bits(is, 64, &bits);
  if (err) return err;
  *dod = m3o_sign_extend(bits, 64);
  return M3O_OK;
}

static int ts_read_delta_of_delta(ts_iter *t, m3o_istream *is, int64_t *dod) { /* :246-305 */
  *dod = 0;
  if (t->unit_changed) return ts_read_full_timestamp(t, is, dod);
  if (!t->scheme) return M3O_ERR_NO_TIME_SCHEME;
  uint64_t cb;
  int err = m3o_istream_read_bits(is, 1, &cb);
  if (err) return err;
  const time_scheme *tes = t->scheme;
  if (cb == tes->zero.opcode) return M3O_OK;
  for (int i = 0; i < tes->n_buckets; i++) {
    uint64_t ncb;
    if (m3o_istream_read_bits(is, 1, &ncb)) return M3O_OK; /* swallowed, :272-274 */
    cb = (cb << 1) | ncb;
    if (cb == tes->buckets[i].opcode) {
      uint64_t bits;
      err = m3o_istream_read_bits(is, tes->buckets[i].n_value_bits, &bits);
      if (err) return err;
      int64_t d = m3o_sign_extend(bits, tes->buckets[i].n_value_bits);
      int64_t u;
      if (unit_value(t->time_unit, &u)) return M3O_OK; /* swallowed, :284-287 */
      *dod = (int64_t)((uint64_t)d * (uint64_t)u); /* FromNormalizedDuration, x/time/time.go:60-62 */
      return M3O_OK;
    }
  }
  int nvb = tes->dflt.n_value_bits;
  uint64_t bits;
  err = m3o_istream_read_bits(is, nvb, &bits);
  if (err) return err;
  int64_t d = m3o_sign_extend(bits, nvb);
  int64_t u;
  if (unit_value(t->time_unit, &u)) return M3O_OK; /* swallowed, :299-302 */
  *dod = (int64_t)((uint64_t)d * (uint64_t)u);
  return M3O_OK;
}

static int ts_read_next_timestamp(ts_iter *t, m3o_istream *is) { /* :164-173 */
  int64_t dod;
  int err = ts_read_marker_or_dod(t, is, &dod);
  if (err) return err;
  t->prev_time_delta = (int64_t)((uint64_t)t->prev_time_delta + (uint64_t)dod);
  t->prev_time = (int64_t)((uint64_t)t->prev_time + (uint64_t)t->prev_time_delta);
  return M3O_OK;
}

static int ts_read_first_timestamp(ts_iter *t, m3o_istream *is) { /* :136-162 */
  uint64_t nt_bits;
  int err = m3o_istream_read_bits(is, 64, &nt_bits);
  if (err) return err;
  int64_t nt = (int64_t)nt_bits;
  if (t->time_unit == M3O_UNIT_NONE) t->time_unit = m3o_initial_time_unit(nt, t->default_unit);
  const time_scheme *s = scheme_for_unit(t->time_unit);
  if (s) t->scheme = s;
  err = ts_read_next_timestamp(t, is);
  if (err) return err;
  t->prev_time = (int64_t)((uint64_t)nt + (uint64_t)t->prev_time_delta);
  return M3O_OK;
}

/* ReadTimestamp :80-113 */
static int ts_read_timestamp(ts_iter *t, m3o_istream *is, int *first, int *done) {
  t->has_ant = 0;
  t->prev_ant_len = 0;
  *first = 0;
  int err;
  if (t->prev_time != 0) {
    int64_t dod;
    err = ts_read_marker_or_dod(t, is, &dod);
    if (!err) {
      t->prev_time_delta = (int64_t)((uint64_t)t->prev_time_delta + (uint64_t)dod);
      t->prev_time = (int64_t)((uint64_t)t->prev_time + (uint64_t)t->prev_time_delta);
    }
  } else {
    *first = 1;
    err = ts_read_first_timestamp(t, is);
  }
  if (err) {
    *first = 0;
    *done = 0;
    return err;
  }
  if (t->unit_changed) {
    t->prev_time_delta = 0;
    t->unit_changed = 0;
  }
  *done = t->done;
  return M3O_OK;
}

/* ------------------------------------------------------------------------- */
/* m3tsz/iterator.go                                                           */
/* ------------------------------------------------------------------------- */
struct m3o_iter {
  m3o_istream is;
  int err;
  double int_val;
  ts_iter ts;
  float_state fl;
  uint8_t mult, sig;
  int64_t cur_ts;
  double cur_val;
  int int_optimized, is_float, closed;
  int default_unit;
};

m3o_iter *m3o_iter_new(const uint8_t *data, size_t len, int int_optimized, int default_unit) { /* :67-78 */
  m3o_iter *it = (m3o_iter *)calloc(1, sizeof(*it));
  is_reset(&it->is, data, len);
  it->default_unit = default_unit;
  ts_iter_init(&it->ts, default_unit, 0);
  it->int_optimized = int_optimized;
  return it;
}
void m3o_iter_free(m3o_iter *it) {
  if (!it) return;
  free(it->ts.prev_ant);
  free(it);
}
void m3o_iter_reset(m3o_iter *it, const uint8_t *data, size_t len) { /* :253-263 */
  is_reset(&it->is, data, len);
  ts_iter_init(&it->ts, it->default_unit, it->ts.skip_markers);
  it->err = M3O_OK;
  it->is_float = 0;
  it->int_val = 0.0;
  it->mult = 0;
  it->sig = 0;
  it->closed = 0;
  memset(&it->fl, 0, sizeof(it->fl)); /* NB: the reference leaves floatIter stale; it is
                                         always overwritten by the first value read */
}

static uint64_t it_read_bits(m3o_iter *it, int n) { /* :221-224: overwrites it.err every call */
  uint64_t res;
  it->err = m3o_istream_read_bits(&it->is, n, &res);
  return res;
}
static void it_read_int_sig_mult(m3o_iter *it) { /* :178-193 */
  if (it_read_bits(it, 1) == OPC_UPDATE_SIG) {
    if (it_read_bits(it, 1) == OPC_ZERO_SIG) {
      it->sig = 0;
    } else {
      it->sig = (uint8_t)(it_read_bits(it, NUM_SIG_BITS) + 1);
    }
  }
  if (it_read_bits(it, 1) == OPC_UPDATE_MULT) {
    it->mult = (uint8_t)it_read_bits(it, NUM_MULT_BITS);
    if (it->mult > MAX_MULT) it->err = M3O_ERR_INVALID_MULT;
  }
}
static void it_read_int_val_diff_slow(m3o_iter *it) { /* :212-219 */
  double sign = -1.0;
  if (it_read_bits(it, 1) == OPC_NEGATIVE) sign = 1.0;
  it->int_val += sign * (double)it_read_bits(it, it->sig);
}
static void it_read_int_val_diff(m3o_iter *it) { /* :195-210 */
  if (it->sig == 64) {
    it_read_int_val_diff_slow(it);
    return;
  }
  uint64_t bits = it_read_bits(it, it->sig + 1);
  double sign = -1.0;
  if ((bits >> it->sig) == OPC_NEGATIVE) {
    sign = 1.0;
    bits ^= ((uint64_t)1) << it->sig;
  }
  it->int_val += sign * (double)bits;
}
static void it_read_first_value(m3o_iter *it) { /* :108-126 */
  if (!it->int_optimized) {
    int err = read_full_float(&it->fl, &it->is);
    if (err) it->err = err;
    return;
  }
  if (it_read_bits(it, 1) == OPC_FLOAT_MODE) {
    int err = read_full_float(&it->fl, &it->is);
    if (err) it->err = err;
    it->is_float = 1;
    return;
  }
  it_read_int_sig_mult(it);
  it_read_int_val_diff(it);
}
void m3o_iter_read_next_value(m3o_iter *it) { /* :128-176 */
  if (!it->int_optimized) {
    int err = read_next_float(&it->fl, &it->is);
    if (err) it->err = err;
    return;
  }
  if (it_read_bits(it, 1) == OPC_UPDATE) {
    if (it_read_bits(it, 1) == OPC_REPEAT) return;
    if (it_read_bits(it, 1) == OPC_FLOAT_MODE) {
      int err = read_full_float(&it->fl, &it->is);
      if (err) it->err = err;
      it->is_float = 1;
      return;
    }
    it_read_int_sig_mult(it);
    it_read_int_val_diff(it);
    it->is_float = 0;
    return;
  }
  if (it->is_float) {
    int err = read_next_float(&it->fl, &it->is);
    if (err) it->err = err;
    return;
  }
  it_read_int_val_diff(it);
}

static inline int it_has_next(const m3o_iter *it) { return it->err == M3O_OK && !it->ts.done; } /* :248-250 */

int m3o_iter_next(m3o_iter *it) { /* :81-106 */
  if (!it_has_next(it)) return 0;
  int first, done;
  int err = ts_read_timestamp(&it->ts, &it->is, &first, &done);
  if (err || done) {
    it->err = err;
    return 0;
  }
  if (!first)
    m3o_iter_read_next_value(it);
  else
    it_read_first_value(it);
  it->cur_ts = it->ts.prev_time;
  if (!it->int_optimized || it->is_float) {
    it->cur_val = f64_from_bits(it->fl.prev_bits);
  } else {
    if (it->mult > MAX_MULT) {
      /* the reference indexes multipliers[mult] out of range here and panics
       * (m3tsz.go:126); the oracle reports a sticky error instead. */
      it->err = M3O_ERR_INVALID_MULT;
      return 0;
    }
    it->cur_val = m3o_convert_from_int_float(it->int_val, it->mult);
  }
  return it_has_next(it);
}
void m3o_iter_current(const m3o_iter *it, int64_t *ts_ns, double *value, int *unit,
                      const uint8_t **ann, size_t *ann_len) { /* :229-231 */
  if (ts_ns) *ts_ns = it->cur_ts;
  if (value) *value = it->cur_val;
  if (unit) *unit = it->ts.time_unit;
  if (ann) *ann = it->ts.has_ant ? it->ts.prev_ant : NULL;
  if (ann_len) *ann_len = it->ts.has_ant ? it->ts.prev_ant_len : 0;
}
int m3o_iter_err(const m3o_iter *it) { return it->err; }
int m3o_iter_done(const m3o_iter *it) { return it->ts.done; }

/* white-box hooks for the field-level goldens (iterator_test.go:44-179) */
void m3o_iter_set_float_state(m3o_iter *it, uint64_t prev_bits, uint64_t prev_xor) {
  it->fl.prev_bits = prev_bits;
  it->fl.prev_xor = prev_xor;
}
void m3o_iter_get_float_state(const m3o_iter *it, uint64_t *prev_bits, uint64_t *prev_xor) {
  *prev_bits = it->fl.prev_bits;
  *prev_xor = it->fl.prev_xor;
}
void m3o_iter_set_ts_state(m3o_iter *it, int unit, int64_t prev_delta) {
  it->ts.time_unit = unit;
  it->ts.prev_time_delta = prev_delta;
  it->ts.scheme = scheme_for_unit(unit);
}
int m3o_iter_read_next_timestamp(m3o_iter *it) { return ts_read_next_timestamp(&it->ts, &it->is); }
int m3o_iter_read_first_timestamp(m3o_iter *it) { return ts_read_first_timestamp(&it->ts, &it->is); }
int64_t m3o_iter_prev_time_delta(const m3o_iter *it) { return it->ts.prev_time_delta; }
int m3o_iter_read_annotation(m3o_iter *it, const uint8_t **ann, size_t *ann_len) {
  int err = ts_read_annotation(&it->ts, &it->is);
  *ann = it->ts.prev_ant;
  *ann_len = it->ts.prev_ant_len;
  return err;
}
int m3o_iter_read_time_unit(m3o_iter *it, int *unit, int *changed) {
  int err = ts_read_time_unit(&it->ts, &it->is);
  *unit = it->ts.time_unit;
  *changed = it->ts.unit_changed;
  return err;
}

/* ------------------------------------------------------------------------- */
/* whole-series helpers                                                        */
/* ------------------------------------------------------------------------- */
int64_t m3o_encode_series(const int64_t *ts, const double *vals, size_t n, int64_t start_ns,
                          int unit, int int_optimized, int default_unit, uint8_t *out,
                          size_t cap) {
  m3o_encoder *e = m3o_encoder_new(0, int_optimized, default_unit);
  m3o_encoder_reset(e, start_ns); /* pool-allocated encoders: NewEncoder(0) then Reset(start), server.go:1786 */
  for (size_t i = 0; i < n; i++) {
    int err = m3o_encoder_encode(e, ts[i], vals[i], unit, NULL, 0);
    if (err) {
      m3o_encoder_free(e);
      return -(int64_t)err;
    }
  }
  size_t len = m3o_encoder_stream(e, out, cap);
  m3o_encoder_free(e);
  if (len > cap) return -1000;
  return (int64_t)len;
}

int64_t m3o_decode_series(const uint8_t *data, size_t len, int int_optimized, int default_unit,
                          int64_t *ts_out, double *val_out, size_t cap, int *err) {
  m3o_iter it;
  memset(&it, 0, sizeof(it));
  is_reset(&it.is, data, len);
  it.default_unit = default_unit;
  ts_iter_init(&it.ts, default_unit, 0);
  it.int_optimized = int_optimized;
  size_t n = 0;
  while (m3o_iter_next(&it)) {
    if (n < cap) {
      ts_out[n] = it.cur_ts;
      val_out[n] = it.cur_val;
    }
    n++;
  }
  if (err) *err = it.err;
  free(it.ts.prev_ant);
  return (int64_t)n;
}

/* ------------------------------------------------------------------------- */
/* multi-threaded batch drivers (CPU baseline)                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int kind; /* 0 decode, 1 encode */
  size_t lo, hi;
  /* decode */
  const uint8_t *streams;
  const uint64_t *off;
  int64_t *ts_out;
  double *val_out;
  size_t cap;
  uint32_t *n_points;
  /* encode */
  const int64_t *ts;
  const double *vals;
  size_t n_per;
  const int64_t *start_ns;
  int unit;
  uint8_t *out;
  size_t out_stride;
  uint64_t *out_len;
  /* common */
  int int_optimized, default_unit;
  int32_t *status;
} batch_task;

static void *batch_worker(void *arg) {
  batch_task *t = (batch_task *)arg;
  /* one pooled encoder / iterator per worker, Reset between series (the reference
   * pools them too: encoder_pool.go / iterator_pool.go) */
  m3o_encoder *enc = NULL;
  m3o_iter *it = NULL;
  if (t->kind == 1) enc = m3o_encoder_new(0, t->int_optimized, t->default_unit);
  if (t->kind == 0) it = m3o_iter_new(NULL, 0, t->int_optimized, t->default_unit);
  for (size_t s = t->lo; s < t->hi; s++) {
    if (t->kind == 0) {
      m3o_iter_reset(it, t->streams + t->off[s], (size_t)(t->off[s + 1] - t->off[s]));
      int64_t *ts_out = t->ts_out + s * t->cap;
      double *val_out = t->val_out + s * t->cap;
      size_t n = 0;
      while (m3o_iter_next(it)) {
        if (n < t->cap) {
          ts_out[n] = it->cur_ts;
          val_out[n] = it->cur_val;
        }
        n++;
      }
      if (t->n_points) t->n_points[s] = (uint32_t)n;
      if (t->status) t->status[s] = it->err;
    } else {
      m3o_encoder_reset(enc, t->start_ns[s]);
      const int64_t *ts = t->ts + s * t->n_per;
      const double *vals = t->vals + s * t->n_per;
      int err = 0;
      for (size_t i = 0; i < t->n_per; i++) {
        err = m3o_encoder_encode(enc, ts[i], vals[i], t->unit, NULL, 0);
        if (err) break;
      }
      size_t len = err ? 0 : m3o_encoder_stream(enc, t->out + s * t->out_stride, t->out_stride);
      if (!err && len > t->out_stride) {
        err = 1000;
        len = 0;
      }
      if (t->out_len) t->out_len[s] = (uint64_t)len;
      if (t->status) t->status[s] = err;
    }
  }
  m3o_encoder_free(enc);
  m3o_iter_free(it);
  return NULL;
}

static int run_batch(batch_task *proto, size_t n_series, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n_series) n_threads = n_series ? (int)n_series : 1;
  pthread_once(&g_schemes_once, init_schemes);
  batch_task *tasks = (batch_task *)malloc(sizeof(batch_task) * (size_t)n_threads);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  size_t per = (n_series + (size_t)n_threads - 1) / (size_t)n_threads;
  for (int i = 0; i < n_threads; i++) {
    tasks[i] = *proto;
    tasks[i].lo = (size_t)i * per;
    tasks[i].hi = tasks[i].lo + per;
    if (tasks[i].lo > n_series) tasks[i].lo = n_series;
    if (tasks[i].hi > n_series) tasks[i].hi = n_series;
  }
  if (n_threads == 1) {
    batch_worker(&tasks[0]);
  } else {
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &tasks[i]);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  }
  free(tasks);
  free(th);
  return 0;
}

int m3o_decode_batch(const uint8_t *streams, const uint64_t *off, size_t n_series,
                     int int_optimized, int default_unit, int64_t *ts_out, double *val_out,
                     size_t cap, uint32_t *n_points, int32_t *status, int n_threads) {
  batch_task t;
  memset(&t, 0, sizeof(t));
  t.kind = 0;
  t.streams = streams;
  t.off = off;
  t.ts_out = ts_out;
  t.val_out = val_out;
  t.cap = cap;
  t.n_points = n_points;
  t.status = status;
  t.int_optimized = int_optimized;
  t.default_unit = default_unit;
  return run_batch(&t, n_series, n_threads);
}

int m3o_encode_batch(const int64_t *ts, const double *vals, size_t n_series, size_t n_per,
                     const int64_t *start_ns, int unit, int int_optimized, int default_unit,
                     uint8_t *out, size_t out_stride, uint64_t *out_len, int32_t *status,
                     int n_threads) {
  batch_task t;
  memset(&t, 0, sizeof(t));
  t.kind = 1;
  t.ts = ts;
  t.vals = vals;
  t.n_per = n_per;
  t.start_ns = start_ns;
  t.unit = unit;
  t.out = out;
  t.out_stride = out_stride;
  t.out_len = out_len;
  t.status = status;
  t.int_optimized = int_optimized;
  t.default_unit = default_unit;
  return run_batch(&t, n_series, n_threads);
}

/* ------------------------------------------------------------------------- */
/* downsample oracle: aggregation.Gauge.updateTotals,                          */
/* /root/reference/src/aggregator/aggregation/gauge.go:73-106; window start =  */
/* timestamp.Truncate(resolution), aggregator/generic_elem.go:220              */
/* ------------------------------------------------------------------------- */
void m3o_downsample_series(const int64_t *ts, const double *vals, size_t n, int64_t range_start_ns,
                           int64_t window_ns, size_t n_windows, double *sum, int64_t *count,
                           double *min, double *max, double *last) {
  const double nan = f64_from_bits(0x7ff8000000000001ULL); /* Go math.NaN() */
  int64_t *last_at = (int64_t *)malloc(sizeof(int64_t) * (n_windows ? n_windows : 1));
  uint8_t *has_last = (uint8_t *)calloc(n_windows ? n_windows : 1, 1);
  for (size_t w = 0; w < n_windows; w++) {
    sum[w] = 0;
    count[w] = 0;
    min[w] = nan;
    max[w] = nan;
    last[w] = 0;
  }
  for (size_t i = 0; i < n; i++) {
    int64_t rel = ts[i] - range_start_ns;
    int64_t w = rel / window_ns;
    if (rel % window_ns != 0 && rel < 0) w--; /* floor: Truncate rounds down */
    if (w < 0 || (uint64_t)w >= (uint64_t)n_windows) continue;
    double v = vals[i];
    if (!has_last[w] || ts[i] > last_at[w]) { /* lastAt.IsZero() || timestamp.After(lastAt) */
      has_last[w] = 1;
      last_at[w] = ts[i];
      last[w] = v;
    }
    count[w]++;
    if (is_nan(v)) continue;
    sum[w] += v;
    if (is_nan(max[w]) || max[w] < v) max[w] = v;
    if (is_nan(min[w]) || min[w] > v) min[w] = v;
  }
  free(last_at);
  free(has_last);
}

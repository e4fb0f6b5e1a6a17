"""Host-side mirror of the reference's per-series codec interface, routed through
the GPU batch kernels (there is no CPU codec in this package).

Reference interface mirrored (paths under /root/reference/src/dbnode/encoding):
  encoding.Encoder        types.go:39-91    -> Encoder
  encoding.ReaderIterator types.go:180-203  -> ReaderIterator
  encoding.Decoder        types.go:342-345  -> Decoder
  m3tsz.NewEncoder / NewReaderIterator / NewDecoder  m3tsz/encoder.go:64-85,
  m3tsz/iterator.go:67-78, m3tsz/decoder.go:33-38

Every method is ONE call into the C ABI's streaming handles (m3tsz_encoder_* /
m3tsz_iter_* / *_pool_*, include/m3tsz_b200.h) - the same entry points the cgo shim of
INTEGRATION.md binds - so these classes are the executable check of that boundary.
The handles buffer on the host: Encoder.encode() validates and buffers, the bitstream is
produced by one GPU launch when stream()/len()/discard()/last_encoded() is called;
ReaderIterator decodes its whole stream with one GPU launch on the first next().  Real
callers should batch many series per launch with m3_b200.codec.BatchCodec - a cgo call
or a kernel launch per datapoint costs more than the reference spends encoding it
(SURVEY.md §7).
"""
import ctypes as C
import struct
from typing import Optional

from . import capi

_UNIT_NS = [0, 10 ** 9, 10 ** 6, 10 ** 3, 1, 60 * 10 ** 9, 3600 * 10 ** 9, 86400 * 10 ** 9,
            365 * 86400 * 10 ** 9]

_ctxs = {}


def _ctx(device=0) -> capi.Context:
    if device not in _ctxs:
        _ctxs[device] = capi.Context(device)
    return _ctxs[device]


# XXH64 (seed 0) == cespare/xxhash/v2 Sum64, used only for LastAnnotationChecksum()
_P1, _P2, _P3, _P4, _P5 = (11400714785074694791, 14029467366897019727, 1609587929392839161,
                           9650029242287828579, 2870177450012600261)
_M = (1 << 64) - 1


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & _M


def _round(acc, inp):
    acc = (acc + inp * _P2) & _M
    return (_rotl(acc, 31) * _P1) & _M


def xxh64(data: bytes) -> int:
    n = len(data)
    p = 0
    if n >= 32:
        v1, v2, v3, v4 = (_P1 + _P2) & _M, _P2, 0, (-_P1) & _M
        while p <= n - 32:
            a, b, c, d = struct.unpack_from("<QQQQ", data, p)
            v1, v2, v3, v4 = _round(v1, a), _round(v2, b), _round(v3, c), _round(v4, d)
            p += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & _M
        for v in (v1, v2, v3, v4):
            h = ((h ^ _round(0, v)) * _P1 + _P4) & _M
    else:
        h = _P5
    h = (h + n) & _M
    while p + 8 <= n:
        (k,) = struct.unpack_from("<Q", data, p)
        h = (_rotl(h ^ _round(0, k), 27) * _P1 + _P4) & _M
        p += 8
    if p + 4 <= n:
        (k,) = struct.unpack_from("<I", data, p)
        h = (_rotl(h ^ (k * _P1) & _M, 23) * _P2 + _P3) & _M
        p += 4
    while p < n:
        h = (_rotl(h ^ (data[p] * _P5) & _M, 11) * _P1) & _M
        p += 1
    h ^= h >> 33
    h = (h * _P2) & _M
    h ^= h >> 29
    h = (h * _P3) & _M
    h ^= h >> 32
    return h


_EMPTY_ANN_CHECKSUM = xxh64(b"")


def initial_time_unit(start_ns: int, unit: int) -> int:
    """m3tsz/timestamp_encoder.go:248-259"""
    if unit < 1 or unit > 8:
        return 0
    return unit if start_ns % _UNIT_NS[unit] == 0 else 0


def _trunc_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _check(rc, what=""):
    if rc != capi.OK:
        raise capi.M3tszError(rc, what)


class Encoder:
    """m3tsz encoder (m3tsz/encoder.go) over m3tsz_encoder_*.  `start_ns` is the encoder
    start (block start), not the first datapoint's time."""

    def __init__(self, start_ns: int, int_optimized: bool = True, default_unit: int = capi.UNIT_S,
                 device: int = 0, _handle=None):
        self._ctx = _ctx(device)
        self._opts = capi.Options(int(bool(int_optimized)), int(default_unit))
        self._owned = _handle is None
        if _handle is None:
            self._h = C.c_void_p()
            _check(capi.lib().m3tsz_encoder_create(self._ctx.handle, C.byref(self._opts), int(start_ns),
                                                   C.byref(self._h)), "m3tsz_encoder_create")
        else:
            self._h = _handle
            _check(capi.lib().m3tsz_encoder_reset(self._h, int(start_ns), 0))

    def __del__(self):
        try:
            if self._owned and self._h:
                capi.lib().m3tsz_encoder_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # encoding.Encoder.Encode, m3tsz/encoder.go:90-110
    def encode(self, ts_ns: int, value: float, unit: int = capi.UNIT_S, annotation: bytes = b""):
        annotation = bytes(annotation or b"")
        rc = capi.lib().m3tsz_encoder_encode(self._h, int(ts_ns), float(value), int(unit),
                                             annotation if annotation else None, len(annotation))
        if rc == capi.ERR_DOD_OVERFLOW:
            # the reference formats the offending value into the message (timestamp_encoder.go:219)
            dod = int(capi.lib().m3tsz_encoder_failed_dod(self._h))
            raise capi.M3tszError(rc, "deltaOfDelta value %d %s overflows 32 bits"
                                  % (dod, "s" if unit == capi.UNIT_S else "ms"))
        _check(rc)

    def num_encoded(self) -> int:  # :299-302
        return int(capi.lib().m3tsz_encoder_num_encoded(self._h))

    def last_encoded(self):  # :305-319, with the reference's scaled-int / zero quirk
        t, v = C.c_int64(), C.c_double()
        _check(capi.lib().m3tsz_encoder_last_encoded(self._h, C.byref(t), C.byref(v)))
        return t.value, v.value

    def last_annotation_checksum(self) -> int:  # :321-327
        c = C.c_uint64()
        _check(capi.lib().m3tsz_encoder_last_annotation_checksum(self._h, C.byref(c)))
        return c.value

    def empty(self) -> bool:  # :330-332
        return bool(capi.lib().m3tsz_encoder_empty(self._h))

    def len(self) -> int:  # :336-354
        n = C.c_uint64()
        _check(capi.lib().m3tsz_encoder_len(self._h, C.byref(n)))
        return n.value

    def _segment(self, fn, *pre):
        n = C.c_uint64()
        _check(capi.lib().m3tsz_encoder_len(self._h, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        ln, tail = C.c_uint64(), C.c_uint64()
        _check(fn(self._h, *pre, buf, n.value, C.byref(ln), C.byref(tail)))
        self._tail_len = tail.value
        return buf.raw[: ln.value]

    def stream(self) -> Optional[bytes]:  # Stream(): (nil, false) when empty, :282-297
        b = self._segment(capi.lib().m3tsz_encoder_stream)
        return b if b else None

    def segment(self):
        """ts.Segment view of the stream: (head, tail) (encoder.go:394-457)."""
        b = self._segment(capi.lib().m3tsz_encoder_stream)
        return b[: len(b) - self._tail_len], b[len(b) - self._tail_len:]

    def reset(self, start_ns: int, capacity: int = 0):  # :262-264
        _check(capi.lib().m3tsz_encoder_reset(self._h, int(start_ns), int(capacity)))

    def close(self):  # :357-370
        _check(capi.lib().m3tsz_encoder_close(self._h))

    def discard(self) -> bytes:  # :374-381
        return self._segment(capi.lib().m3tsz_encoder_discard)

    def discard_reset(self, start_ns: int, capacity: int = 0) -> bytes:  # :385-392
        return self._segment(capi.lib().m3tsz_encoder_discard_reset, int(start_ns), int(capacity))


class ReaderIterator:
    """m3tsz reader iterator (m3tsz/iterator.go:67-278) over m3tsz_iter_*."""

    def __init__(self, data: Optional[bytes], int_optimized: bool = True,
                 default_unit: int = capi.UNIT_S, device: int = 0, _handle=None):
        self._ctx = _ctx(device)
        self._opts = capi.Options(int(bool(int_optimized)), int(default_unit))
        self._owned = _handle is None
        if _handle is None:
            self._h = C.c_void_p()
            _check(capi.lib().m3tsz_iter_create(self._ctx.handle, C.byref(self._opts), C.byref(self._h)),
                   "m3tsz_iter_create")
        else:
            self._h = _handle
        self.reset(data)

    def __del__(self):
        try:
            if self._owned and self._h:
                capi.lib().m3tsz_iter_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self, data: Optional[bytes]):  # :253-263
        data = bytes(data) if data is not None else b""
        self._first_ann = None
        self._seen_first = False
        _check(capi.lib().m3tsz_iter_reset(self._h, data if data else None, len(data)))

    def next(self) -> bool:  # :81-106
        ok = bool(capi.lib().m3tsz_iter_next(self._h))
        if ok and not self._seen_first:
            self._seen_first = True
            self._first_ann = self.current_full()[3] or None
        return ok

    def current_full(self):
        """Current(): (ts_ns, value, unit in force at this datapoint, annotation of this datapoint)."""
        t, v, u = C.c_int64(), C.c_double(), C.c_int32()
        p, n = C.c_void_p(), C.c_uint64()
        _check(capi.lib().m3tsz_iter_current(self._h, C.byref(t), C.byref(v), C.byref(u), C.byref(p), C.byref(n)))
        ann = C.string_at(p.value, n.value) if n.value else b""
        return t.value, v.value, u.value, ann

    def current(self):  # :229-231 -> (ts_ns, value, unit)
        return self.current_full()[:3]

    def first_annotation(self) -> Optional[bytes]:
        """SeriesIterator.FirstAnnotation(): the first datapoint's annotation (None if it has none)."""
        return self._first_ann

    def err(self) -> int:  # :234-236
        return int(capi.lib().m3tsz_iter_err(self._h))

    def close(self):  # :267-278
        _check(capi.lib().m3tsz_iter_close(self._h))


class Decoder:
    """m3tsz.NewDecoder(intOptimized, opts).Decode(reader), m3tsz/decoder.go:27-44."""

    def __init__(self, int_optimized: bool = True, default_unit: int = capi.UNIT_S, device: int = 0):
        self._int_optimized = int_optimized
        self._default_unit = default_unit
        self._device = device

    def decode(self, data: bytes) -> ReaderIterator:
        return ReaderIterator(data, self._int_optimized, self._default_unit, self._device)


class EncoderPool:
    """encoding.EncoderPool (encoder_pool.go:27-48) over m3tsz_encoder_pool_*: get() hands out a
    pooled encoder (the caller reset()s it); Encoder.close() / discard() returns it."""

    def __init__(self, size: int, int_optimized: bool = True, default_unit: int = capi.UNIT_S, device: int = 0):
        self._args = (bool(int_optimized), int(default_unit), device)
        self._ctx = _ctx(device)
        self._opts = capi.Options(int(bool(int_optimized)), int(default_unit))
        self._h = C.c_void_p()
        _check(capi.lib().m3tsz_encoder_pool_create(self._ctx.handle, C.byref(self._opts), int(size),
                                                    C.byref(self._h)))

    def get(self, start_ns: int = 0) -> Encoder:
        h = C.c_void_p()
        _check(capi.lib().m3tsz_encoder_pool_get(self._h, C.byref(h)))
        return Encoder(start_ns, *self._args, _handle=h)

    def __del__(self):
        try:
            if self._h:
                capi.lib().m3tsz_encoder_pool_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ReaderIteratorPool:
    """encoding.ReaderIteratorPool (iterator_pool.go:27-47) over m3tsz_iter_pool_*."""

    def __init__(self, size: int, int_optimized: bool = True, default_unit: int = capi.UNIT_S, device: int = 0):
        self._args = (bool(int_optimized), int(default_unit), device)
        self._ctx = _ctx(device)
        self._opts = capi.Options(int(bool(int_optimized)), int(default_unit))
        self._h = C.c_void_p()
        _check(capi.lib().m3tsz_iter_pool_create(self._ctx.handle, C.byref(self._opts), int(size),
                                                 C.byref(self._h)))

    def get(self, data: Optional[bytes] = None) -> ReaderIterator:
        h = C.c_void_p()
        _check(capi.lib().m3tsz_iter_pool_get(self._h, C.byref(h)))
        return ReaderIterator(data, *self._args, _handle=h)

    def __del__(self):
        try:
            if self._h:
                capi.lib().m3tsz_iter_pool_destroy(self._h)
                self._h = None
        except Exception:
            pass

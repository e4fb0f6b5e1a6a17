"""Host-side mirror of the reference's per-series codec interface, routed through
the GPU batch kernels (there is no CPU codec in this package).

Reference interface mirrored (paths under /root/reference/src/dbnode/encoding):
  encoding.Encoder        types.go:39-91    -> Encoder
  encoding.ReaderIterator types.go:180-203  -> ReaderIterator
  encoding.Decoder        types.go:342-345  -> Decoder
  m3tsz.NewEncoder / NewReaderIterator / NewDecoder  m3tsz/encoder.go:64-85,
  m3tsz/iterator.go:67-78, m3tsz/decoder.go:33-38

The per-datapoint methods are a buffered facade: Encoder.encode() validates and
buffers, and the bitstream is produced by one GPU launch when stream()/len()/
discard() is called; ReaderIterator decodes its whole stream with one GPU
launch on the first next().  Real callers should batch many series per launch
with m3_b200.codec.BatchCodec - a cgo call or a kernel launch per datapoint
costs more than the reference spends encoding it.
"""
import struct
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import capi
from .codec import BatchCodec

_UNIT_NS = [0, 10 ** 9, 10 ** 6, 10 ** 3, 1, 60 * 10 ** 9, 3600 * 10 ** 9, 86400 * 10 ** 9,
            365 * 86400 * 10 ** 9]

_codecs = {}


def _codec(int_optimized, default_unit, device=0) -> BatchCodec:
    key = (bool(int_optimized), int(default_unit), device)
    if key not in _codecs:
        _codecs[key] = BatchCodec(device, int_optimized, default_unit)
    return _codecs[key]


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


class Encoder:
    """m3tsz encoder facade (m3tsz/encoder.go).  `start_ns` is the encoder start
    (block start), not the first datapoint's time."""

    def __init__(self, start_ns: int, int_optimized: bool = True, default_unit: int = capi.UNIT_S,
                 device: int = 0):
        self._int_optimized = bool(int_optimized)
        self._default_unit = int(default_unit)
        self._device = device
        self._codec = _codec(int_optimized, default_unit, device)
        self._closed = False
        self._reset(start_ns)

    def _reset(self, start_ns):
        self._start = int(start_ns)
        self._ts: List[int] = []
        self._vals: List[float] = []
        self._units: List[int] = []
        self._anns: List[Tuple[int, bytes]] = []
        self._unit = initial_time_unit(self._start, self._default_unit)
        self._prev_time = self._start
        self._prev_delta = 0
        self._ann_checksum = _EMPTY_ANN_CHECKSUM
        self._cached: Optional[bytes] = None

    # encoding.Encoder.Encode, m3tsz/encoder.go:90-110
    def encode(self, ts_ns: int, value: float, unit: int = capi.UNIT_S, annotation: bytes = b""):
        if self._closed:
            raise capi.M3tszError(2, "encoder is closed")
        # host-side validation of the timestamp step (timestamp_encoder.go:104-246), so that
        # Encode() fails at the datapoint that the reference fails on
        delta = ts_ns - self._prev_time
        changed = 1 <= unit <= 8 and unit != self._unit
        if not changed:
            if not (1 <= unit <= 8):
                raise capi.M3tszError(capi.ERR_UNRECOGNIZED_UNIT)
            dod = _trunc_div(delta - self._prev_delta, _UNIT_NS[unit])
            if unit in (capi.UNIT_S, capi.UNIT_MS) and not (-2 ** 31 <= dod < 2 ** 31):
                raise capi.M3tszError(
                    capi.ERR_DOD_OVERFLOW,
                    "deltaOfDelta value %d %s overflows 32 bits" % (dod, "s" if unit == 1 else "ms"))
        annotation = bytes(annotation or b"")
        if annotation:
            cs = xxh64(annotation)
            if cs != self._ann_checksum:
                self._ann_checksum = cs
            self._anns.append((len(self._ts), annotation))
        self._prev_time = ts_ns
        if changed:
            self._unit = unit
            self._prev_delta = 0
        else:
            self._prev_delta = delta
        self._ts.append(int(ts_ns))
        self._vals.append(float(value))
        self._units.append(int(unit))
        self._cached = None

    def num_encoded(self) -> int:  # :299-302
        return len(self._ts)

    def last_encoded(self):  # :305-319 (returns the datapoint as written)
        if not self._ts:
            raise capi.M3tszError(3)
        return self._ts[-1], self._vals[-1]

    def last_annotation_checksum(self) -> int:  # :321-327
        if not self._ts:
            raise capi.M3tszError(3)
        return self._ann_checksum

    def empty(self) -> bool:  # :330-332
        return not self._ts

    def _encode_now(self) -> bytes:
        if self._cached is not None:
            return self._cached
        n = len(self._ts)
        if n == 0:
            self._cached = b""
            return self._cached
        dev = self._codec.device
        ts = torch.tensor([self._ts], dtype=torch.int64, device=dev)
        vals = torch.tensor([self._vals], dtype=torch.float64, device=dev)
        units = torch.tensor([self._units], dtype=torch.uint8, device=dev)
        start = torch.tensor([self._start], dtype=torch.int64, device=dev)
        ann = None
        ann_total = 0
        if self._anns:
            ent = np.zeros(len(self._anns), dtype=[("dp", "<u4"), ("len", "<u4"), ("off", "<u8")])
            blob = bytearray()
            for i, (dp, a) in enumerate(self._anns):
                ent[i] = (dp, len(a), len(blob))
                blob += a
            ann_total = len(blob)
            ann = (torch.tensor([0, len(self._anns)], dtype=torch.int64, device=dev),
                   torch.from_numpy(ent.view(np.uint8).reshape(-1, 16).copy()).to(dev),
                   torch.frombuffer(bytes(blob), dtype=torch.uint8).to(dev))
        stride = self._codec.encode_bound(n) + ((ann_total + 16 * len(self._anns) + 15) // 16) * 16
        res = self._codec.encode(ts, vals, start, unit=capi.UNIT_S, units=units, annotations=ann,
                                 out_stride=stride)
        st = int(res.status[0].item())
        if st != capi.OK:
            raise capi.M3tszError(st, "m3tsz_encode_batch status")
        ln = int(res.out_len[0].item())
        self._cached = bytes(res.out[0, :ln].cpu().numpy().tobytes())
        return self._cached

    def stream(self) -> Optional[bytes]:  # Stream(): (nil, false) when empty, :282-297
        b = self._encode_now()
        return b if b else None

    def len(self) -> int:  # :336-354
        return len(self._encode_now())

    def reset(self, start_ns: int, capacity: int = 0):  # :262-264
        self._closed = False
        self._reset(start_ns)

    def close(self):  # :357-370
        self._closed = True
        self._ts, self._vals, self._units, self._anns = [], [], [], []
        self._cached = None

    def discard(self) -> bytes:  # :374-381
        b = self._encode_now()
        self.close()
        return b

    def discard_reset(self, start_ns: int, capacity: int = 0) -> bytes:  # :385-392
        b = self._encode_now()
        self.reset(start_ns, capacity)
        return b


class ReaderIterator:
    """m3tsz reader iterator facade (m3tsz/iterator.go:67-278)."""

    def __init__(self, data: Optional[bytes], int_optimized: bool = True,
                 default_unit: int = capi.UNIT_S, device: int = 0, max_points: int = 4096):
        self._int_optimized = bool(int_optimized)
        self._default_unit = int(default_unit)
        self._codec = _codec(int_optimized, default_unit, device)
        self._max_points = max_points
        self.reset(data)

    def reset(self, data: Optional[bytes]):  # :253-263
        self._data = bytes(data) if data is not None else None
        self._decoded = False
        self._i = -1
        self._n = 0
        self._err = 0
        self._closed = False
        self._ts = self._vals = None
        self._unit = 0
        self._ann = None

    def _decode_now(self):
        if self._decoded:
            return
        self._decoded = True
        data = self._data or b""
        dev = self._codec.device
        padded = data + b"\0" * ((-len(data)) % 16 + 16)
        streams = torch.frombuffer(bytearray(padded), dtype=torch.uint8).to(dev)
        offsets = torch.tensor([0, len(data)], dtype=torch.int64, device=dev)
        cap = self._max_points
        while True:
            r = self._codec.decode(streams, offsets, cap, want_annotations=True)
            n = int(r.n_points[0].item()) & 0xFFFFFFFF
            st = int(r.status[0].item())
            if st == capi.ERR_CAPACITY:
                cap = max(n, cap * 2)
                continue
            break
        self._n = min(n, cap)
        self._err = st
        self._ts = r.ts[0, : self._n].cpu().numpy()
        self._vals = r.values[0, : self._n].cpu().numpy()
        self._unit = int(r.unit[0].item())
        a = r.annotations[0].cpu().numpy().tobytes()
        bit_off, length, count = struct.unpack("<QII", a)
        self._ann = (bit_off, length, count)

    def next(self) -> bool:  # :81-106
        if self._closed:
            return False
        self._decode_now()
        if self._i + 1 < self._n:
            self._i += 1
            return True
        self._i = self._n
        return False

    def current(self):  # :229-231 -> (ts_ns, value, unit)
        return int(self._ts[self._i]), float(self._vals[self._i]), self._unit

    def first_annotation(self) -> Optional[bytes]:
        """Bytes of the first annotation in the stream (None if there is none)."""
        self._decode_now()
        bit_off, length, count = self._ann
        if not count:
            return None
        data = self._data
        bits = int.from_bytes(data, "big")
        total = len(data) * 8
        out = bytearray()
        for i in range(length):
            sh = total - (bit_off + 8 * (i + 1))
            out.append((bits >> sh) & 0xFF)
        return bytes(out)

    def err(self) -> int:  # :234-236
        if self._closed:
            return 10
        self._decode_now()
        return self._err if self._i >= self._n - 1 or self._n == 0 else 0

    def close(self):  # :267-278
        self._closed = True


class Decoder:
    """m3tsz.NewDecoder(intOptimized, opts).Decode(reader), m3tsz/decoder.go:27-44."""

    def __init__(self, int_optimized: bool = True, default_unit: int = capi.UNIT_S, device: int = 0):
        self._int_optimized = int_optimized
        self._default_unit = default_unit
        self._device = device

    def decode(self, data: bytes) -> ReaderIterator:
        return ReaderIterator(data, self._int_optimized, self._default_unit, self._device)

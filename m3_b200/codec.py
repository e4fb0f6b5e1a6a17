"""Batch codec over the C ABI, with torch tensors as device memory.

torch is plumbing here (device allocations, streams, torch.distributed); every
codec operation is one call into libm3tsz_b200.so (hand-written sm_100a
kernels).  Nothing in this module computes on the CPU.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import capi


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _cuda_stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@dataclass
class DecodeResult:
    ts: torch.Tensor        # int64  [S, max_points]
    values: torch.Tensor    # float64 [S, max_points]
    n_points: torch.Tensor  # int32 (uint32 bits) [S]
    status: torch.Tensor    # int32 [S]
    unit: torch.Tensor      # uint8 [S]
    annotations: Optional[torch.Tensor]  # uint8 view of m3tsz_annotation_ref [S, 16] or None
    unit_first: Optional[torch.Tensor] = None   # uint8 [S]: unit in force at the first datapoint
    events: Optional[torch.Tensor] = None       # uint8 [E, 32] view of m3tsz_dp_event (want_events)
    event_count: Optional[torch.Tensor] = None  # int64 [1]: events produced (may exceed E)


@dataclass
class EncodeResult:
    out: torch.Tensor      # uint8 [S, out_stride] slots
    out_len: torch.Tensor  # int64 [S]
    status: torch.Tensor   # int32 [S]


@dataclass
class PackedResult:
    packed: torch.Tensor   # uint8 [capacity]
    offsets: torch.Tensor  # int64 [S] start of every stream (completion order, not monotonic)
    out_len: torch.Tensor  # int64 [S]
    status: torch.Tensor   # int32 [S]
    total: torch.Tensor    # int64 [1] bytes used


@dataclass
class DownsampleResult:
    sum: torch.Tensor    # float64 [W, S] (window-major)
    count: torch.Tensor  # int64 [W, S]
    min: torch.Tensor    # float64 [W, S]
    max: torch.Tensor    # float64 [W, S]
    n_points: torch.Tensor
    status: torch.Tensor
    last: Optional[torch.Tensor] = None     # float64 [W, S] Gauge.Last() (want_last)
    last_at: Optional[torch.Tensor] = None  # int64 [W, S] Gauge.LastAt() in ns


class BatchCodec:
    """Device-resident batch encode / decode (the drop-in for the per-series
    loops at the reference's batch sites)."""

    def __init__(self, device=0, int_optimized=True, default_unit=capi.UNIT_S):
        if not torch.cuda.is_available():
            raise capi.M3tszError(capi.ERR_NO_DEVICE, "torch sees no CUDA device")
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index)
        self.ctx = capi.Context(self.device.index)
        self.opts = capi.Options(int(bool(int_optimized)), int(default_unit))
        self.int_optimized = bool(int_optimized)
        self.default_unit = int(default_unit)

    # ------------------------------------------------------------------ decode
    def decode(self, streams: torch.Tensor, offsets: torch.Tensor, max_points: int,
               want_annotations=False, out: Optional[DecodeResult] = None,
               lengths: Optional[torch.Tensor] = None, want_events=0, point_major=False) -> DecodeResult:
        """streams: uint8 [total] on device; offsets: int64 [S+1] on device (byte offsets), or
        int64 [S] starts + `lengths` int64 [S] (streams placed anywhere, index-entry style).
        want_events = capacity of the per-datapoint unit / annotation event table (0: none).
        point_major: outputs are [max_points, S] (step-major, coalesced stores) instead of [S, max_points]."""
        assert streams.dtype == torch.uint8 and streams.is_cuda and streams.is_contiguous()
        assert offsets.dtype == torch.int64 and offsets.is_cuda and offsets.is_contiguous()
        S = offsets.numel() - 1 if lengths is None else lengths.numel()
        dev = self.device
        if lengths is not None or want_events or point_major:
            if out is None:
                shape = (max_points, S) if point_major else (S, max_points)
                out = DecodeResult(
                    ts=torch.empty(shape, dtype=torch.int64, device=dev),
                    values=torch.empty(shape, dtype=torch.float64, device=dev),
                    n_points=torch.empty(S, dtype=torch.int32, device=dev),
                    status=torch.empty(S, dtype=torch.int32, device=dev),
                    unit=torch.empty(S, dtype=torch.uint8, device=dev),
                    annotations=(torch.empty((S, 16), dtype=torch.uint8, device=dev)
                                 if want_annotations else None))
            ex = capi.DecodeExtras()
            ex.d_lengths = lengths.data_ptr() if lengths is not None else None
            ex.point_major = 1 if point_major else 0
            if want_events:
                if out.unit_first is None:
                    out.unit_first = torch.empty(S, dtype=torch.uint8, device=dev)
                if out.events is None or out.events.shape[0] < want_events:
                    out.events = torch.zeros((int(want_events), 32), dtype=torch.uint8, device=dev)
                if out.event_count is None:
                    out.event_count = torch.zeros(1, dtype=torch.int64, device=dev)
                out.event_count.zero_()
                ex.d_unit_first = out.unit_first.data_ptr()
                ex.d_events = out.events.data_ptr()
                ex.events_capacity = int(out.events.shape[0])
                ex.d_event_count = out.event_count.data_ptr()
            rc = capi.lib().m3tsz_decode_batch_ex(
                self.ctx.handle, C.byref(self.opts), _ptr(streams), streams.numel(), _ptr(offsets), S,
                _ptr(out.ts), _ptr(out.values), max_points, _ptr(out.n_points), _ptr(out.status),
                _ptr(out.unit), _ptr(out.annotations), C.byref(ex), _cuda_stream_ptr(dev))
            self.ctx.check(rc, "m3tsz_decode_batch_ex")
            return out
        if out is None:
            out = DecodeResult(
                ts=torch.empty((S, max_points), dtype=torch.int64, device=dev),
                values=torch.empty((S, max_points), dtype=torch.float64, device=dev),
                n_points=torch.empty(S, dtype=torch.int32, device=dev),
                status=torch.empty(S, dtype=torch.int32, device=dev),
                unit=torch.empty(S, dtype=torch.uint8, device=dev),
                annotations=(torch.empty((S, 16), dtype=torch.uint8, device=dev)
                             if want_annotations else None))
        rc = capi.lib().m3tsz_decode_batch(
            self.ctx.handle, C.byref(self.opts), _ptr(streams), streams.numel(), _ptr(offsets), S,
            _ptr(out.ts), _ptr(out.values), max_points, _ptr(out.n_points), _ptr(out.status),
            _ptr(out.unit), _ptr(out.annotations), _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_decode_batch")
        return out

    def decode_downsample(self, streams, offsets, range_start_ns, window_ns, n_windows,
                          out: Optional[DownsampleResult] = None, want_last=False) -> DownsampleResult:
        assert streams.dtype == torch.uint8 and streams.is_cuda
        assert offsets.dtype == torch.int64 and offsets.is_cuda
        S = offsets.numel() - 1
        dev = self.device
        if want_last:
            if out is None:
                mk = lambda dt: torch.empty((n_windows, S), dtype=dt, device=dev)
                out = DownsampleResult(
                    sum=mk(torch.float64), count=mk(torch.int64), min=mk(torch.float64),
                    max=mk(torch.float64), n_points=torch.empty(S, dtype=torch.int32, device=dev),
                    status=torch.empty(S, dtype=torch.int32, device=dev), last=mk(torch.float64),
                    last_at=mk(torch.int64))
            rc = capi.lib().m3tsz_decode_downsample_last_batch(
                self.ctx.handle, C.byref(self.opts), _ptr(streams), streams.numel(), _ptr(offsets), S,
                int(range_start_ns), int(window_ns), int(n_windows), _ptr(out.sum), _ptr(out.count),
                _ptr(out.min), _ptr(out.max), _ptr(out.last), _ptr(out.last_at), _ptr(out.n_points),
                _ptr(out.status), _cuda_stream_ptr(dev))
            self.ctx.check(rc, "m3tsz_decode_downsample_last_batch")
            return out
        if out is None:
            out = DownsampleResult(
                sum=torch.empty((n_windows, S), dtype=torch.float64, device=dev),
                count=torch.empty((n_windows, S), dtype=torch.int64, device=dev),
                min=torch.empty((n_windows, S), dtype=torch.float64, device=dev),
                max=torch.empty((n_windows, S), dtype=torch.float64, device=dev),
                n_points=torch.empty(S, dtype=torch.int32, device=dev),
                status=torch.empty(S, dtype=torch.int32, device=dev))
        rc = capi.lib().m3tsz_decode_downsample_batch(
            self.ctx.handle, C.byref(self.opts), _ptr(streams), streams.numel(), _ptr(offsets), S,
            int(range_start_ns), int(window_ns), int(n_windows), _ptr(out.sum), _ptr(out.count),
            _ptr(out.min), _ptr(out.max), _ptr(out.n_points), _ptr(out.status),
            _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_decode_downsample_batch")
        return out

    # ------------------------------------------------------------------ encode
    def encode_bound(self, n_points: int) -> int:
        return int(capi.lib().m3tsz_encode_bound(int(n_points)))

    def encode(self, ts: torch.Tensor, values: torch.Tensor, start: torch.Tensor,
               unit=capi.UNIT_S, n_points: Optional[torch.Tensor] = None,
               units: Optional[torch.Tensor] = None, annotations=None,
               out_stride: Optional[int] = None, out: Optional[EncodeResult] = None,
               point_major=False) -> EncodeResult:
        """ts int64 [S,P], values float64 [S,P], start int64 [S] (all on device).
        annotations: optional (series_off int64 [S+1], entries uint8 [E,16], bytes uint8 [B]).
        point_major: ts / values are [P,S] (step-major, what decode(point_major=True) writes)."""
        assert ts.dtype == torch.int64 and values.dtype == torch.float64
        assert ts.is_cuda and values.is_cuda and ts.is_contiguous() and values.is_contiguous()
        assert start.dtype == torch.int64 and start.is_cuda
        S, P = (ts.shape[1], ts.shape[0]) if point_major else ts.shape
        dev = self.device
        if out_stride is None:
            out_stride = self.encode_bound(P)
        if out is None:
            out = EncodeResult(
                out=torch.empty((S, out_stride), dtype=torch.uint8, device=dev),
                out_len=torch.empty(S, dtype=torch.int64, device=dev),
                status=torch.empty(S, dtype=torch.int32, device=dev))
        a_off = a_ent = a_bytes = None
        if annotations is not None:
            a_off, a_ent, a_bytes = annotations
        if point_major:
            ex = capi.EncodeExtras(None, None, 1, 0)
            rc = capi.lib().m3tsz_encode_batch_ex(
                self.ctx.handle, C.byref(self.opts), _ptr(ts), _ptr(values), S, P, _ptr(n_points),
                _ptr(start), int(unit), _ptr(units), _ptr(a_off), _ptr(a_ent), _ptr(a_bytes),
                _ptr(out.out), out.out.shape[1], _ptr(out.out_len), _ptr(out.status), C.byref(ex),
                _cuda_stream_ptr(dev))
            self.ctx.check(rc, "m3tsz_encode_batch_ex")
            return out
        rc = capi.lib().m3tsz_encode_batch(
            self.ctx.handle, C.byref(self.opts), _ptr(ts), _ptr(values), S, P, _ptr(n_points),
            _ptr(start), int(unit), _ptr(units), _ptr(a_off), _ptr(a_ent), _ptr(a_bytes),
            _ptr(out.out), out.out.shape[1], _ptr(out.out_len), _ptr(out.status),
            _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_encode_batch")
        return out

    def encode_packed(self, ts, values, start, unit=capi.UNIT_S, n_points=None, units=None,
                      annotations=None, align=64, capacity: Optional[int] = None, slot_bytes=0,
                      out: Optional["PackedResult"] = None, point_major=False) -> "PackedResult":
        """Encodes straight into one packed buffer (no slots, no compaction pass).  Streams are
        placed in completion order: stream s = packed[offsets[s] : offsets[s] + out_len[s]].
        point_major: ts / values are [P, S] (step-major)."""
        S, P = (ts.shape[1], ts.shape[0]) if point_major else ts.shape
        dev = self.device
        if out is None:
            if capacity is None:
                capacity = S * (self.encode_bound(P) if units is None else
                                int(capi.lib().m3tsz_encode_bound_units(P, 1)))
            out = PackedResult(
                packed=torch.empty(capacity, dtype=torch.uint8, device=dev),
                offsets=torch.empty(S, dtype=torch.int64, device=dev),
                out_len=torch.empty(S, dtype=torch.int64, device=dev),
                status=torch.empty(S, dtype=torch.int32, device=dev),
                total=torch.zeros(1, dtype=torch.int64, device=dev))
        a_off = a_ent = a_bytes = None
        if annotations is not None:
            a_off, a_ent, a_bytes = annotations
        ex = capi.EncodeExtras()
        ex.point_major_input = 1 if point_major else 0
        rc = capi.lib().m3tsz_encode_batch_packed_ex(
            self.ctx.handle, C.byref(self.opts), _ptr(ts), _ptr(values), S, P, _ptr(n_points),
            _ptr(start), int(unit), _ptr(units), _ptr(a_off), _ptr(a_ent), _ptr(a_bytes),
            int(slot_bytes), int(align), _ptr(out.packed), out.packed.numel(), _ptr(out.offsets),
            _ptr(out.out_len), _ptr(out.status), _ptr(out.total), C.byref(ex) if point_major else None,
            _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_encode_batch_packed_ex")
        return out

    def compact(self, enc: EncodeResult, align=64, capacity: Optional[int] = None):
        """Packs slots into (packed uint8 [total], offsets int64 [S+1])."""
        S, stride = enc.out.shape
        dev = self.device
        if capacity is None:
            capacity = int(enc.out_len.sum().item()) + S * align + 16
        packed = torch.empty(capacity, dtype=torch.uint8, device=dev)
        offsets = torch.empty(S + 1, dtype=torch.int64, device=dev)
        rc = capi.lib().m3tsz_compact_streams(
            self.ctx.handle, _ptr(enc.out), stride, _ptr(enc.out_len), S, int(align), _ptr(packed),
            capacity, _ptr(offsets), _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_compact_streams")
        return packed, offsets

    # ------------------------------------------------------------------ merge (row N1)
    def merge_series(self, ts, values, n_points, seq_status, slice_off, replica_off, series_off,
                     out_cap, start_ns=0, end_ns=0, strategy=0, point_major=False):
        """seriesIterator / multiReaderIterator semantics over decoded streams (device tensors;
        offsets int64).  Returns (ts_out [S,out_cap], val_out, n_out int32 [S], status int32 [S]);
        point_major: inputs are [cap, n_seq] and outputs [out_cap, S] (step-major)."""
        S = series_off.numel() - 1
        dev = self.device
        shape = (out_cap, S) if point_major else (S, out_cap)
        ts_out = torch.empty(shape, dtype=torch.int64, device=dev)
        val_out = torch.empty(shape, dtype=torch.float64, device=dev)
        n_out = torch.empty(S, dtype=torch.int32, device=dev)
        status = torch.empty(S, dtype=torch.int32, device=dev)
        cap, n_seq = (ts.shape[0], ts.shape[1]) if point_major else (ts.shape[1], ts.shape[0])
        rc = capi.lib().m3tsz_merge_series_batch_ex(
            self.ctx.handle, _ptr(ts), _ptr(values), cap, _ptr(n_points), _ptr(seq_status),
            _ptr(slice_off), _ptr(replica_off), _ptr(series_off), S, int(start_ns), int(end_ns),
            int(strategy), _ptr(ts_out), _ptr(val_out), out_cap, _ptr(n_out), _ptr(status), n_seq,
            1 if point_major else 0, 1 if point_major else 0, _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_merge_series_batch_ex")
        return ts_out, val_out, n_out, status

    # ------------------------------------------------------------ query-side consumers (rows N3, N4)
    def prom_convert(self, ts, values, n_points, resolution_ns=0, handle_resets=None, tolerance=0.0,
                     tolerance_until_ns=0, out_cap=None):
        """iteratorToPromResult over decoded / merged series in HBM.  Returns
        (ts_ms int64 [S,out_cap], values float64 [S,out_cap], n_out int32 [S], status int32 [S])."""
        S, cap = ts.shape
        dev = self.device
        if out_cap is None:
            out_cap = cap + 1 if handle_resets is not None else cap
        ts_out = torch.empty((S, out_cap), dtype=torch.int64, device=dev)
        val_out = torch.empty((S, out_cap), dtype=torch.float64, device=dev)
        n_out = torch.empty(S, dtype=torch.int32, device=dev)
        status = torch.empty(S, dtype=torch.int32, device=dev)
        rc = capi.lib().m3tsz_prom_convert_batch(
            self.ctx.handle, _ptr(ts), _ptr(values), cap, _ptr(n_points), S, int(resolution_ns),
            _ptr(handle_resets), float(tolerance), int(tolerance_until_ns), _ptr(ts_out), _ptr(val_out),
            out_cap, _ptr(n_out), _ptr(status), _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_prom_convert_batch")
        return ts_out, val_out, n_out, status

    def aggregate_tiles(self, streams, offsets, start_ns, step_ns, n_windows, agg_type=capi.AGG_LAST,
                        out_unit=capi.UNIT_S, lengths=None, align=64, capacity=None,
                        out: Optional["PackedResult"] = None):
        """storage.TileAggregator compute: decode -> Gauge per Step window -> re-encode (packed).
        Returns (PackedResult, n_tiles int32 [S])."""
        S = offsets.numel() - 1 if lengths is None else lengths.numel()
        dev = self.device
        if out is None:
            if capacity is None:
                capacity = S * self.encode_bound(n_windows)
            out = PackedResult(
                packed=torch.empty(capacity, dtype=torch.uint8, device=dev),
                offsets=torch.empty(S, dtype=torch.int64, device=dev),
                out_len=torch.empty(S, dtype=torch.int64, device=dev),
                status=torch.empty(S, dtype=torch.int32, device=dev),
                total=torch.zeros(1, dtype=torch.int64, device=dev))
        n_tiles = torch.empty(S, dtype=torch.int32, device=dev)
        rc = capi.lib().m3tsz_aggregate_tiles_batch(
            self.ctx.handle, C.byref(self.opts), _ptr(streams), streams.numel(), _ptr(offsets),
            _ptr(lengths), S, int(start_ns), int(step_ns), int(n_windows), int(agg_type), int(out_unit),
            int(align), _ptr(out.packed), out.packed.numel(), _ptr(out.offsets), _ptr(out.out_len),
            _ptr(out.status), _ptr(n_tiles), _ptr(out.total), _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_aggregate_tiles_batch")
        return out, n_tiles

    def segment_checksums(self, streams, offsets, lengths=None, expected=None):
        """Adler-32 of every stream (ts.Segment.CalculateChecksum, src/dbnode/ts/segment.go:60-76)
        over the decoder's CSR layout; with `lengths` (int64 [S]) stream s = offsets[s] .. +lengths[s]
        (any placement, `offsets` may then have S entries); `expected` (int32/uint32 [S], the index entries' DataChecksum) turns differences
        into status M3TSZ_ERR_CHECKSUM_MISMATCH.  Returns (checksums int32-viewed-as-uint32 [S],
        status int32 [S])."""
        S = offsets.numel() - 1 if lengths is None else lengths.numel()
        dev = self.device
        out = torch.empty(S, dtype=torch.int32, device=dev)
        status = torch.empty(S, dtype=torch.int32, device=dev)
        rc = capi.lib().m3tsz_checksum_batch(
            self.ctx.handle, _ptr(streams), streams.numel(), _ptr(offsets),
            None if lengths is None else _ptr(lengths), S, None if expected is None else _ptr(expected),
            _ptr(out), _ptr(status), _cuda_stream_ptr(dev))
        self.ctx.check(rc, "m3tsz_checksum_batch")
        return out, status

    # ------------------------------------------------------------ host-buffer calls
    def decode_host(self, h_streams, h_offsets, max_points, h_ts, h_values, h_n_points, h_status):
        """All arguments are host tensors (pinned or pageable); copies are inside the call."""
        S = h_offsets.numel() - 1
        rc = capi.lib().m3tsz_decode_batch_host(
            self.ctx.handle, C.byref(self.opts), _ptr(h_streams), h_streams.numel(), _ptr(h_offsets),
            S, _ptr(h_ts), _ptr(h_values), max_points, _ptr(h_n_points), _ptr(h_status), None, None)
        self.ctx.check(rc, "m3tsz_decode_batch_host")

    def encode_host(self, h_ts, h_values, h_start, unit, h_packed, h_offsets, h_out_len, h_status,
                    align=64):
        """Host tensors in, ONE packed stream buffer + CSR offsets (int64 [S+1]) out."""
        S, P = h_ts.shape
        rc = capi.lib().m3tsz_encode_batch_host(
            self.ctx.handle, C.byref(self.opts), _ptr(h_ts), _ptr(h_values), S, P, None,
            _ptr(h_start), int(unit), None, None, None, None, 0, int(align), _ptr(h_packed),
            h_packed.numel(), _ptr(h_offsets), _ptr(h_out_len), _ptr(h_status))
        self.ctx.check(rc, "m3tsz_encode_batch_host")

    def decode_downsample_host(self, h_streams, h_offsets, range_start_ns, window_ns, n_windows,
                               h_sum, h_count, h_min, h_max, h_n_points, h_status):
        S = h_offsets.numel() - 1
        rc = capi.lib().m3tsz_decode_downsample_batch_host(
            self.ctx.handle, C.byref(self.opts), _ptr(h_streams), h_streams.numel(), _ptr(h_offsets),
            S, int(range_start_ns), int(window_ns), int(n_windows), _ptr(h_sum), _ptr(h_count),
            _ptr(h_min), _ptr(h_max), _ptr(h_n_points), _ptr(h_status))
        self.ctx.check(rc, "m3tsz_decode_downsample_batch_host")

    def fetch_host(self, h_streams, h_offsets, h_slice_off, h_replica_off, h_series_off, max_points, out_cap,
                   h_ts_out, h_val_out, h_n_out, h_status, start_ns=0, end_ns=0, strategy=0):
        """Host tensors in and out: compressed replica streams up, merged series down."""
        n_seq = h_offsets.numel() - 1
        S = h_series_off.numel() - 1
        rc = capi.lib().m3tsz_fetch_batch_host(
            self.ctx.handle, C.byref(self.opts), _ptr(h_streams), h_streams.numel(), _ptr(h_offsets), n_seq,
            _ptr(h_slice_off), _ptr(h_replica_off), _ptr(h_series_off), S, int(max_points), int(start_ns),
            int(end_ns), int(strategy), _ptr(h_ts_out), _ptr(h_val_out), int(out_cap), _ptr(h_n_out),
            _ptr(h_status))
        self.ctx.check(rc, "m3tsz_fetch_batch_host")

    def launch_count(self):
        return self.ctx.launch_count()

"""ctypes binding of the C ABI in include/m3tsz_b200.h (libm3tsz_b200.so).

There is no CPU fallback anywhere in this package: if the shared library has
not been built, or no CUDA device is present, the calls below raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("M3TSZ_B200_LIB") or os.path.join(_HERE, "libm3tsz_b200.so")  # env: tuning builds

OK = 0
ERR_EOF = 1
ERR_DOD_OVERFLOW = 4
ERR_NO_TIME_SCHEME = 5
ERR_UNRECOGNIZED_UNIT = 6
ERR_INVALID_MULT = 7
ERR_OUT_OF_ORDER = 13
ERR_TOO_MANY_ITERATORS = 14
ERR_CAPACITY = 100
ERR_INVALID_ARG = 101
ERR_CUDA = 102
ERR_NO_DEVICE = 103

AGG_LAST, AGG_MIN, AGG_MAX, AGG_MEAN, AGG_COUNT, AGG_SUM = 1, 2, 3, 4, 6, 7  # aggregation.Type ids

UNIT_NONE, UNIT_S, UNIT_MS, UNIT_US, UNIT_NS, UNIT_MIN, UNIT_HOUR, UNIT_DAY, UNIT_YEAR = range(9)


class Options(C.Structure):
    """m3tsz_options: intOptimized flag + encoding.Options.DefaultTimeUnit."""
    _fields_ = [("int_optimized", C.c_int32), ("default_time_unit", C.c_int32)]


class AnnotationRef(C.Structure):
    _fields_ = [("bit_offset", C.c_uint64), ("length", C.c_uint32), ("count", C.c_uint32)]


class DpEvent(C.Structure):
    """m3tsz_dp_event: per-datapoint unit change / annotation (include/m3tsz_b200.h)."""
    _fields_ = [("series", C.c_uint64), ("dp_index", C.c_uint32), ("kind", C.c_uint16),
                ("unit", C.c_uint16), ("bit_offset", C.c_uint64), ("length", C.c_uint32),
                ("reserved", C.c_uint32)]


EVENT_TIME_UNIT, EVENT_ANNOTATION = 1, 2


class DecodeExtras(C.Structure):
    """m3tsz_decode_extras."""
    _fields_ = [("d_lengths", C.c_void_p), ("d_unit_first", C.c_void_p), ("d_events", C.c_void_p),
                ("events_capacity", C.c_uint64), ("d_event_count", C.c_void_p), ("point_major", C.c_int32),
                ("reserved", C.c_int32)]


class EncodeExtras(C.Structure):
    """m3tsz_encode_extras."""
    _fields_ = [("d_last_value", C.c_void_p), ("d_out_bits", C.c_void_p), ("point_major_input", C.c_int32),
                ("reserved", C.c_int32)]


class AnnotationEntry(C.Structure):
    _fields_ = [("dp_index", C.c_uint32), ("length", C.c_uint32), ("byte_offset", C.c_uint64)]


class M3tszError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = status_string(status)
        super().__init__("m3tsz_b200: %s (status %d)%s" % (msg, status, (": " + what) if what else ""))


_lib = None

# every symbol include/m3tsz_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "m3tsz_version", "m3tsz_status_string", "m3tsz_ctx_create", "m3tsz_ctx_destroy",
    "m3tsz_last_cuda_error", "m3tsz_ctx_launch_count", "m3tsz_decode_batch",
    "m3tsz_decode_batch_host", "m3tsz_encode_batch", "m3tsz_encode_bound",
    "m3tsz_compact_streams", "m3tsz_encode_batch_host", "m3tsz_decode_downsample_batch",
    "m3tsz_decode_downsample_batch_host", "m3tsz_merge_series_batch", "m3tsz_checksum_batch",
    "m3tsz_decode_batch_ex", "m3tsz_decode_downsample_last_batch", "m3tsz_encode_bound_units",
    "m3tsz_encode_batch_packed", "m3tsz_encode_batch_packed_ex", "m3tsz_prom_convert_batch", "m3tsz_aggregate_tiles_batch",
    "m3tsz_encode_batch_ex", "m3tsz_fetch_batch_host", "m3tsz_merge_series_batch_ex", "m3tsz_nccl_unique_id", "m3tsz_nccl_comm_create",
    "m3tsz_nccl_comm_destroy", "m3tsz_allgather_decoded",
    "m3tsz_encoder_create", "m3tsz_encoder_destroy", "m3tsz_encoder_reset", "m3tsz_encoder_encode",
    "m3tsz_encoder_num_encoded", "m3tsz_encoder_failed_dod", "m3tsz_encoder_last_encoded", "m3tsz_encoder_last_annotation_checksum",
    "m3tsz_encoder_empty", "m3tsz_encoder_len", "m3tsz_encoder_stream", "m3tsz_encoder_close",
    "m3tsz_encoder_discard", "m3tsz_encoder_discard_reset",
    "m3tsz_iter_create", "m3tsz_iter_destroy", "m3tsz_iter_reset", "m3tsz_iter_next", "m3tsz_iter_current",
    "m3tsz_iter_err", "m3tsz_iter_close",
    "m3tsz_encoder_pool_create", "m3tsz_encoder_pool_get", "m3tsz_encoder_pool_destroy",
    "m3tsz_iter_pool_create", "m3tsz_iter_pool_get", "m3tsz_iter_pool_destroy",
]


def lib():
    """Loads libm3tsz_b200.so.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "m3_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  This package has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, i64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32
    po = C.POINTER(Options)
    L.m3tsz_version.restype = C.c_int
    L.m3tsz_version.argtypes = []
    L.m3tsz_status_string.restype = C.c_char_p
    L.m3tsz_status_string.argtypes = [C.c_int]
    L.m3tsz_ctx_create.restype = C.c_int
    L.m3tsz_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.m3tsz_ctx_destroy.restype = None
    L.m3tsz_ctx_destroy.argtypes = [vp]
    L.m3tsz_last_cuda_error.restype = C.c_char_p
    L.m3tsz_last_cuda_error.argtypes = [vp]
    L.m3tsz_ctx_launch_count.restype = u64
    L.m3tsz_ctx_launch_count.argtypes = [vp]
    L.m3tsz_encode_bound.restype = u64
    L.m3tsz_encode_bound.argtypes = [u64]
    L.m3tsz_decode_batch.restype = C.c_int
    L.m3tsz_decode_batch.argtypes = [vp, po, vp, u64, vp, u64, vp, vp, u64, vp, vp, vp, vp, vp]
    L.m3tsz_encode_bound_units.restype = u64
    L.m3tsz_encode_bound_units.argtypes = [u64, C.c_int]
    L.m3tsz_encode_batch_packed.restype = C.c_int
    L.m3tsz_encode_batch_packed.argtypes = [vp, po, vp, vp, u64, u64, vp, vp, i32, vp, vp, vp, vp, u64, u32,
                                            vp, u64, vp, vp, vp, vp, vp]
    pv, pu64 = C.POINTER(vp), C.POINTER(u64)
    L.m3tsz_encode_batch_ex.restype = C.c_int
    L.m3tsz_encode_batch_ex.argtypes = [vp, po, vp, vp, u64, u64, vp, vp, i32, vp, vp, vp, vp, vp, u64,
                                        vp, vp, C.POINTER(EncodeExtras), vp]
    L.m3tsz_encode_batch_packed_ex.restype = C.c_int
    L.m3tsz_encode_batch_packed_ex.argtypes = [vp, po, vp, vp, u64, u64, vp, vp, i32, vp, vp, vp, vp, u64, u32,
                                               vp, u64, vp, vp, vp, vp, C.POINTER(EncodeExtras), vp]
    L.m3tsz_encoder_create.restype = C.c_int
    L.m3tsz_encoder_create.argtypes = [vp, po, i64, pv]
    L.m3tsz_encoder_destroy.restype = None
    L.m3tsz_encoder_destroy.argtypes = [vp]
    L.m3tsz_encoder_reset.restype = C.c_int
    L.m3tsz_encoder_reset.argtypes = [vp, i64, u64]
    L.m3tsz_encoder_encode.restype = C.c_int
    L.m3tsz_encoder_encode.argtypes = [vp, i64, C.c_double, i32, C.c_char_p, u64]
    L.m3tsz_encoder_failed_dod.restype = i64
    L.m3tsz_encoder_failed_dod.argtypes = [vp]
    L.m3tsz_encoder_num_encoded.restype = u64
    L.m3tsz_encoder_num_encoded.argtypes = [vp]
    L.m3tsz_encoder_last_encoded.restype = C.c_int
    L.m3tsz_encoder_last_encoded.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_double)]
    L.m3tsz_encoder_last_annotation_checksum.restype = C.c_int
    L.m3tsz_encoder_last_annotation_checksum.argtypes = [vp, pu64]
    L.m3tsz_encoder_empty.restype = C.c_int
    L.m3tsz_encoder_empty.argtypes = [vp]
    L.m3tsz_encoder_len.restype = C.c_int
    L.m3tsz_encoder_len.argtypes = [vp, pu64]
    for f in ("m3tsz_encoder_stream", "m3tsz_encoder_discard"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [vp, vp, u64, pu64, pu64]
    L.m3tsz_encoder_discard_reset.restype = C.c_int
    L.m3tsz_encoder_discard_reset.argtypes = [vp, i64, u64, vp, u64, pu64, pu64]
    L.m3tsz_encoder_close.restype = C.c_int
    L.m3tsz_encoder_close.argtypes = [vp]
    L.m3tsz_iter_create.restype = C.c_int
    L.m3tsz_iter_create.argtypes = [vp, po, pv]
    L.m3tsz_iter_destroy.restype = None
    L.m3tsz_iter_destroy.argtypes = [vp]
    L.m3tsz_iter_reset.restype = C.c_int
    L.m3tsz_iter_reset.argtypes = [vp, C.c_char_p, u64]
    L.m3tsz_iter_next.restype = C.c_int
    L.m3tsz_iter_next.argtypes = [vp]
    L.m3tsz_iter_current.restype = C.c_int
    L.m3tsz_iter_current.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_double), C.POINTER(i32), pv, pu64]
    L.m3tsz_iter_err.restype = C.c_int
    L.m3tsz_iter_err.argtypes = [vp]
    L.m3tsz_iter_close.restype = C.c_int
    L.m3tsz_iter_close.argtypes = [vp]
    for f in ("m3tsz_encoder_pool_create", "m3tsz_iter_pool_create"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [vp, po, u64, pv]
    for f in ("m3tsz_encoder_pool_get", "m3tsz_iter_pool_get"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [vp, pv]
    for f in ("m3tsz_encoder_pool_destroy", "m3tsz_iter_pool_destroy"):
        getattr(L, f).restype = None
        getattr(L, f).argtypes = [vp]
    L.m3tsz_nccl_unique_id.restype = C.c_int
    L.m3tsz_nccl_unique_id.argtypes = [vp]
    L.m3tsz_nccl_comm_create.restype = C.c_int
    L.m3tsz_nccl_comm_create.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.m3tsz_nccl_comm_destroy.restype = C.c_int
    L.m3tsz_nccl_comm_destroy.argtypes = [vp]
    L.m3tsz_allgather_decoded.restype = C.c_int
    L.m3tsz_allgather_decoded.argtypes = [vp, po, vp, C.c_int, vp, u64, vp, vp, u64, u64, u64, u64, vp, vp, vp,
                                          vp, vp]
    L.m3tsz_fetch_batch_host.restype = C.c_int
    L.m3tsz_fetch_batch_host.argtypes = [vp, po, vp, u64, vp, u64, vp, vp, vp, u64, u64, i64, i64, i32, vp, vp,
                                         u64, vp, vp]
    L.m3tsz_prom_convert_batch.restype = C.c_int
    L.m3tsz_prom_convert_batch.argtypes = [vp, vp, vp, u64, vp, u64, i64, vp, C.c_double, i64, vp, vp, u64,
                                           vp, vp, vp]
    L.m3tsz_aggregate_tiles_batch.restype = C.c_int
    L.m3tsz_aggregate_tiles_batch.argtypes = [vp, po, vp, u64, vp, vp, u64, i64, i64, u32, i32, i32, u32, vp,
                                              u64, vp, vp, vp, vp, vp, vp]
    L.m3tsz_decode_batch_ex.restype = C.c_int
    L.m3tsz_decode_batch_ex.argtypes = [vp, po, vp, u64, vp, u64, vp, vp, u64, vp, vp, vp, vp,
                                        C.POINTER(DecodeExtras), vp]
    L.m3tsz_decode_downsample_last_batch.restype = C.c_int
    L.m3tsz_decode_downsample_last_batch.argtypes = [vp, po, vp, u64, vp, u64, i64, i64, u32, vp, vp, vp,
                                                     vp, vp, vp, vp, vp, vp]
    L.m3tsz_decode_batch_host.restype = C.c_int
    L.m3tsz_decode_batch_host.argtypes = [vp, po, vp, u64, vp, u64, vp, vp, u64, vp, vp, vp, vp]
    L.m3tsz_encode_batch.restype = C.c_int
    L.m3tsz_encode_batch.argtypes = [vp, po, vp, vp, u64, u64, vp, vp, i32, vp, vp, vp, vp, vp, u64,
                                     vp, vp, vp]
    L.m3tsz_encode_batch_host.restype = C.c_int
    L.m3tsz_encode_batch_host.argtypes = [vp, po, vp, vp, u64, u64, vp, vp, i32, vp, vp, vp, vp, u64,
                                          u32, vp, u64, vp, vp, vp]
    L.m3tsz_compact_streams.restype = C.c_int
    L.m3tsz_compact_streams.argtypes = [vp, vp, u64, vp, u64, u32, vp, u64, vp, vp]
    L.m3tsz_decode_downsample_batch.restype = C.c_int
    L.m3tsz_decode_downsample_batch.argtypes = [vp, po, vp, u64, vp, u64, i64, i64, u32, vp, vp, vp,
                                                vp, vp, vp, vp]
    L.m3tsz_decode_downsample_batch_host.restype = C.c_int
    L.m3tsz_decode_downsample_batch_host.argtypes = [vp, po, vp, u64, vp, u64, i64, i64, u32, vp, vp,
                                                     vp, vp, vp, vp]
    L.m3tsz_merge_series_batch.restype = C.c_int
    L.m3tsz_merge_series_batch.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp, u64, i64, i64, i32, vp, vp,
                                           u64, vp, vp, vp]
    L.m3tsz_merge_series_batch_ex.restype = C.c_int
    L.m3tsz_merge_series_batch_ex.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, vp, u64, i64, i64, i32, vp, vp,
                                              u64, vp, vp, u64, i32, i32, vp]
    L.m3tsz_checksum_batch.restype = C.c_int
    L.m3tsz_checksum_batch.argtypes = [vp, vp, u64, vp, vp, u64, vp, vp, vp, vp]
    _lib = L
    return L


def status_string(status):
    try:
        return lib().m3tsz_status_string(int(status)).decode()
    except Exception:  # library missing: still produce a message
        return "status %d" % status


class Context:
    """m3tsz_ctx: owns the device binding and the scratch used by *_host calls."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().m3tsz_ctx_create(int(device), C.byref(self._h))
        if rc != OK:
            raise M3tszError(rc, "m3tsz_ctx_create(device=%d)" % device)
        self.device = device

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return int(lib().m3tsz_ctx_launch_count(self._h))

    def last_cuda_error(self):
        return lib().m3tsz_last_cuda_error(self._h).decode()

    def check(self, rc, what=""):
        if rc != OK:
            extra = what
            if rc == ERR_CUDA:
                extra = (what + " " + self.last_cuda_error()).strip()
            raise M3tszError(rc, extra)

    def close(self):
        if self._h:
            lib().m3tsz_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

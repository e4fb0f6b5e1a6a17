"""CPU-side checks of the product boundary: the C-ABI library loads, exports every
symbol include/m3tsz_b200.h declares, refuses to run without a GPU (no CPU
fallback), and the product package never touches the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "m3tsz_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(m3tsz_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from m3_b200 import capi
    lib = capi.lib()
    declared = _declared_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    assert lib.m3tsz_version() == 100
    assert lib.m3tsz_encode_bound(1440) % 16 == 0 and lib.m3tsz_encode_bound(1440) >= 1440 * 148 // 8


def test_status_strings_match_reference_errors():
    from m3_b200 import capi
    assert capi.status_string(1) == "EOF"
    assert capi.status_string(2) == "encoder is closed"  # m3tsz/encoder.go:37
    assert capi.status_string(3) == "encoder has no encoded datapoints"  # encoder.go:38
    assert capi.status_string(5) == "time encoding scheme doesn't exist for unit"  # timestamp_iterator.go:33
    assert capi.status_string(7) == "supplied multiplier is invalid"  # m3tsz.go:69
    assert capi.status_string(10) == "iterator is closed"  # iterator.go:33


def test_no_cpu_fallback_without_gpu():
    import torch
    from m3_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert capi.lib().m3tsz_ctx_create(0, C.byref(h)) == capi.ERR_NO_DEVICE
    from m3_b200.codec import BatchCodec
    with pytest.raises(capi.M3tszError):
        BatchCodec(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "m3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), (dirpath, f)


def test_host_helpers():
    from m3_b200.encoding import initial_time_unit, xxh64
    assert xxh64(b"") == 0xEF46DB3751D8E999
    assert xxh64(b"abc") == 0x44BC2CF5AD770999
    assert xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1
    assert xxh64(bytes(range(100))) == xxh64(bytes(range(100)))
    # encoder_test.go:395-410
    assert initial_time_unit(1, 1) == 0
    assert initial_time_unit(10 ** 9, 1) == 1
    assert initial_time_unit(10 ** 9, 0) == 0


def test_xxh64_matches_oracle_on_long_inputs():
    import oracle_lib as O
    from m3_b200.encoding import xxh64
    import random
    r = random.Random(1)
    for n in (1, 3, 4, 7, 8, 31, 32, 33, 63, 64, 100, 1000):
        b = bytes(r.randrange(256) for _ in range(n))
        assert xxh64(b) == O.lib().m3o_xxh64(b, n)

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


def _split_top_level(args):
    """Splits an argument list at top-level commas (parentheses / braces / brackets nest)."""
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(text, open_idx):
    """text[open_idx] == '(' -> the text between it and its matching ')'."""
    depth = 0
    for i in range(open_idx, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return text[open_idx + 1: i]
    raise AssertionError("unbalanced call at %d" % open_idx)


def _header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "m3tsz_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = {}
    for m in re.finditer(r"\b(m3tsz_[a-z_0-9]+)\s*\(", hdr):
        args = _balanced(hdr, m.end() - 1).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top_level(args))
    return protos, hdr


def _header_param_types():
    """function -> [C type of every parameter], qualifiers and parameter names dropped."""
    _, hdr = _header_prototypes()
    out = {}
    for m in re.finditer(r"\b(m3tsz_[a-z_0-9]+)\s*\(", hdr):
        args = _balanced(hdr, m.end() - 1).strip()
        types = []
        if args not in ("", "void"):
            for a in _split_top_level(args):
                a = re.sub(r"\s+", " ", a.strip())
                mm = re.match(r"^(.*?)(\b\w+)$", a)
                t = mm.group(1).strip() if mm and not a.endswith("*") else a
                types.append(re.sub(r"\bconst\b", "", t).replace(" ", ""))
        out[m.group(1)] = types
    return out


def test_go_shim_argument_types_match_the_header():
    """Every argument of the shim that carries an explicit C conversion -- `C.T(x)` for scalars,
    `(*C.T)(&x[0])` / `(*C.T)(unsafe.Pointer(..))` for buffers -- converts to the type the header declares for that
    parameter."""
    types = _header_param_types()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    go = re.sub(r"//[^\n]*", "", "\n".join(re.findall(r"```go\n(.*?)```", md, flags=re.S)))
    checked = 0
    for m in re.finditer(r"\bC\.(m3tsz_[a-z_0-9]+)\s*\(", go):
        name = m.group(1)
        for arg, want in zip(_split_top_level(_balanced(go, m.end() - 1)), types[name]):
            scalar = re.match(r"^C\.(\w+)\(", arg)
            pointer = re.match(r"^\(\*C\.(\w+)\)\(", arg)
            if scalar:
                assert want == scalar.group(1), (name, arg, want)
                checked += 1
            elif pointer:
                assert want == pointer.group(1) + "*", (name, arg, want)
                checked += 1
            elif arg == "nil" or arg.startswith("&"):
                assert want.endswith("*"), (name, arg, want)
                checked += 1
    assert checked >= 70


def test_go_shim_in_integration_md_matches_the_header():
    """INTEGRATION.md's cgo shim cannot be compiled here (no Go toolchain), so it is checked mechanically: every
    C function it calls is declared in include/m3tsz_b200.h with the same number of arguments, every C constant
    and C type it names exists there, and the per-datapoint interface methods of encoding.Encoder /
    ReaderIterator (encoding/types.go:39-91,180-203) are all implemented."""
    protos, hdr = _header_prototypes()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    go = "\n".join(re.findall(r"```go\n(.*?)```", md, flags=re.S))
    assert len(go) > 4000
    go_nc = re.sub(r"//[^\n]*", "", go)
    calls = list(re.finditer(r"\bC\.(m3tsz_[a-z_0-9]+)\s*\(", go_nc))
    assert len(calls) >= 25
    seen = set()
    for m in calls:
        name = m.group(1)
        assert name in protos, "INTEGRATION.md calls undeclared %s" % name
        n_args = len(_split_top_level(_balanced(go_nc, m.end() - 1)))
        assert n_args == protos[name], (name, n_args, protos[name])
        seen.add(name)
    # the streaming handles a shim binds method for method
    for need in ("m3tsz_encoder_create", "m3tsz_encoder_encode", "m3tsz_encoder_stream", "m3tsz_encoder_len",
                 "m3tsz_encoder_reset", "m3tsz_encoder_discard", "m3tsz_encoder_last_encoded",
                 "m3tsz_iter_create", "m3tsz_iter_reset", "m3tsz_iter_next", "m3tsz_iter_current",
                 "m3tsz_iter_err"):
        assert need in protos, need
        assert need in seen, "the shim never calls %s" % need
    for const in set(re.findall(r"\bC\.(M3TSZ_[A-Z_0-9]+)\b", go_nc)):
        assert re.search(r"\b%s\b" % const, hdr), const
    for typ in set(re.findall(r"\bC\.(m3tsz_[a-z_0-9]+)\b(?!\s*\()", go_nc)):
        assert re.search(r"\b%s\b" % typ, hdr), typ
    # interface coverage (method names of encoding.Encoder and encoding.ReaderIterator)
    for method in ("Encode", "Stream", "NumEncoded", "LastEncoded", "LastAnnotationChecksum", "Empty", "Len",
                   "Reset", "Close", "Discard", "DiscardReset", "SetSchema"):
        assert re.search(r"func \(e \*Encoder\) %s\(" % method, go), "Encoder.%s" % method
    for method in ("Next", "Current", "Err", "Close", "Reset"):
        assert re.search(r"func \(\w+ \*ReaderIterator\) %s\(" % method, go), "ReaderIterator.%s" % method

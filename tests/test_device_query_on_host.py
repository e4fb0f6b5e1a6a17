"""The Prometheus-epilogue kernels (m3_b200/csrc/m3tsz_query.cu: `prom_simple_kernel`, one warp per series, and
`prom_general_kernel`, one thread per series) compiled for the host from the CUDA source and run thread by thread
the way `launch_prom` picks and shapes them; compared with the oracle's iteratorToPromResult restatement and the
independent Python one (tests/test_independent_query_rows.py) on random batches: counters with resets, tolerance
clamps, NaN, empty series, output capacity overflow."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from test_device_encoder_on_host import _cut
from test_independent_query_rows import prom_convert

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "m3_b200", "csrc")
SEC = 10 ** 9

SHIM = r"""
#include <cstdint>
#include <cstring>
#include "%s"
#define __device__
#define __global__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
struct Dim3 { unsigned x, y, z; };
static thread_local Dim3 threadIdx, blockIdx, blockDim;
"""

DRIVER = r"""
extern "C" int dev_prom(const int64_t *ts, const double *val, uint64_t cap, const uint32_t *n_points, uint64_t n_series,
                        int64_t resolution, const uint8_t *handle_resets, double tolerance, int64_t tolerance_until,
                        int64_t *ts_out, double *val_out, uint64_t out_cap, uint32_t *n_out, int32_t *status) {
  using namespace m3tsz;
  PromParams p;
  memset(&p, 0, sizeof(p));
  p.ts = ts; p.val = val; p.cap = cap; p.n_points = n_points; p.n_series = n_series; p.resolution = resolution;
  p.handle_resets = handle_resets; p.tolerance = tolerance; p.tolerance_until = tolerance_until;
  p.ts_out = ts_out; p.val_out = val_out; p.out_cap = out_cap; p.n_out = n_out; p.status = status;
  blockDim.x = 256;
  const bool simple = !(p.tolerance > 0) && !p.handle_resets;  // launch_prom
  const uint64_t threads = simple ? n_series * 32ull : n_series;
  for (uint64_t t = 0; t < threads; t++) {
    blockIdx.x = (unsigned)(t / 256);
    threadIdx.x = (unsigned)(t % 256);
    if (simple) prom_simple_kernel(p); else prom_general_kernel(p);
  }
  return simple ? 1 : 0;
}
"""


@pytest.fixture(scope="module")
def dev():
    src = open(os.path.join(CSRC, "m3tsz_query.cu")).read()
    ker = open(os.path.join(CSRC, "m3tsz_kernels.h")).read()
    a = src.index("namespace m3tsz {")
    b = src.index("cudaError_t launch_prom")
    body = src[a:b] + "\n}  // namespace m3tsz\n"
    params = "namespace m3tsz {\n" + _cut(ker, r"struct PromParams") + "\n}\n"
    assert "asm" not in body and "<<<" not in body
    d = tempfile.mkdtemp(prefix="m3dev_query_host_")
    path = os.path.join(d, "dev_query_host.cpp")
    open(path, "w").write(SHIM % os.path.join(ROOT, "include", "m3tsz_b200.h") + params + body + DRIVER)
    so = os.path.join(d, "dev_query_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", path, "-o", so])
    lib = C.CDLL(so)
    lib.dev_prom.restype = C.c_int
    return lib


def _batch(rng, S, cap):
    ts = np.zeros((S, cap), dtype=np.int64)
    vals = np.zeros((S, cap), dtype=np.float64)
    n = rng.integers(0, cap + 1, size=S).astype(np.uint32)
    n[:3] = [0, 1, cap]
    for s in range(S):
        k = int(n[s])
        t0 = 1599955200 * SEC + int(rng.integers(0, 10 ** 9))
        ts[s, :k] = t0 + np.cumsum(rng.integers(1, 120, size=k)) * SEC // 2
        kind = s % 3
        if kind == 0 and k:
            inc = np.abs(np.round(rng.normal(size=k) * 10, 1))
            v = np.cumsum(inc)
            for r in rng.integers(0, k, size=max(1, k // 30)):
                v[r:] -= v[r] - inc[r]
            vals[s, :k] = v
        elif kind == 1 and k:
            v = 1000.0 + np.cumsum(np.abs(rng.normal(size=k)))
            v[rng.integers(0, k, size=max(1, k // 10))] *= (1 - 10.0 ** -rng.integers(2, 7))
            vals[s, :k] = v
        elif k:
            v = rng.normal(size=k) * 100
            v[rng.integers(0, k, size=max(1, k // 20))] = np.nan
            vals[s, :k] = v
    return ts, vals, n


@pytest.mark.parametrize("mode", ["simple", "tolerance", "resets", "both", "small_cap"])
def test_prom_kernels_on_host_match_oracle_and_model(dev, mode):
    rng = np.random.default_rng({"simple": 1, "tolerance": 2, "resets": 3, "both": 4, "small_cap": 5}[mode])
    S, cap = 90, 70
    ts, vals, n = _batch(rng, S, cap)
    resolution = 300 * SEC
    tol = 1e-3 if mode in ("tolerance", "both", "small_cap") else 0.0
    until = int(ts[:, : cap // 2].max()) if tol else 0
    handle = None
    if mode in ("resets", "both", "small_cap"):
        handle = (rng.random(S) < 0.6).astype(np.uint8)
    out_cap = cap + 1 if mode != "small_cap" else 20
    t_out = np.zeros((S, out_cap), dtype=np.int64)
    v_out = np.zeros((S, out_cap), dtype=np.float64)
    n_out = np.zeros(S, dtype=np.uint32)
    st = np.zeros(S, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    was_simple = dev.dev_prom(p(ts), p(vals), C.c_uint64(cap), p(n), C.c_uint64(S), C.c_int64(resolution),
                              None if handle is None else p(handle), C.c_double(tol), C.c_int64(until), p(t_out),
                              p(v_out), C.c_uint64(out_cap), p(n_out), p(st))
    assert was_simple == (1 if mode == "simple" else 0)
    for s in range(S):
        k = int(n[s])
        h = bool(handle[s]) if handle is not None else False
        o_ts, o_v = O.prom_convert_series(ts[s, :k], vals[s, :k], resolution, h, tol, until)
        ref = prom_convert(ts[s, :k].tolist(), vals[s, :k].tolist(), resolution, h, tol, until)
        assert len(ref) == len(o_ts) == int(n_out[s]), (s, len(ref), len(o_ts), int(n_out[s]))
        assert int(st[s]) == (100 if len(ref) > out_cap else 0), s  # M3TSZ_ERR_CAPACITY
        m = min(len(ref), out_cap)
        assert t_out[s, :m].tolist() == o_ts[:m].tolist() == [r[0] for r in ref[:m]], s
        assert (v_out[s, :m].view(np.uint64) == o_v[:m].view(np.uint64)).all(), s

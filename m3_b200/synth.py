"""Synthetic workloads of BASELINE.json's shape: S series x P
points, 60 s cadence from a day-aligned block start, unit = Second, values a
Gaussian random walk x0 = 100, x_i = x_{i-1} + N(0,1).  Generated with torch on
whichever device is asked for; the same tensors feed the GPU codec and (copied
to the host) any CPU-side checker, so both sides see bit-identical inputs."""
import torch

BLOCK_START_S = 1_599_955_200  # day-aligned => initial time unit = Second
SEC = 1_000_000_000


def gaussian_walk(n_series: int, n_points: int, device, seed: int = 1234, chunk: int = 65536):
    """Returns (ts int64 [S,P], values float64 [S,P], start int64 [S])."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    start_ns = BLOCK_START_S * SEC
    ts_row = start_ns + torch.arange(n_points, dtype=torch.int64, device=device) * (60 * SEC)
    ts = ts_row.unsqueeze(0).expand(n_series, n_points).contiguous()
    vals = torch.empty((n_series, n_points), dtype=torch.float64, device=device)
    for lo in range(0, n_series, chunk):
        hi = min(n_series, lo + chunk)
        inc = torch.randn((hi - lo, n_points), dtype=torch.float64, device=device, generator=g)
        inc[:, 0] = 0.0
        torch.cumsum(inc, dim=1, out=inc)
        inc += 100.0
        vals[lo:hi] = inc
    start = torch.full((n_series,), start_ns, dtype=torch.int64, device=device)
    return ts, vals, start

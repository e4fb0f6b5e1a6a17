"""Multi-GPU plumbing: series batches shard trivially (the reference itself
hash-partitions series across shards, src/dbnode/sharding/shardset.go:157-173),
so every rank encodes / decodes its own contiguous series range with NO
collective on the data path.  The only exchange step is on the fetch side, when
a query spans shards: one all-gather of the decoded (or downsampled) blocks
over NCCL / NVLink (gloo on CPU for tests)."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def partition(n_series: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of series owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n_series, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_series: int, world: int) -> List[int]:
    return [partition(n_series, r, world)[1] - partition(n_series, r, world)[0] for r in range(world)]


def all_gather_blocks(local: torch.Tensor, n_series_total: int, group=None) -> torch.Tensor:
    """Gathers per-rank blocks [S_r, ...] (series-major, S_r from partition()) into
    [n_series_total, ...] on every rank.  Uniform shards use one
    all_gather_into_tensor; ragged shards pad to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_series_total, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes[rank])
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((n_series_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(sizes)
    padded = torch.zeros((mx,) + tail, dtype=local.dtype, device=local.device)
    padded[: sizes[rank]] = local
    buf = torch.empty((world * mx,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def all_gather_windows(local: torch.Tensor, n_series_total: int, group=None) -> torch.Tensor:
    """Same for window-major downsample outputs [W, S_r] -> [W, n_series_total]."""
    return all_gather_blocks(local.t().contiguous(), n_series_total, group).t().contiguous()

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


# ---------------------------------------------------------------------------
# Config 5 as written: decode sharded over the GPUs + all-gather of the decoded blocks,
# chunked and overlapped (m3tsz_allgather_decoded, include/m3tsz_b200.h).
# ---------------------------------------------------------------------------
def make_nccl_comm(codec, dist_mod, dev):
    """An ncclComm_t for the C ABI: rank 0 makes the unique id, torch.distributed carries it."""
    import ctypes as C
    from . import capi
    world, rank = dist_mod.get_world_size(), dist_mod.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        codec.ctx.check(capi.lib().m3tsz_nccl_unique_id(buf), "m3tsz_nccl_unique_id")
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.to(dev)
    dist_mod.broadcast(uid, src=0)
    raw = bytes(uid.cpu().tolist())
    comm = C.c_void_p()
    codec.ctx.check(capi.lib().m3tsz_nccl_comm_create(codec.ctx.handle, raw, world, rank, C.byref(comm)),
                    "m3tsz_nccl_comm_create")
    return comm


def allgather_decoded(codec, comm, world, streams, offsets, lengths, gather_series, max_points, chunk_series, out=None):
    """Decode the first `gather_series` local streams and all-gather the decoded blocks; returns
    (ts [K, world, C, P], values [K, world, C, P], n_points [world, G], status [world, G]) on every rank,
    K = gather_series / chunk_series: the datapoints of series s of rank r are [s // C, r, s % C]."""
    import ctypes as C
    from . import capi
    dev = codec.device
    G, P, CH = int(gather_series), int(max_points), int(chunk_series)
    assert G % CH == 0, "gather_series must be a multiple of chunk_series"
    K = G // CH
    if out is None:
        out = (torch.empty((K, world, CH, P), dtype=torch.int64, device=dev),
               torch.empty((K, world, CH, P), dtype=torch.float64, device=dev),
               torch.empty((world, G), dtype=torch.int32, device=dev),
               torch.empty((world, G), dtype=torch.int32, device=dev))
    n_series = lengths.numel() if lengths is not None else offsets.numel() - 1
    rc = capi.lib().m3tsz_allgather_decoded(
        codec.ctx.handle, C.byref(codec.opts), comm, world, C.c_void_p(streams.data_ptr()), streams.numel(),
        C.c_void_p(offsets.data_ptr()), C.c_void_p(lengths.data_ptr()) if lengths is not None else None, n_series, G,
        P, int(chunk_series), C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
        C.c_void_p(out[2].data_ptr()), C.c_void_p(out[3].data_ptr()),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    codec.ctx.check(rc, "m3tsz_allgather_decoded")
    return out


def fetch_allgather_decoded(codec, pk, P, dist_mod, dev, barrier, budget_bytes=None):
    """bench.py side measurement at N > 1: how much of config 5 ("decode sharded across the GPUs +
    all-gather of decoded blocks") fits, how fast the overlapped pipeline runs, and what the
    gather alone costs against the measured NVLink peer figure (770 GB/s per direction)."""
    import ctypes as C
    from . import capi
    world = dist_mod.get_world_size()
    S = pk.offsets.numel()
    free, _ = torch.cuda.mem_get_info(dev)
    budget = int(0.6 * free) if budget_bytes is None else budget_bytes
    per_series = world * P * 16 + 2 * P * 16 // 8  # gathered + staging share
    G = int(max(1024, min(S, budget // max(1, per_series))))
    # every rank must gather the same number of series: agree on the smallest budget
    g = torch.tensor([G], dtype=torch.int64, device=dev)
    dist_mod.all_reduce(g, op=dist_mod.ReduceOp.MIN)
    G = int(g.item())
    chunk = min(G, 32768)
    G = (G // chunk) * chunk
    comm = make_nccl_comm(codec, dist_mod, dev)
    out = allgather_decoded(codec, comm, world, pk.packed, pk.offsets, pk.out_len, G, P, chunk)
    barrier()
    ok = bool((out[3] == 0).all()) and bool((out[2] == P).all())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    allgather_decoded(codec, comm, world, pk.packed, pk.offsets, pk.out_len, G, P, chunk, out=out)
    b.record()
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(ms, op=dist_mod.ReduceOp.MAX)
    # decode alone over the same streams, for the overlap figure
    from .codec import DecodeResult
    d_ts = torch.empty((G, P), dtype=torch.int64, device=dev)
    d_v = torch.empty((G, P), dtype=torch.float64, device=dev)
    dr = DecodeResult(ts=d_ts, values=d_v, n_points=torch.empty(G, dtype=torch.int32, device=dev),
                      status=torch.empty(G, dtype=torch.int32, device=dev),
                      unit=torch.empty(G, dtype=torch.uint8, device=dev), annotations=None)
    codec.decode(pk.packed, pk.offsets[:G], P, out=dr, lengths=pk.out_len[:G])
    torch.cuda.synchronize()
    c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c.record()
    codec.decode(pk.packed, pk.offsets[:G], P, out=dr, lengths=pk.out_len[:G])
    d.record()
    torch.cuda.synchronize()
    dec_ms = c.elapsed_time(d)
    same = torch.equal(out[0][:, dist_mod.get_rank()].reshape(G, P), d_ts)
    recv = (world - 1) * G * P * 16
    capi.lib().m3tsz_nccl_comm_destroy(comm)
    t = float(ms[0])
    return {"series_per_rank_gathered": G, "of_series_per_rank": S, "chunk_series": chunk, "ok": ok and same,
            "pipeline_ms": t, "decode_only_ms": dec_ms, "bytes_received_per_gpu": recv,
            "nvlink_gbs_per_gpu": recv / (t * 1e-3) / 1e9, "nvlink_peer_peak_gbs": 770.0,
            "frac_of_nvlink": recv / (t * 1e-3) / 1e9 / 770.0,
            "gathered_dps": world * world * G * P / (t * 1e-3),
            "api": "m3tsz_allgather_decoded (decode chunk k+1 || in-place ncclAllGather of chunk k, chunk-major result, no staging)",
            "note": "config 5 in full (8 x 1M x 1440 x 16 B = 184 GB per GPU) exceeds HBM: the call gathers the "
                    "first G series of every shard; the pipeline is bound by the gather (16 B/dp over NVLink), "
                    "the decode hides behind it"}

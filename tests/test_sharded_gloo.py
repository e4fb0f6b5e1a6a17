"""world_size-2 gloo test of the multi-GPU host logic (partitioning + the
fetch-side all-gather), on CPU."""
import os
import subprocess
import sys
import textwrap

import pytest


def test_partition_covers_everything():
    from m3_b200.sharded import partition, shard_sizes
    for n in (0, 1, 7, 8, 100_000, 1_000_001):
        for w in (1, 2, 3, 8):
            spans = [partition(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sz = shard_sizes(n, w)
            assert sum(sz) == n and max(sz) - min(sz) <= 1


WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from m3_b200.sharded import partition, all_gather_blocks, all_gather_windows
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for total in (64, 65):
        lo, hi = partition(total, rank, world)
        full_ts = torch.arange(total * 5, dtype=torch.int64).reshape(total, 5)
        full_v = torch.arange(total * 5, dtype=torch.float64).reshape(total, 5) * 0.5
        got_ts = all_gather_blocks(full_ts[lo:hi].clone(), total)
        got_v = all_gather_blocks(full_v[lo:hi].clone(), total)
        assert torch.equal(got_ts, full_ts) and torch.equal(got_v, full_v)
        win = full_v.t().contiguous()  # [W=5, S]
        got_w = all_gather_windows(win[:, lo:hi].contiguous(), total)
        assert torch.equal(got_w, win)
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()
""")


def test_all_gather_world2_gloo(tmp_path):
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % root)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29543", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "GLOO_OK" in out.stdout

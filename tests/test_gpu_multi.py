"""N=2 test of the one exchange step of the path (BASELINE config 5): per-GPU partition
decode + chunked NCCL all-gather of the decoded blocks through m3tsz_allgather_decoded.
Needs 2 GPUs (skipped otherwise): run under `gpurun --gpus 2`."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from m3_b200 import synth
    from m3_b200.codec import BatchCodec
    from m3_b200.sharded import make_nccl_comm, allgather_decoded
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    S, P, G, CH = 5000, 300, 4096, 1024   # 4 chunks of 1024 series
    codec = BatchCodec(lr, True)
    data = [synth.gaussian_walk(S, P, dev, seed=50 + r) for r in range(world)]
    ts, vals, start = data[rank]
    pk = codec.encode_packed(ts, vals, start, unit=1, align=64)
    comm = make_nccl_comm(codec, dist, dev)
    for lengths in (pk.out_len, None):
        if lengths is None:   # CSR variant: slots + compaction
            enc = codec.encode(ts, vals, start, unit=1)
            packed, off = codec.compact(enc, align=64)
            out = allgather_decoded(codec, comm, world, packed, off, None, G, P, CH)
        else:
            out = allgather_decoded(codec, comm, world, pk.packed, pk.offsets, lengths, G, P, CH)
        torch.cuda.synchronize()
        assert bool((out[3] == 0).all()) and bool((out[2] == P).all())
        for r in range(world):  # chunk-major result: series s of rank r = [s // CH, r, s mod CH]
            assert torch.equal(out[0][:, r].reshape(G, P), data[r][0][:G]), (rank, r)
            assert torch.equal(out[1][:, r].reshape(G, P).view(torch.int64), data[r][1][:G].view(torch.int64)), (rank, r)
    dist.barrier()
    if rank == 0:
        print("ALLGATHER_DECODED_OK")
    dist.destroy_process_group()
""")


def test_allgather_decoded_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % root)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29544", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ALLGATHER_DECODED_OK" in out.stdout

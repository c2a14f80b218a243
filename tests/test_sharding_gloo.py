"""N > 1 host logic on CPU: two gloo ranks shard the streams, each runs its shard (here through the
oracle, the GPU path is exercised by bench.py on the box), and the union equals the 1-rank result;
timing is reduced with MAX over ranks exactly as bench.py does."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiorust_b200 import sharding  # noqa: E402


def test_stream_range_partitions_exactly():
    for total in (1, 7, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            got = [sharding.stream_range(r, world, total) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.weak_scaling_streams(3, 8, 4096) == (3 * 4096, 4 * 4096)
    assert sharding.aggregate_throughput(1000, 4, 2.0) == 2000.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c
    from oracle import radiorust_oracle as orc

    lo, hi = sharding.stream_range(rank, world, total_streams)
    sr, n = 2_400_000.0, 1024
    x = np.stack([orc.synth_noise(20260000 + 300000 + s, 6 * n, "f32") for s in range(lo, hi)])
    shifts = [float((s * 577) % 2_400_000 - 1_200_000) for s in range(lo, hi)]
    ys = oracle_c.chain(x, "f32", sr, n, shifts=shifts, freq_resp=orc.lowpass(3000.0), down=(48000.0, 6000.0, 3.0))
    # per-stream checksums gathered on every rank (the optional output gather of SURVEY.md 8e)
    sums = torch.zeros(total_streams, dtype=torch.float64)
    for i, y in enumerate(ys):
        sums[lo + i] = float(np.sum(np.abs(y.astype(np.complex128)) ** 2))
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    elapsed = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((sums.numpy().copy(), float(elapsed.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    from oracle import oracle_c
    from oracle import radiorust_oracle as orc

    total = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    sums, elapsed = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert elapsed == pytest.approx(0.2)  # MAX over ranks
    sr, n = 2_400_000.0, 1024
    x = np.stack([orc.synth_noise(20260000 + 300000 + s, 6 * n, "f32") for s in range(total)])
    shifts = [float((s * 577) % 2_400_000 - 1_200_000) for s in range(total)]
    ys = oracle_c.chain(x, "f32", sr, n, shifts=shifts, freq_resp=orc.lowpass(3000.0), down=(48000.0, 6000.0, 3.0))
    want = np.array([float(np.sum(np.abs(y.astype(np.complex128)) ** 2)) for y in ys])
    assert np.allclose(sums, want, rtol=1e-12)

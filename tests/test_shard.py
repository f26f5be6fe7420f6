"""Batch sharding and the multi-rank timing reduction, world_size 2 over gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smfft_b200 import shard_ffts


@pytest.mark.parametrize("n_ffts", [0, 1, 7, 100, 131072, 16777216, 12345])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("gran", [1, 4, 64])
def test_shards_partition_the_batch(n_ffts, world, gran):
    spans = [shard_ffts(n_ffts, world, r, gran) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n_ffts
    for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert hi == lo2 and lo <= hi
    for lo, hi in spans[:-1]:
        assert (lo % gran == 0 or lo == n_ffts) and (hi % gran == 0 or hi == n_ffts)
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) < 2 * gran or n_ffts < world * gran


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench

    # each rank owns a slice of a global batch; job time is the max over ranks, units are summed
    lo, hi = shard_ffts(1000, world, rank, 4)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    units = torch.tensor([float(hi - lo)], dtype=torch.float64)
    tmax, total = bench.reduce_job(t, units)
    q.put((rank, lo, hi, tmax, total))
    dist.destroy_process_group()


def test_two_rank_job_reduction_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1000
    for _, _, _, tmax, total in res:
        assert tmax == 11.0 and total == 1000.0

"""CPU tests (gloo, world_size 2) of the window sharding: partition arithmetic and the
scatter -> per-rank model -> gather round trip.  The model here is a stand-in (nearest x4 of the
centre frame) because the real one is CUDA-only; the point is the plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realvsr_b200 import dist as D


def test_shard_bounds_partition():
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert D.shard_bounds(32, 3, 8) == (12, 16)  # BASELINE cfg3: 4 windows per GPU
    with pytest.raises(ValueError):
        D.shard_bounds(4, 2, 2)


def _fake_model(x):  # [b, N, C, H, W] -> [b, C, 4H, 4W]
    c = x[:, x.shape[1] // 2]
    return c.repeat_interleave(4, 2).repeat_interleave(4, 3) + 1.0


def _worker(rank, world, port, n_windows, q, src=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (n_windows, 5, 3, 8, 12)
        clips = torch.arange(float(torch.tensor(shape).prod())).view(shape) if rank == src else None
        out = D.sr_windows(_fake_model, clips, shape=shape, dtype=torch.float32, device="cpu", src=src)
        if rank == src:
            q.put(bool(torch.equal(out, _fake_model(clips))))
        else:
            q.put(out is None)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_windows,src", [(4, 0), (5, 0), (1, 0), (1, 1), (3, 1)])   # (1, 1): the source rank gets no window itself
def test_scatter_model_gather_roundtrip_gloo(n_windows, src):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_windows, q, src)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)


def _sharded_worker(rank, world, port, total, src, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tail_in, tail_out = (3, 2, 4, 6), (2, 16, 24)

        def fwd(x, out):
            out.copy_(_fake_model(x))

        sh = D.ShardedSR(fwd, total, tail_in, tail_out, torch.float32, "cpu", src=src, chunk=2)
        n_jobs = 5   # > 2: both buffer sets are reused
        if rank == src:
            g = torch.Generator().manual_seed(3)
            jobs = [(torch.rand((total,) + tail_in, generator=g), torch.zeros((total,) + tail_out)) for _ in range(n_jobs)]
            sh.run(jobs)
            q.put(all(bool(torch.equal(o, _fake_model(i))) for i, o in jobs))
        else:
            sh.run(n_jobs)
            q.put(True)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,src", [(8, 0), (5, 1), (1, 0)])
def test_sharded_sr_pipeline_gloo(total, src):
    """cfg3's scatter -> forward -> gather pipeline (ShardedSR) on two gloo ranks: every job's frames arrive complete and
    in window order, with the double-buffered schedule reusing both buffer sets."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, total, src, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)

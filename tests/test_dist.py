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


# ---------------------------------------------------------------- NCCL, real model (needs two GPUs)
def _nccl_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p_ in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from helpers import load_case
        from realvsr_b200.archs import EDVR_arch as E
        from synth import synth_input
        c = load_case("edvr_nf64_crop")
        net = E.EDVR(**c["kwargs"]).eval()
        net.load_state_dict(c["sd"], strict=True)
        net = net.to(dev).half()
        net.exec_path = "engine"
        total, (H, W) = 5, c["x"].shape[-2:]
        clips = synth_input((total, 5, 3, H, W), 77).half()
        with torch.no_grad():
            want = net(clips.to(dev)) if rank == 0 else None          # all windows on one GPU
        # 1. blocking scatter -> model -> gather
        got = D.sr_windows(net, clips.to(dev) if rank == 0 else None, shape=tuple(clips.shape), dtype=torch.float16, device=dev, src=0)
        ok = bool(torch.equal(got, want)) if rank == 0 else got is None
        # 2. the double-buffered host pipeline of bench.py's cfg3 section
        eng = net._get_engine(clips[:1].to(dev))
        sh = D.ShardedSR(lambda xc, yc: eng.forward(xc, out=yc), total, (5, 3, H, W), (3, 4 * H, 4 * W), torch.float16, dev, src=0, chunk=2)
        if rank == 0:
            jobs = [(clips.pin_memory(), torch.empty(total, 3, 4 * H, 4 * W, dtype=torch.float16).pin_memory()) for _ in range(3)]
            sh.run(jobs)
            ok = ok and all(bool(torch.equal(o, want.cpu())) for _, o in jobs) and sh.bytes_scatter > 0 and sh.bytes_gather > 0
        else:
            sh.run(3)
        q.put(ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_window_sharding_over_nccl_with_the_real_model():
    """BASELINE cfg3's plumbing on real hardware: 5 windows scattered from rank 0 over NCCL, super-resolved by the fp16 engine
    on two GPUs, gathered back -- bit-identical to all windows on one GPU, for the blocking helper (dist.sr_windows) and for
    the double-buffered host pipeline (dist.ShardedSR) that bench.py's `cfg3` key times."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res)


# ---------------------------------------------------------------- DDP training step on the bf16 training path (needs two GPUs)
def _ddp_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p_ in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import torch.nn.functional as F
        from helpers import load_case
        from realvsr_b200.archs import EDVR_arch as E
        from synth import synth_input
        c = load_case("edvr_nf64_crop")
        net = E.EDVR(**c["kwargs"]).train()
        net.load_state_dict(c["sd"], strict=True)
        net = net.to(dev)
        H, W = c["x"].shape[-2:]
        x = synth_input((2, 5, 3, H, W), 100 + rank).to(dev)       # a different batch on every rank
        gt = synth_input((2, 3, 4 * H, 4 * W), 200 + rank).to(dev)

        def step(model):
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):     # -> EDVR._forward_c8 (train_c8 path)
                loss = F.l1_loss(model(x).float(), gt)
            loss.backward()

        step(net)                                                  # local gradients, then their average over the ranks by hand
        want = []
        for p_ in net.parameters():
            g = p_.grad.detach().clone()
            dist.all_reduce(g)
            want.append(g / world)
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank])  # VideoSR_AllPair_model_YCbCr_Split.py:33-34
        step(ddp)
        worst = 0.0
        for p_, w in zip(net.parameters(), want):
            worst = max(worst, float((p_.grad - w).abs().max() / w.abs().max().clamp_min(1e-12)))
        # the same kernels ran on the same data; what differs is the order of the fp32 atomics inside dcn_bwd_tc_kernel (its
        # grad_input is then rounded to bf16, and a flipped last bit travels on through the data gradients) and NCCL's sum order
        print("rank %d: DDP vs hand-averaged gradients, worst relative difference %.2e" % (rank, worst), flush=True)
        q.put(worst < 2e-2)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ddp_training_step_on_the_c8_path():
    """SURVEY 8(e) training row: replicas + gradient all-reduce.  DistributedDataParallel around the module, bf16 autocast step on
    the train_c8 path on two GPUs with different batches: DDP's gradients equal the hand-averaged local gradients."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res)

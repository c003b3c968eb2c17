"""Window sharding across ranks (one process per GPU).

The hot path has no cross-window state (the reference slides its window one frame at a time and
recomputes everything, test_RealVSR_wi_GT.py:114-119), so multi-GPU inference is a plain partition
of the window list: rank r takes windows [r*B/G, (r+1)*B/G) and there is NO collective on the data
path.  The only communication is the scatter of LQ clips from the rank that holds the video and the
gather of SR frames back to it (1.73 MB in / 5.53 MB out per cfg2 window in fp16) -- NCCL over NVLink
on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_windows, rank, world):
    """Contiguous, balanced partition: the first n % world ranks get one extra window."""
    if world <= 0 or not (0 <= rank < world) or n_windows < 0:
        raise ValueError("bad shard request n=%d rank=%d world=%d" % (n_windows, rank, world))
    base, extra = divmod(n_windows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def scatter_clips(clips, shape=None, dtype=None, device=None, src=0, group=None):
    """Root holds `clips` [B, N, C, H, W]; every rank returns its shard [b_r, N, C, H, W].
    Non-root ranks pass clips=None and (shape, dtype, device) of the full tensor."""
    rank, world = _world()
    if world == 1:
        return clips
    if rank == src:
        shape, dtype, device = tuple(clips.shape), clips.dtype, clips.device
    B = shape[0]
    lo, hi = shard_bounds(B, rank, world)
    mine = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=device)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_bounds(B, r, world)
            if r == src:
                mine.copy_(clips[a:b])
            elif b > a:
                reqs.append(dist.isend(clips[a:b].contiguous(), dst=r, group=group))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=src, group=group)
    return mine


def gather_frames(frames, total, dst=0, group=None):
    """Every rank passes its SR frames [b_r, C, sH, sW]; rank `dst` returns [total, C, sH, sW] in window
    order, the others None."""
    rank, world = _world()
    if world == 1:
        return frames
    if rank == dst:
        out = torch.empty((total,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
        for r in range(world):
            a, b = shard_bounds(total, r, world)
            if r == dst:
                out[a:b].copy_(frames)
            elif b > a:
                dist.recv(out[a:b], src=r, group=group)
        return out
    if frames.shape[0] > 0:
        dist.send(frames.contiguous(), dst=dst, group=group)
    return None


def sr_windows(model, clips, shape=None, dtype=None, device=None, src=0):
    """Super-resolve a batch of independent windows across all ranks: scatter -> model -> gather."""
    rank, world = _world()
    total = clips.shape[0] if clips is not None else shape[0]
    mine = scatter_clips(clips, shape, dtype, device, src)
    with torch.no_grad():
        out = model(mine) if mine.shape[0] > 0 else None
    if out is None:  # this rank got no windows: still needs the output geometry for an empty send
        return gather_frames(torch.empty((0, 1, 1, 1), dtype=mine.dtype, device=mine.device), total, src) \
            if rank != src else None
    return gather_frames(out, total, src)

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


def sr_windows(model, clips, shape=None, dtype=None, device=None, src=0, out_shape=None):
    """Super-resolve a batch of independent windows across all ranks: scatter -> model -> gather.
    Rank `src` returns [total, C, sH, sW], the others None.  shard_bounds gives the extra windows to the LOWEST ranks,
    so `src` itself may receive none while others do (src != 0, few windows): it must still run the gather, and then
    needs the frame geometry from `out_shape` = (C, sH, sW) (default: C of the clips, x4 upscale)."""
    rank, world = _world()
    full = tuple(clips.shape) if clips is not None else tuple(shape)
    total = full[0]
    mine = scatter_clips(clips, shape, dtype, device, src)
    with torch.no_grad():
        out = model(mine) if mine.shape[0] > 0 else None
    if out is None:
        if rank == src:
            geom = tuple(out_shape) if out_shape is not None else (full[2], 4 * full[3], 4 * full[4])
            out = torch.empty((0,) + geom, dtype=mine.dtype, device=mine.device)
        else:  # nothing to send, nothing to receive
            return None
    return gather_frames(out, total, src)


class ShardedSR:
    """BASELINE cfg3 as a pipeline: rank `src` holds every job's clips in (pinned) HOST memory; per job
    H2D on src -> scatter of the shards (NCCL send/recv) -> forward on every rank -> gather of the frames to src ->
    D2H on src.  Two buffer sets and three side streams (copy-in, communication, copy-out) let job j + 1's H2D and
    scatter and job j - 1's gather and D2H run under job j's forward; every rank issues its p2p operations in the same
    order (scatter(0), scatter(1), gather(0), scatter(2), gather(1), ...), one look-ahead job.

        fwd(x, out)   writes the frames of the windows x [b, N, C, H, W] into out [b, C, sH, sW] (both on `device`)
        run(jobs)     jobs: on src a list of (host_in [total, N, C, H, W], host_out [total, C, sH, sW]); elsewhere the
                      job count.  Returns when every job's frames are in its host_out (src).
    On CPU tensors (gloo, the unit tests) the same schedule runs without streams."""

    def __init__(self, fwd, total, clip_shape, out_shape, dtype, device, src=0, chunk=4, group=None):
        self.fwd, self.total, self.src, self.chunk, self.group = fwd, int(total), src, max(1, int(chunk)), group
        self.rank, self.world = _world()
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        lo, hi = shard_bounds(self.total, self.rank, self.world)
        self.lo, self.hi = lo, hi
        mk = lambda n, tail: torch.empty((n,) + tuple(tail), dtype=dtype, device=self.device)  # noqa: E731
        self.is_src = self.rank == src
        # src computes its shard in place inside the full buffers; the others have shard-sized buffers
        n_in = self.total if self.is_src else hi - lo
        self.buf_in = [mk(n_in, clip_shape) for _ in range(2)]
        self.buf_out = [mk(n_in, out_shape) for _ in range(2)]
        self.s_in = self.s_comm = self.s_out = None
        if self.cuda:
            self.s_in, self.s_comm, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.bytes_scatter = self.bytes_gather = 0

    # -- small stream helpers (no-ops on CPU)
    def _on(self, stream):
        import contextlib
        return torch.cuda.stream(stream) if self.cuda else contextlib.nullcontext()

    def _event(self, stream):
        if not self.cuda:
            return None
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    def _wait(self, stream, ev):
        if self.cuda and ev is not None:
            stream.wait_event(ev)

    def _p2p(self, ops):
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()   # CUDA: orders the current (communication) stream behind NCCL's; CPU: blocks

    def _mine(self, bufs, j):
        b = bufs[j % 2]
        return b[self.lo:self.hi] if self.is_src else b

    def _scatter(self, j):
        ops = []
        if self.world > 1:
            if self.is_src:
                for r in range(self.world):
                    a, b = shard_bounds(self.total, r, self.world)
                    if r != self.src and b > a:
                        ops.append(dist.P2POp(dist.isend, self.buf_in[j % 2][a:b], r, self.group))
                        self.bytes_scatter += (b - a) * self.buf_in[0][0].numel() * self.buf_in[0].element_size()
            elif self.hi > self.lo:
                ops.append(dist.P2POp(dist.irecv, self.buf_in[j % 2], self.src, self.group))
        self._p2p(ops)

    def _gather(self, j):
        ops = []
        if self.world > 1:
            if self.is_src:
                for r in range(self.world):
                    a, b = shard_bounds(self.total, r, self.world)
                    if r != self.src and b > a:
                        ops.append(dist.P2POp(dist.irecv, self.buf_out[j % 2][a:b], r, self.group))
                        self.bytes_gather += (b - a) * self.buf_out[0][0].numel() * self.buf_out[0].element_size()
            elif self.hi > self.lo:
                ops.append(dist.P2POp(dist.isend, self.buf_out[j % 2], self.src, self.group))
        self._p2p(ops)

    def run(self, jobs):
        n = len(jobs) if self.is_src else int(jobs)
        cur = torch.cuda.current_stream(self.device) if self.cuda else None
        ev_h2d, ev_scat, ev_fwd, ev_gath, ev_out = ({} for _ in range(5))

        def stage_in(j):       # H2D (src) + scatter of job j
            if j >= n:
                return
            with self._on(self.s_in):
                if self.is_src:
                    self._wait(self.s_in, ev_scat.get(j - 2))      # buf_in[j % 2] was last read by scatter(j - 2)
                    self.buf_in[j % 2].copy_(jobs[j][0], non_blocking=True)
                    ev_h2d[j] = self._event(self.s_in)
            with self._on(self.s_comm):
                self._wait(self.s_comm, ev_h2d.get(j))
                self._wait(self.s_comm, ev_fwd.get(j - 2))         # shard buffer last read by forward(j - 2)
                self._scatter(j)
                ev_scat[j] = self._event(self.s_comm)

        stage_in(0)
        for j in range(n):
            stage_in(j + 1)
            # ---- forward of this rank's shard, `chunk` windows per engine call
            self._wait(cur, ev_scat.get(j))
            self._wait(cur, ev_gath.get(j - 2))                    # buf_out[j % 2] last read by gather(j - 2) ...
            self._wait(cur, ev_out.get(j - 2))                     # ... and, on src, by its D2H
            xin, yout = self._mine(self.buf_in, j), self._mine(self.buf_out, j)
            with torch.no_grad():
                for a in range(0, xin.shape[0], self.chunk):
                    self.fwd(xin[a:a + self.chunk], yout[a:a + self.chunk])
            ev_fwd[j] = self._event(cur)
            with self._on(self.s_comm):
                self._wait(self.s_comm, ev_fwd[j])
                self._wait(self.s_comm, ev_out.get(j - 2))
                self._gather(j)
                ev_gath[j] = self._event(self.s_comm)
            if self.is_src:
                with self._on(self.s_out):
                    self._wait(self.s_out, ev_gath[j])
                    jobs[j][1].copy_(self.buf_out[j % 2], non_blocking=True)
                    ev_out[j] = self._event(self.s_out)
            for d in (ev_h2d, ev_scat, ev_fwd, ev_gath, ev_out):
                d.pop(j - 3, None)
        if self.cuda:
            for s in (self.s_in, self.s_comm, self.s_out):
                s.synchronize()
            cur.synchronize()

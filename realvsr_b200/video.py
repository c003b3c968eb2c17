"""Video-level callers of the hot path: the reference's inference loop pieces.

  index_generation   data/util.py:169-214      temporal window indices with 4 padding modes
  single_forward     utils/util.py:222-237     no_grad forward, result .float().cpu()
  flipx4_forward     utils/util.py:240-261     x4 flip self-ensemble (flipx4_forward_batched: the same as one batch-4 call)
  sr_sequence        test_RealVSR_wi_GT.py:114-119 (the per-frame loop), re-designed: the per-frame
                     feature pyramid is extracted ONCE per frame into a cache and windows are batched,
                     instead of recomputing all N pyramids for every output frame.
  tiled_forward      (no reference counterpart: the reference's test scripts run whole frames) spatial tiling
                     with a halo for frames whose activations would not fit, e.g. BASELINE cfg4's 540x960 -> 4K.
"""
import torch


def index_generation(crt_i, max_n, N, padding='reflection'):
    """Indices of the N frames read for output frame crt_i of a max_n-frame sequence.
    crt_i = 0, N = 5:  replicate [0,0,0,1,2]  reflection [2,1,0,1,2]  new_info [4,3,0,1,2]  circle [3,4,0,1,2]"""
    if padding not in ('replicate', 'reflection', 'new_info', 'circle'):
        raise ValueError('Wrong padding mode')
    last, half = max_n - 1, N // 2
    out = []
    for i in range(crt_i - half, crt_i + half + 1):
        if i < 0:
            j = {'replicate': 0, 'reflection': -i, 'new_info': crt_i + half - i, 'circle': N + i}[padding]
        elif i > last:
            j = {'replicate': last, 'reflection': 2 * last - i, 'new_info': crt_i - half - (i - last),
                 'circle': i - N}[padding]
        else:
            j = i
        out.append(j)
    return out


def pad_to_multiple(x, m=4):
    """Replicate-pad the last two dims of x up to multiples of m (bottom / right).  The network halves the resolution
    twice and doubles it back (EDVR_arch.py:279-287, :111-124), so H and W must be multiples of 4 -- the reference fails
    on e.g. 270x480 (-> 1080p); pad, run, and crop the result to [..., :s*H, :s*W].  Returns (padded, (H, W))."""
    import torch.nn.functional as F
    H, W = x.shape[-2:]
    ph, pw = (-H) % m, (-W) % m
    if ph == 0 and pw == 0:
        return x, (H, W)
    lead = x.shape[:-3]
    y = F.pad(x.reshape((-1,) + tuple(x.shape[-3:])), (0, pw, 0, ph), mode="replicate")
    return y.reshape(tuple(lead) + tuple(y.shape[-3:])), (H, W)


def single_forward(model, inp):
    """model(inp) without autograd; first element if the model returns a list/tuple; float, on the CPU."""
    with torch.no_grad():
        y = model(inp)
    if isinstance(y, (list, tuple)):
        y = y[0]
    return y.data.float().cpu()


def flipx4_forward(model, inp):
    """Average of the forward on the input and its W-, H- and HW-flipped versions (flipped back)."""
    acc = single_forward(model, inp)
    for dims in ((-1,), (-2,), (-2, -1)):
        acc = acc + torch.flip(single_forward(model, torch.flip(inp, dims)), dims)
    return acc / 4


def flipx4_forward_batched(model, inp):
    """flipx4_forward as ONE batch-4 call: the four flipped copies of every window go through the model together (the
    reference runs four forwards, utils/util.py:240-261).  Same result: the engine is batch-invariant."""
    B = inp.shape[0]
    dims = (None, (-1,), (-2,), (-2, -1))
    x = torch.cat([inp if d is None else torch.flip(inp, d) for d in dims], 0)
    y = single_forward(model, x)
    acc = y[:B]
    for k, d in enumerate(dims[1:], 1):
        acc = acc + torch.flip(y[k * B:(k + 1) * B], d)
    return acc / 4


def sr_sequence(model, frames, padding='replicate', batch=4, cache=True):
    """Super-resolve every frame of a clip.  frames: [T, C, H, W] CUDA tensor (fp16 or fp32);
    model: realvsr_b200 EDVR / EDVR_NoUp.  Returns [T, C, sH, sW] on the same device.

    Window for output t = index_generation(t, T, nframes, padding) (as the reference's test loop).
    cache=True: extract each frame's feature pyramid once (rvsr_engine_extract_features) and run
    alignment + fusion + reconstruction on batches of `batch` windows given as cache slots
    (rvsr_engine_forward_cached); cache=False: plain batched forwards on gathered windows.
    Both produce bit-identical frames."""
    T = frames.shape[0]
    N = model._cfg["nframes"]
    windows = [index_generation(t, T, N, padding) for t in range(T)]
    outs = []
    with torch.no_grad():
        if not cache:
            for t0 in range(0, T, batch):
                idx = torch.tensor(windows[t0:t0 + batch], device=frames.device)
                outs.append(model(frames[idx.view(-1)].view(idx.shape[0], N, *frames.shape[1:])))
            return torch.cat(outs, 0)
        eng = model._get_engine(frames[:1].unsqueeze(0).expand(1, N, *frames.shape[1:]))
        H, W = frames.shape[-2:]
        buf = eng.make_cache(T, H, W)
        for t0 in range(0, T, 4 * batch):  # extract in chunks to bound the workspace
            eng.extract_features(frames[t0:t0 + 4 * batch], buf, T, t0)
        for t0 in range(0, T, batch):
            outs.append(eng.forward_cached(buf, T, windows[t0:t0 + batch], frames))
    return torch.cat(outs, 0)


def tiled_forward(model, x, tile=(180, 320), halo=16, scale=None):
    """model(x) computed on overlapping spatial tiles.  x: [B, N, C, H, W] CUDA tensor; tile = LQ tile size (h, w),
    halo = LQ pixels of context on every side (cropped from the result).  Tile origins, sizes and halos are kept
    multiples of 4 (the pyramid halves the resolution twice).  The network's receptive field is larger than any
    practical halo, so pixels near tile seams differ slightly from a whole-frame forward (they are exact when one
    tile covers the frame); with halo = 16 the seam error is of the order of the fp16 storage error.
    Returns [B, C, s*H, s*W] on x's device."""
    B, N, C, H, W = x.shape
    th, tw = (min(tile[0], H) // 4 * 4, min(tile[1], W) // 4 * 4)
    halo = max(0, int(halo)) // 4 * 4
    if H % 4 or W % 4 or th <= 0 or tw <= 0:
        raise RuntimeError("tiled_forward: H, W and the tile must be multiples of 4 (got %dx%d, tile %s)" % (H, W, tile))
    out = None
    with torch.no_grad():
        for y0 in range(0, H, th):
            for x0 in range(0, W, tw):
                y1, x1 = min(y0 + th, H), min(x0 + tw, W)
                ya, xa = max(y0 - halo, 0), max(x0 - halo, 0)
                yb, xb = min(y1 + halo, H), min(x1 + halo, W)
                yt = model(x[..., ya:yb, xa:xb].contiguous())
                if isinstance(yt, (list, tuple)):
                    yt = yt[0]
                s = scale or yt.shape[-1] // (xb - xa)
                if out is None:
                    out = torch.empty(B, yt.shape[1], s * H, s * W, dtype=yt.dtype, device=yt.device)
                out[..., s * y0:s * y1, s * x0:s * x1] = yt[..., s * (y0 - ya):s * (y1 - ya), s * (x0 - xa):s * (x1 - xa)]
    return out

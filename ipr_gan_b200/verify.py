"""On-device watermark / signature verification sweep (BASELINE config 5; the loop of
experiments/image_generation.py:185-223 and sign_flip.py:59-77 without the .cpu() hops and the per-image
Python loop):

    z -> x = G(z);  zwm = fn_inp(z);  xwm = G(zwm);  ywm = fn_out(x)
    wm_x = postproc(apply_mask(xwm));  wm_y = postproc(apply_mask(ywm))          postproc = (clamp(-1,1)+1)/2
    q = ssim(wm_x, wm_y) per sample;  p = PDQ-hash matching p-value;  match = p < p_thres

All per-sample quantities stay on the GPU; one small all-reduce merges the per-rank sums when the sweep is sharded
over ranks (independent shards, no data-path collective).
"""
import torch

from . import dist, ops


_STAGING = {}


def _staging(rows):
    """Pinned host buffer for the latents of one sweep (page-locking 5 MB costs ~2 ms: done once per size)."""
    buf = _STAGING.get(rows)
    if buf is None:
        buf = _STAGING[rows] = torch.empty(rows, 128).pin_memory()
    return buf


def _postproc(x):
    return (x.clamp(-1, 1) + 1.0) / 2.0


@torch.no_grad()
def verification_sweep(G, fn_inp, fn_out, n_samples, batch=500, p_thres=0.01, seed=1234, sign_model=None,
                       return_per_sample=False):
    """G, fn_inp, fn_out: the (drop-in) generator and trigger modules on a CUDA device.  The watermark window is taken
    with ``opaque=True`` semantics like evaluate() does (a pure crop).  -> dict with Q_WM, P, MATCH, N (+ WBOX)."""
    dev = next(G.parameters()).device
    was_training = G.training
    G.eval()
    rank, world = dist.rank(), dist.world()
    per_rank = (n_samples + world - 1) // world
    lo, hi = rank * per_rank, min(n_samples, (rank + 1) * per_rank)
    gen = torch.Generator().manual_seed(seed + rank)
    mod = fn_out.module if hasattr(fn_out, "module") else fn_out
    size, pos = mod.config.size, mod.position
    crop_bg = torch.zeros(1, 1, size, size, device=dev)             # evaluate() forces opaque=True for apply_mask
    sums = torch.zeros(4, device=dev, dtype=torch.float64)
    keep = {"q": [], "p": [], "r": []}
    # latents: drawn batch by batch from the CPU generator (experiments/image_generation.py:191-196) straight into ONE
    # pinned buffer, so every copy is truly asynchronous and the host draws batch i+1 while the GPU works on batch i
    # (a pageable 1.3 MB source made each copy wait for the stream: ~1 ms of idle GPU per 2 500 samples)
    staged = _staging(max(hi - lo, 1))
    count = torch.zeros(4, dtype=torch.float64)
    i = lo
    while i < hi:
        b = min(batch, hi - i)
        z_host = staged[i - lo:i - lo + b]
        torch.randn(b, 128, generator=gen, out=z_host)
        z = z_host.to(dev, non_blocking=True)
        x = G(z)
        xwm = G(fn_inp(z))
        ywm = fn_out(x)
        wm_x = ops.crop_patch(xwm, crop_bg, pos, size, postproc=True)     # crop + clamp + (x + 1) / 2 in one launch
        wm_y = ops.crop_patch(ywm, crop_bg, pos, size, postproc=True)
        q = ops.ssim_per_sample(wm_x, wm_y)
        p, r = ops.matching_prob(wm_x, wm_y)
        sums[:3] += torch.stack([q.double().sum(), p.double().sum(), (p < p_thres).double().sum()])
        count[3] += b
        if return_per_sample:
            keep["q"].append(q), keep["p"].append(p), keep["r"].append(r)
        i += b
    sums += count.to(dev)
    if world > 1:
        torch.distributed.all_reduce(sums)
    s = sums.tolist()
    out = {"Q_WM": s[0] / s[3], "P": s[1] / s[3], "MATCH": int(round(s[2])), "N": int(round(s[3]))}
    if sign_model is not None:
        wrong, total = sign_model.compute_ber_counts(G)
        out["WBOX"] = wrong / total
        out["WBOX_counts"] = (wrong, total)
    if return_per_sample:
        out["per_sample"] = {k: torch.cat(v) for k, v in keep.items()}
    if was_training:
        G.train()
    return out

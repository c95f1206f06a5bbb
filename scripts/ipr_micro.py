"""Micro-benchmarks of the HBM-bound IPR kernels (CUDA events on the launching stream, L2 flushed
between iterations): achieved algorithmic GB/s vs MEASURED_PEAKS.json.  Run on the B200 box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipr_gan_b200 import ops  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    peak, src = peak_gbs()
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")   # 256 MiB > 126 MB L2
    rows = []
    for B, H in ((512, 32), (4096, 32), (16384, 32), (65536, 32), (16, 96), (1024, 96), (64, 256)):
        x = torch.rand(B, 3, H, H, device="cuda")
        y = torch.rand(B, 3, H, H, device="cuda")
        t = timeit(lambda: ops.ssim_loss_fwd_bwd(x, y, True), flush=flush)
        nbytes = 36.0 * B * H * H
        rows.append({"kernel": "ssim_fwd_bwd", "B": B, "H": H, "us": t * 1e6, "GBs": nbytes / t / 1e9,
                     "frac": nbytes / t / 1e9 / peak})
        t = timeit(lambda: ops.ssim_loss_fwd_bwd(x, y, True, need_grad=False), flush=flush)
        rows.append({"kernel": "ssim_fwd_only", "B": B, "H": H, "us": t * 1e6, "GBs": 24.0 * B * H * H / t / 1e9,
                     "frac": 24.0 * B * H * H / t / 1e9 / peak})
    for B in (512, 65536):
        x = torch.rand(B, 3, 32, 32, device="cuda")
        fg, bg = torch.rand(1, 3, 16, 16, device="cuda"), torch.zeros(1, 1, 16, 16, device="cuda")
        t = timeit(lambda: ops.paste_patch(x, fg, bg, "tl", 16), flush=flush)
        nbytes = 24.0 * B * 32 * 32
        rows.append({"kernel": "paste_patch", "B": B, "H": 32, "us": t * 1e6, "GBs": nbytes / t / 1e9,
                     "frac": nbytes / t / 1e9 / peak})
    for B, s in ((10000, 16), (10000, 32), (100000, 32)):
        x = torch.rand(B, 3, s, s, device="cuda")
        y = torch.rand(B, 3, s, s, device="cuda")
        t = timeit(lambda: ops.ssim_per_sample(x, y), flush=flush)
        rows.append({"kernel": "ssim_per_sample", "B": B, "H": s, "us": t * 1e6, "GBs": 24.0 * B * s * s / t / 1e9,
                     "frac": 24.0 * B * s * s / t / 1e9 / peak})
        t = timeit(lambda: ops.matching_prob(x, y), flush=flush)
        rows.append({"kernel": "phash_pvalue", "B": B, "H": s, "us": t * 1e6, "GBs": 24.0 * B * s * s / t / 1e9,
                     "frac": 24.0 * B * s * s / t / 1e9 / peak})
    for r in rows:
        print(json.dumps(r))
    print(json.dumps({"peak_gbs": peak, "peak_source": src}))


if __name__ == "__main__":
    main()

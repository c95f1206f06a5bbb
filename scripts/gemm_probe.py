"""Representative tap-GEMM / wgrad launches of the batch-512 DCGAN step, for `ncu --set full` and quick timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
cases = [("conv3", 64, 128, 16), ("conv4s2", 128, 128, 16), ("conv3", 256, 512, 4), ("convT4s2", 256, 128, 8),
         ("conv4s2_dgrad", 128, 128, 8), ("convT4s2_dgrad", 128, 256, 16)]
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")


def bench(fn, flops, name):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    t = ts[len(ts) // 2]
    print("%-28s %8.1f us  %7.1f TFLOP/s" % (name, t * 1e3, flops / (t * 1e-3) / 1e12))


for kind, C, O, H in cases:
    plan = dense.Plan(kind, C, O)
    x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
    ksz = 3 if "3" in kind.split("_")[0][-1:] or kind.startswith("conv3") else 4
    if kind in ("conv3", "conv4s2"):
        w = torch.randn(O, C, ksz, ksz, device="cuda") * 0.05
    elif kind in ("convT4s2",):
        w = torch.randn(C, O, 4, 4, device="cuda") * 0.05
    elif kind == "conv4s2_dgrad":
        w = torch.randn(C, O, 4, 4, device="cuda") * 0.05
    else:
        w = torch.randn(O, C, 4, 4, device="cuda") * 0.05
    wp = plan.pack(w)
    oh, ow = plan.out_hw(H, H)
    taps_total = plan.n_taps * plan.n_phases
    q = (H // 2) if plan.a_parity else H
    flops = 2.0 * B * q * q * plan.n_phases * plan.n_taps * C * O
    bench(lambda: plan.run(x, wp), flops, "tapgemm:%s %d->%d @%d" % (kind, C, O, H))
    if kind in ("conv3", "conv4s2", "convT4s2"):
        wg = dense.WGradPlan(plan, tuple(w.shape))
        dy = torch.randn(B, oh, ow, O, device="cuda").to(torch.bfloat16)
        g = torch.empty_like(w)
        bench(lambda: wg.run(dy, x, g), flops, "wgrad:%s %d->%d @%d" % (kind, C, O, H))

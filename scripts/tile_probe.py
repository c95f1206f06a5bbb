import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense
B, H = 512, 16
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
def bench(fn, name, flush_l2):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        if flush_l2: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for C, N in ((64, 128), (64, 64), (128, 128), (256, 128)):
    x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
    for kind, taps in (("linear", 1), ("conv3", 9)):
        plan = dense.Plan(kind, C, N)
        w = torch.randn(N, C, device="cuda") if kind == "linear" else torch.randn(N, C, 3, 3, device="cuda")
        wp = plan.pack(w)
        tiles = B * H * H // 128
        for fl in (True, False):
            t = bench(lambda: plan.run(x, wp), kind, fl)
            kb = taps * C // 64
            print("%-7s C=%3d N=%3d L2flush=%d: %7.1f us  tiles/CTA %.1f  cycles/tile %6.0f  cycles/k-block %5.0f" %
                  (kind, C, N, fl, t, tiles / 148, t * 1.9e3 / (tiles / 148), t * 1.9e3 / (tiles / 148) / kb))

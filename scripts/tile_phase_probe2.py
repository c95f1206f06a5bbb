"""Per-tile phase timestamps (CTA 0) of the low-K layers: conv4s2_dgrad 64->64 with mask(+stats), and the 64->64 patch GEMM."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dbg = torch.zeros(64 + 2 * 148, dtype=torch.int64, device="cuda")
os.environ["IPR_TG_DBG_PTR"] = hex(dbg.data_ptr())
from ipr_gan_b200 import dense
B = 512


def report(name, fn):
    for _ in range(3):
        dbg.zero_(); fn(); torch.cuda.synchronize()
    g = dbg[64:].view(148, 2).cpu(); g = g[g[:, 1] > 0]
    print("%s: CTAs %d: kernel span %.1f us; CTA durations min %.1f mean %.1f max %.1f us" % (name, g.shape[0], (g[:, 1].max() - g[:, 0].min()).item() / 1e3, (g[:, 1] - g[:, 0]).min().item() / 1e3, (g[:, 1] - g[:, 0]).float().mean().item() / 1e3, (g[:, 1] - g[:, 0]).max().item() / 1e3))
    d = dbg[:64].view(8, 8).cpu(); t0 = int(d[0, 0])
    print(" tile | mma: start  acc_free  issued | epi: wait_from  acc_ready  done")
    for i in range(7):
        print("  %d   | %8d %8d %8d | %8d %8d %8d" % tuple([i] + [int(d[i, j]) - t0 for j in range(6)]))


plan = dense.Plan("conv4s2_dgrad", 64, 64)
w = torch.randn(64, 64, 4, 4, device="cuda") * 0.05
wp = plan.pack(w)
dy = torch.randn(B, 16, 16, 64, device="cuda").to(torch.bfloat16)
act = torch.randn(B, 32, 32, 64, device="cuda").to(torch.bfloat16)
sig = torch.ones(1, device="cuda")
report("dgrad64 mask+stats", lambda: plan.run(dy, wp, epi=dense.EPI_MASK, slope=0.1, mask=act, sigma=sig, want_stats=True))
report("dgrad64 mask", lambda: plan.run(dy, wp, epi=dense.EPI_MASK, slope=0.1, mask=act, sigma=sig))
report("dgrad64 linear", lambda: plan.run(dy, wp))
lin = dense.Plan("linear", 64, 64)
wl = lin.pack(torch.randn(64, 64, device="cuda"))
xl = torch.randn(B, 32, 32, 64, device="cuda").to(torch.bfloat16)
bias = torch.zeros(64, device="cuda")
report("linear64 bias_lrelu", lambda: lin.run(xl, wl, epi=dense.EPI_BIAS_LRELU, slope=0.1, sigma=sig, bias=bias))
report("linear64 linear", lambda: lin.run(xl, wl))

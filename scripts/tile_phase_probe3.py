"""Per-tile phase timestamps (CTA 0) of the 3-channel patch GEMM (D's first conv at batch 512)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dbg = torch.zeros(64 + 2 * 148, dtype=torch.int64, device="cuda")
os.environ["IPR_TG_DBG_PTR"] = hex(dbg.data_ptr())
from ipr_gan_b200 import dense
B = 512
for name, cin, cout, a_shape, epi in (("patch 32->64 bias+lrelu", 32, 64, (B, 32, 32, 32), dense.EPI_BIAS_LRELU),
                                      ("64->64 linear", 64, 64, (B, 32, 32, 64), dense.EPI_BIAS_LRELU),
                                      ("64->32 f32", 64, 32, (B, 32, 32, 64), dense.EPI_LINEAR_F32)):
    plan = dense.Plan("linear", cin, cout)
    wp = (torch.randn(1, cout, cin, device="cuda") * 0.05).to(torch.bfloat16)
    x = torch.randn(*a_shape, device="cuda").to(torch.bfloat16)
    bias = torch.randn(cout, device="cuda")
    kw = dict(epi=epi, slope=0.1, bias=bias) if epi == dense.EPI_BIAS_LRELU else dict(epi=epi, n_valid=32)
    for _ in range(3):
        dbg.zero_(); plan.run(x, wp, **kw); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); plan.run(x, wp, **kw); ev[1].record(); torch.cuda.synchronize()
    g = dbg[64:].view(148, 2).cpu(); g = g[g[:, 1] > 0]
    print("%s: event %.1f us; CTAs %d span %.1f us; CTA durations min %.1f mean %.1f max %.1f us; start skew %.1f us" % (
        name, ev[0].elapsed_time(ev[1]) * 1e3, g.shape[0], (g[:, 1].max() - g[:, 0].min()).item() / 1e3,
        (g[:, 1] - g[:, 0]).min().item() / 1e3, (g[:, 1] - g[:, 0]).float().mean().item() / 1e3,
        (g[:, 1] - g[:, 0]).max().item() / 1e3, (g[:, 0].max() - g[:, 0].min()).item() / 1e3))
    d = dbg[:64].view(8, 8).cpu(); t0 = int(d[0, 0])
    print(" tile | mma: start  acc_free  issued | epi: wait_from  acc_ready  done")
    for i in range(8):
        print("  %d   | %8d %8d %8d | %8d %8d %8d" % tuple([i] + [int(d[i, j]) - t0 for j in range(6)]))

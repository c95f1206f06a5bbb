import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dbg = torch.zeros(64 + 2 * 148, dtype=torch.int64, device="cuda")
os.environ["IPR_TG_DBG_PTR"] = hex(dbg.data_ptr())
from ipr_gan_b200 import dense
B, H, C, N = 512, 16, 64, 128
x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
for kind in ("linear", "conv3"):
    plan = dense.Plan(kind, C, N)
    w = torch.randn(N, C, device="cuda") if kind == "linear" else torch.randn(N, C, 3, 3, device="cuda")
    wp = plan.pack(w)
    for _ in range(3):
        dbg.zero_(); plan.run(x, wp); torch.cuda.synchronize()
    g = dbg[64:].view(148, 2).cpu(); g = g[g[:, 1] > 0]
    print("CTAs %d: kernel span %.1f us; CTA durations min %.1f mean %.1f max %.1f us; start skew %.1f us" % (g.shape[0], (g[:, 1].max() - g[:, 0].min()).item() / 1e3, (g[:, 1] - g[:, 0]).min().item() / 1e3, (g[:, 1] - g[:, 0]).float().mean().item() / 1e3, (g[:, 1] - g[:, 0]).max().item() / 1e3, (g[:, 0].max() - g[:, 0].min()).item() / 1e3))
    d = dbg[:64].view(8, 8).cpu(); t0 = int(d[0, 0])
    print(kind, "(cycles relative to first tile start)")
    print(" tile | mma: start  acc_free  issued | epi: wait_from  acc_ready  done")
    for i in range(7):
        print("  %d   | %8d %8d %8d | %8d %8d %8d" % tuple([i] + [int(d[i, j]) - t0 for j in range(6)]))

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dbg = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
os.environ["IPR_WGRAD_DBG_PTR"] = hex(dbg.data_ptr())
from ipr_gan_b200 import dense
B, C, O, H = 512, 64, 128, 16
plan = dense.Plan("conv3", C, O)
x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(B, H, H, O, device="cuda").to(torch.bfloat16)
w = torch.randn(O, C, 3, 3, device="cuda")
wg = dense.WGradPlan(plan, tuple(w.shape))
g = torch.empty_like(w)
for _ in range(3):
    dbg.zero_()
    wg.run(dy, x, g)
    torch.cuda.synchronize()
d = dbg.view(-1, 8).cpu()
d = d[d[:, 3] > 0]
print("CTAs", d.shape[0], "k-blocks/CTA", d[:, 4].float().mean().item())
for i, name in enumerate(("setup done", "mma issue done", "accumulator ready", "epilogue done")):
    print("%-18s mean %8.0f  min %8.0f  max %8.0f cycles" % (name, d[:, i].float().mean(), d[:, i].min(), d[:, i].max()))

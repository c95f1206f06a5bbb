"""One MMA-bound launch (ConvTranspose2d 512->256 k4s2 at 4x4 -> 8x8, batch 512) for `ncu --set full`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense
B = 512
plan = dense.Plan("convT4s2", 512, 256)
wp = plan.pack(torch.randn(512, 256, 4, 4, device="cuda") * 0.02)
x = torch.randn(B, 4, 4, 512, device="cuda").to(torch.bfloat16)
for _ in range(4):
    plan.run(x, wp, want_stats=True)
torch.cuda.synchronize()

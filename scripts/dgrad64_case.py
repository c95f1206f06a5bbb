"""One low-K masked data-gradient launch (conv4s2_dgrad 64->64 @16->32, batch 512) for ncu."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense
B = 512
plan = dense.Plan("conv4s2_dgrad", 64, 64)
wp = plan.pack(torch.randn(64, 64, 4, 4, device="cuda") * 0.05)
dy = torch.randn(B, 16, 16, 64, device="cuda").to(torch.bfloat16)
act = torch.randn(B, 32, 32, 64, device="cuda").to(torch.bfloat16)
sig = torch.ones(1, device="cuda")
for _ in range(4):
    plan.run(dy, wp, epi=dense.EPI_MASK, slope=0.1, mask=act, sigma=sig, want_stats=True)
torch.cuda.synchronize()

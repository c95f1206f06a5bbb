"""One L2-stream-bound launch (Conv2d 64->128 k3s1 at 16x16 with bias + LeakyReLU epilogue, batch 512) for `ncu --set full`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense
B = 512
plan = dense.Plan("conv3", 64, 128)
wp = plan.pack(torch.randn(128, 64, 3, 3, device="cuda") * 0.02)
x = torch.randn(B, 16, 16, 64, device="cuda").to(torch.bfloat16)
sig, bias = torch.ones(1, device="cuda"), torch.randn(128, device="cuda")
for _ in range(4):
    plan.run(x, wp, epi=dense.EPI_BIAS_LRELU, slope=0.1, sigma=sig, bias=bias)
torch.cuda.synchronize()

"""One 3-channel patch GEMM (D's first conv: patches [B*1024][32] x W[64][32], bias + LeakyReLU epilogue, batch 512) for ncu."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense
B = 512
plan = dense.Plan("linear", 32, 64)
wp = (torch.randn(1, 64, 32, device="cuda") * 0.05).to(torch.bfloat16)
col = torch.randn(B, 32, 32, 32, device="cuda").to(torch.bfloat16)
sig, bias = torch.ones(1, device="cuda"), torch.randn(64, device="cuda")
for _ in range(4):
    plan.run(col, wp, epi=dense.EPI_BIAS_LRELU, slope=0.1, sigma=sig, bias=bias)
torch.cuda.synchronize()

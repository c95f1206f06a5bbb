import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
x = torch.rand(B, 3, 32, 32, device="cuda"); y = torch.rand(B, 3, 32, 32, device="cuda")
for _ in range(3):
    ops.ssim_loss_fwd_bwd(x, y, True)
torch.cuda.synchronize()

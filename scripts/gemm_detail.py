"""Per-launch table of the tcgen05 GEMMs of one eager batch-B DCGAN step (CUDA events, kernels queued back to back)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipr_gan_b200 import dense, engine  # noqa: E402
from ipr_gan_b200.trainer import ProtectedDCGANTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(1234)
dev = torch.device("cuda:0")
tr = ProtectedDCGANTrainer(batch=B, device=dev)
real = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
z = torch.randn(B, 128, device=dev)
tr.real.copy_(real); tr.latent.copy_(z)
for _ in range(2):
    tr._step()
torch.cuda.synchronize()
engine._USE_SIDE = False
dense.PROFILE_DETAIL = True
torch.cuda._sleep(300_000_000)
dense.PROFILE = []
dense.PROFILE_L2 = []
tr._step()
torch.cuda.synchronize()
rows = {}
for (kind, f, a, b), (_, l2) in zip(dense.PROFILE, dense.PROFILE_L2):
    e = rows.setdefault(kind, [0.0, 0.0, 0, 0.0])
    e[0] += f; e[1] += a.elapsed_time(b); e[2] += 1; e[3] += l2
tot = sum(e[1] for e in rows.values())
l2_tot = sum(e[3] for e in rows.values())
print("total GEMM time %.3f ms over %d launches; operand stream through L2 %.2f GB = %.2f TB/s (cap ~12.4 TB/s = 6300 B/clk)"
      % (tot, len(dense.PROFILE), l2_tot / 1e9, l2_tot / (tot * 1e-3) / 1e12))
for k, (f, ms, n, l2) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us x%d  %7.1f TFLOP/s  %5.1f TB/s L2  %5.1f%%  %s" % (ms / n * 1e3, n, f / ms / 1e9, l2 / (ms * 1e-3) / 1e12,
                                                                    100 * ms / tot, k))

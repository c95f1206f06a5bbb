"""One eager protected IPR-DCGAN step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipr_gan_b200.trainer import ProtectedDCGANTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(B, dev, use_graph=False)
g = torch.Generator().manual_seed(1234)
tr.set_inputs(torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1), torch.randn(B, 128, generator=g))
for _ in range(3):
    tr._step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr._step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

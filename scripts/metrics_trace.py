"""Metrics of the first steps of the batch-B protected DCGAN run (for comparing schedules / builds)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200.trainer import ProtectedDCGANTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(B, dev, use_graph=graph)
g = torch.Generator().manual_seed(1234)
if graph:
    tr.capture(warmup=0) if False else None
for i in range(n):
    real = torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1)
    z = torch.randn(B, 128, generator=g)
    tr.set_inputs(real, z)
    tr._step()
    m = tr.model.get_metrics()
    print(i, " ".join("%s=%.6f" % (k, v) for k, v in m.items()))

"""Eager (no CUDA graph) step times through the reference-API calls: what a user's own training loop gets."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from ipr_gan_b200.trainer import ProtectedDCGANTrainer
dev = torch.device("cuda", 0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B in (512, 64):
    tr = ProtectedDCGANTrainer(B, dev, use_graph=False)
    tr.capture(3)
    print("DCGAN batch %d eager: %.2f ms/step (%d library launches)" % (B, timed(tr.step), tr.launches_per_step))
    del tr
for w in ("srgan", "cyclegan"):
    tr, name, host = bench.build_family(w, use_graph=False)
    tr.capture(3)
    print("%s eager: %.2f ms/step (%d library launches)" % (w, timed(tr.step), tr.launches_per_step))
    del tr

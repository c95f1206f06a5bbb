"""Dump the captured step graph (DOT, verbose) for critical-path analysis: scripts/graph_critical_path.py reads it."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200 import _lib, engine
from ipr_gan_b200.trainer import ProtectedDCGANTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "step_graph_b%d.dot" % B)
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(B, dev, use_graph=False)
g = torch.Generator().manual_seed(1234)
tr.set_inputs(torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1), torch.randn(B, 128, generator=g))
s = torch.cuda.Stream(device=dev)
s.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(s):
    for _ in range(3):
        tr._step()
torch.cuda.current_stream(dev).wait_stream(s)
torch.cuda.synchronize()
engine.reset_caches()
graph = torch.cuda.CUDAGraph(keep_graph=True)
graph.enable_debug_mode()
with torch.cuda.graph(graph, stream=s, capture_error_mode="thread_local"):
    tr._step()
os.makedirs(os.path.dirname(out), exist_ok=True)
import warnings
with warnings.catch_warnings(record=True) as wlist:
    warnings.simplefilter('always')
    graph.debug_dump(out)
for w in wlist:
    print('WARNING:', str(w.message)[:300])
print("dumped", out, os.path.getsize(out))

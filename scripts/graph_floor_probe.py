"""How long does a captured graph of N trivial dependent kernels take on this GPU?  (launch-latency floor of the step)"""
import torch, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 254
x = torch.zeros(32, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        for _ in range(n): x.add_(1.0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(n): x.add_(1.0)
for _ in range(5): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): g.replay()
b.record(); torch.cuda.synchronize()
print({"n": n, "us_per_graph": a.elapsed_time(b) / 20 * 1e3, "us_per_kernel": a.elapsed_time(b) / 20 * 1e3 / n})

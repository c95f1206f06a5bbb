"""Host time of torch.empty inside eager protected DCGAN steps, by allocation size (is the caching allocator the cost?)."""
import os, sys, time, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ipr_gan_b200.trainer import ProtectedDCGANTrainer
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(64, dev, use_graph=False)
tr.capture(3)
real_empty = torch.empty
acc = collections.defaultdict(lambda: [0, 0.0, 0.0])
def timed_empty(*a, **k):
    t0 = time.perf_counter()
    r = real_empty(*a, **k)
    dt = time.perf_counter() - t0
    nb = r.numel() * r.element_size()
    b = 0 if nb < 1 << 20 else (1 if nb < 16 << 20 else 2)
    e = acc[b]; e[0] += 1; e[1] += dt; e[2] = max(e[2], dt)
    return r
torch.empty = timed_empty
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    tr.step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
torch.empty = real_empty
print("20 eager steps: host enqueue %.1f ms/step, wall %.1f ms/step" % (t_host * 50, t_all * 50))
for b, name in ((0, "< 1 MB"), (1, "1-16 MB"), (2, ">= 16 MB")):
    n, t, mx = acc[b]
    if n: print("torch.empty %-9s: %5d calls/step, %.1f us avg, %.1f us max, %.2f ms/step" % (name, n / 20, t / n * 1e6, mx * 1e6, t / 20 * 1e3))
print(torch.cuda.memory_stats()["num_alloc_retries"], "alloc retries;", torch.cuda.memory_stats()["num_device_alloc"], "cudaMallocs;",
      torch.cuda.memory_stats()["reserved_bytes.all.current"] >> 20, "MB reserved")

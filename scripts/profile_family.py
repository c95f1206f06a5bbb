"""SRGAN / CycleGAN protected step (BASELINE configs 3 / 4): eager vs CUDA-graph step time, and an ncu-friendly mode.

    python scripts/profile_family.py srgan|cyclegan            # eager ms/step, graph ms/step, launches/step
    ncu --metrics gpu__time_duration.sum --profile-from-start off ... python scripts/profile_family.py srgan one
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "srgan"
one = len(sys.argv) > 2 and sys.argv[2] == "one"


def timed(tr, host, n=5):
    for _ in range(2):
        tr.step_from_host(*host)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        last = tr.step_from_host(*host)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, last


tr, name, host = bench.build_family(workload, use_graph=False)
tr.capture(3)
if one:
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
ms_eager, m_eager = timed(tr, host)
print("%s eager: %.2f ms/step, %d library launches/step" % (workload, ms_eager, tr.launches_per_step))
del tr
tr, name, host = bench.build_family(workload, use_graph=True)
tr.capture(3)
ms_graph, m_graph = timed(tr, host)
print("%s graph: %.2f ms/step, %d library launches/step" % (workload, ms_graph, tr.launches_per_step))
print("metrics eager", {k: round(v, 4) for k, v in m_eager.items()})
print("metrics graph", {k: round(v, 4) for k, v in m_graph.items()})

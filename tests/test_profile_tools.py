"""The profiling tools that produce the evidence under profiles/ run on the committed artefacts (no GPU): the
critical-path walk of the captured step graph must match every kernel node with a launch of the ncu list and report a
dependency chain no longer than the summed kernel time."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("batch", [512, 64])
def test_graph_critical_path_on_committed_profiles(batch):
    dot = os.path.join(ROOT, "profiles", "r2_step_graph_b%d.dot.gz" % batch)
    csv = os.path.join(ROOT, "profiles", "r2_launches_step_b%d.csv" % batch)
    if not (os.path.exists(dot) and os.path.exists(csv)):
        pytest.skip("profiles not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "graph_critical_path.py"), dot, csv],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"graph: (\d+) nodes \((\d+) kernels\), (\d+) edges; launch list: (\d+) kernels", r.stdout)
    assert m, r.stdout[:500]
    nodes, kernels, edges, launches = map(int, m.groups())
    assert kernels >= 200 and edges >= kernels and abs(kernels - launches) <= 2
    m = re.search(r"summed kernel time ([\d.]+) us; longest dependency chain ([\d.]+) us over (\d+) kernels", r.stdout)
    assert m, r.stdout[:800]
    total, chain, n_chain = float(m.group(1)), float(m.group(2)), int(m.group(3))
    assert 0 < chain <= total and 50 <= n_chain <= kernels
    miss = re.search(r"WARNING: (\d+) graph kernels without a duration", r.stdout)
    assert miss is None or int(miss.group(1)) <= 2          # the one-off packing launch of the capture
    assert "tapgemm_kernel" in r.stdout

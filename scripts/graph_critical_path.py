"""Critical path of the captured step graph.

    python scripts/graph_critical_path.py gpurun_out/step_graph_b512.dot.gz profiles/r2_launches_step_b512.csv

Nodes and dependency edges come from torch.cuda.CUDAGraph.debug_dump (scripts/graph_dump.py); per-kernel durations
from the ncu launch list of the same (eager) step -- both enumerate the launches in issue order, names are checked.
Prints: kernels / summed time of the graph, length of its longest dependency chain, and which kernels sit on it.
ncu durations are cold-cache and serialised, so the absolute path length is an upper bound; what matters is the SHARE
of each kernel family on the chain and how much of the total work is off it (what the extra streams can hide).
"""
import csv, gzip, re, sys
from collections import defaultdict

dot, launches = sys.argv[1], sys.argv[2]
txt = (gzip.open(dot, "rt") if dot.endswith(".gz") else open(dot)).read()
nodes = {}
for m in re.finditer(r'"(graph_\d+_node_(\d+))"\[[^\]]*?label="\{(\w+)\s*\n\| \{ID \| \d+ \(topoId: \d+\) \| ([^\\}]*)', txt):
    nodes[m.group(1)] = (int(m.group(2)), m.group(3), m.group(4))
edges = re.findall(r'"(graph_\d+_node_\d+)" -> "(graph_\d+_node_\d+)"', txt)
kern = sorted([(i, n, name) for n, (i, kind, name) in nodes.items() if kind == "KERNEL"])


def short(mangled):
    return re.sub(r"<.*", "", _gk.get(mangled, mangled))[:40]


rows = []
with open(launches) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        rows.append((r["Kernel Name"], v / 1000.0 if u.startswith("n") else (v * 1000.0 if u.startswith("m") else v)))
print("graph: %d nodes (%d kernels), %d edges; launch list: %d kernels" % (len(nodes), len(kern), len(edges), len(rows)))
# the k-th launch of a kernel in the graph takes the duration of the k-th launch of the same kernel in the list (the two
# enumerate the same Python issue order; matching per name is robust against the one-off packing launches)
import subprocess


def norm(demangled):
    """kernel name + template arguments, the same string from c++filt output and from ncu's demangled names"""
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", demangled.strip())
    head = d.split("(")[0]
    head = re.sub(r"\((int|bool)\)", "", head).replace(" ", "")
    head = head.replace("true", "1").replace("false", "0")
    return head[:80]


mangled = [name for _, _, name in kern]
dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")


def key(name):
    return norm(name)


_gk = {m: norm(d) for m, d in zip(mangled, dem)}


def graph_key(m):
    return _gk[m]


by_name = defaultdict(list)
for lname, us in rows:
    by_name[key(lname)].append(us)
seen = defaultdict(int)
dur, missing = {}, 0
for i, n, name in kern:
    k = graph_key(name)
    lst = by_name.get(k)
    if not lst:
        missing += 1
        dur[n] = 0.0
        continue
    j = seen[k]; seen[k] += 1
    dur[n] = lst[j] if j < len(lst) else sum(lst) / len(lst)
if missing:
    print("WARNING: %d graph kernels without a duration" % missing)
for n in nodes:
    dur.setdefault(n, 0.0)
succ, indeg = defaultdict(list), defaultdict(int)
for a, b in edges:
    succ[a].append(b); indeg[b] += 1
order, stack = [], [n for n in nodes if indeg[n] == 0]
deg = dict(indeg)
while stack:
    n = stack.pop()
    order.append(n)
    for m in succ[n]:
        deg[m] -= 1
        if deg[m] == 0:
            stack.append(m)
best, prev = {}, {}
for n in order:
    best.setdefault(n, dur[n])
    for m in succ[n]:
        if best[n] + dur[m] > best.get(m, -1.0):
            best[m] = best[n] + dur[m]; prev[m] = n
end = max(best, key=best.get)
path = []
while end is not None:
    path.append(end); end = prev.get(end)
path.reverse()
total = sum(dur.values())
print("summed kernel time %.1f us; longest dependency chain %.1f us over %d kernels (%.0f %% of the work is on it)" %
      (total, best[path[-1]], sum(1 for n in path if nodes[n][1] == "KERNEL"), 100 * best[path[-1]] / total))
fam = defaultdict(lambda: [0.0, 0])
for n in path:
    if nodes[n][1] == "KERNEL":
        f = fam[short(nodes[n][2])]; f[0] += dur[n]; f[1] += 1
print("on the chain:")
for k, (us, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print("  %8.1f us %5.1f%% x%-3d %s" % (us, 100 * us / best[path[-1]], c, k))
off = defaultdict(lambda: [0.0, 0])
onset = set(path)
for n, (i, kind, name) in nodes.items():
    if kind == "KERNEL" and n not in onset:
        f = off[short(name)]; f[0] += dur[n]; f[1] += 1
print("off the chain (hidden if the other streams find free SMs):")
for k, (us, c) in sorted(off.items(), key=lambda kv: -kv[1][0])[:12]:
    print("  %8.1f us x%-3d %s" % (us, c, k))

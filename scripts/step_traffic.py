"""ncu launch list with DRAM bytes -> per-kernel table (time, DRAM bytes, achieved GB/s vs the measured HBM peak) and
profiles/r2_step_traffic.json (what bench.py's roofline.traffic reads).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/r2_launches_b512.csv python scripts/profile_step.py 512
    python scripts/step_traffic.py gpurun_out/r2_launches_b512.csv gpurun_out/r2_step_algorithmic.json 512
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, algo_path, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
with open(src) as fh:
    lines = [l for l in fh if not l.startswith("==")]
launch = OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<unnamed>::|at::native::|void ", "", name)
    e = launch.setdefault(key, {"name": name[:100]})
    val = float(r["Metric Value"].replace(",", ""))
    unit, metric = r["Metric Unit"], r["Metric Name"]
    if metric == "gpu__time_duration.sum":
        e["us"] = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        e[metric] = val * mult
agg = defaultdict(lambda: [0.0, 0.0, 0])
for e in launch.values():
    a = agg[e["name"]]
    a[0] += e.get("us", 0.0)
    a[1] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    a[2] += 1
tot_us = sum(a[0] for a in agg.values())
print("one eager step, batch %d: %d launches, %.1f us summed kernel time (cold-cache, serialised: use the shares)" %
      (batch, len(launch), tot_us))
print("%9s %6s %5s %10s %9s %6s  %s" % ("time us", "share", "n", "DRAM MB", "GB/s", "of pk", "kernel"))
for n, (us, by, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    gbs = by / (us * 1e-6) / 1e9 if us > 0 else 0.0
    print("%9.1f %5.1f%% %5d %10.2f %9.1f %5.1f%%  %s" % (us, 100 * us / tot_us, c, by / 1e6, gbs, 100 * gbs / peak, n))
gemm = [(us, by, c) for n, (us, by, c) in agg.items() if n.startswith("tapgemm_kernel") or n.startswith("wgrad_kernel")]
algo = json.load(open(algo_path)) if os.path.exists(algo_path) else {}
out = {"batch": batch, "gemm_launches": int(sum(c for _, _, c in gemm)), "gemm_dram_bytes_per_step": sum(b for _, b, _ in gemm),
       "gemm_time_us_ncu": sum(u for u, _, _ in gemm), "gemm_share_of_kernel_time": sum(u for u, _, _ in gemm) / tot_us,
       "gemm_algorithmic_bytes_per_step": algo.get("gemm_algorithmic_bytes", 0.0), "all_launches": len(launch),
       "all_dram_bytes_per_step": sum(a[1] for a in agg.values()), "kernel_time_us_ncu": tot_us,
       "source": "profiles/r2_launches_step_b%d.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                 "dram__bytes_write.sum --clock-control none, one eager step via scripts/profile_step.py)" % batch}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_step_traffic.json"), "w"), indent=1)
print(json.dumps(out))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<unnamed>::|at::native::|void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("ns", "nsecond"):
        val /= 1000.0
    elif unit in ("ms", "msecond"):
        val *= 1000.0
    rows.append((name[:90], val))
agg = defaultdict(lambda: [0.0, 0])
for n, v in rows:
    agg[n][0] += v
    agg[n][1] += 1
tot = sum(v for _, v in rows)
print("total %.1f us over %d launches" % (tot, len(rows)))
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%8.1f us %5.1f%% %4d  %s" % (v, 100 * v / tot, c, n))

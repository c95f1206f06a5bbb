#!/bin/bash
# A/B of one environment switch of the library: GPU tests, per-layer GEMM table and bench (batch 512 / 64) with the
# variable unset ("on") and set to 1 ("off").   usage: scripts/ab_env.sh IPR_TG_GENERIC_EPI
VAR=$1
mkdir -p gpurun_out/ab
python -m pytest tests/test_gpu_dense.py tests/test_gpu_dcgan.py tests/test_gpu_seqnet.py -x -q 2>&1 | tail -3
for v in on off; do
  unset $VAR
  [ $v = off ] && export $VAR=1
  python scripts/gemm_detail.py 512 > gpurun_out/ab/detail_$v.txt 2>&1
  python bench.py --skip-cpu-baseline --skip-eager-baseline > gpurun_out/ab/bench_$v.json 2>/dev/null
  python bench.py --batch 64 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/ab/bench64_$v.json 2>/dev/null
done
python - <<P
import json
for v in ("on","off"):
    for f in ("bench","bench64"):
        d=json.loads(open("gpurun_out/ab/%s_%s.json"%(f,v)).read().strip().splitlines()[-1]); print(v,f,d["ms_per_step"],d["e2e"]["value"],d["roofline"]["gemm_ms_per_step"])
P

#!/bin/bash
# A/B of two builds / env switches of the tap GEMM: per-layer table + bench at batch 512 and 64.
mkdir -p gpurun_out/ab
for v in cur noocc prev; do
  unset IPR_B200_LIB IPR_TG_NO_OCC2
  [ $v = noocc ] && export IPR_TG_NO_OCC2=1
  [ $v = prev ] && export IPR_B200_LIB=$PWD/scripts/ab_prev.so
  python scripts/gemm_detail.py 512 > gpurun_out/ab/detail_$v.txt 2>&1
  python bench.py --skip-cpu-baseline --skip-eager-baseline > gpurun_out/ab/bench_$v.json 2>/dev/null
  python bench.py --batch 64 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/ab/bench64_$v.json 2>/dev/null
done
python - <<P
import json
for v in ("cur","noocc","prev"):
    for f in ("bench","bench64"):
        d=json.loads(open("gpurun_out/ab/%s_%s.json"%(f,v)).read().strip().splitlines()[-1]); print(v,f,d["ms_per_step"],d["e2e"]["value"],d["roofline"]["gemm_ms_per_step"])
P

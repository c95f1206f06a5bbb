#!/bin/bash
# 1 -> 8 GPU curve of the step (global batch 512) and of the verification sweep, plus the 2-rank NCCL parity test.
mkdir -p gpurun_out/scale
[ -z "$SKIP_NCCL_TEST" ] && python -m pytest tests/test_gpu_dist_nccl.py -q 2>&1 | tail -2
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  $L bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu-baseline --skip-eager-baseline > gpurun_out/scale/bench_n$n.json 2> gpurun_out/scale/bench_n$n.err
  $L bench.py --gpus $n --workload verify > gpurun_out/scale/verify_n$n.json 2> gpurun_out/scale/verify_n$n.err
done
python - <<P
import json
import os
for n in [int(x) for x in os.environ.get('SCALE_NS', '1 2 4 8').split()]:
    for f in ("bench","verify"):
        try:
            d=json.loads(open("gpurun_out/scale/%s_n%d.json"%(f,n)).read().strip().splitlines()[-1]); print(n,f,round(d["value"],1),d["unit"],round(d["ms_per_step"],3),"e2e",round(d.get("e2e",{}).get("value",0),1))
        except Exception as e: print(n,f,"failed",e)
P

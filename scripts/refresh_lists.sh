#!/bin/bash
# launch lists (time + DRAM bytes) and graph dumps of the current build, batch 512 and 64
O=gpurun_out/final; mkdir -p $O
for b in 512 64; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file $O/launches_b$b.csv python scripts/profile_step.py $b > /dev/null 2>&1
  cp gpurun_out/r2_step_algorithmic.json $O/algorithmic_b$b.json
  python scripts/graph_dump.py $b /tmp/g$b.dot 2>&1 | grep dumped; gzip -c /tmp/g$b.dot > $O/step_graph_b$b.dot.gz
done

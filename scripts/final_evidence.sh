#!/bin/bash
# Round-end evidence, part 1: bench line (with CPU + eager-GPU baselines), reference arm, launch lists with DRAM bytes at
# batch 512 / 64.  Everything lands in gpurun_out/final/.
O=gpurun_out/final; mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for b in 512 64; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file $O/launches_b$b.csv python scripts/profile_step.py $b > /dev/null 2>&1
  cp gpurun_out/r2_step_algorithmic.json $O/algorithmic_b$b.json
done
tail -c 300 $O/bench_n1.json

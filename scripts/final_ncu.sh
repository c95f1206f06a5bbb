#!/bin/bash
# Round-end evidence, part 2: ncu --set full of three tap-GEMM regimes, exported to text on the box (the reports
# themselves are too large to travel back).
O=gpurun_out/final; mkdir -p $O
for c in conv3 convT linear; do
  ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 3 -c 1 -o /tmp/prof_$c python scripts/${c}_case.py > $O/prof_$c.log 2>&1
  ncu -i /tmp/prof_$c.ncu-rep --page details > $O/prof_${c}_details.txt 2>/dev/null
  ncu -i /tmp/prof_$c.ncu-rep --page raw --csv 2>/dev/null | grep -E 'dram__bytes_(read|write)\.sum|dram__cycles_active|gpu__dram_throughput|sm__pipe_tensor_cycles_active|sm__warps_active|launch__registers_per_thread|lts__t_bytes|gpu__time_duration|sm__inst_executed_pipe_tensor|smsp__inst_executed.sum|lts__throughput' > $O/prof_${c}_raw.csv
  ncu -i /tmp/prof_$c.ncu-rep --page source --csv 2>/dev/null | gzip > $O/prof_${c}_source.csv.gz
done
ls -la $O

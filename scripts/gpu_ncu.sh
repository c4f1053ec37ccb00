#!/bin/bash
# ncu evidence for the step kernel (B200_PROFILING.md recipe): launch list, then one full capture of the top kernel.
# Run only after the plain command exited 0; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
CMD=${CMD:-"python bench.py --steps 5 --warmup 3 --no-cpu-baseline"}
KERNEL=${KERNEL:-cbx_pipe_kernel}
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 16 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/prof_src.csv 40 | cut -c1-200 > gpurun_out/prof_lines.txt
ls -la gpurun_out/ | tail -12

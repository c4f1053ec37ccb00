#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_pipe_kernel -s 10 -c 1 -f -o gpurun_out/pipe $CMD > gpurun_out/ncu_pipe.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_pipe.log
ncu -i gpurun_out/pipe.ncu-rep --page raw --csv > gpurun_out/pipe_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/pipe_raw.csv > gpurun_out/pipe_summary.txt 2>&1
ncu -i gpurun_out/pipe.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/pipe_src.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/pipe_src.csv 70 | cut -c1-220 > gpurun_out/pipe_lines.txt
head -75 gpurun_out/pipe_lines.txt
gzip -f gpurun_out/pipe_src.csv

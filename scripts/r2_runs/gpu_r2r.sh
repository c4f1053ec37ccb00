#!/bin/bash
mkdir -p gpurun_out
for w in chain100 random16; do echo "== phases $w"; WORKLOAD=$w ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -14; done
echo "== pipe default now"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q.log 2>gpurun_out/q.err; python scripts/bench_line.py b20 < gpurun_out/q.log
CMD3="python bench.py --workload random16 --envs-per-gpu 131072 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_wide_kernel -s 5 -c 1 -f -o gpurun_out/wide16 $CMD3 > gpurun_out/ncu_wide16.log 2>&1
ncu -i gpurun_out/wide16.ncu-rep --page raw --csv > gpurun_out/wide16_raw.csv 2>/dev/null; python scripts/ncu_summary.py gpurun_out/wide16_raw.csv > gpurun_out/wide16_summary.txt 2>&1
ncu -i gpurun_out/wide16.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/wide16_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/wide16_src.csv 70 | cut -c1-220 > gpurun_out/wide16_lines.txt; rm -f gpurun_out/wide16_src.csv gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload chain100 --envs-per-gpu 131072 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_wide_kernel -s 5 -c 1 -f -o gpurun_out/wide $CMD > gpurun_out/ncu_wide.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_wide.log
ncu -i gpurun_out/wide.ncu-rep --page raw --csv > gpurun_out/wide_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/wide_raw.csv > gpurun_out/wide_summary.txt 2>&1; grep -E "dram__bytes|time_duration|inst_executed|issue_active|warps_active|registers" gpurun_out/wide_summary.txt
ncu -i gpurun_out/wide.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/wide_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/wide_src.csv 50 | cut -c1-200 > gpurun_out/wide_lines.txt
gzip -f gpurun_out/wide_src.csv
head -55 gpurun_out/wide_lines.txt
echo "== phases chain100"; WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -10

#!/bin/bash
# Round 2, first GPU pass: smoke, every GPU test (new: dynamic tile order, term_* arrays, escalation / Chain-100 / masked tapes),
# the bench at the driver's 20 steps and at the default 1000, the other workloads, then dram traffic of the wide kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader > gpurun_out/gpu.txt
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench 20 steps (driver settings)"; timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.log 2> gpurun_out/bench_20.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_20.err
python scripts/bench_line.py b20 < gpurun_out/bench_20.log
echo "== bench default"; timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench.err
python scripts/bench_line.py b1000 < gpurun_out/bench.log
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref.log
for w in chain100 random16; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?"
  python scripts/bench_line.py $w < gpurun_out/bench_$w.log
done
echo "== ncu: dram traffic + store sectors of cbx_wide_kernel (chain100, 131072 envs)"
CMD="python bench.py --workload chain100 --envs-per-gpu 131072 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_wide_kernel -s 5 -c 1 -f -o gpurun_out/wide $CMD > gpurun_out/ncu_wide.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_wide.log
ncu -i gpurun_out/wide.ncu-rep --page raw --csv > gpurun_out/wide_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/wide_raw.csv > gpurun_out/wide_summary.txt 2>&1; head -40 gpurun_out/wide_summary.txt
ncu -i gpurun_out/wide.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/wide_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/wide_src.csv 40 | cut -c1-200 > gpurun_out/wide_lines.txt
rm -f gpurun_out/wide_src.csv
ls -la gpurun_out | tail -15

#!/bin/bash
# Round 2 confirmation pass on one B200: smoke, every GPU test, the bench at the driver's settings and at the default, the
# reference arm, the other workloads, then the ncu evidence (launch list + one full capture per step kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader > gpurun_out/gpu.txt
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench 20 steps (driver settings)"; timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.log 2> gpurun_out/bench_20.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_20.err
python scripts/bench_line.py b20 < gpurun_out/bench_20.log
echo "== bench default"; timeout 500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench.err
python scripts/bench_line.py b1000 < gpurun_out/bench.log
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
for w in chain100 random16; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?"
  python scripts/bench_line.py $w < gpurun_out/bench_$w.log | head -1
done
for w in chain100_scan; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; python scripts/bench_line.py $w < gpurun_out/bench_$w.log | head -1
done
timeout 200 python bench.py --workload toyctf_live --no-cpu-baseline --no-e2e > gpurun_out/bench_toyctf_live.log 2> gpurun_out/bench_toyctf_live.err; python scripts/bench_line.py toyctf_live < gpurun_out/bench_toyctf_live.log | head -1
timeout 200 python bench.py --workload toyctf_scan --no-cpu-baseline --no-e2e > gpurun_out/bench_toyctf_scan.log 2> gpurun_out/bench_toyctf_scan.err; python scripts/bench_line.py toyctf_scan < gpurun_out/bench_toyctf_scan.log | head -1
timeout 300 python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 1048576 --steps 100 > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; python scripts/bench_line.py 1m < gpurun_out/bench_1m.log | head -1
echo "== ncu launch list (default workload)"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pipe.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full: cbx_pipe_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_pipe_kernel -s 10 -c 1 -f -o gpurun_out/pipe $CMD > gpurun_out/ncu_pipe.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/pipe.ncu-rep --page raw --csv > gpurun_out/pipe_raw.csv 2>/dev/null; python scripts/ncu_summary.py gpurun_out/pipe_raw.csv > gpurun_out/pipe_summary.txt 2>&1
ncu -i gpurun_out/pipe.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/pipe_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/pipe_src.csv 50 | cut -c1-220 > gpurun_out/pipe_lines.txt; rm -f gpurun_out/pipe_src.csv
echo "== ncu launch list + full: cbx_wide_kernel (chain100)"
CMD2="python bench.py --workload chain100 --envs-per-gpu 131072 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wide.csv $CMD2 > gpurun_out/ncu_launches2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_wide_kernel -s 5 -c 1 -f -o gpurun_out/wide $CMD2 > gpurun_out/ncu_wide.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/wide.ncu-rep --page raw --csv > gpurun_out/wide_raw.csv 2>/dev/null; python scripts/ncu_summary.py gpurun_out/wide_raw.csv > gpurun_out/wide_summary.txt 2>&1
ncu -i gpurun_out/wide.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/wide_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/wide_src.csv 50 | cut -c1-220 > gpurun_out/wide_lines.txt; rm -f gpurun_out/wide_src.csv
CMD3="python bench.py --workload random16 --envs-per-gpu 131072 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cbx_wide_kernel -s 5 -c 1 -f -o gpurun_out/wide16 $CMD3 > gpurun_out/ncu_wide16.log 2>&1
ncu -i gpurun_out/wide16.ncu-rep --page raw --csv > gpurun_out/wide16_raw.csv 2>/dev/null; python scripts/ncu_summary.py gpurun_out/wide16_raw.csv > gpurun_out/wide16_summary.txt 2>&1
ncu -i gpurun_out/wide16.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/wide16_src.csv 2>/dev/null; python scripts/ncu_lines.py gpurun_out/wide16_src.csv 50 | cut -c1-220 > gpurun_out/wide16_lines.txt; rm -f gpurun_out/wide16_src.csv
rm -f gpurun_out/*.ncu-rep
grep -E "dram__bytes_(read|write).sum  |time_duration" gpurun_out/pipe_summary.txt gpurun_out/wide_summary.txt gpurun_out/wide16_summary.txt
ls -la gpurun_out | tail -30

#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu (pipelined-kernel subset)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "dynamic or toyctf or chain10 or tape or reproducible or step_host or static" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for wl in 6 5 4; do
  echo "== toyctf WL=$wl"
  CBX_PIPE_WL=$wl timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_wl$wl.log 2> gpurun_out/bench_wl$wl.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_wl$wl.err
  python scripts/bench_line.py wl$wl < gpurun_out/bench_wl$wl.log
done
echo "== WL=6 WE=6";  CBX_PIPE_WL=6 CBX_PIPE_WE=6 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/bench_line.py wl6we6
echo "== phases (default)"; timeout 300 python scripts/gpu_phases.py 2>&1 | tail -10

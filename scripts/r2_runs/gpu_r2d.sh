#!/bin/bash
# Round 2, pass d: pipelined kernel with gather-then-expand small fields and 4 / 5 / 6 logic warps.
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for wl in 6 5 4; do
  echo "== toyctf WL=$wl"
  CBX_PIPE_WL=$wl timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_wl$wl.log 2> gpurun_out/bench_wl$wl.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_wl$wl.err
  python scripts/bench_line.py wl$wl < gpurun_out/bench_wl$wl.log
done
echo "== WL=6 WE=6";  CBX_PIPE_WL=6 CBX_PIPE_WE=6 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/bench_line.py wl6we6
echo "== WL=4 WE=10"; CBX_PIPE_WL=4 CBX_PIPE_WE=10 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python scripts/bench_line.py wl4we10
echo "== phases (default)"; timeout 300 python scripts/gpu_phases.py 2>&1 | tail -10
echo "== vecenv profile"; timeout 300 python scripts/gpu_vecenv_profile.py 2>&1 | head -16 | cut -c1-180

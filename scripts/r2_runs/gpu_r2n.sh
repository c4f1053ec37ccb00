#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu (live binding)"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "live" > gpurun_out/pytest_live.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_live.log | cut -c1-300
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench default"; timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench.err
python scripts/bench_line.py b1000 < gpurun_out/bench.log

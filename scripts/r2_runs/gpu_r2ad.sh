#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --config5 --no-cpu-baseline --no-e2e > gpurun_out/u.log 2> gpurun_out/u.err; echo "rc=$?"; tail -c 1500 gpurun_out/u.err
python scripts/bench_line.py c5 < gpurun_out/u.log
timeout 300 python -m pytest tests/test_ppo.py -q -m gpu 2>&1 | tail -2

#!/bin/bash
# First contact with a B200: smoke (TMA and plain-copy kernels), GPU parity tests, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke (TMA)"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_tma.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke_tma.log
echo "== smoke (plain copies)"; CBX_NO_TMA=1 timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke_plain.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err

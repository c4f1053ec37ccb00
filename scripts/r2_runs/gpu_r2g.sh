#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu (pipelined-kernel subset)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "dynamic or toyctf or chain10 or tape or reproducible or step_host or static" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for ov in 1 0; do
  CBX_PIPE_OVERLAP=$ov timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/bench_ov$ov.err | tee gpurun_out/bench_ov$ov.log | python scripts/bench_line.py ov$ov; tail -c 400 gpurun_out/bench_ov$ov.err
done
echo "== 262144 envs"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 262144 --steps 300 2>/dev/null | python scripts/bench_line.py 256k
echo "== 1048576 envs"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 1048576 --steps 100 2>/dev/null | python scripts/bench_line.py 1m
echo "== phases"; timeout 300 python scripts/gpu_phases.py 2>&1 | tail -10

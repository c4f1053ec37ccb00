#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "live or tape" 2>&1 | tail -5
echo "-- toyctf_live"; timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --workload toyctf_live 2>gpurun_out/y.err | python scripts/bench_line.py q | head -1; tail -3 gpurun_out/y.err
echo "-- toyctf"; timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>gpurun_out/y.err | python scripts/bench_line.py q | head -1; tail -3 gpurun_out/y.err

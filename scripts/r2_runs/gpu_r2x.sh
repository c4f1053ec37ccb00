#!/bin/bash
mkdir -p gpurun_out
for h in 0 1; do echo "-- toyctf_live hints=$h"; CBX_L2_HINTS=$h timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --workload toyctf_live 2>gpurun_out/x.err | python scripts/bench_line.py q | head -1; tail -3 gpurun_out/x.err; done
for h in 0 1; do echo "-- toyctf on the fused kernel hints=$h"; CBX_PIPE=0 CBX_L2_HINTS=$h timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>gpurun_out/x.err | python scripts/bench_line.py q | head -1; tail -3 gpurun_out/x.err; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "live or fused or launch_modes" 2>&1 | tail -3

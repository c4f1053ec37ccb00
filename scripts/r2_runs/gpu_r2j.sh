#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (pipelined-kernel subset)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "pipe_kernel or launch or overlapped or toyctf or chain10 or tape or reproducible or step_host" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
run() { echo "== $1"; env $2 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/h.err | python scripts/bench_line.py "$1" | head -1; tail -c 300 gpurun_out/h.err; }
run "default" "X=1"
run "serial static" "CBX_PIPE_OVERLAP=0"

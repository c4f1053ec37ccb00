#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu (pipelined-kernel subset)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "pipe_kernel or launch or overlapped or toyctf or chain10 or tape or reproducible or step_host" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
run() { echo "== $1"; env $2 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/h.err | python scripts/bench_line.py "$1" | head -1; tail -c 300 gpurun_out/h.err; }
run "default (4,8)x1 overlap dynamic" "X=1"
run "(2,4)x2" "CBX_PIPE_WL=2 CBX_PIPE_WE=4 CBX_PIPE_CTAS=2"
run "(2,3)x2" "CBX_PIPE_WL=2 CBX_PIPE_WE=3 CBX_PIPE_CTAS=2"
run "(3,8)x1" "CBX_PIPE_WL=3 CBX_PIPE_WE=8"
run "(4,6)x1" "CBX_PIPE_WL=4 CBX_PIPE_WE=6"
run "(4,7)x1" "CBX_PIPE_WL=4 CBX_PIPE_WE=7"
run "default again" "X=1"
run "serial static" "CBX_PIPE_OVERLAP=0"

#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; env $2 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/h.err | python scripts/bench_line.py "$1" | head -1; tail -c 300 gpurun_out/h.err; }
run "default(4,8)x1" "X=1"
run "dynamic" "CBX_PIPE_DYNAMIC=1"
run "(2,4)x2" "CBX_PIPE_WL=2 CBX_PIPE_WE=4 CBX_PIPE_CTAS=2"
run "(2,4)x2 dynamic" "CBX_PIPE_WL=2 CBX_PIPE_WE=4 CBX_PIPE_CTAS=2 CBX_PIPE_DYNAMIC=1"
run "(2,4)x2 no overlap" "CBX_PIPE_WL=2 CBX_PIPE_WE=4 CBX_PIPE_CTAS=2 CBX_PIPE_OVERLAP=0"
run "(3,4)x2" "CBX_PIPE_WL=3 CBX_PIPE_WE=4 CBX_PIPE_CTAS=2"
run "(2,3)x2" "CBX_PIPE_WL=2 CBX_PIPE_WE=3 CBX_PIPE_CTAS=2"
run "default again" "X=1"
run "no overlap" "CBX_PIPE_OVERLAP=0"

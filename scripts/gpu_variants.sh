#!/bin/bash
# phase breakdown of the pipelined kernel under tuning knobs
mkdir -p gpurun_out; : > gpurun_out/variants.log
run() { echo "=== $*" >> gpurun_out/variants.log; env "$@" timeout 300 python scripts/gpu_phases.py >> gpurun_out/variants.log 2>&1; }
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 >> gpurun_out/variants.log
run CBX_PIPE_LOGIC_TMA=0
run CBX_PIPE_LOGIC_TMA=1
run CBX_PIPE_LOGIC_TMA=0 CBX_PIPE_WE=6
run CBX_PIPE_LOGIC_TMA=0 CBX_PIPE_WE=4
run CBX_PIPE_LOGIC_TMA=0 ENVS=262144
cat gpurun_out/variants.log

#!/bin/bash
# phase breakdown of the pipelined kernel under tuning knobs
mkdir -p gpurun_out; : > gpurun_out/variants.log
run() { echo "=== $*" >> gpurun_out/variants.log; env "$@" timeout 300 python scripts/gpu_phases.py >> gpurun_out/variants.log 2>&1; }
run CBX_DEBUG_SKIP=0
run CBX_DEBUG_SKIP=32
run CBX_DEBUG_SKIP=64
run CBX_DEBUG_SKIP=96
run CBX_DEBUG_SKIP=128
run CBX_DEBUG_SKIP=256
run CBX_DEBUG_SKIP=352
cat gpurun_out/variants.log

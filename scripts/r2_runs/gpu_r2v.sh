#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_e2e_breakdown.py 2>&1 | tail -8
echo "-- overlap off"; CBX_PIPE_OVERLAP=0 timeout 300 python scripts/gpu_e2e_breakdown.py 2>&1 | tail -4

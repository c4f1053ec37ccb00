#!/bin/bash
mkdir -p gpurun_out
echo "== phases (overlap on)"; timeout 300 python scripts/gpu_phases.py 2>&1 | tail -12
echo "== phases (overlap off)"; CBX_PIPE_OVERLAP=0 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -12
echo "== vecenv profile"; timeout 300 python scripts/gpu_vecenv_profile.py 2>&1 | head -50 | cut -c1-180

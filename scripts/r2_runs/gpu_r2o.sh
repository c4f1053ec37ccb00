#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "live or ppo or rollout" > gpurun_out/pytest_live.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_live.log | cut -c1-300
echo "== phases random16"; WORKLOAD=random16 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -7

#!/bin/bash
# warp-per-tile kernel timeline: one round on every SM (66 304 envs) and one round on half of the SMs (33 152 envs)
for n in 66304 33152; do
  echo "== ENVS=$n"; CBX_LIB=marlon_b200/libcbx_trace.so WORKLOAD=chain100 ENVS=$n timeout 300 python scripts/gpu_wide_trace.py 2>&1 | grep -E "^cta +(0|37) warp +(0|3|6|9|12)|last stamp|warps per"
done

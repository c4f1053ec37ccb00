#!/bin/bash
# warp-per-tile kernel: per-phase cycles with and without the batched in-place loads (libcbx_wb.so = the batching patch)
for lib in "" marlon_b200/libcbx_wb.so; do
  echo "== phases chain100 lib=$lib"; env ${lib:+CBX_LIB=$lib} WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -7
done

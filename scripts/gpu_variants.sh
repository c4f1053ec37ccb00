#!/bin/bash
# Compare experiment builds (libcbx_<name>.so next to libcbx.so) on the bench workload; each must pass oracle parity first.
mkdir -p gpurun_out
for v in "" $VARIANTS; do
  if [ -z "$v" ]; then unset CBX_LIB; name=default; else export CBX_LIB=$PWD/marlon_b200/libcbx_$v.so; name=$v; fi
  echo "=== $name"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "marlon_pair or odd_bounds" 2>&1 | tail -1
  timeout 300 python scripts/gpu_phases.py 2>&1 | grep -E "kernel|logic|encode|state_" | tr '\n' ';'; echo
done

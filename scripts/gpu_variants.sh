#!/bin/bash
# Compare experiment builds (tile size / CTAs per SM) on the bench workload; each must pass the oracle parity first.
mkdir -p gpurun_out
for v in "" t32c6 t64 t128 t128r64; do
  if [ -z "$v" ]; then unset CBX_LIB; name=default; else export CBX_LIB=$PWD/marlon_b200/libcbx_$v.so; name=$v; fi
  echo "=== $name"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider -k "marlon_pair or odd_bounds or chain10_attacker" 2>&1 | tail -1
  timeout 300 python scripts/gpu_phases.py 2>&1 | grep -E "kernel|logic|encode|state_" | tr '\n' ';'; echo
done

#!/bin/bash
# Timing experiment: which output sections cost what (CBX_DEBUG_SKIP drops stores of a section; results are then invalid).
for sk in 0 1 2 4 8 16 23 31 15; do
  echo -n "skip=$sk  "; CBX_DEBUG_SKIP=$sk timeout 300 python scripts/gpu_phases.py 2>&1 | grep -E "kernel|encode|logic" | tr '\n' ';' | cut -c1-330; echo
done

#!/bin/bash
# dynamic tile order in the warp-per-tile kernel: parity, then A/B (short timeouts: a hang must not eat the budget)
mkdir -p gpurun_out
CBX_WIDE_DYNAMIC=1 timeout 150 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "chain100_factored or multi_scenario or int16" 2>&1 | tail -2
run() {  # workload dynamic
  CBX_WIDE_DYNAMIC=$2 timeout 60 python bench.py --no-cpu-baseline --workload $1 --envs-per-gpu 131072 --steps 100 --warmup 5 > gpurun_out/wdyn_$1_$2.log 2> gpurun_out/wdyn.err; rc=$?
  if [ $rc -ne 0 ]; then echo "$1 dynamic=$2 rc=$rc"; tail -3 gpurun_out/wdyn.err; return; fi
  python scripts/bench_line.py "$1 dynamic=$2" < gpurun_out/wdyn_$1_$2.log
}
run chain100 1; run chain100 0; run random16 1; run random16 0

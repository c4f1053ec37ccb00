#!/bin/bash
# dynamic tile order in the pipelined kernel: bounded lookahead sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "toyctf_marlon_pair_vs_oracle or step_host" 2>&1 | tail -2
CBX_PIPE_DYNAMIC=2 CBX_PIPE_LOOKAHEAD=4 timeout 300 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "toyctf_marlon_pair or step_host or tape" 2>&1 | tail -2
run() {  # n dynamic lookahead
  st=$((200 * 65536 / $1 + 20))
  CBX_PIPE_DYNAMIC=$2 CBX_PIPE_LOOKAHEAD=$3 timeout 300 python bench.py --no-cpu-baseline --envs-per-gpu $1 --steps $st --warmup 5 > gpurun_out/dyn_$1_$2_$3.log 2> gpurun_out/dyn.err; rc=$?
  python scripts/bench_line.py "envs=$1 dynamic=$2 lookahead=$3 rc=$rc" < gpurun_out/dyn_$1_$2_$3.log
}
for n in 65536 131072; do
  run $n 0 0; run $n 2 0
  for la in 4 5 6 8; do run $n 2 $la; done
done
run 524288 2 0; run 524288 2 4; run 524288 2 6

#!/bin/bash
# warp-per-tile kernel: does it speed up per env when a warp runs many tiles (phases of the warps drift apart)?
mkdir -p gpurun_out
for n in 131072 262144 524288 1048576; do
  for w in chain100 random16; do
    echo "-- $w envs=$n"; timeout 300 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu $n 2>/dev/null | python scripts/bench_line.py q | head -1
  done
done

#!/bin/bash
# warp-per-tile kernel: odd warps start one game-logic phase late (so that one half of an SM's warps emits while the other plays)
export CBX_LIB=marlon_b200/libcbx_trace.so
for sg in 0 1; do
  for w in chain100 random16; do
    echo "-- $w stagger=$sg"; CBX_WIDE_STAGGER=$sg timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1
  done
done
echo "-- 1M chain100 stagger=1"; CBX_WIDE_STAGGER=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload chain100 --envs-per-gpu 1048576 2>/dev/null | python scripts/bench_line.py q | head -1
CBX_WIDE_STAGGER=1 WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_wide_trace.py 2>&1 | grep -E "^cta +(0|74) warp|last stamp"

#!/bin/bash
# warp-per-tile kernel: independent in-place loads batched (countdowns, snapshot, actions), next group's property lines prefetched
mkdir -p gpurun_out
for lib in "" marlon_b200/libcbx_w00.so marlon_b200/libcbx_w10.so marlon_b200/libcbx_w01.so ""; do
  for w in chain100 random16; do
    echo "-- $w lib=$lib"; env ${lib:+CBX_LIB=$lib} timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1
  done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or chain100 or multi or random or factored" 2>&1 | tail -3

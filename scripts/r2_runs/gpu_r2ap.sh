#!/bin/bash
# warp-per-tile kernel: L2 policies -- the tile's state prefetched with evict_last, the 16-byte row stores with evict_first
for lib in "" marlon_b200/libcbx_l2w.so; do
for w in chain100 random16 chain100_scan; do echo "-- $w lib=$lib"; env ${lib:+CBX_LIB=$lib} timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done
done
CBX_LIB=marlon_b200/libcbx_l2w.so WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -6
echo "-- 1M chain100 l2w"; CBX_LIB=marlon_b200/libcbx_l2w.so timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload chain100 --envs-per-gpu 1048576 2>/dev/null | python scripts/bench_line.py q | head -1

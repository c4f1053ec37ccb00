#!/bin/bash
# pipelined kernel: L2 cache policies on the bulk copies. bit0 state evict_last, bit1 masks evict_first, bit2 other
# observation stores + action loads evict_first, bit3 tables evict_last
mkdir -p gpurun_out
run() { echo "== $*"; env "${@:2}" timeout 300 python bench.py --steps $STEPS --warmup 20 --no-e2e --no-cpu-baseline --envs-per-gpu $1 > gpurun_out/q.log 2> gpurun_out/q.err || tail -3 gpurun_out/q.err; python scripts/bench_line.py q < gpurun_out/q.log | head -1; }
STEPS=200
for h in 3 7 15 6; do run 65536 CBX_L2_HINTS=$h; done
for h in 0 2 3 7; do run 131072 CBX_L2_HINTS=$h; done
STEPS=100
for h in 0 2 3 7; do run 262144 CBX_L2_HINTS=$h; done
STEPS=40
for h in 0 2 6 7; do run 1048576 CBX_L2_HINTS=$h; done
STEPS=200
for h in 0 7; do echo "-- toyctf_scan $h"; CBX_L2_HINTS=$h timeout 300 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --workload toyctf_scan 2>/dev/null | python scripts/bench_line.py q | head -1; done
# wide / fused kernels: st.global.cs for every observation store (libcbx_cs.so = -DCBX_STREAMING_STORES=1)
for lib in "" marlon_b200/libcbx_cs.so; do
  for w in chain100 random16; do
    echo "-- $w lib=$lib"; env ${lib:+CBX_LIB=$lib} timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1
  done
done

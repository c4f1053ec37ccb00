#!/bin/bash
# countdown tick + observation snapshot with four in-place loads in flight (on top of the bulk-copied actions)
for w in chain100 random16 chain100_scan; do echo "-- $w"; timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done
echo "-- toyctf"; timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py q | head -1
echo "-- 1m"; timeout 300 python bench.py --steps 60 --warmup 10 --no-e2e --no-cpu-baseline --envs-per-gpu 1048576 2>/dev/null | python scripts/bench_line.py q | head -1
WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -6

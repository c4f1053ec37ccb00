#!/bin/bash
for w in chain100 random16 chain100_scan; do echo "-- $w"; timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done

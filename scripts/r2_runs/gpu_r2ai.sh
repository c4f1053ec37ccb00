#!/bin/bash
# warp-per-tile kernel: the tile's actions as two bulk copies on a per-warp mbarrier
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_faces.py -x -q -k "wide or chain100 or multi or random or factored or tape or scenario" 2>&1 | tail -3
for w in chain100 random16 chain100_scan; do echo "-- $w"; timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done
WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -6

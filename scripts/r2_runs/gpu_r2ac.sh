#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do echo "-- 1M"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 1048576 --steps 60 2>/dev/null | python scripts/bench_line.py q | head -1; done
echo "-- 65536"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 200 2>/dev/null | python scripts/bench_line.py q | head -1
echo "-- live"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 200 --workload toyctf_live 2>/dev/null | python scripts/bench_line.py q | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "live or tape or overlap or dynamic" 2>&1 | tail -3

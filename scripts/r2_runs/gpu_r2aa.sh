#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
for w in random16 chain100; do echo "-- $w"; timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done
echo "-- toyctf"; timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py q | head -1
echo "-- toyctf_live"; timeout 300 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline --workload toyctf_live 2>/dev/null | python scripts/bench_line.py q | head -1
python - <<'PY'
import bench
from marlon_b200.batch import Batch
comp, cfg = bench.workload_config(workload="random16")
b = Batch(comp, cfg, [8192]*16); print(b.kernel_info()); b.close()
PY

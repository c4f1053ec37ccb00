#!/bin/bash
# warp-per-tile kernel: the wrappers' state words of the tile cached in shared memory for the duration of the tile
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_faces.py -x -q -k "wide or chain100 or multi or random or factored or tape or scenario" 2>&1 | tail -3
for hot in 1 0; do
for w in chain100 random16 chain100_scan; do echo "-- $w hot=$hot"; CBX_WIDE_HOT=$hot timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --workload $w --envs-per-gpu 131072 2>/dev/null | python scripts/bench_line.py q | head -1; done
done
WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -6
python - <<'PY'
import bench
from marlon_b200.batch import Batch
for w in ("chain100", "random16"):
    comp, cfg = bench.workload_config(workload=w)
    n = 131072 if not isinstance(comp, list) else [8192] * 16
    b = Batch(comp, cfg, n); print(w, b.kernel_info()["threads"], b.kernel_info()["smem_bytes"]); b.close()
PY

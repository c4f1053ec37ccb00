#!/bin/bash
# Quick iteration on a B200: smoke, the oracle-parity tests, a short bench (no CPU baseline).
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/smoke.log; if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_iter.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_iter.log
echo "== bench"; timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_iter.log 2> gpurun_out/bench_iter.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_iter.log').read().strip().splitlines()[-1])
    print("value %.4g env-steps/s  ms/step %.4f  kernel_ms %.4f  achieved %.1f GB/s  frac %.3f  e2e %.4g  clocks %s" % (
        d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_iter.err').read()[-2000:])
PY
echo "== phases"; timeout 300 python scripts/gpu_phases.py 2>&1 | tee gpurun_out/phases.log | tail -14

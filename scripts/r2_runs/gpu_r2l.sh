#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (wide-kernel subset)"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "chain100 or random or factored or generated" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
for w in chain100 random16; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.log | python scripts/bench_line.py $w | head -1; tail -c 600 gpurun_out/bench_$w.err
done
echo "== phases chain100"; WORKLOAD=chain100 ENVS=131072 timeout 300 python scripts/gpu_phases.py 2>&1 | tail -6

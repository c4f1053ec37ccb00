#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "chain100 or multi_scenario or random or universe" 2>&1 | tail -15
timeout 400 python bench.py --workload chain100 --envs-per-gpu 131072 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_chain100.log 2> gpurun_out/bench_chain100.err; echo rc=$?; tail -c 300 gpurun_out/bench_chain100.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_chain100.log").read().strip().splitlines()[-1])
print("chain100 value %.4g ms/step %.4f kernel %s %.4f ms achieved %.1f frac %.3f e2e %.4g launch %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"]["kernel_launch"]))
PY
timeout 400 python bench.py --workload random16 --envs-per-gpu 131072 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_random16.log 2> gpurun_out/bench_random16.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_random16.log").read().strip().splitlines()[-1])
print("random16 value %.4g ms/step %.4f kernel %s %.4f ms achieved %.1f frac %.3f launch %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["kernel_launch"]))
PY

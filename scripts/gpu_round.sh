#!/bin/bash
# One full confirmation pass on a B200: smoke, every GPU test, the default bench (with the CPU baseline leg), the
# reference arm, then the ncu evidence.
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
[ $rc -ne 0 ] && exit 1
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 600 gpurun_out/bench.err
echo "== bench --impl reference"; timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.log | cut -c1-600
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print("value %.4g env-steps/s  ms/step %.4f  kernel %s %.4f ms  achieved %.1f GB/s  frac %.3f  e2e %.4g  cpu %s  clocks %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"],
    d["e2e"]["value"], d.get("cpu_baseline"), d["clocks"]))
PY
bash scripts/gpu_ncu.sh
echo "== 1M envs on one GPU (index-width sanity + large-batch throughput)"
timeout 200 python bench.py --envs-per-gpu 1048576 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m.log 2> gpurun_out/bench_1m.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_1m.log').read().strip().splitlines()[-1])
    print("1M envs: value %.4g  kernel %.4f ms  achieved %.1f GB/s  frac %.3f  env_steps %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"], d["episode_stats"]["env_steps"]))
except Exception as e:
    print("1M run failed", e); print(open('gpurun_out/bench_1m.err').read()[-1500:])
PY
echo "== the other workloads (131 072 envs per GPU)"
for w in chain100 random16; do
  timeout 100 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?"
  python scripts/bench_line.py $w < gpurun_out/bench_$w.log
done
timeout 100 python bench.py --workload toyctf_scan --no-cpu-baseline > gpurun_out/bench_toyctf_scan.log 2> gpurun_out/bench_toyctf_scan.err; echo "rc=$?"
python scripts/bench_line.py toyctf_scan < gpurun_out/bench_toyctf_scan.log
rm -f gpurun_out/prof.ncu-rep gpurun_out/prof_src.csv

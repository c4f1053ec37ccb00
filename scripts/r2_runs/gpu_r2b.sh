#!/bin/bash
# Round 2, second GPU pass: the rewritten wide-kernel field phase and the overlapped launches of the pipelined kernel.
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for ov in 1 0; do
  echo "== toyctf, overlapped launches = $ov"
  CBX_PIPE_OVERLAP=$ov timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_ov$ov.log 2> gpurun_out/bench_ov$ov.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_ov$ov.err
  python scripts/bench_line.py ov$ov < gpurun_out/bench_ov$ov.log
done
echo "== toyctf 262144 envs, overlapped"; timeout 300 python bench.py --no-cpu-baseline --no-e2e --envs-per-gpu 262144 --steps 300 > gpurun_out/bench_256k.log 2> gpurun_out/bench_256k.err; python scripts/bench_line.py 256k < gpurun_out/bench_256k.log
for w in chain100 random16; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_$w.err
  python scripts/bench_line.py $w < gpurun_out/bench_$w.log
done
echo "== full default bench (e2e legs)"; timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -c 800 gpurun_out/bench.err
python scripts/bench_line.py full < gpurun_out/bench.log
ls -la gpurun_out | tail -8

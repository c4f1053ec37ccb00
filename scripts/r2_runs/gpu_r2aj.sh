#!/bin/bash
# the driver's scaling launch at N=8 and N=2 with everything in the line (config4, config5, all e2e variants)
mkdir -p gpurun_out
for N in 8 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/bench_${N}gpu.err | tail -c 1200
  python scripts/bench_line.py ${N}gpu < gpurun_out/bench_${N}gpu.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>/dev/null | cut -c1-400

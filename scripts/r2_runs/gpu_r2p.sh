#!/bin/bash
# 8 GPUs: the driver's scaling launch at N=8 (headline + config4 block + timed collective + e2e legs), then N=4
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 | cut -c1-150
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/bench_${N}gpu.err | tail -c 1200
  python scripts/bench_line.py ${N}gpu < gpurun_out/bench_${N}gpu.log
done

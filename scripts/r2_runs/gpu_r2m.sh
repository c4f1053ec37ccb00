#!/bin/bash
# 2 GPUs: the driver's scaling launch (bench.py under torchrun): main workload + config4 block + timed collective + e2e legs
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_2gpu.err
python scripts/bench_line.py 2gpu < gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 5 --warmup 3 > gpurun_out/bench_2gpu_ref.log 2> gpurun_out/bench_2gpu_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_2gpu_ref.log
echo "== 1 GPU chain100 / random16 with unrolled row expansion"
for w in chain100 random16; do
  timeout 200 python bench.py --workload $w --envs-per-gpu 131072 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.log | python scripts/bench_line.py $w | head -1; tail -c 600 gpurun_out/bench_$w.err
done

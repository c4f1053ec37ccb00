#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/w.log 2> gpurun_out/w.err; echo "rc=$?"; tail -c 2500 gpurun_out/w.err
python scripts/bench_line.py w < gpurun_out/w.log

#!/bin/bash
# host-buffer step: parity of the page-locked / int16 paths, then the e2e leg of the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "step_host or int16" 2>&1 | tail -8
for w in toyctf chain100; do
  extra=""; [ $w = chain100 ] && extra="--envs-per-gpu 131072 --steps 200"
  timeout 300 python bench.py --workload $w $extra --no-cpu-baseline > gpurun_out/host_$w.log 2> gpurun_out/host_$w.err; echo rc=$?
  tail -c 400 gpurun_out/host_$w.err
  python scripts/bench_line.py "$w" < gpurun_out/host_$w.log
  python -c "import json,sys; d=json.loads(open('gpurun_out/host_$w.log').read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['int32_actions'])"
done
CBX_HOST_RESULTS=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/host_nomirror.log 2>/dev/null
python -c "import json,sys; d=json.loads(open('gpurun_out/host_nomirror.log').read().strip().splitlines()[-1]); print('no mirror', d['e2e']['value'], d['e2e']['int32_actions'])"

#!/bin/bash
# Quick GPU pass: parity tests + bench lines of the given configs (no CPU baseline, no ncu).
# Usage: bash tools/gpu_quick.sh <tag> [configs...]
tag=${1:-q}; shift
cfgs=${@:-"cfg2 cfg1 cfg3 cfg5"}
out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for c in $cfgs; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_$c.json 2>> $out/bench.err
  python - $out/bench_$c.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['config']['workload'],'value %.4g'%d['value'],'ms %.4f'%d['ms_per_step'],'tiles/s %.4g'%d['tiles_per_s'],'e2e %.4g'%d['e2e']['value'],{k:round(v,4) for k,v in d['kernel_ms'].items()})
P
done

#!/bin/bash
# A/B of the claim policies (library variants) on the single-capture configurations: device time with cold L2 (bench) and acq_search latency, hot.
tag=${1:-r2tail}; shift; out=gpurun_out/$tag; mkdir -p $out
for v in "$@"; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 400 python bench.py --captures 16 --steps 5 --only cfg1,cfg3,cfg4 --no-cpu-baseline --no-cufft > $out/bench_$v.json 2>> $out/bench.err
  python - $out/bench_$v.json $v <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], ' '.join('%s cold %.2f us e2e %.2f us |' % (k, v['ms_per_step']*1e3, v['e2e']['ms_per_step']*1e3) for k,v in d.get('configs',{}).items()))
P
done
python tools/e2e_latency.py "$@" 2>&1 | tee $out/e2e.txt
tail -3 $out/bench.err

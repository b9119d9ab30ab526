#!/bin/bash
# A/B of K = 1 C/A search variants against the product: bitwise equality, cfg5 (128 captures) bench lines, acq_search latency.
# Usage: bash tools/gpu_ab_sp.sh <tag> <variant...>
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for v in "$@"; do
  timeout 100 python tools/ab_equal.py $v > $out/equal_$v.log 2>&1; echo "$v equal rc=$?"; grep -c "equal: True" $out/equal_$v.log; grep -v "equal: True" $out/equal_$v.log | tail -5
done
show() { python - "$@" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'cfg5 %d caps: ms %.4f tiles/s %.4g e2e_ms %.4f search_ms %.4f' % (d['config']['captures_total'], d['ms_per_step'], d['tiles_per_s'], d['e2e']['ms_per_step'], d['kernel_ms']['search']))
    for k,v in d.get('configs',{}).items():
        print('   ', k, 'ms %.5f tiles/s %.4g e2e_ms %.5f' % (v['ms_per_step'], v['tiles_per_s'], v['e2e']['ms_per_step']), {a:round(b,4) for a,b in v['kernel_ms'].items()})
except Exception as e:
    print(sys.argv[1], 'failed', e)
P
}
for v in product "$@"; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 150 python bench.py --captures 128 --steps 20 --no-cpu-baseline --no-cufft --only ${ONLY:-cfg1,cfg4} > $out/bench_$v.json 2>> $out/bench.err
  show $out/bench_$v.json $v
done
tail -3 $out/bench.err

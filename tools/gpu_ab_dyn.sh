#!/bin/bash
# A/B of the claimed-tile feed against the static stride (variant static_tiles): parity suite, cfg1..cfg4 bench entries, timeline.
tag=${1:-r2dyn}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for v in product static_tiles; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 400 python bench.py --captures 128 --steps 10 --no-cpu-baseline --no-cufft > $out/bench_$v.json 2>> $out/bench.err
  python - $out/bench_$v.json $v <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'cfg5 %d caps: ms %.4f tiles/s %.4g' % (d['config']['captures_total'], d['ms_per_step'], d['tiles_per_s']))
for k,v in d.get('configs',{}).items():
    print('   ', k, 'ms %.5f tiles/s %.4g e2e_ms %.5f' % (v['ms_per_step'], v['tiles_per_s'], v['e2e']['ms_per_step']), {a:round(b,4) for a,b in v['kernel_ms'].items()})
P
done
TRACE_DUMP=1 python tools/trace_timeline.py cfg3 cfg2 > $out/trace.txt 2>&1; grep -v "by CTA" $out/trace.txt
tail -3 $out/bench.err

#!/bin/bash
# Round 2: parity suite, the default bench line, quick A/B lines of library variants (reduced farm), optional reference arm.
# Usage: bash tools/gpu_round.sh <tag> [variants...]
tag=${1:-r2c}; shift
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
show() { python - "$@" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'cfg5 %d caps: value %.4g ms %.4f tiles/s %.4g e2e_ms %.4f same %s search_ms %.4f' % (d['config']['captures_total'], d['value'], d['ms_per_step'], d['tiles_per_s'], d['e2e']['ms_per_step'], d['device_equals_host_path'], d['kernel_ms']['search']))
    for k,v in d.get('configs',{}).items():
        print('   ', k, 'ms %.5f tiles/s %.4g e2e_ms %.5f frac %.3f' % (v['ms_per_step'], v['tiles_per_s'], v['e2e']['ms_per_step'], v['roofline']['frac']), {a:round(b,4) for a,b in v['kernel_ms'].items()}, v.get('device_equals_host_path'))
except Exception as e:
    print(sys.argv[1], 'failed', e)
P
}
if [ -z "$NO_DEFAULT" ]; then
  timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"; tail -2 $out/bench_default.err
  show $out/bench_default.json default
fi
for v in "$@"; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 400 python bench.py --captures 128 --steps 20 --no-cpu-baseline --no-cufft > $out/bench_$v.json 2>> $out/bench.err
  show $out/bench_$v.json $v
done
if [ -n "$REFARM" ]; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; cut -c1-300 $out/bench_reference.json
fi
tail -5 $out/bench.err 2>/dev/null

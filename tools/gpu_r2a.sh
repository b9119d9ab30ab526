#!/bin/bash
# Round-2 first light of the rewritten search kernels (folded best-Doppler pick, balanced K = 1 launch, mapped records):
# a short guarded parity subset first (a deadlock must not eat the box), then the suite, then bench lines.
tag=${1:-r2a}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native or golden or cfg1" > $out/pytest_first.log 2>&1
echo "first rc=$?"; tail -3 $out/pytest_first.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for c in cfg1 cfg5 cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_$c.json 2>> $out/bench.err
  python - $out/bench_$c.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['config']['workload'],'value %.4g'%d['value'],'ms %.5f'%d['ms_per_step'],'tiles/s %.4g'%d['tiles_per_s'],'e2e %.4g'%d['e2e']['value'],'e2e_ms %.5f'%d['e2e']['ms_per_step'],{k:round(v,4) for k,v in d['kernel_ms'].items()}, d['device_equals_host_path'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
P
done
for v in l1_strided devrec pdl0; do
  for c in cfg1 cfg4 cfg5; do
    ACQ_B200_LIB=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_${c}_$v.json 2>> $out/bench.err
    python - $out/bench_${c}_$v.json $v <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], d['config']['workload'],'value %.4g'%d['value'],'ms %.5f'%d['ms_per_step'],'e2e_ms %.5f'%d['e2e']['ms_per_step'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
P
  done
done
tail -5 $out/bench.err

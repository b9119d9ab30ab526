out=gpurun_out/r2mst; mkdir -p $out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_kernel or three_cta or capture_resident or cfg2" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for v in product l1_multi_cta; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 150 python bench.py --captures 128 --steps 10 --no-cpu-baseline --no-cufft --only cfg2 > $out/bench_$v.json 2>> $out/bench.err
  python - $out/bench_$v.json $v <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k,v in d.get('configs',{}).items():
    print(sys.argv[2], k, 'ms %.5f tiles/s %.4g e2e_ms %.5f' % (v['ms_per_step'], v['tiles_per_s'], v['e2e']['ms_per_step']), {a:round(b,4) for a,b in v['kernel_ms'].items()})
P
done
tail -3 $out/bench.err

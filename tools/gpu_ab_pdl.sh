for i in 1 2; do for v in product pdl1; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  ACQ_B200_LIB=$lib timeout 400 python bench.py --captures 128 --steps 20 --only cfg2,cfg3_k4 --no-cpu-baseline --no-cufft 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'cfg5x128 ms %.4f e2e %.4f |' % (d['ms_per_step'], d['e2e']['ms_per_step']), ' '.join('%s ms %.4f e2e %.4f |' % (k, v['ms_per_step'], v['e2e']['ms_per_step']) for k,v in d['configs'].items()))"
done; done

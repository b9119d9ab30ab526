#!/bin/bash
# Quick A/B of library variants on the reduced farm (cfg5 with 128 captures) and chosen configs, two passes.
# Usage: bash tools/gpu_ab_quick.sh "<configs or none>" variant...
only=$1; shift
for i in 1 2; do for v in "$@"; do
  lib=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  [ "$v" = product ] && lib=$PWD/flydog_sdr_gps_b200/csrc/libacq_b200.so
  extra="--only $only"; [ "$only" = none ] && extra="--no-configs"
  ACQ_B200_LIB=$lib timeout 400 python bench.py --captures 128 --steps 20 $extra --no-cpu-baseline --no-cufft 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'cfg5x128 ms %.4f e2e %.4f tiles/s %.4g |' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['tiles_per_s']), ' '.join('%s ms %.4f e2e %.4f |' % (k, v['ms_per_step'], v['e2e']['ms_per_step']) for k,v in d.get('configs',{}).items()))"
done; done

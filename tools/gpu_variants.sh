#!/bin/bash
# A/B of library variants built by tools/build_variants.py: parity subset + bench lines per variant.
# Usage: [CFGS="cfg3 cfg4"] [KSEL="e1b or cfg3"] bash tools/gpu_variants.sh <tag> <variant> [<variant> ...]
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for v in "$@"; do
  export ACQ_B200_LIB=$PWD/flydog_sdr_gps_b200/csrc/variants/libacq_b200_$v.so
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${KSEL:-cfg1 or cfg2 or cfg5 or golden or sweep}" > $out/pytest_$v.log 2>&1
  echo "$v: $(tail -1 $out/pytest_$v.log)"
  for cfg in ${CFGS:-cfg2 cfg5 cfg1}; do
    timeout 300 python bench.py --config $cfg --no-cpu-baseline > $out/bench_${cfg}_$v.json 2>> $out/bench.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${cfg}_$v.json").read().strip().splitlines()[-1])
    print("  $cfg $v", round(d["tiles_per_s"]/1e6,2), "Mtiles/s", round(d["value"]/1e9,2), "Gcells/s search_ms", round(d["kernel_ms"]["search"],4), "same", d["device_equals_host_path"])
except Exception as e:
    print("  $cfg $v failed", e)
PY
  done
done

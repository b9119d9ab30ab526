#!/bin/bash
# A/B of the C/A search kernels on one box: parity tests, then bench lines for both forms.
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for cfg in cfg2 cfg1 cfg5; do
  for k in tma ldg; do
    ACQ_L1_KERNEL=$k timeout 300 python bench.py --config $cfg --no-cpu-baseline > $out/bench_${cfg}_$k.json 2>> $out/bench.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${cfg}_$k.json").read().strip().splitlines()[-1])
    print("$cfg $k", round(d["tiles_per_s"]/1e6,2), "Mtiles/s", round(d["value"]/1e9,2), "Gcells/s search_ms", round(d["kernel_ms"]["search"],4), "same", d["device_equals_host_path"])
except Exception as e:
    print("$cfg $k failed", e)
PY
  done
done

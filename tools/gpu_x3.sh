#!/bin/bash
# A/B of the C/A search kernels: two CTAs per SM (default) against three (ACQ_L1_KERNEL=x3).
tag=${1:-x3}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "three_cta" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for k in tma x3 tma x3; do
  for cfg in cfg2 cfg5 cfg1; do
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

#!/bin/bash
# A/B of programmatic dependent launch between the kernels of a search (ACQ_PDL=1|0): full parity suite with PDL
# on, then bench lines for the single-capture configurations where the launch gaps matter.
tag=${1:-pdl_ab}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for pdl in 1 0; do
  for c in ${CFGS:-cfg1 cfg3 cfg4 cfg2}; do
    ACQ_PDL=$pdl timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_${c}_pdl$pdl.json 2>> $out/bench.err
    python - <<PY
import json
try:
    d=json.loads(open("$out/bench_${c}_pdl$pdl.json").read().strip().splitlines()[-1])
    print("$c pdl=$pdl", "%.2f Gcells/s" % (d["value"]/1e9), "%.4f ms/step" % d["ms_per_step"], "(with events %.4f)" % d["kernel_ms_pass"]["ms_per_step"],
          "e2e %.2f G %.4f ms" % (d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]), "same", d["device_equals_host_path"])
except Exception as e:
    print("$c pdl=$pdl failed", e)
PY
  done
done
tail -5 $out/bench.err

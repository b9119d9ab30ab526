#!/bin/bash
# A/B of the one-CTA E1B search kernels on the GPU box: parity tests, then cfg3 / cfg4 bench lines per kernel.
tag=${1:-e1b_ab}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "e1b or cfg3 or cfg4 or golden" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for k in tma ldg; do
  for c in cfg3 cfg4; do
    ACQ_E1B_CTA_KERNEL=$k timeout 300 python bench.py --config $c --no-cpu-baseline > $out/bench_${c}_$k.json 2>> $out/bench.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*/bench_cfg[34]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1f Gcells/s" % (d["value"] / 1e9), "%.4f ms" % d["ms_per_step"], d["kernel_ms"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(f, "unreadable", e)
PY

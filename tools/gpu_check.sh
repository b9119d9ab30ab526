#!/bin/bash
# Parity suite + default bench line (no CPU baseline) in one call.  Usage: bash tools/gpu_check.sh <tag>
tag=${1:-chk}; out=gpurun_out/$tag; mkdir -p $out
timeout 700 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-cufft > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -2 $out/bench.err
python - $out/bench.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg5 ms %.3f tiles/s %.4g e2e %.3f frac %.3f pipe %s"%(d["ms_per_step"],d["tiles_per_s"],d["e2e"]["ms_per_step"],d["roofline"]["frac"],d["roofline"].get("frac_pipe_measured")), d["roofline"]["kernel"])
for k,v in d["configs"].items(): print(k,"ms %.5f e2e %.5f tiles/s %.4g"%(v["ms_per_step"],v["e2e"]["ms_per_step"],v["tiles_per_s"]), {a:round(b,4) for a,b in v["kernel_ms"].items()})
P

#!/bin/bash
# Full ncu capture of the one-CTA E1B search kernel on cfg3.  Usage (under gpurun): bash tools/gpu_ncu_e1b.sh <tag>
tag=${1:-ncu_e1b}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search_e1b -s 3 -c 1 -f -o $out/prof_search_e1b \
    python bench.py --config cfg3 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log

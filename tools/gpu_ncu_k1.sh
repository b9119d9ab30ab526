#!/bin/bash
# Full ncu captures of the K = 1 C/A kernel (cfg5) and the one-CTA E1B kernel (cfg3) at HEAD.
tag=${1:-ncu_k1}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search_l1 -s 3 -c 1 -f -o $out/prof_search_l1_k1 \
    python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_k1.log 2>&1; tail -2 $out/ncu_k1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search_e1b -s 3 -c 1 -f -o $out/prof_search_e1b \
    python bench.py --config cfg3 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_e1b.log 2>&1; tail -2 $out/ncu_e1b.log

#!/bin/bash
# ncu evidence of a round: launch list of the default bench command, one --set full capture per product search kernel.
# Usage: bash tools/gpu_ncu.sh <tag>
tag=${1:-ncu}; out=gpurun_out/$tag; mkdir -p $out
# launch list of the default bench command (our kernels + NCCL + the L2 flush fill; the synthetic-capture generator's
# torch kernels are left out by the name filter)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|ncclDevKernel|FillFunctor" -c 2000 --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cufft > $out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
cap() { # name regex cfg [captures]
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o $out/prof_$1 \
      python tools/ncu_driver.py $3 $4 > $out/ncu_$1.log 2>&1; echo "$1 rc=$?"; tail -1 $out/ncu_$1.log
}
cap l1_k1_cfg5 'k_search_l1' cfg5 128
cap l1_multi_cfg2 'k_search_l1' cfg2
cap e1b_cfg3 'k_search_e1b' cfg3
cap l1_k1_cfg1 'k_search_l1' cfg1
cap e1b_multi_cfg3k4 'k_search_e1b_multi' cfg3_k4
ls -la $out | head -20

#!/bin/bash
# One GPU-box pass: parity tests, default bench, launch list and a full ncu capture of the hot kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 600 python bench.py --cufft > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --config cfg1 --no-cpu-baseline > $out/bench_cfg1.json 2>> $out/bench_default.err
timeout 300 python bench.py --config cfg3 --no-cpu-baseline > $out/bench_cfg3.json 2>> $out/bench_default.err
timeout 300 python bench.py --config cfg5 --no-cpu-baseline > $out/bench_cfg5.json 2>> $out/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_search_l1 -s 3 -c 1 -f -o $out/prof_search_l1 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out

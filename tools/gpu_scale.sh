#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): default bench, cfg5 and cfg4 at N ranks, and the reference arm under torchrun.
# Usage: bash tools/gpu_scale.sh <N> <tag>
n=${1:-2}; tag=${2:-scale}
out=gpurun_out/$tag
mkdir -p $out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
run 29511 --steps 20 --warmup 5 > $out/bench_default_n$n.json 2> $out/err_default.log; echo "default rc=$?"
run 29512 --steps 20 --warmup 5 --config cfg5 --no-cpu-baseline > $out/bench_cfg5_n$n.json 2> $out/err_cfg5.log; echo "cfg5 rc=$?"
run 29513 --steps 20 --warmup 5 --config cfg4 --no-cpu-baseline > $out/bench_cfg4_n$n.json 2> $out/err_cfg4.log; echo "cfg4 rc=$?"
run 29514 --steps 2 --warmup 1 --impl reference > $out/bench_reference_n$n.json 2> $out/err_ref.log; echo "ref rc=$?"
tail -c 600 $out/bench_default_n$n.json; echo; tail -c 300 $out/err_default.log

#!/bin/bash
# Multi-GPU pass (under gpurun --gpus N): the default bench line at N ranks (cfg5 strong-scaled, cfg1..4 PRN-sharded)
# and the reference arm under torchrun.
# Usage: bash tools/gpu_scale.sh <N> <tag>
n=${1:-2}; tag=${2:-scale}
out=gpurun_out/$tag
mkdir -p $out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
run 29511 > $out/bench_default_n$n.json 2> $out/err_default.log; echo "default rc=$?"
run 29514 --steps 2 --warmup 1 --impl reference > $out/bench_reference_n$n.json 2> $out/err_ref.log; echo "ref rc=$?"
tail -c 600 $out/bench_default_n$n.json; echo; tail -c 300 $out/err_default.log

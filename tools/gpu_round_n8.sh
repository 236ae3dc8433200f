#!/bin/bash
# 8-GPU evidence (one call): default bench (cfg2 DP + cfg5 sharded + ViT-SOM DP + parity check), DP timeline, cfg4 / cfg3 DP
set -u
TAG=${1:-r02n8}; N=${2:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N" > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "bench n$N rc=$?"
tail -c 7000 gpurun_out/bench_n${N}_${TAG}.json; grep -v "^W\|^$\|^\*\|OMP" gpurun_out/bench_n${N}_${TAG}.err | tail -15
timeout 300 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --no-extras --dp-overlap split" > gpurun_out/bench_n${N}_split_${TAG}.json 2> gpurun_out/bench_n${N}_split_${TAG}.err; echo "bench split rc=$?"
tail -c 1200 gpurun_out/bench_n${N}_split_${TAG}.json
for w in cfg4 cfg3; do
  timeout 300 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --no-extras --workload $w" > gpurun_out/bench_n${N}_${w}_${TAG}.json 2> gpurun_out/bench_n${N}_${w}_${TAG}.err; echo "bench $w rc=$?"
  tail -c 1200 gpurun_out/bench_n${N}_${w}_${TAG}.json; grep -v "^W\|^$\|^\*\|OMP" gpurun_out/bench_n${N}_${w}_${TAG}.err | tail -5
done

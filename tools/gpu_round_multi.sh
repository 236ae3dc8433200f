#!/bin/bash
# Multi-GPU evidence:  gpurun --gpus N -- 'bash tools/gpu_round_multi.sh <tag> N [stage ...]'    stages: tests bench modes cfg
set -u
TAG=${1:-r02m}; N=${2:-2}; shift 2 || true
STAGES=${*:-tests bench}
mkdir -p gpurun_out
has() { [[ " $STAGES " == *" $1 "* ]]; }
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
if has tests; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/pytest_multi_${TAG}.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/pytest_multi_${TAG}.log
fi
if has bench; then
  timeout 900 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N" > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "bench n$N rc=$?"
  tail -c 7000 gpurun_out/bench_n${N}_${TAG}.json; grep -v "^W\|^$" gpurun_out/bench_n${N}_${TAG}.err | tail -15
fi
if has modes; then
  for m in split after; do
    timeout 600 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --no-extras --dp-overlap $m" > gpurun_out/bench_n${N}_${m}_${TAG}.json 2> gpurun_out/bench_n${N}_${m}_${TAG}.err; echo "bench $m rc=$?"
    tail -c 1500 gpurun_out/bench_n${N}_${m}_${TAG}.json; grep -v "^W\|^$" gpurun_out/bench_n${N}_${m}_${TAG}.err | tail -5
  done
fi
if has cfg; then
  for w in cfg3 cfg4; do
    timeout 600 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --no-extras --workload $w" > gpurun_out/bench_n${N}_${w}_${TAG}.json 2> gpurun_out/bench_n${N}_${w}_${TAG}.err; echo "bench $w rc=$?"
    tail -c 1500 gpurun_out/bench_n${N}_${w}_${TAG}.json; grep -v "^W\|^$" gpurun_out/bench_n${N}_${w}_${TAG}.err | tail -5
  done
fi
ls -la gpurun_out | tail -6

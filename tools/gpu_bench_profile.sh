#!/bin/bash
# Round bench + ncu evidence, run on the GPU box:  gpurun -- 'bash tools/gpu_bench_profile.sh r01'
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:som_gemm3x -s 9 -c 3 -f -o gpurun_out/prof_gemm_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12

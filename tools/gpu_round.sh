#!/bin/bash
# Round evidence on the GPU box:  gpurun -- 'bash tools/gpu_round.sh <tag>'
# tests -> bench -> per-kernel step timing -> ncu launch list -> ncu full capture of the hot kernels.
set -u
TAG=${1:-r01b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}.json
python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg2_${TAG}.log
python tools/step_time.py 4096 128 128 256 > gpurun_out/steptime_cfg5_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg5_${TAG}.log
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:som_gemm3x|loss_coeffs|prep_rows' -s 16 -c 6 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -8

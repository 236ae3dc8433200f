#!/bin/bash
# ncu evidence of one round:  gpurun -- 'bash tools/ncu_round.sh <tag>'   (each capture only after the same command exited 0 without ncu)
set -u
TAG=${1:-r02v}
mkdir -p gpurun_out
# the fused loss kernel at a config-5 chunk (HBM-bound): full capture of one launch
python tools/step_time.py 8192 128 128 256 > gpurun_out/steptime_cfg5_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:loss_coeffs' -s 3 -c 1 -f -o gpurun_out/prof_loss_cfg5_${TAG} python tools/step_time.py 8192 128 128 256 > gpurun_out/ncu_loss_${TAG}.log 2>&1
echo "ncu loss rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:som_gemm3x|loss_coeffs|prep_rows|bmu_decode' -s 20 -c 5 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -6

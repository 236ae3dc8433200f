set -u
TAG=${1:-r01e}
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:som_gemm3x|loss_coeffs|prep_rows' -s 16 -c 6 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | grep ${TAG}

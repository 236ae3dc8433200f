set -u
TAG=${1:-v24}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
timeout 120 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg2_${TAG}.log
timeout 120 python tools/step_time.py 4096 128 128 256 > gpurun_out/steptime_cfg5_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg5_${TAG}.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac_of_3xtf32_bound'], d['e2e'])
PY
timeout 200 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; echo "bench c5 rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac_of_3xtf32_bound'], d['e2e'])
PY

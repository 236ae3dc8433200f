set -u
TAG=${1:-v34}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${TAG}.log
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
timeout 200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_${TAG}.json
timeout 120 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1
timeout 100 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; echo "bench c5 rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac_of_3xtf32_bound'])
PY

set -u
TAG=${1:-v28}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
timeout 120 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; grep -v "^$" gpurun_out/steptime_cfg2_${TAG}.log | cut -c1-150
timeout 200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac'], d['roofline']['frac_of_3xtf32_bound'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
PY
timeout 200 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err; echo "bench c5 rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_c5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac_of_3xtf32_bound'], d['e2e']['value'])
PY
timeout 100 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_${TAG}_ref.json

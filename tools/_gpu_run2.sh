set -u
TAG=${1:-v18}
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py --cg 2 --sk 1 --only p_ --out gpurun_out/probe_sk_${TAG}.json > gpurun_out/probe_sk_${TAG}.log 2>&1; echo "probe sk rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/probe_sk_${TAG}.json'))
print({k:(v.get('ok'), v.get('rel_fro')) for k,v in d.items()})
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_${TAG}.log
timeout 120 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg2_${TAG}.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['per_gemm_ms'], d['roofline']['frac_of_3xtf32_bound'], d['e2e']['value'])
PY

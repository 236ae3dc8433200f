set -u
TAG=${1:-v15}
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py --cg 2 --sk 1 --only p_ --out gpurun_out/probe_sk_${TAG}.json > gpurun_out/probe_sk_${TAG}.log 2>&1; echo "probe sk rc=$?"
timeout 300 python tools/gpu_probe.py --cg 2 --sk -1 --only p_ --out gpurun_out/probe_nosk_${TAG}.json > gpurun_out/probe_nosk_${TAG}.log 2>&1; echo "probe nosk rc=$?"
python - <<PY
import json
for f in ('gpurun_out/probe_sk_${TAG}.json','gpurun_out/probe_nosk_${TAG}.json'):
    d=json.load(open(f))
    for k,v in d.items(): print(f[-14:], k, v.get('ok'), v.get('rel_fro'), v.get('mean_signed_rel'), (v.get('error') or '')[:200] if isinstance(v.get('error'),str) else '')
PY
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_${TAG}.log
timeout 120 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg2_${TAG}.log
timeout 120 python tools/step_time.py 4096 128 128 256 > gpurun_out/steptime_cfg5_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg5_${TAG}.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err

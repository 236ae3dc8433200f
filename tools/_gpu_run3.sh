set -u
TAG=${1:-v31}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 -k "data_parallel" > gpurun_out/pytest_multi_${TAG}.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/pytest_multi_${TAG}.log
for sms in 132 136; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --gemm-sms $sms > gpurun_out/bench_${TAG}_n${N}_sms$sms.json 2> gpurun_out/bench_${TAG}_n${N}_sms$sms.err; echo "bench n$N sms=$sms rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${TAG}_n${N}_sms$sms.json') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_gemm_ms'])
except Exception as e: print('no json', e)
PY
done

set -u
TAG=${1:-v26}
N=${2:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s --timeout 300 > gpurun_out/pytest_multi_${TAG}.log 2>&1; echo "pytest multi rc=$?"; grep -E "exchange path|passed|failed|Error|error" gpurun_out/pytest_multi_${TAG}.log | head -20
for nv in 1 0; do
SOM_DP_NVLS=$nv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c5_n${N}_nvls$nv.json 2> gpurun_out/bench_${TAG}_c5_n${N}_nvls$nv.err; echo "bench c5 n$N nvls=$nv rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${TAG}_c5_n${N}_nvls$nv.json') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('no json', e)
PY
grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_${TAG}_c5_n${N}_nvls$nv.err | tail -3
done

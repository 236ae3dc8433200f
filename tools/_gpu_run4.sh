set -u
TAG=${1:-v29}
N=${2:-8}
mkdir -p gpurun_out
run() { # name, env, args
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${TAG}_${name}.json 2> gpurun_out/bench_${TAG}_${name}.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_${TAG}_${name}.json') if l.startswith('{')][-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_gemm_ms'])
except Exception as e: print('no json', e)
PY
  grep -v "OMP_NUM\|\*\*\*\*\|^$" gpurun_out/bench_${TAG}_${name}.err | tail -3
}
run c2_n${N}_nvls SOM_DP_NVLS=1 --steps 50 --warmup 5 --no-cpu-baseline
run c5_n${N}_nvls SOM_DP_NVLS=1 --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline

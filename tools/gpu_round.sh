#!/bin/bash
# Round evidence on the GPU box:  gpurun -- 'bash tools/gpu_round.sh <tag> [stage ...]'
# stages: smoke tests bench steptime small tf32 ncu   (default: all but tf32 and ncu)
set -u
TAG=${1:-r02a}; shift || true
STAGES=${*:-smoke tests bench steptime small}
mkdir -p gpurun_out
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_${TAG}.log
fi
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_${TAG}.log
fi
if has bench; then
  timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
fi
if has steptime; then
  timeout 300 python tools/step_time.py > gpurun_out/steptime_cfg2_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg2_${TAG}.log
  timeout 300 python tools/step_time.py 1024 40 40 3136 euclidean tf32x3 > gpurun_out/steptime_cfg2_tf32_${TAG}.log 2>&1; grep -E "GEMM|sum" gpurun_out/steptime_cfg2_tf32_${TAG}.log
  timeout 300 python tools/step_time.py 8192 128 128 256 > gpurun_out/steptime_cfg5chunk_${TAG}.log 2>&1; cat gpurun_out/steptime_cfg5chunk_${TAG}.log
  timeout 300 python tools/step_time.py 128 4 4 12288 cosine > gpurun_out/steptime_cfg3_${TAG}.log 2>&1
fi
if has tf32; then
  timeout 600 python bench.py --precision tf32x3 --no-vit --no-cpu-baseline > gpurun_out/bench_tf32_${TAG}.json 2> gpurun_out/bench_tf32_${TAG}.err; echo "bench tf32x3 rc=$?"; tail -c 1500 gpurun_out/bench_tf32_${TAG}.json
fi
if has small; then
  for w in cfg1 cfg3 cfg4; do
    timeout 600 python bench.py --workload $w --no-extras > gpurun_out/bench_${w}_${TAG}.json 2> gpurun_out/bench_${w}_${TAG}.err; echo "bench $w rc=$?"; tail -c 1800 gpurun_out/bench_${w}_${TAG}.json
  done
fi
if has ncu; then
  SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras"
  $SHORT > gpurun_out/plain_${TAG}.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $SHORT > gpurun_out/ncu_list_${TAG}.log 2>&1
  echo "ncu list rc=$?"
  $SHORT > gpurun_out/plain2_${TAG}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k 'regex:som_gemm3x|loss_coeffs|prep_rows|bmu_decode' -s 20 -c 5 -f -o gpurun_out/prof_${TAG} $SHORT > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "ncu full rc=$?"
fi
ls -la gpurun_out | tail -8

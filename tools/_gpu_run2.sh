set -u
TAG=${1:-v32}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
timeout 200 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_${TAG}.json

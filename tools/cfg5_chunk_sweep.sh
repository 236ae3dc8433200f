set -u
mkdir -p gpurun_out
for c in 4096 8192 16384 32768; do
  timeout 300 python bench.py --workload cfg5 --no-extras --no-cpu-baseline --steps 10 --row-chunk $c > gpurun_out/bench_cfg5_chunk${c}_r02p.json 2> gpurun_out/bench_cfg5_chunk${c}_r02p.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg5_chunk${c}_r02p.json"))
print("chunk", ${c}, "ms/step", round(d["ms_per_step"],3), "samples/s", round(d["value"]), "whole", round(d["roofline"]["whole_step"]["frac_of_3xtf32_bound"],3), d["roofline"]["per_gemm_ms"], "e2e", round(d["e2e"]["value"]))
PY
done

#!/bin/bash
# quick check after scheduler changes: tests, bench, and the grid sizes of the GEMM launches (148 = all 74 CTA pairs)
bash tools/gpu_round.sh ${1:-r02u} smoke tests bench
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${1:-r02u}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras > /dev/null 2>&1
grep gemm3x gpurun_out/launches_${1:-r02u}.csv | tail -2 | cut -d, -f5-9

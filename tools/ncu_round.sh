set -u
mkdir -p gpurun_out
python tools/step_time.py 128 4 4 12288 > gpurun_out/steptime_cfg3_r02h.log 2>&1; cat gpurun_out/steptime_cfg3_r02h.log
python tools/step_time.py 4096 128 128 256 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:loss_coeffs|adamw' -c 4 -f -o gpurun_out/prof_loss_cfg5_r02h python tools/step_time.py 4096 128 128 256 > gpurun_out/ncu_loss_r02h.log 2>&1
echo "ncu loss rc=$?"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-extras"
$SHORT > gpurun_out/plain_r02h.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02h.csv $SHORT > gpurun_out/ncu_list_r02h.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:som_gemm3x|loss_coeffs|prep_rows' -s 16 -c 6 -f -o gpurun_out/prof_r02h $SHORT > gpurun_out/ncu_full_r02h.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -5

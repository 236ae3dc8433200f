for args in "counter 136" "counter 132" "counter 128" "split 136"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) tools/dp_probe.py $args 2>&1 | grep "^rank 0"
done

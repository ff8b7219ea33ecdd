set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/dist_check8.log 2>&1; grep -E "rank|DIST|Error|error" gpurun_out/dist_check8.log | head -12
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; tail -1 gpurun_out/scale_$n.err | cut -c1-200
done

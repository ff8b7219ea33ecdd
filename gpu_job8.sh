# 8-GPU bench line of the C4 workload (the sharded = unsharded check is tests/test_gpu_dist.py, run by pytest -m gpu)
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err; tail -2 gpurun_out/scale_8.err

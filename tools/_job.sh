set -x
python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/pytest_dist_oz.log 2>&1; tail -12 gpurun_out/pytest_dist_oz.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/b_oz_2gpu.json 2> gpurun_out/b_oz_2gpu.err; tail -2 gpurun_out/b_oz_2gpu.err

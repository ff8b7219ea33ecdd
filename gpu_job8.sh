# 8-GPU check: sharded = unsharded (mini, MRSBlurred, C4) and the C4 bench line at 8 ranks
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err; tail -2 gpurun_out/scale_8.err
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/pytest_dist8.log 2>&1; tail -14 gpurun_out/pytest_dist8.log

set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err; tail -2 gpurun_out/scale_8.err
$TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err; tail -2 gpurun_out/scale_2.err
$TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 > gpurun_out/scale_4.json 2> gpurun_out/scale_4.err; tail -2 gpurun_out/scale_4.err
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --impl reference --steps 1 --warmup 0 > gpurun_out/scale_ref8.json 2> gpurun_out/scale_ref8.err; tail -2 gpurun_out/scale_ref8.err
python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/pytest_dist8.log 2>&1; tail -14 gpurun_out/pytest_dist8.log

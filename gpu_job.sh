set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s1_c4.json 2> gpurun_out/s1_c4.err; tail -2 gpurun_out/s1_c4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c4 --steps 5 --warmup 3 > gpurun_out/s2_c4.json 2> gpurun_out/s2_c4.err; tail -5 gpurun_out/s2_c4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_check.py > gpurun_out/dist_check.log 2>&1; tail -5 gpurun_out/dist_check.log

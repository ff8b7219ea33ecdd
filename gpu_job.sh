set -x
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/t_c4.json 2> gpurun_out/t_c4.err; tail -2 gpurun_out/t_c4.err

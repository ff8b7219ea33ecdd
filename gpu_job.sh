set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v6_c4.json 2> gpurun_out/v6_c4.err; tail -2 gpurun_out/v6_c4.err

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v5_c4.json 2> gpurun_out/v5_c4.err; tail -2 gpurun_out/v5_c4.err
SURFH_FFT_PRUNE=0 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v5_c4_noprune.json 2> gpurun_out/v5_c4_noprune.err; tail -2 gpurun_out/v5_c4_noprune.err
python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v5_c2.json 2> gpurun_out/v5_c2.err; tail -2 gpurun_out/v5_c2.err

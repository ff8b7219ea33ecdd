python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_oz8.log 2>&1; tail -3 gpurun_out/pytest_oz8.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz8.json 2> gpurun_out/b_oz8.err; tail -1 gpurun_out/b_oz8.err
SURFH_OZAKI_MMAJOR=1 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz8m.json 2> gpurun_out/b_oz8m.err; tail -1 gpurun_out/b_oz8m.err

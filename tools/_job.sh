tools/_build/ozaki_test 1050 1764 3440 8 10 2 2>&1 | tail -4
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_oz3.log 2>&1; tail -3 gpurun_out/pytest_oz3.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz3.json 2> gpurun_out/b_oz3.err; tail -2 gpurun_out/b_oz3.err

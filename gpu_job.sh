set -x
python -m pytest tests/test_gpu_fft.py tests/test_gpu_parity.py tests/test_blind.py -q -x -m gpu > gpurun_out/pytest_par.log 2>&1; tail -4 gpurun_out/pytest_par.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_rows.json 2> gpurun_out/b_rows.err; tail -1 gpurun_out/b_rows.err
python bench.py --config c5 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_rows_c5.json 2> gpurun_out/b_rows_c5.err; tail -1 gpurun_out/b_rows_c5.err

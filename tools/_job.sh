python -m pytest tests/test_gpu_parity.py tests/test_blind.py -m gpu -q -x > gpurun_out/pytest_oz4.log 2>&1; tail -3 gpurun_out/pytest_oz4.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_lanes.json 2> gpurun_out/b_lanes.err; tail -2 gpurun_out/b_lanes.err
SURFH_STREAMS=1 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_lanes1.json 2> gpurun_out/b_lanes1.err; tail -2 gpurun_out/b_lanes1.err

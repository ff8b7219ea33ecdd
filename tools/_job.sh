set -x
python -m pytest tests/test_gpu_parity.py tests/test_blind.py tests/test_gpu_cg.py -m gpu -q -x > gpurun_out/pytest_oz2.log 2>&1; tail -15 gpurun_out/pytest_oz2.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz2.json 2> gpurun_out/b_oz2.err; tail -2 gpurun_out/b_oz2.err
python bench.py --dtype float32 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz2_f32.json 2> gpurun_out/b_oz2_f32.err; tail -2 gpurun_out/b_oz2_f32.err

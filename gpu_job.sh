set -x
python -m pytest tests/test_distorsion_correction.py tests/test_gpu_fft.py -q -m gpu > gpurun_out/pytest_dc.log 2>&1; tail -4 gpurun_out/pytest_dc.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_chk.json 2> gpurun_out/b_chk.err; tail -1 gpurun_out/b_chk.err

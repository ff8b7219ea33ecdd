set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/pytest_par.log 2>&1; tail -6 gpurun_out/pytest_par.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_tma.json 2> gpurun_out/b_tma.err; tail -1 gpurun_out/b_tma.err
SURFH_F64_GEMM=mma timeout 300 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_mma.json 2> gpurun_out/b_mma.err; tail -1 gpurun_out/b_mma.err
timeout 300 python bench.py --config c2 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b2_tma.json 2> gpurun_out/b2_tma.err; tail -1 gpurun_out/b2_tma.err

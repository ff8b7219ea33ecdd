set -x
python -m pytest tests/test_gpu_parity.py tests/test_blind.py -q -x -m gpu > gpurun_out/pytest_par.log 2>&1; tail -4 gpurun_out/pytest_par.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_merge.json 2> gpurun_out/b_merge.err; tail -1 gpurun_out/b_merge.err
SURFH_B200_LIB=$PWD/surfh_b200/libsurfh_nomerge.so python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_nomerge.json 2> gpurun_out/b_nomerge.err; tail -1 gpurun_out/b_nomerge.err
python bench.py --dtype float32 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_merge_f32.json 2> gpurun_out/b_merge_f32.err; tail -1 gpurun_out/b_merge_f32.err

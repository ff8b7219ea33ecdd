# quick GPU sanity of the current tree: parity suites + smoke + a short bench line
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py tests/test_blind.py -m gpu -q -x > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --no-cpu-baseline --solve-iters 0 > gpurun_out/b_quick.json 2> gpurun_out/b_quick.err; tail -1 gpurun_out/b_quick.err

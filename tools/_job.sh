T=tools/_build/ozaki_test
for args in "300 500 700 8 1 2" "4200 7056 3440 8 5 2" "13760 7056 1050 8 5 2"; do
  echo "=== $args"; timeout 120 $T $args 2>&1 | grep -v "^digits"; echo "rc=$?"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_oz7.log 2>&1; tail -3 gpurun_out/pytest_oz7.log
python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/b_oz7.json 2> gpurun_out/b_oz7.err; tail -2 gpurun_out/b_oz7.err

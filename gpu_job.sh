# the round's standard check: GPU parity suite, smoke, default bench line, reference arm
set -x
python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/pytest_gpu.log 2>&1; tail -16 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err

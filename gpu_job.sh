# the round's standard check: GPU parity suite, smoke, default bench line, reference arm, the other configurations,
# one ncu --set full capture of the tcgen05 contraction (every step under its own timeout)
set -x
timeout 900 python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err
timeout 300 python bench.py --dtype float32 --no-cpu-baseline > gpurun_out/bench_c4_f32.json 2> gpurun_out/bench_c4_f32.err; tail -1 gpurun_out/bench_c4_f32.err
for c in c2 c3 c5; do
timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -1 gpurun_out/bench_$c.err
done

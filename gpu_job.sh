# the round's standard check: GPU parity suite, smoke, default bench line, reference arm; then the ncu passes of profiles/
set -x
python -m pytest tests -m gpu -q -s --durations=5 > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 601 -c 114 --csv --log-file gpurun_out/r02_app_metrics_c4.csv python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/ncu_app.log 2>&1
tail -2 gpurun_out/ncu_app.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/ncu_l.log 2>&1

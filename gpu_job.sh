set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -s 589 -c 102 --csv --log-file gpurun_out/r02_app_metrics_c4.csv python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --solve-iters 0 > gpurun_out/ncu_app.log 2>&1
tail -2 gpurun_out/ncu_app.log
ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 12 -c 2 -o gpurun_out/r02_fft_c2r_final python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 22 -c 2 -o gpurun_out/r02_fft_r2c_final python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_f2.log 2>&1
for c in c2 c3 c5; do python bench.py --config $c --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; tail -1 gpurun_out/bench_$c.err; done
python bench.py --dtype float32 --no-cpu-baseline > gpurun_out/bench_c4_f32.json 2> gpurun_out/bench_c4_f32.err; tail -1 gpurun_out/bench_c4_f32.err
python -m pytest tests/test_gpu_cg.py -q -m gpu -s -k "huber" > gpurun_out/pytest_huber.log 2>&1; tail -5 gpurun_out/pytest_huber.log

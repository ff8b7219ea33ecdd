set -x
CMD="python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
$CMD > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err && timeout 600 ncu --metrics $M --clock-control none -s 589 -c 102 --csv --log-file gpurun_out/app_metrics_c4.csv $CMD > gpurun_out/ncu_m.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 110 -c 2 -o gpurun_out/prof_c4_fft_c2r -f $CMD > gpurun_out/ncu_f1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 126 -c 2 -o gpurun_out/prof_c4_fft_r2c -f $CMD > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/ | tail -8

set -x
ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 12 -c 2 -o gpurun_out/r02_fft_c2r_v2 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 22 -c 2 -o gpurun_out/r02_fft_r2c_v2 python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
python tools/precond_sweep.py c2 150 > gpurun_out/precond_c2.txt 2>&1; cat gpurun_out/precond_c2.txt
python tools/precond_sweep.py c2 150 5 > gpurun_out/precond_c2_mu5.txt 2>&1; cat gpurun_out/precond_c2_mu5.txt

set -x
python -m pytest tests/test_gpu_cg.py -q -m gpu -s -k "precond or state or main_fusion" > gpurun_out/pytest_pre.log 2>&1; tail -12 gpurun_out/pytest_pre.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_probe.csv python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_l.log 2>&1
grep -n "fft_pass" gpurun_out/r02_launches_probe.csv | head -12
ncu --set full --clock-control none --import-source on -k regex:fft_pass -s 10 -c 4 -o gpurun_out/r02_fft_stageA python bench.py --no-cpu-baseline --no-e2e --solve-iters 0 --steps 1 --warmup 1 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log

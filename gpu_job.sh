set -x
python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -5
python tools/fft_bench.py 501 512 float64 5
python tools/fft_bench.py 251 512 float64 5
python tools/fft_bench.py 501 512 float32 5
python tools/fft_bench.py 501 128 float64 1 > gpurun_out/plain_fft.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_ -s 8 -c 4 -o gpurun_out/prof_fft3 -f python tools/fft_bench.py 501 128 float64 1 > gpurun_out/ncu_fft.log 2>&1
tail -3 gpurun_out/ncu_fft.log

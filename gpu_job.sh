set -x
python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -1
python tools/fft_bench.py 501 512 float64 5
python tools/fft_bench.py 251 512 float64 5
python tools/fft_bench.py 501 512 float32 5

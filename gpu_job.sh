set -x
python -m pytest tests/test_gpu_full_size.py tests/test_gpu_fft.py -x -q --durations=8 2>&1 | tail -25

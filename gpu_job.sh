set -x
python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -25

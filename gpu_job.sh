set -x
python -m pytest tests/test_gpu_fft.py -x -q 2>&1 | tail -3
python tools/fft_bench.py 501 512 float64 5
python tools/fft_bench.py 251 512 float64 5
SURFH_B200_LIB=$PWD/surfh_b200/libsurfh_b200_nochain.so python tools/fft_bench.py 501 512 float64 5
SURFH_B200_LIB=$PWD/surfh_b200/libsurfh_b200_nochain.so python tools/fft_bench.py 251 512 float64 5

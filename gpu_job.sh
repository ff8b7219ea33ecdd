set -x
python -m pytest tests/test_gpu_cg.py -x -q 2>&1 | tail -25

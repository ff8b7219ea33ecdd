set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in c2 c4; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/d_$c.json 2> gpurun_out/d_$c.err; tail -2 gpurun_out/d_$c.err; done

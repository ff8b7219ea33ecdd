"""Times the hand-written 2-D FFT pair alone (CUDA events, inputs larger than L2) and, optionally,
cuFFT through torch for comparison.  usage: python tools/fft_bench.py [N] [batch] [dtype] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surfh_b200 import fft  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 501
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dtype = getattr(torch, sys.argv[3]) if len(sys.argv) > 3 else torch.float64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
x = torch.randn((batch, n, n), dtype=dtype, device="cuda")
s = fft.rfft2(x)
y = fft.irfft2(s, (n, n))
err = float((y / (n * n) - x).abs().max())
torch.cuda.synchronize()


def time_it(fn):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


es = x.element_size()
bytes_alg = batch * (n * n * es + n * (n // 2 + 1) * 2 * es)
t_f = time_it(lambda: fft.rfft2(x))
t_i = time_it(lambda: fft.irfft2(s, (n, n)))
print(f"own   N={n} batch={batch} {dtype}: r2c {t_f:.3f} ms ({bytes_alg / t_f / 1e6:.0f} GB/s alg), "
      f"c2r {t_i:.3f} ms ({bytes_alg / t_i / 1e6:.0f} GB/s alg), round-trip max err {err:.2e}")
if "--cufft" in sys.argv:
    t_f = time_it(lambda: torch.fft.rfft2(x))
    t_i = time_it(lambda: torch.fft.irfft2(s, s=(n, n)))
    print(f"cufft N={n} batch={batch}: r2c {t_f:.3f} ms, c2r {t_i:.3f} ms")

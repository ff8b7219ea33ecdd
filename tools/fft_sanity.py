"""Smallest FFT exercise for compute-sanitizer: a few shapes through both directions, checked vs numpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surfh_b200 import fft  # noqa: E402

for shape, batch in (((16, 9), 2), ((251, 251), 2), ((129, 300), 1)):
    for dt in (torch.float64, torch.float32):
        x = np.random.default_rng(1).standard_normal((batch,) + shape)
        got = fft.rfft2(torch.as_tensor(x, device="cuda", dtype=dt))
        back = fft.irfft2(got, shape).cpu().numpy() / (shape[0] * shape[1])
        e1 = np.linalg.norm(got.cpu().numpy() - np.fft.rfft2(x)) / np.linalg.norm(np.fft.rfft2(x))
        e2 = np.linalg.norm(back - x) / np.linalg.norm(x)
        print(shape, dt, f"{e1:.1e} {e2:.1e}")
        assert e1 < (1e-12 if dt == torch.float64 else 1e-5) and e2 < (1e-12 if dt == torch.float64 else 1e-5)
print("FFT SANITY OK")

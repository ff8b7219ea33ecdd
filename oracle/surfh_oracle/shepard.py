"""TEST INFRASTRUCTURE ONLY -- numpy (float32) restatement of the reference's exponential modified-Shepard
interpolation, surfh/ToolsDir/shepard_interpolation.pyx:20-141 (`pixel_distance`, `exponential_weight`,
`exponential_modified_shepard`): every intermediate is float32 like the Cython `float` locals, and the sums run
over the samples in index order (np.cumsum-free: float32 accumulation by np.add.reduce on a row is pairwise, so
the accumulation is done sample-block by sample-block in a float32 loop over blocks of ONE sample column).
Pinned by tests/golden/shepard.npz, produced by the reference's own compiled .pyx (oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np


def exponential_modified_shepard(alpha_coord, lambda_coord, values, alpha_mesh, lambda_mesh, p=2.0, alpha=2.0,
                                 pixel_cutoff=1.0, alpha_res=1.0, lambda_res=1.0, epsilon=1e-6):
    f = np.float32
    a = np.asarray(alpha_coord, dtype=f)
    lam = np.asarray(lambda_coord, dtype=f)
    v = np.asarray(values, dtype=f)
    qa = np.asarray(alpha_mesh, dtype=f).ravel()
    ql = np.asarray(lambda_mesh, dtype=f).ravel()
    inv_a, inv_l = f(1) / f(alpha_res), f(1) / f(lambda_res)
    num = np.zeros(qa.shape, dtype=f)
    den = np.zeros(qa.shape, dtype=f)
    for k in range(len(v)):          # sample order = the reference's summation order
        d1 = (a[k] - qa) * inv_a
        d2 = (lam[k] - ql) * inv_l
        dist = np.sqrt(d1 * d1 + d2 * d2, dtype=f) + f(epsilon)
        near = dist <= f(pixel_cutoff)
        if not near.any():
            continue
        w = np.exp(f(-alpha) * np.power(dist[near], f(p), dtype=f), dtype=f)
        num[near] = num[near] + w * v[k]
        den[near] = den[near] + w
    out = np.where(den != 0, num / np.where(den != 0, den, f(1)), f(0)).astype(f)
    return out.reshape(np.shape(alpha_mesh))

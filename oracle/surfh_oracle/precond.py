"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the Fourier-domain block preconditioner.

Follows the per-frequency Hessian of the reference's Fourier-domain mixing model,
    hess_spec_freq[k1, k2](f) = sum_l T[k1, l] T[k2, l] |psfs_freq[l](f)|^2     (di = dj = 1)
(surfh/Models/mixing.py:131-207, `Model_WCT.__init__`), plus the regulariser, inverted bin by bin like
`Inv_Regul_Fusion_Model3` / algorithms.make_iHtH_spectro (surfh/ToolsDir/fusion_mixing.py:401-438).  The
wavelength weights w_l (mean gain of the MRS detector sampling) are this project's addition: the reference's
Fourier model has no slit / detector stage.  Parity unpinned: no reference test stores an output of those classes."""
from __future__ import annotations

import numpy as np


def laplacian_eigenvalues(shape):
    """Eigenvalues of the circular D_r^T D_r + D_c^T D_c (fusion_CT.py:16-43) on the rfft2 grid."""
    na, nb = shape
    di = 2.0 - 2.0 * np.cos(2.0 * np.pi * np.arange(na) / na)
    dj = 2.0 - 2.0 * np.cos(2.0 * np.pi * np.arange(nb // 2 + 1) / nb)
    return di[:, None] + dj[None, :]


def apply(sotf, templates, weights, mu_s, mu_r, r, shape, joint=False):
    """z = P r for r of shape [K, Na, Nb]."""
    k = templates.shape[0]
    power = (np.abs(np.asarray(sotf)) ** 2) * np.asarray(weights)[:, None, None]
    gram = np.einsum("lij,kl,ml->ijkm", power, templates, templates, optimize=True)
    d = laplacian_eigenvalues(shape)
    if joint:
        d = d ** 2
    blocks = mu_s * gram + mu_r * d[:, :, None, None] * np.eye(k)[None, None]
    rhat = np.fft.rfft2(np.asarray(r, dtype=np.float64).reshape((k,) + tuple(shape)))
    zhat = np.linalg.solve(blocks, np.moveaxis(rhat, 0, -1)[..., None])[..., 0]
    return np.fft.irfft2(np.moveaxis(zhat, -1, 0), s=shape)

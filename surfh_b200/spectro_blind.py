"""`MRSBlurred`: the reference's single-wavelength (non-LMM) MRS operator
(surfh/Models/spectro_blind.py:27-323) -- spatial blur C, bilinear gridding S, box-sum, slit weights L
and a plain beta-sum onto the detector rows; no templates, no spectral response -- plus the batched
form BASELINE.json's configuration 5 asks for (the same operator applied to every wavelength of a
cube, each with its own OTF).

    MRSBlurred(sotf[N, N//2+1], ...)          ishape (N, N)        oshape (P*S*na,)       as the reference
    MRSBlurred(sotf[L, N, N//2+1], ...)       ishape (L, N, N)     oshape (L*P*S*na,)     [L][P,S,na] C-order

Same CUDA library and kernels as `spectroSigRLSCT` (no-LMM pipeline + SURFH_SPECTRAL_BETA_SUM band);
the geometry follows MRSBlurred's own copies of the slicing rules (no even-na adjustment, instrument
and pointings used as given)."""
from __future__ import annotations

import numpy as np

from . import instru
from .model import _is_torch, spectroSigRLSCT


class MRSBlurred(spectroSigRLSCT):
    _rules = "blind"

    def __init__(self, sotf, alpha_axis, beta_axis, instr: instru.IFU, step_degree: float,
                 pointings: instru.CoordList, **kwargs):
        self.single = (not callable(sotf)) and sotf.ndim == 2
        if callable(sotf) and "n_lambda" not in kwargs:
            raise ValueError("a callable sotf needs n_lambda=")
        n_lambda = kwargs.pop("n_lambda", None)
        if n_lambda is None:
            n_lambda = 1 if self.single else int(sotf.shape[0])
        if self.single:
            sotf = sotf[None, ...]
        super().__init__(sotf, None, alpha_axis, beta_axis, np.arange(n_lambda, dtype=np.float64), [instr],
                         step_degree, [pointings], **kwargs)
        t = self.band_tables[0]
        self.instr = instr
        self.srf = t.srf
        self.local_alpha_axis, self.local_beta_axis = t.local_alpha_axis, t.local_beta_axis
        self.local_im_shape = t.local_shape
        self.slices_shape = (t.n_pointing, t.n_slit, t.na)
        self.npix_slit_alpha_width = t.npix_slit_alpha_width
        self.npix_slit_beta_width = t.nb
        if self.single:
            self.ishape = self.imshape
        self.oshape = (int(n_lambda * np.prod(self.slices_shape)),)

    def _shape_in(self, x):
        return x.reshape(self.cube_shape)

    def forward(self, x):
        return super().forward(self._shape_in(x))

    def adjoint(self, data):
        out = super().adjoint(data)
        return out.reshape(self.ishape)

    def fwadj(self, x, out=None):
        res = super().fwadj(self._shape_in(x), out=None if out is None else out.reshape(self.cube_shape))
        return out if out is not None else res.reshape(self.ishape)

    fwback = fwadj

    def get_slit_slices(self, slit_idx: int):
        return self.band_tables[0].slices[slit_idx]

    def get_slit_weights(self, slit_idx: int, slices=None):
        return self.channels[0].slicer.get_slit_weights(slit_idx)

    def real_data_janskySR_to_jansky(self, data: np.ndarray) -> np.ndarray:
        """spectro_blind.py: per-slit factor (sum of beta weights) * srf on [.., P, S, na] data."""
        t = self.band_tables[0]
        scale = t.weights.sum(axis=1) * t.srf
        block = np.asarray(data, dtype=np.float64).reshape((-1,) + self.slices_shape)
        return (block * scale[None, None, :, None]).reshape(np.shape(data))

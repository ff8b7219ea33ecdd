"""TEST INFRASTRUCTURE ONLY -- CPU (numpy fp64) restatement of the reference's single-wavelength
operator `MRSBlurred` (surfh/Models/spectro_blind.py:27-323): C, S, Sum, L and a plain beta-sum; no
templates, no spectral response.  Pinned by tests/golden/blind_*.npz (oracle/make_golden.py).
Never imported by the product."""
from __future__ import annotations

from math import ceil, floor

import numpy as np

from . import instrument as ins
from .model import dft, idft, interpn_cube2local, interpn_local2cube, scatter_local2cube
from .thirdparty import LinOp, ir2fr


class MRSBlurred(LinOp):
    def __init__(self, sotf, alpha_axis, beta_axis, instr, step_degree, pointings, adjoint_mode="reference"):
        """spectro_blind.py:28-74.  Unlike Channel, the instrument and the pointings are used as given
        (no .pix()) and the slit rules are the class's own copies (no even-na adjustment)."""
        self.sotf = np.asarray(sotf)
        self.alpha_axis = np.asarray(alpha_axis, dtype=np.float64)
        self.beta_axis = np.asarray(beta_axis, dtype=np.float64)
        self.band = ins.Band.from_ifu(instr)
        self.pointings = [(p.alpha, p.beta) if hasattr(p, "alpha") else (p[0], p[1]) for p in pointings]
        self.adjoint_mode = adjoint_mode
        self.srf = ins.get_srf([self.band.det_pix_size], step_degree * 3600)[0]
        self.la, self.lb = ins.local_axes(self.band.alpha_width, self.band.beta_width, step_degree, 5 * step_degree)
        self.local_im_shape = (len(self.la), len(self.lb))
        self._otf_sr = ir2fr(np.ones((self.srf, 1)), self.local_im_shape)
        step = self.la[1] - self.la[0]
        w = self.band.alpha_width
        self.npix_slit_alpha_width = int(ceil(w / 2 / step)) - int(floor(-w / 2 / step))  # :89-98
        self.npix_slit_beta_width = int(ceil(self.band.slit_beta_width / (self.beta_axis[1] - self.beta_axis[0])))
        self.slices_shape = (len(self.pointings), self.band.n_slit, ceil(self.npix_slit_alpha_width / self.srf))
        super().__init__((len(self.alpha_axis), len(self.beta_axis)), (int(np.prod(self.slices_shape)),))
        decal = np.zeros(self.local_im_shape)
        decal[-int((self.srf - 1) / 2), 0] = np.sqrt(self.local_im_shape[0] * self.local_im_shape[1])
        self.decalf = dft(decal)
        # geometry helper without the Channel-only alpha adjustment: reuse SlitGeometry's pieces
        self._geo = ins.SlitGeometry(self.band, self.beta_axis, self.la, self.lb, self.srf)

    def get_slit_slices(self, s):
        """:120-145 -- to_slices + the one-pixel beta trim; the alpha adjustment is commented out there."""
        bounds = self._geo.slit_bounds(s)
        sa, sb = self._geo._to_slices(bounds)
        if (sb.stop - sb.start) > self.npix_slit_beta_width:
            if abs(self.lb[sb.stop] - bounds[3]) > abs(self.lb[sb.start] - bounds[2]):
                sb = slice(sb.start, sb.stop - 1)
            else:
                sb = slice(sb.start + 1, sb.stop)
        return sa, sb

    def get_slit_weights(self, s, slices):
        """:148-167 -- note the second guard compares with npix_slit_beta_width, not the slit count."""
        _, _, b_start, b_end = self._geo.slit_bounds(s)
        sa, sb = slices
        db = self.lb[1] - self.lb[0]
        sel = self.lb[sb]
        w = np.ones((sa.stop - sa.start, sb.stop - sb.start))
        if sel[0] - db / 2 < b_start:
            w[:, 0] = 1 - abs(sel[0] - db / 2 - b_start) / db
        if sel[-1] + db / 2 > b_end:
            w[:, -1] = 1 - abs(sel[-1] + db / 2 - b_end) / db
        if s > 0 and self.get_slit_slices(s - 1)[1].stop - 1 != sb.start:
            w[:, 0] = 1
        if s < self.npix_slit_beta_width - 1 and sb.stop - 1 != self.get_slit_slices(s + 1)[1].start:
            w[:, -1] = 1
        return w[np.newaxis, ...]

    def _origin(self, pointing):
        return (self.band.origin[0] + pointing[0], self.band.origin[1] + pointing[1])

    def forward(self, x):
        """:191-207"""
        out = np.zeros(self.slices_shape)
        x = np.asarray(x, dtype=np.float64).reshape(self.ishape)
        blurred = idft(dft(x) * self.sotf, self.ishape)
        na, srf = self.slices_shape[2], self.srf
        for p_idx, pointing in enumerate(self.pointings):
            ga, gb = ins.local2global(self.la, self.lb, self._origin(pointing), self.band.angle)
            gridded = interpn_cube2local(self.alpha_axis, self.beta_axis, blurred[np.newaxis], ga, gb)[0]
            summed = idft(dft(gridded) * (self._otf_sr * self.decalf), self.local_im_shape)
            for s in range(self.band.n_slit):
                sl = self.get_slit_slices(s)
                sliced = summed[sl[0], sl[1]] * self.get_slit_weights(s, sl)[0]
                out[p_idx, s] = np.sum(sliced[: na * srf: srf], axis=1)
        return out.ravel()

    def adjoint(self, data):
        """:210-235"""
        data = np.reshape(data, self.slices_shape)
        na, srf = self.slices_shape[2], self.srf
        img = np.zeros(self.ishape)
        for p_idx, pointing in enumerate(self.pointings):
            local = np.zeros(self.local_im_shape)
            for s in range(self.band.n_slit):
                sl = self.get_slit_slices(s)
                over = np.repeat(data[p_idx, s][:, np.newaxis], self.npix_slit_beta_width, axis=1)
                sl0 = self.get_slit_slices(0)
                placed = np.zeros((sl0[0].stop - sl0[0].start, sl0[1].stop - sl0[1].start))
                placed[: na * srf: srf, :] = over
                tmp = np.zeros(self.local_im_shape)
                tmp[sl[0], sl[1]] = placed * self.get_slit_weights(s, sl)[0]
                local += tmp
            sum_t = idft(dft(local) * self._otf_sr.conj() * self.decalf.conj(), self.local_im_shape)
            if self.adjoint_mode == "exact":
                ga, gb = ins.local2global(self.la, self.lb, self._origin(pointing), self.band.angle)
                img += scatter_local2cube(self.alpha_axis, self.beta_axis, sum_t[np.newaxis], ga, gb)[0]
            else:
                la, lb = ins.global2local(self.alpha_axis, self.beta_axis, self._origin(pointing), self.band.angle)
                img += interpn_local2cube(self.la, self.lb, np.array(sum_t, dtype=np.float64)[np.newaxis], la, lb)[0]
        return idft(dft(img) * self.sotf.conj(), self.ishape)

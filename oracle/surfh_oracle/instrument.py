"""TEST INFRASTRUCTURE ONLY -- CPU (numpy fp64) restatement of the reference's host-side
instrument geometry: FoV axes, rotations, slit index ranges and edge weights, wavelength
windows and the spectral line-spread table.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Pinned against the reference's own code through tests/golden/*.npz (made by
oracle/make_golden.py, which runs the unmodified reference arithmetic in a scratch copy).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import ceil, floor
from typing import List, Sequence, Tuple

import numpy as np


def rotmatrix(degree: float) -> np.ndarray:
    """2x2 rotation by `degree` degrees (surfh/Models/instru.py:36-45)."""
    t = np.radians(degree)
    return np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])


def get_srf(det_pix_sizes: Sequence[float], step_arcsec: float) -> List[int]:
    """Super-resolution factor per band = det_pix_size // step (instru.py:67-84)."""
    return [int(d // step_arcsec) for d in det_pix_sizes]


def pix(value: float, step: float) -> float:
    """Coord.pix for one coordinate: python round() to the grid (instru.py:143-145)."""
    return round(value / step) * step


@dataclass
class Band:
    """Plain description of one IFU band -- the attributes of instru.IFU (instru.py:576-610)
    that the hot path reads, flattened.  `origin` is (alpha, beta) in degrees."""

    alpha_width: float
    beta_width: float
    origin: Tuple[float, float]
    angle: float
    det_pix_size: float
    n_slit: int
    grating_resolution: float
    wavel_axis: np.ndarray
    name: str = "_"

    @classmethod
    def from_ifu(cls, ifu) -> "Band":
        """Accept any duck-typed IFU (.fov.alpha_width/.beta_width/.origin/.angle,
        .det_pix_size, .n_slit, .w_blur.grating_resolution, .wavel_axis, .name)."""
        if isinstance(ifu, Band):
            return ifu
        return cls(
            float(ifu.fov.alpha_width), float(ifu.fov.beta_width),
            (float(ifu.fov.origin.alpha), float(ifu.fov.origin.beta)), float(ifu.fov.angle),
            float(ifu.det_pix_size), int(ifu.n_slit), float(ifu.w_blur.grating_resolution),
            np.asarray(ifu.wavel_axis, dtype=np.float64), getattr(ifu, "name", "_"),
        )

    def pixelised(self, step: float) -> "Band":
        """IFU.pix: same band with the FoV origin rounded to the grid (instru.py:681-697)."""
        o = (pix(self.origin[0], step), pix(self.origin[1], step))
        return Band(self.alpha_width, self.beta_width, o, self.angle, self.det_pix_size,
                    self.n_slit, self.grating_resolution, self.wavel_axis, self.name)

    @property
    def slit_beta_width(self) -> float:
        """instru.py:660-663."""
        return self.beta_width / self.n_slit

    @property
    def wavel_step(self) -> float:
        """instru.py:638-641."""
        return self.wavel_axis[1] - self.wavel_axis[0]


def wslice(band: Band, wavel_input_axis: np.ndarray, margin: float) -> slice:
    """IFU.wslice (instru.py:649-658): cube wavelengths seen by the band, `margin` um wider."""
    lo = max(band.wavel_axis[0] - margin, wavel_input_axis.min())
    hi = min(band.wavel_axis[-1] + margin, wavel_input_axis.max())
    return slice(int(np.flatnonzero(wavel_input_axis <= lo)[-1]),
                 int(np.flatnonzero(wavel_input_axis >= hi)[0]))


def local_axes(alpha_width: float, beta_width: float, step: float, margin: float):
    """FOV.local_coords (instru.py:283-304): regular axes in the FoV's own frame with a
    margin on each side; the start is floored onto the step grid."""

    def axis(start, length):
        round_start = int(floor(start / step)) * step
        num = int(ceil((length + (start - round_start)) / step))
        return np.arange(num + 1) * step + round_start

    return (axis(-alpha_width / 2 - margin, alpha_width + 2 * margin),
            axis(-beta_width / 2 - margin, beta_width + 2 * margin))


def local2global(la, lb, origin, angle):
    """FOV.local2global (instru.py:306-321): rotate the local grid by +angle, add origin."""
    na, nb = len(la), len(lb)
    a = np.tile(la.reshape((-1, 1)), [1, nb])
    b = np.tile(lb.reshape((1, -1)), [na, 1])
    c = rotmatrix(angle) @ np.vstack((a.ravel(), b.ravel()))
    return c[0].reshape((na, nb)) + origin[0], c[1].reshape((na, nb)) + origin[1]


def global2local(ga, gb, origin, angle):
    """FOV.global2local (instru.py:323-340): subtract origin, rotate by -angle."""
    na, nb = len(ga), len(gb)
    a = np.tile((ga - origin[0]).reshape((-1, 1)), [1, nb])
    b = np.tile((gb - origin[1]).reshape((1, -1)), [na, 1])
    c = rotmatrix(-angle) @ np.vstack((a.ravel(), b.ravel()))
    return c[0].reshape((na, nb)), c[1].reshape((na, nb))


def lsf_table(grating_resolution, out_axis, beta, wavelength, scale, n_margin=15):
    """SpectralBlur.psfs, type 'mrs' (instru.py:484-572): W[lambda', lambda, beta], a sinc^2
    line-spread function normalised over lambda on an axis extended by n_margin-1 samples
    each side, margins dropped afterwards."""
    grating_len = 2 * 0.44245 / np.pi * grating_resolution
    wavelength = np.asarray(wavelength)
    dw = min(np.diff(wavelength))
    beta = np.asarray(beta).reshape((1, 1, -1))
    out_axis = np.asarray(out_axis).reshape((-1, 1, 1))
    ext = np.concatenate([
        np.linspace(wavelength.min() - n_margin * dw, wavelength.min() - dw, n_margin - 1),
        wavelength,
        np.linspace(wavelength.max() + dw, wavelength.max() + n_margin * dw, n_margin - 1),
    ]).reshape((1, -1, 1))
    out = (np.pi * grating_len / ext
           * np.sinc(np.pi * grating_len * ((out_axis - scale * beta) / ext - 1)) ** 2)
    out /= np.sum(out, axis=1, keepdims=True)
    return out[:, n_margin - 1: -n_margin + 1, :]


class SlitGeometry:
    """Slicer (surfh/Models/slicer.py:14-244) for one band: slit index ranges and weights in
    the local grid.  Like the reference, slices and weights are re-derived on every call."""

    def __init__(self, band: Band, beta_axis, local_alpha_axis, local_beta_axis, srf: int):
        self.band = band
        self.beta_axis = beta_axis
        self.la = local_alpha_axis
        self.lb = local_beta_axis
        self.srf = srf
        # slicer.py:31
        self.slices_shape = (band.n_slit, ceil(self.npix_slit_alpha_width / srf))

    @property
    def slit_beta_width(self):  # slicer.py:39-42
        return self.band.beta_width / self.band.n_slit

    @property
    def npix_slit_beta_width(self):  # slicer.py:44-47
        return int(ceil(self.slit_beta_width / (self.beta_axis[1] - self.beta_axis[0])))

    @property
    def npix_slit_alpha_width(self):  # slicer.py:53-62
        step = self.la[1] - self.la[0]
        w = self.band.alpha_width
        return int(ceil(w / 2 / step)) - int(floor(-w / 2 / step))

    def slit_bounds(self, s: int):
        """Local-frame extent of slit s: slicer.py:87-90 with instru.py:612-626 (slit_shift)
        and instru.py:416-434 (LocalFOV start/end; beta rounded to 9 decimals)."""
        b = self.band
        shift_beta = (-b.beta_width / 2 + b.slit_beta_width / 2) + s * b.slit_beta_width
        origin_beta = 0 + shift_beta
        return (0.0 - b.alpha_width / 2, 0.0 + b.alpha_width / 2,
                round(origin_beta - b.slit_beta_width / 2, 9),
                round(origin_beta + b.slit_beta_width / 2, 9))

    def _to_slices(self, bounds):
        """LocalFOV.to_slices (instru.py:436-459)."""
        a_start, a_end, b_start, b_end = bounds
        da = self.la[1] - self.la[0]
        db = self.lb[1] - self.lb[0]
        return (slice(int(np.flatnonzero(a_start < self.la + da / 2)[0]),
                      int(np.flatnonzero(self.la - da / 2 < a_end)[-1]) + 1),
                slice(int(np.flatnonzero(b_start < self.lb + db / 2)[0]),
                      int(np.flatnonzero(self.lb - db / 2 < b_end)[-1]) + 1))

    def slit_slices(self, s: int):
        """get_slit_slices (slicer.py:118-145): index ranges, one-pixel trim when the beta
        range is longer than the slit, and the even-na<28 alpha adjustment."""
        bounds = self.slit_bounds(s)
        sa, sb = self._to_slices(bounds)
        if (sb.stop - sb.start) > self.npix_slit_beta_width:
            if abs(self.lb[sb.stop] - bounds[3]) > abs(self.lb[sb.start] - bounds[2]):
                sb = slice(sb.start, sb.stop - 1)
            else:
                sb = slice(sb.start + 1, sb.stop)
        if self.slices_shape[1] % 2 == 0 and self.slices_shape[1] < 28:
            if (sa.stop - sa.start) > self.npix_slit_alpha_width:
                sa = slice(sa.start, sa.stop - 1)
            elif (sa.stop - sa.start) < self.npix_slit_alpha_width:
                sa = slice(sa.start - 2, sa.stop)
        return sa, sb

    def slit_weights(self, s: int, slices):
        """get_slit_weights + fov_weight (slicer.py:148-168, 187-244): ones, except the first
        / last beta column carries the covered fraction when it sticks out of the slit and the
        neighbouring slit shares that column."""
        _, _, b_start, b_end = self.slit_bounds(s)
        sa, sb = slices
        db = self.lb[1] - self.lb[0]
        sel = self.lb[sb]
        w = np.ones((sa.stop - sa.start, sb.stop - sb.start))
        if sel[0] - db / 2 < b_start:
            w0 = 1 - abs(sel[0] - db / 2 - b_start) / db
            assert 0 <= w0 <= 1, f"first beta weight out of [0, 1] ({w0:.2f})"
            w[:, 0] = w0
        if sel[-1] + db / 2 > b_end:
            w1 = 1 - abs(sel[-1] + db / 2 - b_end) / db
            assert 0 <= w1 <= 1, f"last beta weight out of [0, 1] ({w1:.2f})"
            w[:, -1] = w1
        if s > 0 and self.slit_slices(s - 1)[1].stop - 1 != sb.start:
            w[:, 0] = 1
        if s < self.slices_shape[0] - 1 and sb.stop - 1 != self.slit_slices(s + 1)[1].start:
            w[:, -1] = 1
        return w[np.newaxis, ...]

    def slicing(self, gridded, s: int):
        """Slicer.slicing (slicer.py:64-68)."""
        sl = self.slit_slices(s)
        return gridded[:, sl[0], sl[1]] * self.slit_weights(s, sl)

    def slicing_t(self, slit, s: int, local_shape):
        """Slicer.slicing_t (slicer.py:72-84)."""
        out = np.zeros(local_shape)
        sl = self.slit_slices(s)
        out[:, sl[0], sl[1]] = slit * self.slit_weights(s, sl)
        return out

    def slit_shape(self, n_wave: int):
        """get_slit_shape_t (slicer.py:179-185)."""
        sl = self.slit_slices(0)
        return (n_wave, sl[0].stop - sl[0].start, sl[1].stop - sl[1].start)

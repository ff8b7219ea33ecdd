"""Seeded synthetic inputs shaped like the reference's MIRI-MRS fusion problem.

No file from the reference's private data directories is needed: the instrument table is the
12-band table of scripts/main_fusion.py:107-120 (slits, resolving power, detector pixel, FoV),
the detector wavelength axes are regular grids with the first / last sample and length of
surfh/Others/global_variables.py, the PSF is the reference's Gaussian fall-back
(surfh/ToolsDir/utils.py:41-51: FWHM = lambda / 6.5 m), and the dither pattern is the one of
test/test_fw_ad.py:736-741.  Configurations C1..C5 are those of BASELINE.md section 3.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import instru

STEP_ARCSEC = 0.025
FOV_ANGLE = 8.2

# name: (n_slit, R_min, R_max, det_pix_size ["], fov_alpha ["], fov_beta ["])
MRS_BANDS: Dict[str, tuple] = {
    "1a": (21, 3320, 3710, 0.196, 3.2, 3.7), "1b": (21, 3190, 3750, 0.196, 3.2, 3.7),
    "1c": (21, 3100, 3610, 0.196, 3.2, 3.7), "2a": (17, 2990, 3110, 0.196, 4.0, 4.8),
    "2b": (17, 2750, 3170, 0.196, 4.0, 4.8), "2c": (17, 2860, 3300, 0.196, 4.0, 4.8),
    "3a": (16, 2530, 2880, 0.245, 5.2, 6.2), "3b": (16, 1790, 2640, 0.245, 5.2, 6.2),
    "3c": (16, 1980, 2790, 0.245, 5.2, 6.2), "4a": (12, 1460, 1930, 0.273, 6.6, 7.7),
    "4b": (12, 1680, 1760, 0.273, 6.6, 7.7), "4c": (12, 1630, 1330, 0.273, 6.6, 7.7),
}

# name: (first sample [um], last sample [um], number of samples)
MRS_DETECTOR_AXES: Dict[str, tuple] = {
    "1a": (4.9004001, 5.73960007, 1050), "1b": (5.66039985, 6.62999982, 1213),
    "1c": (6.53040021, 7.64960018, 1400), "2a": (7.51065023, 8.77035023, 970),
    "2b": (8.67065008, 10.13055008, 1124), "2c": (10.01065023, 11.69935023, 1300),
    "3a": (11.55125019, 13.47125015, 769), "3b": (13.34125015, 15.5687501, 892),
    "3c": (15.41124985, 17.97874979, 1028), "4a": (17.70300076, 20.94900079, 542),
    "4b": (20.69300053, 24.47900057, 632), "4c": (24.40299962, 28.69899966, 717),
}
ALL_BANDS = list(MRS_BANDS)


def mrs_detector_axis(name: str) -> np.ndarray:
    lo, hi, n = MRS_DETECTOR_AXES[name]
    return np.linspace(lo, hi, n)


def cube_wavelength_axis(lo: float, hi: float, ratio: float = 1.0005) -> np.ndarray:
    """Log-spaced cube axis lambda_{i+1} / lambda_i = ratio (R ~ 2000 for 1.0005)."""
    n = int(np.floor(np.log(hi / lo) / np.log(ratio)))
    return lo * ratio ** np.arange(n)


def make_band(name: str, angle: float = FOV_ANGLE) -> instru.IFU:
    n_slit, r_min, r_max, det_pix, fov_a, fov_b = MRS_BANDS[name]
    return instru.IFU(
        fov=instru.FOV(fov_a / 3600, fov_b / 3600, origin=instru.Coord(0, 0), angle=angle),
        det_pix_size=det_pix, n_slit=n_slit,
        w_blur=instru.SpectralBlur(float(np.mean([r_min, r_max]))),
        pce=None, wavel_axis=mrs_detector_axis(name), name=name.upper())


def dither_pointings(ifu: instru.IFU, step_degree: float, n_pointings: int) -> instru.CoordList:
    """1 pointing at the origin, or the first n of the 4-point dither, rounded to the grid."""
    if n_pointings == 1:
        return instru.CoordList([instru.Coord(0, 0)]).pix(step_degree)
    da = (ifu.det_pix_size / 3600) / 4
    db = ifu.slit_beta_width / 4
    pts = [instru.Coord(da, db), instru.Coord(-da, db), instru.Coord(da, -db), instru.Coord(-da, -db)]
    return instru.CoordList(pts[:n_pointings]).pix(step_degree)


def gaussian_psf(wavel_axis: np.ndarray, step_arcsec: float, size: int = 41, D: float = 6.5):
    """Unit-sum Gaussian stamp per wavelength, FWHM = lambda / D."""
    half = size // 2
    ax = np.arange(-half, half + 1, dtype=np.float64)
    r2 = ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    fwhm = (np.asarray(wavel_axis) * 1e-6 / D) * 206265.0
    sigma = (fwhm / (step_arcsec * 2.354))[:, None, None]
    psf = np.exp(-r2 / (2 * sigma ** 2))
    return psf / psf.sum(axis=(1, 2), keepdims=True)


def ir2fr(imp_resp: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """Impulse response -> rFFT frequency response on `shape` (origin-centred, zero-padded,
    non-normalised), i.e. what the reference obtains from `udft.ir2fr(spsf, imshape)`
    (scripts/main_fusion.py:98)."""
    imp_resp = np.asarray(imp_resp)
    nd = len(shape)
    buf = np.zeros(imp_resp.shape[:-nd] + tuple(shape), dtype=imp_resp.dtype)
    buf[(Ellipsis,) + tuple(slice(0, n) for n in imp_resp.shape[-nd:])] = imp_resp
    shifts = [-(n // 2) for n in imp_resp.shape[-nd:]]
    buf = np.roll(buf, shifts, axis=tuple(range(-nd, 0)))
    return np.fft.rfftn(buf, axes=tuple(range(-nd, 0)))


def ir2fr_device(psf: np.ndarray, shape: Sequence[int], device, dtype, chunk: int = 256):
    """Same as `ir2fr` for a [L, h, w] stamp stack, computed chunk-wise on `device` so that a
    multi-GB OTF never exists on the host.  Returns a complex torch tensor [L, N, N//2+1]."""
    import torch

    cdtype = torch.complex128 if dtype in (torch.float64, torch.complex128) else torch.complex64
    n_l, h, w = psf.shape
    out = torch.empty((n_l, shape[0], shape[1] // 2 + 1), dtype=cdtype, device=device)
    for lo in range(0, n_l, chunk):
        hi = min(n_l, lo + chunk)
        buf = torch.zeros((hi - lo, shape[0], shape[1]), dtype=torch.float64, device=device)
        buf[:, :h, :w] = torch.as_tensor(psf[lo:hi], device=device)
        buf = torch.roll(buf, shifts=(-(h // 2), -(w // 2)), dims=(1, 2))
        out[lo:hi] = torch.fft.rfft2(buf).to(cdtype)
    return out


@dataclass
class Config:
    """Everything `spectroSigRLSCT(...)` takes, plus seeded maps."""

    name: str
    templates: Optional[np.ndarray]
    alpha_axis: np.ndarray
    beta_axis: np.ndarray
    wavelength_axis: np.ndarray
    instrs: List[instru.IFU]
    step_degree: float
    pointings: List[instru.CoordList]
    psf: np.ndarray
    maps: np.ndarray
    band_names: List[str] = field(default_factory=list)

    @property
    def imshape(self):
        return (len(self.alpha_axis), len(self.beta_axis))

    def sotf(self) -> np.ndarray:
        return ir2fr(self.psf, self.imshape)

    def model_args(self, sotf=None) -> dict:
        return dict(sotf=self.sotf() if sotf is None else sotf, templates=self.templates,
                    alpha_axis=self.alpha_axis, beta_axis=self.beta_axis,
                    wavelength_axis=self.wavelength_axis, instrs=self.instrs,
                    step_degree=self.step_degree, pointings=self.pointings)


def _axes(n_pix: int, step_degree: float):
    ax = np.arange(n_pix) * step_degree
    return ax - np.mean(ax)


def _assemble(name, instrs, band_names, n_pix, n_templates, n_pointings, wavel, seed, psf_size=41,
              lmm=True) -> Config:
    step_degree = STEP_ARCSEC / 3600
    rng = np.random.default_rng(seed)
    templates = (0.5 + rng.random((n_templates, len(wavel)))) if lmm else None
    shape = (n_templates, n_pix, n_pix) if lmm else (len(wavel), n_pix, n_pix)
    maps = rng.random(shape)
    pointings = [dither_pointings(i, step_degree, n_pointings) for i in instrs]
    return Config(name, templates, _axes(n_pix, step_degree), _axes(n_pix, step_degree), wavel,
                  instrs, step_degree, pointings, gaussian_psf(wavel, STEP_ARCSEC, psf_size), maps,
                  band_names)


def mrs_config(bands: Sequence[str], n_pix: int, n_templates: int, n_pointings: int,
               seed: int = 0, name: str = "mrs", wavel: Optional[np.ndarray] = None,
               lmm: bool = True) -> Config:
    """A configuration over real MRS bands.  The cube axis spans the bands +-0.15 um."""
    instrs = [make_band(b) for b in bands]
    if wavel is None:
        lo = min(MRS_DETECTOR_AXES[b][0] for b in bands) - 0.15
        hi = max(MRS_DETECTOR_AXES[b][1] for b in bands) + 0.15
        wavel = cube_wavelength_axis(lo, hi)
    return _assemble(name, instrs, list(bands), n_pix, n_templates, n_pointings, wavel, seed, lmm=lmm)


def mini_config(n_bands: int = 1, n_pointings: int = 1, n_templates: int = 3, n_pix: int = 96,
                seed: int = 7, lmm: bool = True, n_slit_a: int = 5) -> Config:
    """Small instruments (a few slits, tens of wavelengths) exercising every geometry rule:
    band A has an even number of detector pixels per slit (the even-na alpha adjustment),
    band B an odd one and a different super-resolution factor."""
    det_a = np.linspace(5.00, 5.10, 48)
    det_b = np.linspace(5.12, 5.30, 40)
    band_a = instru.IFU(instru.FOV(1.0 / 3600, 1.1 / 3600, instru.Coord(0, 0), FOV_ANGLE), 0.196, n_slit_a,
                        instru.SpectralBlur(300.0), None, det_a, "MINIA")
    band_b = instru.IFU(instru.FOV(1.4 / 3600, 1.2 / 3600, instru.Coord(0, 0), -11.0), 0.245, 4,
                        instru.SpectralBlur(260.0), None, det_b, "MINIB")
    instrs = [band_a, band_b][:n_bands]
    hi = 5.25 if n_bands == 1 else 5.45
    wavel = cube_wavelength_axis(4.85, hi, 1.002)
    return _assemble(f"mini{n_bands}b{n_pointings}p", instrs, ["minia", "minib"][:n_bands], n_pix,
                     n_templates, n_pointings, wavel, seed, psf_size=15, lmm=lmm)


def baseline_config(which: str, seed: int = 0) -> Config:
    """C1..C5 of BASELINE.md section 3."""
    which = which.lower()
    if which == "c1":
        return mrs_config(["1a"], 251, 4, 1, seed, "c1")
    if which == "c2":
        return mrs_config(["1a"], 251, 4, 4, seed, "c2")
    if which == "c3":
        return mrs_config(["1a", "2a", "3a", "4a"], 501, 4, 4, seed, "c3",
                          wavel=cube_wavelength_axis(4.75, 28.9))
    if which == "c4":
        return mrs_config(ALL_BANDS, 501, 6, 4, seed, "c4", wavel=cube_wavelength_axis(4.75, 28.9))
    if which == "c5":
        # the per-wavelength (non-LMM) MRSBlurred operator batched over a full-size cube: ~3000 wavelengths
        return mrs_config(["1c"], 301, 0, 4, seed, "c5", lmm=False, wavel=cube_wavelength_axis(4.75, 21.3))
    raise ValueError(f"unknown baseline configuration {which!r}")

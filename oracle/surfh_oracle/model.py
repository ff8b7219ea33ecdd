"""TEST INFRASTRUCTURE ONLY -- CPU (numpy/scipy fp64) restatement of the reference's LMM
instrument operator `spectroSigRLSCT` (forward / adjoint) and of its numeric kernels.

It follows the reference's own numpy/scipy path (surfh/ToolsDir/python_utils.py) stage by
stage, including the FFT-based "Sum" stage and the reference's `gridding_t` "adjoint", which is
an interpolation and NOT the transpose of the forward gridding.
`adjoint_mode="exact"` replaces only that stage by the true transpose (scatter-add).

Pinned: tests/test_oracle_golden.py checks this file against golden vectors produced by the
reference's own code (oracle/make_golden.py).  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this package; the product never does.
All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

from math import ceil
from typing import List, Sequence

import numpy as np
import scipy.fft

from . import instrument as ins
from .thirdparty import LinOp, ir2fr

FFT_WORKERS = -1  # python_utils.py:55-57,71 use workers=-1


# ------------------------------------------------------------------ numeric kernels
def lmm_maps2cube(maps, tpls):
    """cube[l,i,j] = sum_k maps[k,i,j] * tpls[k,l]  (python_utils.py:11-24).
    einsum instead of the broadcast-sum: same definition, no [K,L,N,N] temporary."""
    return np.einsum("kij,kl->lij", maps, tpls, optimize=True)


def lmm_cube2maps(cube, tpls):
    """maps[k,i,j] = sum_l cube[l,i,j] * tpls[k,l]  (python_utils.py:27-35)."""
    return np.einsum("lij,kl->kij", cube, tpls, optimize=True)


def dft(a):
    """Ortho rFFT over the last two axes (python_utils.py:60-71)."""
    return scipy.fft.rfftn(a, axes=(-2, -1), norm="ortho", workers=FFT_WORKERS)


def idft(a, shape):
    """Ortho irFFT to `shape` over the last two axes (python_utils.py:41-57)."""
    return scipy.fft.irfftn(a, s=shape, axes=(-2, -1), norm="ortho", workers=FFT_WORKERS)


def wblur_subsampling(arr, wpsf):
    """out[l',a] = sum_l sum_b arr[l,a,b] * wpsf[l',l,b]  (jax_utils.py:72-80; einsum form of
    the broadcast-sum)."""
    return np.einsum("lab,mlb->ma", arr, wpsf, optimize=True)


def wblur_t(arr, wpsf):
    """out[l,a,b] = sum_l' arr[l',a,b] * wpsf[l',l,b]  (jax_utils.py:83-91 /
    python_utils.py:160-180)."""
    return np.einsum("mab,mlb->lab", arr, wpsf, optimize=True)


def _intervals(grid, x):
    """find_interval_ascending + the distance rule of find_indices
    (surfh/ToolsDir/cythons_files.pyx:20-154): i with grid[i] <= x < grid[i+1], clamped to
    [0, n-2] outside the grid (extrapolate=1), x == grid[-1] -> n-2."""
    idx = np.clip(np.searchsorted(grid, x, side="right") - 1, 0, len(grid) - 2)
    return idx, (x - grid[idx]) / (grid[idx + 1] - grid[idx])


def bilinear_tables(grid_a, grid_b, pts_a, pts_b):
    """Indices and the four corner weights of solve_2D_hypercube (cythons_files.pyx:163-193)
    for every sample point."""
    i0, y0 = _intervals(grid_a, pts_a.ravel())
    i1, y1 = _intervals(grid_b, pts_b.ravel())
    w = np.stack([(1 - y0) * (1 - y1), (1 - y0) * y1, y0 * (1 - y1), y0 * y1])
    return i0, i1, w


def interpn_cube2local(alpha_axis, beta_axis, cube, pts_a, pts_b):
    """Bilinear sample of cube[l] at the points; any point outside the cube raises ValueError
    (cython_utils.py:10-30; bounds check cython_2D_interpolation.py:472-478)."""
    for d, (g, p) in enumerate(((alpha_axis, pts_a), (beta_axis, pts_b))):
        if not (np.all(g[0] <= p) and np.all(p <= g[-1])):
            raise ValueError("One of the requested xi is out of bounds in dimension %d" % d)
    i0, i1, w = bilinear_tables(alpha_axis, beta_axis, pts_a, pts_b)
    out = cube[:, i0, i1] * w[0]
    out = out + cube[:, i0, i1 + 1] * w[1]
    out = out + cube[:, i0 + 1, i1] * w[2]
    out = out + cube[:, i0 + 1, i1 + 1] * w[3]
    return out.reshape((cube.shape[0],) + pts_a.shape)


def interpn_local2cube(local_alpha, local_beta, local_cube, pts_a, pts_b):
    """Bilinear sample of the LOCAL cube at the points; points outside the local grid give 0
    (cython_utils.py:33-58; cython_2D_interpolation.py:316-323, 369-376).  The reference evaluates the
    extrapolated value everywhere and then overwrites the out-of-bounds points with 0; only the in-bounds
    points are evaluated here (same values, a third of the work at N = 501)."""
    pa, pb = pts_a.ravel(), pts_b.ravel()
    inb = ~((pa < local_alpha[0]) | (pa > local_alpha[-1]) | (pb < local_beta[0]) | (pb > local_beta[-1]))
    i0, i1, w = bilinear_tables(local_alpha, local_beta, pa[inb], pb[inb])
    val = local_cube[:, i0, i1] * w[0]
    val = val + local_cube[:, i0, i1 + 1] * w[1]
    val = val + local_cube[:, i0 + 1, i1] * w[2]
    val = val + local_cube[:, i0 + 1, i1 + 1] * w[3]
    out = np.zeros((local_cube.shape[0], pa.size))
    out[:, inb] = val
    return out.reshape((local_cube.shape[0],) + pts_a.shape)


def scatter_local2cube(alpha_axis, beta_axis, local_cube, pts_a, pts_b):
    """True transpose of interpn_cube2local: scatter-add every local sample onto the four cube
    pixels it was interpolated from (adjoint_mode="exact"; not in the reference)."""
    i0, i1, w = bilinear_tables(alpha_axis, beta_axis, pts_a, pts_b)
    n_l = local_cube.shape[0]
    flat = local_cube.reshape(n_l, -1)
    out = np.zeros((n_l, len(alpha_axis) * len(beta_axis)))
    nb = len(beta_axis)
    for (da, db), wk in zip(((0, 0), (0, 1), (1, 0), (1, 1)), w):
        cols = (i0 + da) * nb + (i1 + db)
        # np.add.at is slow but exact-in-order; group by column with bincount per wavelength
        for l in range(n_l):
            out[l] += np.bincount(cols, weights=flat[l] * wk, minlength=out.shape[1])
    return out.reshape((n_l, len(alpha_axis), len(beta_axis)))


# ------------------------------------------------------------------------- channel
class Channel:
    """One band: `Channel` of surfh/Models/spectroModelChannel.py:26-264."""

    def __init__(self, band, alpha_axis, beta_axis, wavel_axis, srf, pointings, step_degree,
                 adjoint_mode="reference"):
        self.alpha_axis = np.asarray(alpha_axis, dtype=np.float64)
        self.beta_axis = np.asarray(beta_axis, dtype=np.float64)
        self.step_degree = step_degree
        self.global_wavelength_axis = np.asarray(wavel_axis, dtype=np.float64)
        self.srf = int(srf)
        self.adjoint_mode = adjoint_mode
        # spectroModelChannel.py:43-44
        self.band = ins.Band.from_ifu(band).pixelised(step_degree)
        self.pointings = [(ins.pix(float(p[0]), step_degree), ins.pix(float(p[1]), step_degree))
                          for p in pointings]
        # :46-50
        self.local_alpha_axis, self.local_beta_axis = ins.local_axes(
            self.band.alpha_width, self.band.beta_width, step_degree, 5 * step_degree)
        # :53-59
        self.slicer = ins.SlitGeometry(self.band, self.beta_axis, self.local_alpha_axis,
                                       self.local_beta_axis, self.srf)
        self.wslice = ins.wslice(self.band, self.global_wavelength_axis, 0.1)  # :123-127
        self.n_wave = self.wslice.stop - self.wslice.start
        na = ceil(self.slicer.npix_slit_alpha_width / self.srf)
        # :63-65
        self.oshape = (len(self.pointings), self.band.n_slit, len(self.band.wavel_axis), na)
        self.local_im_shape = (len(self.local_alpha_axis), len(self.local_beta_axis))
        # :81-83  box kernel of srf pixels along alpha, non-normalised
        self._otf_sr = ir2fr(np.ones((self.srf, 1)), self.local_im_shape)[np.newaxis, ...]
        # :85-90, 133-143  LSF table; beta offsets in DEGREES against a um/arcsec scale
        nbw = self.slicer.npix_slit_beta_width
        beta_in_slit = np.arange(0, nbw) * (self.beta_axis[1] - self.beta_axis[0])
        self.wpsf = ins.lsf_table(self.band.grating_resolution, self.band.wavel_axis,
                                  beta_in_slit - np.mean(beta_in_slit),
                                  self.global_wavelength_axis[self.wslice],
                                  self.band.wavel_step / self.band.det_pix_size)
        # :104-108  shift that puts the box sum on its first pixel
        decal = np.zeros(self.local_im_shape)
        dsi = int((self.srf - 1) / 2)
        decal[-dsi, 0] = np.sqrt(self.local_im_shape[0] * self.local_im_shape[1])
        self.decalf = dft(decal)

    def _fov_origin(self, pointing):
        # (self.instr.fov + pointing): instru.py:409-410
        return (self.band.origin[0] + pointing[0], self.band.origin[1] + pointing[1])

    def gridding(self, blurred_cube, pointing):
        """:158-177"""
        ga, gb = ins.local2global(self.local_alpha_axis, self.local_beta_axis,
                                  self._fov_origin(pointing), self.band.angle)
        return interpn_cube2local(self.alpha_axis, self.beta_axis, blurred_cube, ga, gb)

    def gridding_t(self, local_cube, pointing):
        """:180-199 (reference) or the scatter transpose of `gridding` (exact)."""
        if self.adjoint_mode == "exact":
            ga, gb = ins.local2global(self.local_alpha_axis, self.local_beta_axis,
                                      self._fov_origin(pointing), self.band.angle)
            return scatter_local2cube(self.alpha_axis, self.beta_axis, local_cube, ga, gb)
        la, lb = ins.global2local(self.alpha_axis, self.beta_axis,
                                  self._fov_origin(pointing), self.band.angle)
        return interpn_local2cube(self.local_alpha_axis, self.local_beta_axis, local_cube, la, lb)

    def forward(self, blurred_cube):
        """:215-231"""
        out = np.zeros(self.oshape)
        na, srf = self.oshape[3], self.srf
        for p_idx, pointing in enumerate(self.pointings):
            gridded = self.gridding(blurred_cube[self.wslice], pointing)
            sum_cube = idft(dft(gridded) * (self._otf_sr * self.decalf), self.local_im_shape)
            for s in range(self.band.n_slit):
                sliced = self.slicer.slicing(sum_cube, s)
                out[p_idx, s] = wblur_subsampling(sliced[:, : na * srf: srf], self.wpsf)
        return out.ravel()

    def adjoint(self, inarray):
        """:234-264"""
        na, srf = self.oshape[3], self.srf
        y = np.reshape(inarray, self.oshape)
        nbw = self.slicer.npix_slit_beta_width
        inter = np.zeros((self.n_wave, len(self.alpha_axis), len(self.beta_axis)))
        local_shape = (self.n_wave,) + self.local_im_shape
        for p_idx, pointing in enumerate(self.pointings):
            local_cube = np.zeros(local_shape)
            for s in range(self.band.n_slit):
                over = np.repeat(y[p_idx, s][:, :, np.newaxis], nbw, axis=2)
                placed = np.zeros(self.slicer.slit_shape(self.n_wave))
                placed[:, : na * srf: srf, :] = wblur_t(over, self.wpsf.conj())
                local_cube += self.slicer.slicing_t(placed, s, local_shape)
            sum_t = idft(dft(local_cube) * self._otf_sr.conj() * self.decalf.conj(),
                         self.local_im_shape)
            inter += self.gridding_t(np.array(sum_t, dtype=np.float64), pointing)
        return inter


# --------------------------------------------------------------------------- model
class SpectroLMM(LinOp):
    """`spectroSigRLSCT` of surfh/Models/spectroModel.py:39-185:
    y = sum_bands SigR . L . Sum . S . C . T x   (templates=None -> no T, input is the cube)."""

    def __init__(self, sotf, templates, alpha_axis, beta_axis, wavelength_axis, instrs,
                 step_degree, pointings, adjoint_mode="reference"):
        self.sotf = np.asarray(sotf)
        self.templates = None if templates is None else np.asarray(templates, dtype=np.float64)
        self.alpha_axis = np.asarray(alpha_axis, dtype=np.float64)
        self.beta_axis = np.asarray(beta_axis, dtype=np.float64)
        self.wavelength_axis = np.asarray(wavelength_axis, dtype=np.float64)
        self.step_degree = step_degree
        self.lmm = self.templates is not None
        bands = [ins.Band.from_ifu(i) for i in instrs]
        # :67-70
        self.srfs = ins.get_srf([b.det_pix_size for b in bands], step_degree * 3600)
        # :122-133
        self.channels: List[Channel] = [
            Channel(b, self.alpha_axis, self.beta_axis, self.wavelength_axis, srf,
                    _as_pairs(pointings[it]), step_degree, adjoint_mode)
            for it, (srf, b) in enumerate(zip(self.srfs, bands))
        ]
        n_point = len(pointings[0])  # :98 -- every band is sized with len(pointings[0])
        self.instrs_oshape = [(n_point,) + ch.oshape[1:] for ch in self.channels]
        self._idx = np.cumsum([0] + [int(np.prod(s)) for s in self.instrs_oshape])  # :103
        self.list_wslice = [ch.wslice for ch in self.channels]
        self.cube_shape = (len(self.wavelength_axis), len(self.alpha_axis), len(self.beta_axis))
        self.imshape = self.cube_shape[1:]
        ishape = ((self.templates.shape[0],) + self.imshape) if self.lmm else self.cube_shape
        super().__init__(ishape=ishape, oshape=(int(self._idx[-1]),))

    def blurred_cube(self, maps):
        """T then C (spectroModel.py:160-166)."""
        cube = lmm_maps2cube(maps, self.templates) if self.lmm else maps
        return idft(dft(cube) * self.sotf, self.imshape)

    def forward(self, maps):
        """:158-170"""
        blurred = self.blurred_cube(np.asarray(maps, dtype=np.float64).reshape(self.ishape))
        out = np.zeros(self.oshape)
        for c, ch in enumerate(self.channels):
            out[self._idx[c]: self._idx[c + 1]] = ch.forward(blurred)
        return out

    def adjoint_cube(self, y):
        """Sum over bands into the global cube (:174-176)."""
        y = np.asarray(y, dtype=np.float64).ravel()
        cube = np.zeros(self.cube_shape)
        for c, ch in enumerate(self.channels):
            cube[self.list_wslice[c]] += ch.adjoint(y[self._idx[c]: self._idx[c + 1]])
        return cube

    def adjoint(self, y):
        """:173-185"""
        cube = self.adjoint_cube(y)
        blurred_t = idft(dft(cube) * self.sotf.conj(), self.imshape)
        return lmm_cube2maps(blurred_t, self.templates) if self.lmm else blurred_t

    def mapsToCube(self, maps):
        """Result export in float32 (spectroModel.py:190-192 -> matrix_op.py:198-202 ->
        cythons_files.pyx:424-440)."""
        m = np.asarray(maps, dtype=np.float32)
        return np.einsum("kij,kl->lij", m, self.templates.astype(np.float32)).astype(np.float32)

    def cubeTomaps(self, cube):
        """:187-188"""
        return lmm_cube2maps(cube, self.templates)


def _as_pairs(coord_list) -> Sequence:
    """CoordList of Coord(.alpha,.beta) or plain (alpha, beta) pairs -> list of pairs."""
    out = []
    for c in coord_list:
        out.append((c.alpha, c.beta) if hasattr(c, "alpha") else (c[0], c[1]))
    return out


# ----------------------------------------------------- regulariser and criterion
def diff_r(x):
    """NpDiff_r.forward (surfh/Simulation/fusion_CT.py:23-25): x[k,i-1,j] - x[k,i,j], circular."""
    return -np.diff(np.pad(x, ((0, 0), (1, 0), (0, 0)), "wrap"), axis=1)


def diff_r_t(y):
    """NpDiff_r.adjoint (fusion_CT.py:27-29)."""
    return np.diff(np.pad(y, ((0, 0), (0, 1), (0, 0)), "wrap"), axis=1)


def diff_c(x):
    """NpDiff_c.forward (fusion_CT.py:38-40)."""
    return -np.diff(np.pad(x, ((0, 0), (0, 0), (1, 0)), "wrap"), axis=2)


def diff_c_t(y):
    """NpDiff_c.adjoint (fusion_CT.py:42-43)."""
    return np.diff(np.pad(y, ((0, 0), (0, 0), (0, 1)), "wrap"), axis=2)


def criterion(model, y, x, mu_spectro, mu_reg):
    """get_crit_val, 'separated' gradients (fusion_CT.py:242-265)."""
    data_term = mu_spectro * np.sum((y - model.forward(x)) ** 2)
    reg = mu_reg * np.sum(diff_r(x) ** 2 + diff_c(x) ** 2)
    return (data_term + reg) / 2


def solve_lcg(model, y, mu_spectro, mu_reg, niter, value_init=0.0, tol=1e-12, callback=None,
              refresh=50):
    """QuadCriterion_MRS.run_method('lcg', ...) with 'separated' gradients
    (fusion_CT.py:118-238) on top of the restated qmm.lcg."""
    from .thirdparty import QuadObjective, lcg

    shape = model.ishape
    init = (np.ones(shape) * value_init) if np.isscalar(value_init) else np.asarray(value_init)
    objs = [QuadObjective(model.forward, model.adjoint, data=y, hyper=mu_spectro, name="Spectro"),
            QuadObjective(diff_r, diff_r_t, hyper=mu_reg),
            QuadObjective(diff_c, diff_c_t, hyper=mu_reg)]
    return lcg(objs, init, tol=tol, max_iter=niter, callback=callback, refresh=refresh)


def criterion_joint(model, y, x, mu_spectro, mu_reg):
    """get_crit_val with gradient='joint' (fusion_CT.py:249-250)."""
    from .thirdparty import laplacian2_circular
    data_term = mu_spectro * np.sum((y - model.forward(x)) ** 2)
    return (data_term + mu_reg * np.sum(laplacian2_circular(x) ** 2)) / 2


def objectives(model, y, mu_spectro, mu_reg, gradient="separated"):
    """The QuadObjective list QuadCriterion_MRS.run_method builds (fusion_CT.py:130-162)."""
    from .thirdparty import QuadObjective, laplacian2_circular
    objs = [QuadObjective(model.forward, model.adjoint, data=y, hyper=mu_spectro, name="Spectro")]
    if gradient == "joint":
        dtd = lambda v: laplacian2_circular(laplacian2_circular(v))  # noqa: E731
        objs.append(QuadObjective(laplacian2_circular, laplacian2_circular, dtd, hyper=mu_reg, name="Reg joint"))
    else:
        objs += [QuadObjective(diff_r, diff_r_t, hyper=mu_reg), QuadObjective(diff_c, diff_c_t, hyper=mu_reg)]
    return objs


def solve(model, y, mu_spectro, mu_reg, niter, method="lcg", gradient="separated", value_init=0.0, tol=1e-12,
          callback=None, refresh=50):
    """QuadCriterion_MRS.run_method(method, ...) (fusion_CT.py:118-238): 'lcg' or, for any other name,
    mmmg; 'separated' or 'joint' gradients."""
    from .thirdparty import lcg, mmmg
    init = (np.ones(model.ishape) * value_init) if np.isscalar(value_init) else np.asarray(value_init)
    objs = objectives(model, y, mu_spectro, mu_reg, gradient)
    if method == "lcg":
        return lcg(objs, init, tol=tol, max_iter=niter, callback=callback, refresh=refresh)
    return mmmg(objs, init, tol=tol, max_iter=niter, callback=callback)


def solve_huber(model, y, mu_spectro, spat_reg, spat_th, niter, value_init=0.0, tol=1e-12, callback=None):
    """`lmm_reconstruction` of surfh/ToolsDir/algorithms.py:71-106 on this operator: quadratic data term,
    Huber priors of threshold `spat_th` on the row and column differences, qmm.mmmg.  The differences are the
    circular NpDiff_r / NpDiff_c of fusion_CT.py:16-43 (the reference routine uses aljabr.Diff, absent here)."""
    from .thirdparty import Huber, Objective, QuadObjective, mmmg
    init = (np.ones(model.ishape) * value_init) if np.isscalar(value_init) else np.asarray(value_init)
    objs = [QuadObjective(model.forward, model.adjoint, data=y, hyper=mu_spectro, name="Data adeq"),
            Objective(diff_r, diff_r_t, Huber(spat_th), hyper=spat_reg, name="Row prior"),
            Objective(diff_c, diff_c_t, Huber(spat_th), hyper=spat_reg, name="Col prior")]
    return mmmg(objs, init, tol=tol, max_iter=niter, callback=callback)


def criterion_huber(model, y, x, mu_spectro, spat_reg, spat_th):
    from .thirdparty import Huber
    h = Huber(spat_th)
    return (mu_spectro * np.sum((y - model.forward(x)) ** 2) / 2
            + spat_reg * float(np.sum(h.value(diff_r(x)) + h.value(diff_c(x)))))

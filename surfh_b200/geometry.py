"""Host-side precompute: instrument description -> the flat tables the CUDA kernels consume.

Everything here runs once per model, in numpy fp64, and reproduces the reference's geometry
rules exactly (float comparisons in degrees, 9-decimal rounding of slit edges, the one-pixel
trim and the even-na adjustment), because a one-pixel disagreement moves the forward result at
the 1e-2 level.  Unlike the reference, which re-derives slices and weights inside every
forward/adjoint call, the result is a set of immutable tables:

  * slit tables       first local row / column of each slit, beta edge weights
  * gather tables     per pointing, per local grid point: upper-left cube pixel and the two
                      bilinear fractions (reference: find_indices, cythons_files.pyx:109-154)
  * adjoint tables    one CSR per flavour (exact transpose / the reference's gridding_t) mapping
                      each cube pixel to weighted entries of the slit-space vector
  * LSF table         W[lambda', lambda, beta]  (reference: SpectralBlur.psfs, instru.py:499-572)

Reference rules restated (paths relative to /root/reference):
  local axes            surfh/Models/instru.py:283-304
  local<->global        surfh/Models/instru.py:306-340
  slit extents          surfh/Models/instru.py:416-459, 612-626; surfh/Models/slicer.py:87-145
  slit weights          surfh/Models/slicer.py:148-168, 187-244
  wavelength window     surfh/Models/instru.py:649-658
  box-sum / decimation  surfh/Models/spectroModelChannel.py:81-83, 104-108, 220-229
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import instru

N_MARGIN_PIX = 5      # local grid margin, pixels each side (spectroModelChannel.py:46-48)
WAVE_MARGIN_UM = 0.1  # wavelength window margin (spectroModelChannel.py:123-127)
LSF_MARGIN = 15       # SpectralBlur._n_margin (instru.py:491-493)


@dataclass
class Csr:
    row_pixel: np.ndarray  # int32 [n_rows]
    row_ptr: np.ndarray    # int64 [n_rows + 1]
    col: np.ndarray        # int32 [nnz]
    val: np.ndarray        # float64 [nnz]

    @property
    def n_rows(self) -> int:
        return int(self.row_pixel.shape[0])

    @property
    def nnz(self) -> int:
        return int(self.col.shape[0])


@dataclass
class BandTables:
    name: str
    instr: instru.IFU                 # pixelised band
    pointings: instru.CoordList       # pixelised pointings
    srf: int
    local_alpha_axis: np.ndarray
    local_beta_axis: np.ndarray
    slices: List[Tuple[slice, slice]]
    weights: np.ndarray               # [S, nb]
    wslice: slice                     # the band's full wavelength window in the cube
    wave_local: slice                 # the part of it this process computes (lambda sharding)
    n_det: int
    na: int
    nb: int
    slit_a0: np.ndarray
    slit_b0: np.ndarray
    lsf: Optional[np.ndarray]         # [n_det, n_wave (local), nb]; None = beta-sum (MRSBlurred)
    grid_base: np.ndarray             # int32 [P, A*B]
    grid_frac: np.ndarray             # float64 [P, A*B, 2]
    adj_exact: Optional[Csr] = None
    adj_reference: Optional[Csr] = None
    out_offset: int = 0
    npix_slit_alpha_width: int = 0

    @property
    def n_pointing(self) -> int:
        return len(self.pointings)

    @property
    def n_slit(self) -> int:
        return int(self.instr.n_slit)

    @property
    def local_shape(self) -> Tuple[int, int]:
        return (len(self.local_alpha_axis), len(self.local_beta_axis))

    @property
    def n_wave(self) -> int:
        """Wavelengths handled locally (== the whole window unless lambda-sharded)."""
        return self.wave_local.stop - self.wave_local.start

    @property
    def is_local(self) -> bool:
        return self.n_wave > 0

    @property
    def oshape(self) -> Tuple[int, int, int, int]:
        return (self.n_pointing, self.n_slit, self.n_det, self.na)

    @property
    def ncol(self) -> int:
        return self.n_pointing * self.n_slit * self.na * self.nb


# ----------------------------------------------------------------------------- axes
def local_axes(fov: instru.FOV, step: float, margin: float) -> Tuple[np.ndarray, np.ndarray]:
    def axis(start: float, length: float) -> np.ndarray:
        first = int(math.floor(start / step)) * step
        count = int(math.ceil((length + (start - first)) / step))
        return np.arange(count + 1) * step + first

    return (axis(-fov.alpha_width / 2 - margin, fov.alpha_width + 2 * margin),
            axis(-fov.beta_width / 2 - margin, fov.beta_width + 2 * margin))


def _rot(angle_deg: float) -> Tuple[float, float]:
    t = np.radians(angle_deg)
    return float(np.cos(t)), float(np.sin(t))


def local_to_global(la: np.ndarray, lb: np.ndarray, origin: instru.Coord, angle: float):
    c, s = _rot(angle)
    a, b = la[:, None], lb[None, :]
    return c * a - s * b + origin.alpha, s * a + c * b + origin.beta


def global_to_local(ga: np.ndarray, gb: np.ndarray, origin: instru.Coord, angle: float):
    c, s = _rot(-angle)
    a, b = (ga - origin.alpha)[:, None], (gb - origin.beta)[None, :]
    return c * a - s * b, s * a + c * b


def _interval(grid: np.ndarray, x: np.ndarray):
    """Lower index i (clamped to [0, n-2]) and normalised distance (x - g[i]) / (g[i+1] - g[i])."""
    idx = np.clip(np.searchsorted(grid, x, side="right") - 1, 0, len(grid) - 2)
    return idx, (x - grid[idx]) / (grid[idx + 1] - grid[idx])


# ---------------------------------------------------------------------------- slits
def slit_layout(ifu: instru.IFU, beta_axis: np.ndarray, la: np.ndarray, lb: np.ndarray, srf: int,
                rules: str = "channel"):
    """Index ranges and beta weights of every slit of a band in its local grid.

    rules="channel": Slicer as used by Channel (slicer.py:118-168).  rules="blind": the private copies
    in MRSBlurred (spectro_blind.py:120-167): no even-na alpha adjustment, and the "next slit shares my
    last column" reset is only evaluated for slit_idx < npix_slit_beta_width - 1 (sic)."""
    n_slit = ifu.n_slit
    aw, bw = ifu.fov.alpha_width, ifu.fov.beta_width
    slit_w = bw / n_slit
    da, db = la[1] - la[0], lb[1] - lb[0]
    nbw = int(math.ceil(slit_w / (beta_axis[1] - beta_axis[0])))
    npix_alpha = int(math.ceil(aw / 2 / da)) - int(math.floor(-aw / 2 / da))
    na = int(math.ceil(npix_alpha / srf))

    a_start, a_end = 0.0 - aw / 2, 0.0 + aw / 2
    a_lo = int(np.flatnonzero(a_start < la + da / 2)[0])
    a_hi = int(np.flatnonzero(la - da / 2 < a_end)[-1]) + 1
    if rules == "channel" and na % 2 == 0 and na < 28:
        if a_hi - a_lo > npix_alpha:
            a_hi -= 1
        elif a_hi - a_lo < npix_alpha:
            a_lo -= 2

    edges = []
    b_ranges = []
    for s in range(n_slit):
        centre = 0 + ((-bw / 2 + slit_w / 2) + s * slit_w)
        b_start, b_end = round(centre - slit_w / 2, 9), round(centre + slit_w / 2, 9)
        lo = int(np.flatnonzero(b_start < lb + db / 2)[0])
        hi = int(np.flatnonzero(lb - db / 2 < b_end)[-1]) + 1
        if hi - lo > nbw:
            if abs(lb[hi] - b_end) > abs(lb[lo] - b_start):
                hi -= 1
            else:
                lo += 1
        edges.append((b_start, b_end))
        b_ranges.append((lo, hi))

    widths = {hi - lo for lo, hi in b_ranges}
    if widths != {nbw}:
        raise ValueError(f"band {ifu.name}: slits span {sorted(widths)} beta pixels, expected {nbw} "
                         "(the reference cannot broadcast such a band either)")
    if a_lo < 0 or (na - 1) * srf >= a_hi - a_lo:
        raise ValueError(f"band {ifu.name}: alpha range [{a_lo}, {a_hi}) cannot hold {na} detector rows of {srf} pixels")

    weights = np.ones((n_slit, nbw))
    for s, ((b_start, b_end), (lo, hi)) in enumerate(zip(edges, b_ranges)):
        if lb[lo] - db / 2 < b_start:
            w0 = 1 - abs(lb[lo] - db / 2 - b_start) / db
            assert 0 <= w0 <= 1, f"Weight of first beta observed pixel in slit must be in [0, 1] ({w0:.2f})"
            weights[s, 0] = w0
        if lb[hi - 1] + db / 2 > b_end:
            w1 = 1 - abs(lb[hi - 1] + db / 2 - b_end) / db
            assert 0 <= w1 <= 1, f"Weight of last beta observed pixel in slit must be in [0, 1] ({w1:.2f})"
            weights[s, -1] = w1
        if s > 0 and b_ranges[s - 1][1] - 1 != lo:
            weights[s, 0] = 1
        last_checked = (n_slit - 1) if rules == "channel" else min(n_slit - 1, nbw - 1)
        if s < last_checked and hi - 1 != b_ranges[s + 1][0]:
            weights[s, -1] = 1
    slices = [(slice(a_lo, a_hi), slice(lo, hi)) for lo, hi in b_ranges]
    return slices, weights, na, nbw, npix_alpha


# ------------------------------------------------------------------------------ LSF
def lsf_table(ifu: instru.IFU, cube_wavelengths: np.ndarray, nbw: int, beta_step: float) -> np.ndarray:
    """W[l', l, b]: sinc^2 response centred on l' - scale*beta_b, unit sum over l (the sum
    includes LSF_MARGIN-1 virtual samples on each side, which are then dropped)."""
    beta = np.arange(0, nbw) * beta_step
    beta = beta - np.mean(beta)
    scale = ifu.wavel_step / ifu.det_pix_size
    lam = np.asarray(cube_wavelengths, dtype=np.float64)
    dw = min(np.diff(lam))
    ext = np.concatenate([
        np.linspace(lam.min() - LSF_MARGIN * dw, lam.min() - dw, LSF_MARGIN - 1), lam,
        np.linspace(lam.max() + dw, lam.max() + LSF_MARGIN * dw, LSF_MARGIN - 1)])
    glen = ifu.w_blur.grating_len
    out = np.empty((ifu.n_wavel, len(lam), nbw))
    det = np.asarray(ifu.wavel_axis, dtype=np.float64)[:, None]
    for b in range(nbw):  # one beta at a time keeps the temporary at [L', L+28]
        w = np.pi * glen / ext[None, :] * np.sinc(np.pi * glen * ((det - scale * beta[b]) / ext[None, :] - 1)) ** 2
        w /= np.sum(w, axis=1, keepdims=True)
        out[:, :, b] = w[:, LSF_MARGIN - 1: -LSF_MARGIN + 1]
    return out


# ------------------------------------------------------------------- sparse adjoints
def _cover_tables(tb: "BandTables", p: int):
    """For every local grid point q = (i, j): the (at most two) slit-space columns whose box
    contains it and the slit weight.  -1 marks an empty slot."""
    A, B = tb.local_shape
    S, na, nb, srf = tb.n_slit, tb.na, tb.nb, tb.srf
    col = np.full((A, B, 2), -1, dtype=np.int64)
    wgt = np.zeros((A, B, 2))
    fill = np.zeros((A, B), dtype=np.int64)
    i = np.arange(A)
    for s in range(S):
        a0, b0 = int(tb.slit_a0[s]), int(tb.slit_b0[s])
        a = ((i - a0) % A) // srf
        rows = np.flatnonzero(a < na)
        for b in range(nb):
            j = b0 + b
            slot = fill[rows, j]
            if np.any(slot > 1):
                raise ValueError("more than two slits share a local column")
            col[rows, j, slot] = ((p * S + s) * na + a[rows]) * nb + b
            wgt[rows, j, slot] = tb.weights[s, b]
            fill[rows, j] += 1
    return col.reshape(A * B, 2), wgt.reshape(A * B, 2)


def _coo_to_csr(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray) -> Csr:
    keep = vals != 0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    if len(rows) == 0:
        return Csr(np.zeros(0, np.int32), np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0))
    new = np.ones(len(rows), dtype=bool)
    new[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    starts = np.flatnonzero(new)
    vals = np.add.reduceat(vals, starts)
    rows, cols = rows[starts], cols[starts]
    row_new = np.ones(len(rows), dtype=bool)
    row_new[1:] = rows[1:] != rows[:-1]
    row_starts = np.flatnonzero(row_new)
    row_ptr = np.concatenate([row_starts, [len(rows)]]).astype(np.int64)
    return Csr(rows[row_starts].astype(np.int32), row_ptr, cols.astype(np.int32), vals.astype(np.float64))


def build_adjoint_tables(tb: "BandTables", alpha_axis: np.ndarray, beta_axis: np.ndarray) -> None:
    """Fill tb.adj_exact and tb.adj_reference."""
    n_b = len(beta_axis)
    A, B = tb.local_shape
    ex_r, ex_c, ex_v = [], [], []
    rf_r, rf_c, rf_v = [], [], []
    for p, pointing in enumerate(tb.pointings):
        cov_col, cov_w = _cover_tables(tb, p)
        # exact: transpose of the gather.  tap t of local point q lands on pixel base[q] + off[t]
        base = tb.grid_base[p].astype(np.int64)
        y0, y1 = tb.grid_frac[p, :, 0], tb.grid_frac[p, :, 1]
        taps = ((0, (1 - y0) * (1 - y1)), (1, (1 - y0) * y1), (n_b, y0 * (1 - y1)), (n_b + 1, y0 * y1))
        for slot in range(2):
            ok = cov_col[:, slot] >= 0
            for off, w in taps:
                ex_r.append(base[ok] + off)
                ex_c.append(cov_col[ok, slot])
                ex_v.append(w[ok] * cov_w[ok, slot])
        # reference: bilinear sample of the local cube at every cube pixel, 0 outside the local grid
        origin = tb.instr.fov.origin + pointing
        lx, ly = global_to_local(alpha_axis, beta_axis, origin, tb.instr.fov.angle)
        lx, ly = lx.ravel(), ly.ravel()
        la, lb = tb.local_alpha_axis, tb.local_beta_axis
        inside = (lx >= la[0]) & (lx <= la[-1]) & (ly >= lb[0]) & (ly <= lb[-1])
        pix = np.flatnonzero(inside)
        i0, t0 = _interval(la, lx[pix])
        i1, t1 = _interval(lb, ly[pix])
        q00 = i0 * B + i1
        rtaps = ((0, (1 - t0) * (1 - t1)), (1, (1 - t0) * t1), (B, t0 * (1 - t1)), (B + 1, t0 * t1))
        for off, w in rtaps:
            q = q00 + off
            for slot in range(2):
                ok = cov_col[q, slot] >= 0
                rf_r.append(pix[ok])
                rf_c.append(cov_col[q[ok], slot])
                rf_v.append(w[ok] * cov_w[q[ok], slot])
    tb.adj_exact = _coo_to_csr(np.concatenate(ex_r), np.concatenate(ex_c), np.concatenate(ex_v))
    tb.adj_reference = _coo_to_csr(np.concatenate(rf_r), np.concatenate(rf_c), np.concatenate(rf_v))


# ---------------------------------------------------------------------------- bands
def build_band(ifu: instru.IFU, alpha_axis: np.ndarray, beta_axis: np.ndarray, wavelength_axis: np.ndarray,
               srf: int, pointings: Sequence[instru.Coord], step_degree: float,
               with_adjoint: bool = True, lambda_range: Optional[Tuple[int, int]] = None,
               rules: str = "channel") -> BandTables:
    """All tables of one band (reference: Channel.__init__, spectroModelChannel.py:27-108).

    `lambda_range` = (l0, l1) restricts the band to the cube wavelengths [l0, l1) (sharding of the
    wavelength axis across GPUs): the LSF keeps only those columns (its normalisation still runs
    over the band's full window), so forward yields this shard's partial sum of y."""
    alpha_axis = np.asarray(alpha_axis, dtype=np.float64)
    beta_axis = np.asarray(beta_axis, dtype=np.float64)
    wavelength_axis = np.asarray(wavelength_axis, dtype=np.float64)
    blind = rules == "blind"  # MRSBlurred: instrument and pointings used as given, every wavelength, no LSF
    band = ifu if blind else ifu.pix(step_degree)
    points = instru.CoordList(pointings) if blind else instru.CoordList(pointings).pix(step_degree)
    la, lb = local_axes(band.fov, step_degree, N_MARGIN_PIX * step_degree)
    slices, weights, na, nbw, npix_alpha = slit_layout(band, beta_axis, la, lb, srf, rules)
    wsl = slice(0, len(wavelength_axis)) if blind else band.wslice(wavelength_axis, WAVE_MARGIN_UM)
    if lambda_range is None:
        local = wsl
    else:
        lo, hi = max(wsl.start, int(lambda_range[0])), min(wsl.stop, int(lambda_range[1]))
        local = slice(lo, max(lo, hi))
    if blind:
        lsf = None
    elif local.stop > local.start:
        lsf = lsf_table(band, wavelength_axis[wsl], nbw, beta_axis[1] - beta_axis[0])
        lsf = np.ascontiguousarray(lsf[:, local.start - wsl.start: local.stop - wsl.start, :])
    else:
        lsf = np.zeros((band.n_wavel, 0, nbw))

    n_b = len(beta_axis)
    base = np.empty((len(points), len(la) * len(lb)), dtype=np.int32)
    frac = np.empty((len(points), len(la) * len(lb), 2))
    for p, pointing in enumerate(points):
        ga, gb = local_to_global(la, lb, band.fov.origin + pointing, band.fov.angle)
        for dim, (axis, pts) in enumerate(((alpha_axis, ga), (beta_axis, gb))):
            if not (np.all(axis[0] <= pts) and np.all(pts <= axis[-1])):
                raise ValueError("One of the requested xi is out of bounds in dimension %d" % dim)
        i0, t0 = _interval(alpha_axis, ga.ravel())
        i1, t1 = _interval(beta_axis, gb.ravel())
        base[p] = i0 * n_b + i1
        frac[p, :, 0], frac[p, :, 1] = t0, t1

    tb = BandTables(name=ifu.name, instr=band, pointings=points, srf=int(srf), local_alpha_axis=la,
                    local_beta_axis=lb, slices=slices, weights=weights, wslice=wsl, wave_local=local,
                    n_det=(wsl.stop - wsl.start) if blind else band.n_wavel,
                    na=na,
                    nb=nbw, slit_a0=np.array([s[0].start for s in slices], dtype=np.int32),
                    slit_b0=np.array([s[1].start for s in slices], dtype=np.int32), lsf=lsf, grid_base=base,
                    grid_frac=frac, npix_slit_alpha_width=npix_alpha)
    if with_adjoint and tb.is_local:
        build_adjoint_tables(tb, alpha_axis, beta_axis)
    return tb

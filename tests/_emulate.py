"""Test helper: apply the product's host-built TABLES with plain numpy, so that the geometry
precompute (surfh_b200.geometry) can be checked against the oracle without a GPU.  It mirrors
what the CUDA kernels do with the tables (gather, spectral contraction, CSR scatter); it is not
part of the product and never imported by it."""
import numpy as np


def gather(tb, cube_band, n_beta):
    """G[l, p, s, a, b] from the band's slice of the blurred cube [L, Na, Nb]."""
    L = cube_band.shape[0]
    flat = cube_band.reshape(L, -1)
    A, B = tb.local_shape
    P, S, na, nb, srf = tb.n_pointing, tb.n_slit, tb.na, tb.nb, tb.srf
    G = np.zeros((L, P, S, na, nb))
    for p in range(P):
        base = tb.grid_base[p].astype(np.int64)
        y0, y1 = tb.grid_frac[p, :, 0], tb.grid_frac[p, :, 1]
        grid = (flat[:, base] * ((1 - y0) * (1 - y1)) + flat[:, base + 1] * ((1 - y0) * y1)
                + flat[:, base + n_beta] * (y0 * (1 - y1)) + flat[:, base + n_beta + 1] * (y0 * y1))
        grid = grid.reshape(L, A, B)
        for s in range(S):
            a0, b0 = int(tb.slit_a0[s]), int(tb.slit_b0[s])
            for a in range(na):
                rows = (a0 + a * srf + np.arange(srf)) % A
                G[:, p, s, a, :] = grid[:, rows, b0:b0 + nb].sum(axis=1) * tb.weights[s][None, :]
    return G


def forward(model_tables, blurred, n_beta):
    """Detector vector from the blurred cube, bands concatenated like the reference."""
    out = []
    for tb in model_tables:
        G = gather(tb, blurred[tb.wave_local], n_beta)
        y = np.einsum("mlb,lpsab->psma", tb.lsf, G, optimize=True)
        out.append(y.ravel())
    return np.concatenate(out)


def adjoint_cube(model_tables, y, cube_shape, mode):
    """Global cube (before C^T T^T) from the detector vector, using the CSR tables."""
    cube = np.zeros(cube_shape)
    off = 0
    for tb in model_tables:
        n = int(np.prod(tb.oshape))
        yb = y[off:off + n].reshape(tb.oshape)
        off += n
        if tb.n_wave == 0:
            continue
        Gt = np.einsum("mlb,psma->lpsab", tb.lsf, yb, optimize=True).reshape(tb.n_wave, -1)
        csr = tb.adj_exact if mode == "exact" else tb.adj_reference
        contrib = np.zeros((tb.n_wave, cube_shape[1] * cube_shape[2]))
        counts = np.diff(csr.row_ptr)
        rows = np.repeat(np.arange(csr.n_rows), counts)
        for l in range(tb.n_wave):
            acc = np.bincount(rows, weights=csr.val * Gt[l, csr.col], minlength=csr.n_rows)
            contrib[l, csr.row_pixel] = acc
        cube[tb.wave_local] += contrib.reshape((tb.n_wave,) + tuple(cube_shape[1:]))
    return cube

"""The product's host precompute (surfh_b200.geometry) against the reference golden geometry and,
through a numpy emulation of what the kernels do with the tables, against the oracle.  No GPU."""
import os

import numpy as np
import pytest

import _emulate
from cases import CASES
from surfh_b200 import geometry, instru, synthetic
from surfh_oracle import model as om


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))


def tables_for(cfg, with_adjoint=True):
    srfs = instru.get_srf([i.det_pix_size for i in cfg.instrs], cfg.step_degree * 3600)
    return [geometry.build_band(i, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, s, cfg.pointings[k],
                                cfg.step_degree, with_adjoint=with_adjoint)
            for k, (i, s) in enumerate(zip(cfg.instrs, srfs))]


@pytest.mark.parametrize("name", ["mini_1band_1p", "mini_2band_4p", "c1_band1a", "band2a_4p"])
def test_tables_match_reference_geometry(golden_dir, name):
    cfg = CASES[name]()
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    for c, tb in enumerate(tables_for(cfg, with_adjoint=False)):
        sl = np.array([[a.start, a.stop, b.start, b.stop] for a, b in tb.slices])
        assert np.array_equal(sl, gold[f"b{c}_slices"])
        assert np.array_equal(tb.weights, gold[f"b{c}_weights"][:, : tb.nb])
        assert tb.oshape == tuple(gold[f"b{c}_oshape"])
        assert (tb.wslice.start, tb.wslice.stop) == tuple(gold[f"b{c}_wslice"])
        assert tb.local_shape == tuple(gold[f"b{c}_local_shape"])
        assert np.array_equal(tb.local_alpha_axis, gold[f"b{c}_local_alpha"])
        assert np.array_equal(tb.local_beta_axis, gold[f"b{c}_local_beta"])
        assert tb.srf == int(gold[f"b{c}_srf"]) and tb.nb == int(gold[f"b{c}_nbw"])
        assert np.allclose(tb.lsf[::7, ::5, :], gold[f"b{c}_wpsf_sample"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_tables_reproduce_oracle(mode):
    cfg = CASES["mini_2band_4p"]()
    tabs = tables_for(cfg)
    oracle = om.SpectroLMM(**cfg.model_args(), adjoint_mode=mode)
    blurred = oracle.blurred_cube(cfg.maps)
    assert rel(_emulate.forward(tabs, blurred, len(cfg.beta_axis)), oracle.forward(cfg.maps)) < 1e-13
    v = np.random.default_rng(1).standard_normal(oracle.osize)
    assert rel(_emulate.adjoint_cube(tabs, v, oracle.cube_shape, mode), oracle.adjoint_cube(v)) < 1e-13


def test_exact_csr_is_the_transpose_of_the_gather():
    cfg = CASES["mini_1band_1p"]()
    tb = tables_for(cfg)[0]
    n_b = len(cfg.beta_axis)
    rng = np.random.default_rng(2)
    cube = rng.standard_normal((1, len(cfg.alpha_axis), n_b))
    g = rng.standard_normal(tb.ncol)
    G = _emulate.gather(tb, cube, n_b).reshape(-1)
    csr = tb.adj_exact
    back = np.zeros(cube.size)
    rows = np.repeat(np.arange(csr.n_rows), np.diff(csr.row_ptr))
    back[csr.row_pixel] = np.bincount(rows, weights=csr.val * g[csr.col], minlength=csr.n_rows)
    assert abs(np.dot(G, g) - np.dot(cube.ravel(), back)) < 1e-12 * abs(np.dot(G, g))


def test_pointing_outside_the_cube_raises_like_the_reference():
    cfg = synthetic.mini_config(1, 1, n_pix=48)  # FoV larger than the cube
    with pytest.raises(ValueError, match="out of bounds"):
        tables_for(cfg, with_adjoint=False)
    with pytest.raises(ValueError, match="out of bounds"):
        om.SpectroLMM(**cfg.model_args()).forward(cfg.maps)


def test_instru_api_surface():
    c = instru.Coord(1.0, 2.0) + instru.Coord(0.5, -1.0)
    assert (c.alpha, c.beta) == (1.5, 1.0)
    with pytest.raises(ValueError):
        c + 3
    step = 0.025 / 3600
    pts = instru.CoordList([instru.Coord(1.3 * step, -2.6 * step)]).pix(step)
    assert pts[0].alpha == round(1.3) * step and pts[0].beta == round(-2.6) * step
    ifu = synthetic.make_band("2a")
    assert ifu.slit_beta_width == ifu.fov.beta_width / 17
    assert instru.get_srf([0.196, 0.245, 0.273], 0.025) == [7, 9, 10]
    wl = synthetic.cube_wavelength_axis(4.75, 28.9)
    assert len(wl) == 3612
    sl = ifu.wslice(wl, 0.1)
    assert wl[sl.start] <= ifu.wavel_min - 0.1 and wl[sl.stop] >= ifu.wavel_max + 0.1
    assert ifu.pix(step).name.endswith("_pix")

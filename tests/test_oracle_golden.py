"""Pins the CPU oracle (oracle/surfh_oracle) against golden vectors produced by the REFERENCE'S OWN
code (oracle/make_golden.py ran surfh/Models/spectroModel.py etc. from a scratch copy).  No GPU."""
import os

import numpy as np
import pytest

from cases import CASES, MINI
from surfh_oracle import instrument as oins
from surfh_oracle import model as om
from surfh_oracle import thirdparty as tp


def rel(a, b):
    return float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b)))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def check_geometry(model, gold):
    assert np.array_equal(model._idx, gold["idx"])
    for c, ch in enumerate(model.channels):
        sl = np.array([[a.start, a.stop, b.start, b.stop]
                       for a, b in (ch.slicer.slit_slices(s) for s in range(ch.band.n_slit))])
        assert np.array_equal(sl, gold[f"b{c}_slices"])
        for s in range(ch.band.n_slit):
            w = ch.slicer.slit_weights(s, ch.slicer.slit_slices(s))[0, 0, :]
            assert np.array_equal(w, gold[f"b{c}_weights"][s][: len(w)])
        assert tuple(gold[f"b{c}_oshape"]) == ch.oshape
        assert tuple(gold[f"b{c}_wslice"]) == (ch.wslice.start, ch.wslice.stop)
        assert tuple(gold[f"b{c}_local_shape"]) == ch.local_im_shape
        assert np.array_equal(ch.local_alpha_axis, gold[f"b{c}_local_alpha"])
        assert np.array_equal(ch.local_beta_axis, gold[f"b{c}_local_beta"])
        assert int(gold[f"b{c}_srf"]) == ch.srf
        assert np.allclose(ch.wpsf[::7, ::5, :], gold[f"b{c}_wpsf_sample"], rtol=1e-13, atol=0)
        assert np.allclose(ch.wpsf.sum(axis=(1, 2)), gold[f"b{c}_wpsf_sum"], rtol=1e-13)


@pytest.mark.parametrize("name", MINI)
def test_oracle_matches_reference_mini(golden_dir, name):
    cfg = CASES[name]()
    gold = load(golden_dir, name)
    model = om.SpectroLMM(**cfg.model_args(), adjoint_mode="reference")
    assert tuple(gold["ishape"]) == model.ishape
    check_geometry(model, gold)
    y = model.forward(cfg.maps)
    assert rel(y, gold["fwd"]) < 1e-13
    v = np.random.default_rng(1234).standard_normal(model.osize)
    assert rel(model.adjoint(v), gold["adj"]) < 1e-13
    blurred = model.blurred_cube(cfg.maps)
    assert rel(blurred[::5, ::3, ::3], gold["blurred_sample"]) < 1e-13
    ch0 = model.channels[0]
    assert rel(ch0.gridding(blurred[ch0.wslice], ch0.pointings[0])[::4], gold["gridded0"]) < 1e-13


@pytest.mark.parametrize("name", ["c1_band1a", "band3a_n501_4p", "band4a_n501_4p"])
def test_oracle_matches_reference_full_size(golden_dir, name):
    """Full-size single bands run by the reference's own code: 1A (C1, N = 251) and, at the north-star
    map size N = 501 with 4 dithers, one band of channel 3 (srf 9, nb 16) and one of channel 4 (srf 10, nb 26)."""
    cfg = CASES[name]()
    gold = load(golden_dir, name)
    model = om.SpectroLMM(**cfg.model_args(), adjoint_mode="reference")
    check_geometry(model, gold)
    y = model.forward(cfg.maps)
    assert rel(y[:: int(gold["fwd_stride"])], gold["fwd_sample"]) < 1e-13
    assert abs(np.linalg.norm(y) - float(gold["fwd_norm"])) < 1e-12 * float(gold["fwd_norm"])
    v = np.random.default_rng(1234).standard_normal(model.osize)
    x = model.adjoint(v)
    assert rel(x[:, ::7, ::7], gold["adj_sample"]) < 1e-13


def test_oracle_matches_reference_c3_blocks(golden_dir):
    """BASELINE.json configs[2] (C3: 1A, 2A, 3A, 4A on the 3612-plane axis, N = 501, 4 dithers) as run by the
    reference's own code: block offsets and geometry of the 4-band model (tables only, no cube is built), and
    band 2A's block of the reference output -- the band no single-band N = 501 fixture covers -- from a
    one-band oracle on that band's window of the same axis."""
    from cases import band_subconfig
    from surfh_b200 import geometry, instru
    cfg = CASES["c3"]()
    gold = load(golden_dir, "c3")
    idx = [0]
    srfs = instru.get_srf([i.det_pix_size for i in cfg.instrs], cfg.step_degree * 3600)
    for c, (ifu, srf) in enumerate(zip(cfg.instrs, srfs)):
        ws = ifu.wslice(cfg.wavelength_axis, 0.1)
        assert tuple(gold[f"b{c}_wslice"]) == (ws.start, ws.stop)
        assert int(gold[f"b{c}_srf"]) == srf
        idx.append(idx[-1] + int(np.prod(gold[f"b{c}_oshape"])))
    assert np.array_equal(idx, gold["idx"])
    band = 1
    model = om.SpectroLMM(**band_subconfig(cfg, band), adjoint_mode="reference")
    assert tuple(gold[f"b{band}_oshape"]) == model.channels[0].oshape
    y = model.forward(cfg.maps)
    stride = int(gold["fwd_stride"])
    lo, hi = int(gold["idx"][band]), int(gold["idx"][band + 1])
    first = -(-lo // stride) * stride          # first multiple of the stride inside the block
    want = gold["fwd_sample"][first // stride: (hi - 1) // stride + 1]
    assert rel(y[first - lo:: stride], want) < 1e-13


def test_geometry_all_twelve_bands_match_survey_table():
    """S, srf, local grid, na, nb per band as measured with the reference's Slicer."""
    from surfh_b200 import synthetic
    expect = {"1a": (21, 7, (139, 159), 129, 8, 19), "2a": (17, 7, (171, 203), 163, 12, 24),
              "3a": (16, 9, (219, 259), 211, 16, 24), "4a": (12, 10, (275, 319), 265, 26, 27)}
    step = synthetic.STEP_ARCSEC / 3600
    beta = np.arange(501) * step
    beta -= beta.mean()
    for name, (n_slit, srf, shape, a_len, nb, na) in expect.items():
        band = oins.Band.from_ifu(synthetic.make_band(name)).pixelised(step)
        assert oins.get_srf([band.det_pix_size], synthetic.STEP_ARCSEC) == [srf]
        la, lb = oins.local_axes(band.alpha_width, band.beta_width, step, 5 * step)
        assert (len(la), len(lb)) == shape
        geo = oins.SlitGeometry(band, beta, la, lb, srf)
        sa, sb = geo.slit_slices(0)
        assert band.n_slit == n_slit and sa.stop - sa.start == a_len and sb.stop - sb.start == nb
        assert geo.slices_shape[1] == na


def test_cg_criterion_and_regulariser_match_reference(golden_dir):
    """The reference's QuadCriterion_MRS (fusion_CT.py) driven by the restated lcg."""
    cfg = CASES["mini_1band_1p"]()
    gold = load(golden_dir, "mini_1band_1p")
    model = om.SpectroLMM(**cfg.model_args(), adjoint_mode="reference")
    fwd = model.forward(cfg.maps)
    y = fwd + 0.01 * np.sqrt(np.mean(fwd ** 2)) * np.random.default_rng(99).standard_normal(fwd.shape)
    assert abs(om.criterion(model, y, cfg.maps, 1, 5.0) - float(gold["crit_at_maps"])) < 1e-12 * float(gold["crit_at_maps"])
    assert np.allclose(om.diff_r(cfg.maps)[:, ::9, ::9], gold["diff_r"], rtol=0, atol=1e-15)
    assert np.allclose(om.diff_c_t(cfg.maps)[:, ::9, ::9], gold["diff_c_t"], rtol=0, atol=1e-15)
    crit = []

    def cb(res):
        if (len(res.grad_norm)) % 5 == 2:  # fusion_CT.py:172-175: self.it % 5 == 2 after increment
            crit.append(om.criterion(model, y, res.x.reshape(model.ishape), 1, 5.0))

    res = om.solve_lcg(model, y, 1, 5.0, 6, value_init=0, callback=cb)
    assert rel(res.x, gold["cg_x"]) < 1e-10
    assert np.allclose(res.grad_norm, gold["cg_grad_norm"], rtol=1e-9)
    assert np.allclose(crit, gold["cg_crit"], rtol=1e-10)


def test_ir2fr_consistent_with_in_tree_evidence():
    """udft.ir2fr is unpinned third-party arithmetic; check the property the reference relies on:
    the srf box kernel through ir2fr times the sqrt(A*B)-scaled delta equals an integer box-sum."""
    A, B, srf = 23, 17, 7
    g = np.random.default_rng(0).standard_normal((3, A, B))
    otf_sr = tp.ir2fr(np.ones((srf, 1)), (A, B))[np.newaxis]
    decal = np.zeros((A, B))
    decal[-int((srf - 1) / 2), 0] = np.sqrt(A * B)
    out = om.idft(om.dft(g) * otf_sr * om.dft(decal), (A, B))
    box = sum(np.roll(g, -m, axis=1) for m in range(srf))
    assert rel(out, box) < 1e-13


def test_dot_test_pairs_without_gridding_are_exact():
    """The reference asserts dot-tests for every stage pair except S (test/test_fw_ad.py); in exact
    mode the whole oracle operator must pass it too, in reference mode it must NOT (fact 3)."""
    cfg = CASES["mini_1band_1p"]()
    exact = om.SpectroLMM(**cfg.model_args(), adjoint_mode="exact")
    assert tp.dottest(exact, num=2, rtol=1e-10, seed=3)
    ref = om.SpectroLMM(**cfg.model_args(), adjoint_mode="reference")
    assert not tp.dottest(ref, num=1, rtol=1e-6, seed=3)

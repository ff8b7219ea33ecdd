"""`MRSBlurred` (surfh/Models/spectro_blind.py, BASELINE config 5): the oracle restatement against the
golden vectors made by the reference's own class (CPU), and the CUDA path against both (GPU), single
wavelength as the reference and batched over a cube as configuration 5 asks."""
import os

import numpy as np
import pytest

from cases import BLIND


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))


def blind_args(cfg, l_idx=None):
    sotf = cfg.sotf()
    return dict(sotf=sotf if l_idx is None else sotf[l_idx], alpha_axis=cfg.alpha_axis, beta_axis=cfg.beta_axis,
                instr=cfg.instrs[0], step_degree=cfg.step_degree, pointings=cfg.pointings[0])


@pytest.mark.parametrize("name", list(BLIND))
def test_oracle_blind_matches_reference(golden_dir, name):
    from surfh_oracle import blind
    factory, l_idx = BLIND[name]
    cfg = factory()
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    assert int(gold["l_idx"]) == l_idx
    m = blind.MRSBlurred(**blind_args(cfg, l_idx))
    assert tuple(gold["slices_shape"]) == tuple(m.slices_shape)
    sl = [m.get_slit_slices(s) for s in range(m.band.n_slit)]
    assert np.array_equal(np.array([[a.start, a.stop, b.start, b.stop] for a, b in sl]), gold["slices"])
    w = np.array([m.get_slit_weights(s, sl[s])[0, 0, :] for s in range(m.band.n_slit)])
    assert np.array_equal(w, gold["weights"])
    y = m.forward(cfg.maps[l_idx])
    assert rel(y, gold["fwd"]) < 1e-13
    v = np.random.default_rng(1234).standard_normal(y.shape[0])
    assert rel(m.adjoint(v), gold["adj"]) < 1e-13


def test_host_blind_geometry_matches_reference(golden_dir):
    """The product's host tables under rules="blind" reproduce MRSBlurred's own slicing copies."""
    from surfh_b200 import geometry, instru
    for name, (factory, _) in BLIND.items():
        cfg = factory()
        gold = np.load(os.path.join(golden_dir, name + ".npz"))
        srf = instru.get_srf([cfg.instrs[0].det_pix_size], cfg.step_degree * 3600)[0]
        tb = geometry.build_band(cfg.instrs[0], cfg.alpha_axis, cfg.beta_axis, np.arange(3.0), srf,
                                 cfg.pointings[0], cfg.step_degree, with_adjoint=False, rules="blind")
        got = np.array([[a.start, a.stop, b.start, b.stop] for a, b in tb.slices])
        assert np.array_equal(got, gold["slices"])
        assert np.array_equal(tb.weights, gold["weights"])
        assert (tb.n_pointing, tb.n_slit, tb.na) == tuple(gold["slices_shape"])
        assert tb.lsf is None and tb.n_det == 3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", list(BLIND))
def test_gpu_blind_single_wavelength(golden_dir, name, dtype):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200.spectro_blind import MRSBlurred
    from surfh_oracle import blind
    tol = {"float64": 1e-10, "float32": 1e-5}[dtype]
    factory, l_idx = BLIND[name]
    cfg = factory()
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    v = np.random.default_rng(1234).standard_normal(gold["fwd"].shape[0])
    for mode in ("reference", "exact"):
        gpu = MRSBlurred(**blind_args(cfg, l_idx), dtype=dtype, adjoint_mode=mode)
        cpu = blind.MRSBlurred(**blind_args(cfg, l_idx), adjoint_mode=mode)
        assert gpu.ishape == cpu.ishape and gpu.oshape == cpu.oshape
        y = gpu.forward(cfg.maps[l_idx])
        x = gpu.adjoint(v)
        assert x.shape == cpu.ishape
        assert rel(y, cpu.forward(cfg.maps[l_idx])) <= tol
        assert rel(x, cpu.adjoint(v)) <= tol
        if mode == "reference":
            assert rel(y, gold["fwd"]) <= tol
            assert rel(x, gold["adj"]) <= tol
            assert rel(gpu.real_data_janskySR_to_jansky(gold["fwd"].copy()), gold["jansky"]) <= 1e-14
        else:
            rng = np.random.default_rng(0)
            u, w = rng.standard_normal(gpu.isize), rng.standard_normal(gpu.osize)
            left, right = float(np.vdot(gpu.rmatvec(w), u)), float(np.vdot(w, gpu.matvec(u)))
            assert abs(left - right) <= (1e-11 if dtype == "float64" else 1e-4) * abs(right)


@pytest.mark.gpu
def test_gpu_blind_batched_equals_per_wavelength(golden_dir):
    """Configuration 5: every wavelength of a cube through its own OTF = L single-wavelength operators."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200.spectro_blind import MRSBlurred
    from surfh_oracle import blind
    factory, _ = BLIND["blind_mini_2p"]
    cfg = factory()
    n_l = len(cfg.wavelength_axis)
    gpu = MRSBlurred(**blind_args(cfg), dtype="float64", adjoint_mode="reference")
    assert gpu.ishape == (n_l,) + cfg.imshape
    y = gpu.forward(cfg.maps).reshape(n_l, -1)
    v = np.random.default_rng(5).standard_normal(y.shape)
    x = gpu.adjoint(v.ravel())
    for l in range(0, n_l, 5):
        cpu = blind.MRSBlurred(**blind_args(cfg, l))
        assert rel(y[l], cpu.forward(cfg.maps[l])) <= 1e-10
        assert rel(x[l], cpu.adjoint(v[l])) <= 1e-10

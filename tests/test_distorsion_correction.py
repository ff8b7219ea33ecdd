"""Distortion-correction pre-processing: the exponential modified-Shepard interpolation.
CPU: the numpy oracle against the output of the reference's own compiled Cython (tests/golden/shepard.npz, made by
oracle/make_golden.py::run_shepard).  GPU: `surfh_shepard` against both, and the slit loop of
`mrs_slices_distrorsion_correction`.  float32 arithmetic: tolerance 1e-5 relative L2."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {"p2": (2, 2.0, 2), "p15": (1.5, 1.0, 3)}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))


def inputs():
    import make_golden  # the seeded inputs live beside the script that ran the reference on them
    return make_golden.shepard_inputs()


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_matches_reference_cython(golden_dir, tag):
    from surfh_oracle import shepard
    d = inputs()
    gold = np.load(os.path.join(golden_dir, "shepard.npz"))[tag]
    p, a_exp, cut = CASES[tag]
    got = shepard.exponential_modified_shepard(d["alpha"], d["lam"], d["val"], d["alpha_mesh"], d["lambda_mesh"], p, a_exp,
                                               cut, d["alpha_res"], d["lambda_res"])
    assert got.shape == gold.shape and got.dtype == np.float32
    assert rel(got, gold) <= 2e-6
    assert np.count_nonzero(gold) > 0.9 * gold.size


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_gpu_shepard_matches_reference_and_oracle(golden_dir, tag):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200 import distorsion_correction as dc
    from surfh_oracle import shepard
    d = inputs()
    gold = np.load(os.path.join(golden_dir, "shepard.npz"))[tag]
    p, a_exp, cut = CASES[tag]
    got = dc.perform_shepard_interpolation(d["alpha"], d["lam"], d["val"], d["alpha_mesh"], d["lambda_mesh"], p, a_exp, cut,
                                           d["alpha_res"], d["lambda_res"])
    assert got.shape == gold.shape and got.dtype == np.float32
    assert rel(got, gold) <= 1e-5
    want = shepard.exponential_modified_shepard(d["alpha"], d["lam"], d["val"], d["alpha_mesh"], d["lambda_mesh"], p, a_exp,
                                                cut, d["alpha_res"], d["lambda_res"])
    assert rel(got, want) <= 1e-5
    # a grid point with no sample inside the cutoff is 0, like the reference
    far = dc.perform_shepard_interpolation(d["alpha"], d["lam"], d["val"], d["alpha_mesh"] + 1e3, d["lambda_mesh"], p, a_exp,
                                           cut, d["alpha_res"], d["lambda_res"])
    assert np.all(far == 0)


@pytest.mark.gpu
def test_gpu_slit_loop_like_the_reference_driver():
    """mrs_slices_distrorsion_correction (distorsion_correction.py:106-181) on a synthetic 3-slit detector image:
    labels sorted by centroid, every slit interpolated onto the channel's [L', na] grid, NaN samples dropped."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from types import SimpleNamespace
    from surfh_b200 import distorsion_correction as dc
    from surfh_oracle import shepard
    n_rows, width, gap, n_slit, na, n_det = 96, 12, 4, 3, 7, 40
    mask = np.zeros((n_rows, n_slit * (width + gap)), dtype=bool)
    for s in range(n_slit):
        mask[:, s * (width + gap) + 1: s * (width + gap) + 1 + width] = True
    labels = dc.sort_labels_by_centroid(dc.generate_label_image(mask))
    assert labels.max() == n_slit and labels[0, 2] == 1 and labels[0, (width + gap) * 2 + 2] == 3
    rng = np.random.default_rng(0)
    data = rng.standard_normal(mask.shape)
    data[5, 3] = np.nan

    def detector2world(x, y):  # columns, rows -> (alpha, beta, lambda): a sheared lattice per slit
        x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
        return 0.1 * (x % (width + gap)) + 0.002 * y, 0.0 * x, 5.0 + 0.01 * y + 0.0005 * x

    chan_wavelength = np.linspace(5.02, 5.9, n_det)
    channel = SimpleNamespace(oshape=(1, n_slit, n_det, na))
    out = dc.mrs_slices_distrorsion_correction(channel, labels, detector2world, data, chan_wavelength, mode=0)
    assert out.shape == (n_slit, n_det, na) and np.all(np.isfinite(out))
    # slit 1 by hand through the oracle
    rows, cols = np.where(labels == 1)
    alpha, _, lam = detector2world(cols, rows)
    val = data[rows, cols]
    ok = ~np.isnan(val)
    grid_alpha = np.linspace(alpha.min(), alpha.max(), na)
    amesh, lmesh = np.meshgrid(grid_alpha, chan_wavelength)
    ares = (grid_alpha.max() - grid_alpha.min()) / amesh.shape[1]
    lres = (chan_wavelength.max() - chan_wavelength.min()) / lmesh.shape[0]
    want = shepard.exponential_modified_shepard(alpha[ok], lam[ok], val[ok], amesh, lmesh, 2, 2.0, 2, ares, lres)
    assert rel(out[0], want) <= 1e-5

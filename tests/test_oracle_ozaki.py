"""CPU checks of the integer-sliced arithmetic the tcgen05 contraction uses (oracle/surfh_oracle/ozaki.py restates
csrc/kernels_ozaki.cuh): the digits reproduce the operands, the sliced product matches an extended-precision product to
the advertised accuracy on a real line-spread function, and the leading digits of that function vanish outside a band."""
import numpy as np

from surfh_b200 import geometry, synthetic
from surfh_oracle import ozaki


def _lsf_and_data(nb=6, seed=0):
    ifu = synthetic.make_band("1a")
    lam = synthetic.cube_wavelength_axis(4.85, 5.8)
    w = geometry.lsf_table(ifu, lam, nb, 1e-5)[::4]                 # [L'/4, L, nb]
    w = np.ascontiguousarray(w.reshape(w.shape[0], -1))
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((96, w.shape[1])) * np.exp(rng.standard_normal((96, w.shape[1])))
    return w, g


def test_digits_reproduce_the_rows():
    w, _ = _lsf_and_data()
    for digits in (4, 6, 8):
        planes, scale = ozaki.cut_rows(w, digits)
        assert planes.dtype == np.int8 and np.abs(planes.astype(int)).max() <= 64
        back = np.zeros_like(w, dtype=np.longdouble)
        for p in range(digits - 1, -1, -1):
            back = back / 128 + planes[p]
        err = np.abs(back * scale[:, None] - w).max(axis=1) / (64 * scale)      # relative to 2^e
        assert err.max() <= 2.0 ** (-7 - 7 * (digits - 1))
    zero = ozaki.cut_rows(np.zeros((3, 10)), 8)
    assert not zero[0].any() and np.all(zero[1] == 2.0 ** -6)


def test_sliced_product_against_extended_precision():
    w, g = _lsf_and_data()
    exact = np.asarray(w.astype(np.longdouble) @ g.astype(np.longdouble).T, dtype=np.longdouble)
    rel = lambda y: float(np.linalg.norm((y - exact).astype(np.float64)) / np.linalg.norm(exact.astype(np.float64)))  # noqa: E731
    e8, e7, e6, e4 = (rel(ozaki.product(w, g, d)) for d in (8, 7, 6, 4))
    e_fp64 = rel(w @ g.T)
    print(f"relative L2 vs long double: 8 digits {e8:.1e}, 7 {e7:.1e}, 6 {e6:.1e}, 4 {e4:.1e}; numpy fp64 {e_fp64:.1e}")
    # the digits are cut relative to each ROW's maximum, so on rows of high dynamic range (a peaked response against
    # log-normal data) 8 digits sit a few ulp above a straight fp64 product; on the operator's own data it is below it
    assert e8 <= 2e-14 and e7 <= 2e-12 and e6 <= 2e-10 and e4 <= 5e-6
    assert e_fp64 <= 2e-15


def test_leading_digits_of_the_response_are_banded():
    w, _ = _lsf_and_data(nb=8)
    planes, _ = ozaki.cut_rows(w, 8)
    mask = ozaki.tile_mask(planes)
    present = [float(np.mean((mask[:, 1:] >> p) & 1)) for p in range(8)]
    assert present[0] < 0.5 and present[1] < 0.9 and present[7] == 1.0
    assert np.all(mask[:, 0] == 0xff)

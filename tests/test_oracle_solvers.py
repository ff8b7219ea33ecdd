"""CPU checks of the restated third-party solvers (oracle/surfh_oracle/thirdparty.py): qmm.lcg and
qmm.mmmg on a small dense quadratic whose minimiser is known in closed form, and the joint
(Laplacian) regulariser of fusion_CT.py:45-63.  These pin the restatements to the mathematics; parity
with the qmm / udft packages themselves stays unpinned (not installable here)."""
import numpy as np

from surfh_oracle import thirdparty as tp


def _problem(seed=0, n=(2, 6, 5), m=57):
    rng = np.random.default_rng(seed)
    size = int(np.prod(n))
    a = rng.standard_normal((m, size))
    data = rng.standard_normal(m)
    fwd = lambda x: a @ np.ravel(x)                      # noqa: E731
    adj = lambda y: (a.T @ y).reshape(n)                 # noqa: E731
    return a, data, fwd, adj, n


def test_laplacian_is_the_circular_five_point_stencil():
    x = np.random.default_rng(1).standard_normal((3, 9, 8))
    want = 4 * x - np.roll(x, 1, 1) - np.roll(x, -1, 1) - np.roll(x, 1, 2) - np.roll(x, -1, 2)
    got = tp.laplacian2_circular(x)
    assert np.allclose(got, want, atol=1e-13)
    # D_r^T D_r + D_c^T D_c (the 'separated' prior of fusion_CT.py:16-43) is the same operator
    dr = lambda v: np.roll(v, 1, 1) - v                  # noqa: E731
    drt = lambda v: np.roll(v, -1, 1) - v                # noqa: E731
    dc = lambda v: np.roll(v, 1, 2) - v                  # noqa: E731
    dct = lambda v: np.roll(v, -1, 2) - v                # noqa: E731
    assert np.allclose(drt(dr(x)) + dct(dc(x)), want, atol=1e-13)


def test_lcg_and_mmmg_reach_the_closed_form_minimiser():
    a, data, fwd, adj, n = _problem()
    mu = 0.3
    objs = [tp.QuadObjective(fwd, adj, data=data, hyper=2.0),
            tp.QuadObjective(tp.laplacian2_circular, tp.laplacian2_circular, hyper=mu)]
    size = a.shape[1]
    lap = np.stack([tp.laplacian2_circular(e.reshape(n)).ravel() for e in np.eye(size)], axis=1)
    q = 2.0 * a.T @ a + mu * lap.T @ lap
    x_star = np.linalg.solve(q, 2.0 * a.T @ data).reshape(n)
    r_cg = tp.lcg(objs, np.zeros(n), tol=1e-14, max_iter=200)
    r_mm = tp.mmmg(objs, np.zeros(n), tol=1e-26, max_iter=200)
    assert np.allclose(r_cg.x, x_star, atol=1e-9)
    assert np.allclose(r_mm.x, x_star, atol=1e-9)
    assert r_mm.x.shape == n and r_mm.grad_norm[-1] < r_mm.grad_norm[0] * 1e-20


def test_mmmg_first_iterations_equal_cg_in_exact_arithmetic():
    """3MG on a quadratic searches span{-grad, previous move}: the same plane as CG's, so the iterates
    coincide (to rounding) with linear CG started at the same point."""
    a, data, fwd, adj, n = _problem(seed=3)
    objs = [tp.QuadObjective(fwd, adj, data=data, hyper=1.0)]
    for k in (1, 2, 5):
        r_cg = tp.lcg(objs, np.zeros(n), tol=0.0, max_iter=k, refresh=0)
        r_mm = tp.mmmg(objs, np.zeros(n), tol=0.0, max_iter=k)
        assert np.allclose(r_cg.x, r_mm.x, rtol=1e-8, atol=1e-10)

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of third-party arithmetic on the surfh hot path.

The reference imports three un-vendored packages whose arithmetic sits on the path
(SURVEY.md section 8c).  None is present under /root/reference, none is installed here and
there is no network, so their *published* algorithms are restated below from their public
documentation.  PARITY UNPINNED for everything in this file: the reference tree holds no test
or golden vector that pins `udft.ir2fr`, `aljabr.dottest` or `qmm.lcg` outputs.  What *is*
checked (tests/test_oracle_golden.py): with these stubs plugged in, the reference's own files
produce the committed golden vectors, and `ir2fr` is consistent with the in-tree evidence
cited in its docstring.

  udft   3.4.0   (PyPI;           /root/reference/poetry.lock:3803-3810)
  aljabr 0.4.0   (forieux/aljabr; /root/reference/poetry.lock:4-23)
  qmm    0.18.2  (forieux/qmm;    /root/reference/poetry.lock:3328-3344)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product (surfh_b200) never does.
"""
from __future__ import annotations

import time

import numpy as np


# ----------------------------------------------------------------------------- udft
def ir2fr(imp_resp, shape, origin=None, real=True):
    """udft.ir2fr: impulse response -> frequency response.

    Zero-pad `imp_resp` to `shape` over the last len(shape) axes (leading axes are batch),
    circularly shift so that the element at `origin` (default: floor(size/2) on every padded
    axis) lands on index 0, then take the NON-normalised (r)fftn over those axes.

    Call sites in the reference: scripts/main_fusion.py:98 (PSF -> OTF),
    surfh/Models/spectroModelChannel.py:81-83 (srf box kernel).  In-tree evidence that the
    transform is non-normalised: the companion delta in spectroModelChannel.py:104-108 is
    scaled by sqrt(A*B) precisely because it goes through the *ortho* dft while `_otf_sr`
    does not.
    """
    imp_resp = np.asarray(imp_resp)
    nd = len(shape)
    if origin is None:
        origin = [n // 2 for n in imp_resp.shape[-nd:]]
    full = imp_resp.shape[:-nd] + tuple(shape)
    padded = np.zeros(full, dtype=imp_resp.dtype)
    padded[(Ellipsis,) + tuple(slice(0, n) for n in imp_resp.shape[-nd:])] = imp_resp
    for ax, off in zip(range(-nd, 0), origin):
        padded = np.roll(padded, -off, axis=ax)
    axes = tuple(range(-nd, 0))
    return np.fft.rfftn(padded, axes=axes) if real else np.fft.fftn(padded, axes=axes)


def rdft2(arr):
    """udft.rdft2: ortho rfft over the last two axes."""
    return np.fft.rfftn(arr, axes=(-2, -1), norm="ortho")


def irdftn(arr, shape):
    """udft.irdftn: ortho irfft over the last len(shape) axes."""
    return np.fft.irfftn(arr, s=shape, axes=tuple(range(-len(shape), 0)), norm="ortho")


# --------------------------------------------------------------------------- aljabr
class LinOp:
    """aljabr.LinOp protocol: forward/adjoint on shaped arrays, matvec/rmatvec on flat ones.

    Seen through its use in the reference: ctor keywords surfh/Models/spectroModel.py:116,
    positional form surfh/Models/mixing.py:300; flat wrappers as used by
    test/sandbox_dottest.py:16-27; fwadj default noted at surfh/Models/mixing.py:270-272.
    """

    def __init__(self, ishape, oshape, name="_", dtype=np.float64):
        self.ishape = tuple(int(v) for v in ishape)
        self.oshape = tuple(int(v) for v in oshape)
        self.name = name
        self.dtype = dtype

    @property
    def isize(self):
        return int(np.prod(self.ishape))

    @property
    def osize(self):
        return int(np.prod(self.oshape))

    @property
    def shape(self):
        return (self.osize, self.isize)

    def forward(self, x):
        raise NotImplementedError

    def adjoint(self, y):
        raise NotImplementedError

    def matvec(self, x):
        return np.ravel(self.forward(np.reshape(x, self.ishape)))

    def rmatvec(self, y):
        return np.ravel(self.adjoint(np.reshape(y, self.oshape)))

    def fwadj(self, x):
        return self.adjoint(self.forward(x))

    def __call__(self, x):
        return self.forward(x)


def dottest_values(linop, rng):
    """One draw of the two inner products <A^T v, u> and <v, A u> with randn vectors."""
    u = rng.standard_normal(linop.isize)
    v = rng.standard_normal(linop.osize)
    return float(np.vdot(linop.rmatvec(v), u)), float(np.vdot(v, linop.matvec(u)))


def dottest(linop, num=1, rtol=1e-5, atol=1e-8, echo=False, seed=None):
    """aljabr.dottest: True when <A^T v, u> == <v, A u> (np.allclose) for `num` random draws."""
    rng = np.random.default_rng(seed)
    ok = True
    for _ in range(num):
        left, right = dottest_values(linop, rng)
        if echo:
            print(f"(A^T v)^T u = {left} ~= {right} = v^T (A u), rel = {abs(left - right) / abs(right):.3e}")
        ok = ok and bool(np.allclose(left, right, rtol=rtol, atol=atol))
    return ok


# ------------------------------------------------------------------------------ qmm
class OptimizeResult(dict):
    """Attribute-style dict (scipy.optimize.OptimizeResult look-alike) returned by lcg."""

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class QuadObjective:
    """qmm.QuadObjective: J(x) = hyper/2 * ||V x - data||^2.

    Constructed as in surfh/Simulation/fusion_CT.py:130-162:
    QuadObjective(forward, adjoint[, hessp], data=..., hyper=..., name=...).
    hessp(x) = V^T V x (composed from adjoint(forward(x)) when not given);
    the objective's share of the CG right-hand side is hyper * V^T data.
    """

    def __init__(self, operator, adjoint, hessp=None, data=None, hyper=1.0, name=""):
        self.operator = operator
        self.adjoint = adjoint
        self._hessp = hessp
        self.data = data
        self.hyper = hyper
        self.name = name
        self.ht_data = None if data is None else adjoint(data)

    def hessp(self, x):
        if self._hessp is not None:
            return self._hessp(x)
        return self.adjoint(self.operator(x))

    def value(self, x):
        r = self.operator(x) if self.data is None else self.operator(x) - self.data
        return self.hyper * float(np.sum(np.abs(r) ** 2)) / 2

    def gradient(self, x):
        g = self.hessp(x)
        return self.hyper * (g if self.ht_data is None else g - self.ht_data)

    def __repr__(self):
        return f"QuadObjective(name={self.name!r}, hyper={self.hyper})"


class Huber:
    """qmm.Huber(delta): phi(u) = u^2 / 2 for |u| <= delta, delta |u| - delta^2 / 2 beyond; convex, and
    majorised at any point by a quadratic of curvature phi'(u) / u = min(1, delta / |u|) (Geman-Reynolds
    coefficients, `gr_coeffs`).  PARITY UNPINNED (qmm 0.18.2 is not in the reference tree); used by
    surfh/ToolsDir/algorithms.py:27-106 (`vox_reconstruction`, `lmm_reconstruction`)."""

    def __init__(self, delta):
        self.delta = float(delta)
        self.inf = 1.0

    def value(self, u):
        a = np.abs(u)
        return np.where(a <= self.delta, u ** 2 / 2, self.delta * a - self.delta ** 2 / 2)

    def gradient(self, u):
        return np.where(np.abs(u) <= self.delta, u, self.delta * np.sign(u))

    def gr_coeffs(self, u):
        a = np.abs(u)
        return np.where(a <= self.delta, 1.0, self.delta / np.maximum(a, np.finfo(float).tiny))


class Objective:
    """qmm.Objective(operator, adjoint, loss, data=None, hyper=1): J(x) = hyper * sum loss(V x - data);
    gradient hyper * V^T loss'(V x - data); majorant curvature in a subspace spanned by `vecs` (given as
    V vecs): hyper * (V vecs)^T diag(gr_coeffs(V x - data)) (V vecs)."""

    def __init__(self, operator, adjoint, loss, data=None, hyper=1.0, name=""):
        self.operator, self.adjoint, self.loss = operator, adjoint, loss
        self.data = 0.0 if data is None else data
        self.hyper, self.name = hyper, name

    def value(self, x):
        return self.hyper * float(np.sum(self.loss.value(self.operator(x) - self.data)))

    def gradient(self, x):
        return self.hyper * self.adjoint(self.loss.gradient(self.operator(x) - self.data))

    def norm_mat_major(self, op_vecs, x):
        w = self.loss.gr_coeffs(self.operator(x) - self.data).reshape((-1, 1))
        return self.hyper * (op_vecs.T @ (w * op_vecs))


def lcg(objv_list, x0, tol=1e-4, max_iter=500, min_iter=0, callback=None, refresh=50):
    """qmm.lcg: unpreconditioned linear conjugate gradient on  Q x = b,
    Q = sum_i hyper_i V_i^T V_i,  b = sum_i hyper_i V_i^T data_i   (call site
    surfh/Simulation/fusion_CT.py:194-232).

    Textbook recurrences: r0 = b - Q x0, d0 = r0; a = <r,r>/<d,Qd>; x += a d;
    r -= a Qd (recomputed exactly as b - Q x when iteration % refresh == 0, iteration counted
    from 0); beta = <r+,r+>/<r,r>; d = r+ + beta d.  `grad_norm` collects <r,r>; the loop
    stops once sqrt(grad_norm[-1]) < x0.size * tol (after min_iter) or at max_iter.
    The callback receives the running result (fields x, grad_norm), as
    fusion_CT.py:164-175 expects.
    """
    if isinstance(objv_list, QuadObjective):
        objv_list = [objv_list]
    shape = np.shape(x0)

    def hessian(flat):
        arr = np.reshape(flat, shape)
        acc = None
        for obj in objv_list:
            term = obj.hyper * np.asarray(obj.hessp(arr))
            acc = term if acc is None else acc + term
        return np.reshape(acc, (-1,))

    second = np.zeros(int(np.prod(shape)))
    for obj in objv_list:
        if obj.ht_data is not None:
            second = second + obj.hyper * np.reshape(obj.ht_data, (-1,))

    res = OptimizeResult()
    res["x"] = np.array(x0, dtype=np.float64).reshape((-1,)).copy()
    res["success"] = True
    res["status"] = 99
    res["nit"] = max_iter
    res["grad_norm"] = []
    res["time"] = [time.time()]

    residual = second - hessian(res["x"])
    direction = residual.copy()
    res["grad_norm"].append(float(np.vdot(residual, residual)))

    for iteration in range(max_iter):
        hess_dir = hessian(direction)
        step = res["grad_norm"][-1] / float(np.vdot(direction, hess_dir))
        res["x"] += step * direction
        if refresh and iteration % refresh == 0:
            residual = second - hessian(res["x"])
        else:
            residual = residual - step * hess_dir
        res["grad_norm"].append(float(np.vdot(residual, residual)))
        direction = residual + (res["grad_norm"][-1] / res["grad_norm"][-2]) * direction
        res["time"].append(time.time())
        if callback is not None:
            callback(res)
        if np.sqrt(res["grad_norm"][-1]) < res["x"].size * tol and iteration >= min_iter:
            res["status"] = 1
            res["nit"] = iteration + 1
            break
    res["x"] = res["x"].reshape(shape)
    return res


def mmmg(objv_list, x0, tol=1e-4, max_iter=500, min_iter=0, callback=None):
    """qmm.mmmg (Majorize-Minimize Memory Gradient, "3MG") as called by
    surfh/Simulation/fusion_CT.py:196-197 for any method other than 'lcg'.  PARITY UNPINNED: qmm is
    not installed here and the reference holds no solve trace; restated from the published algorithm
    (Chouzenoux, Idier, Moussaoui 2011; qmm 0.18 `mmmg`):

        move = 0; op_directions_i = [V_i move, V_i move]; step = (1, 1)
        each iteration:
            grad = sum_i gradient_i(x)                      (hyper_i V_i^T (V_i x - data_i))
            grad_norm.append(|grad|^2); stop when grad_norm[-1] < x.size * tol (after min_iter)
            directions = [-grad, move]                      (n x 2)
            op_directions_i = [V_i(-grad), op_directions_i @ step]
            step = -lstsq(sum_i hyper_i op_directions_i^T op_directions_i, directions^T grad)
            move = directions @ step;  x += move;  callback(res)

    For QuadObjective the majorant curvature `norm_mat_major` is hyper * vecs^T vecs, i.e. the exact
    2 x 2 Hessian restricted to the plane, so every iteration is an exact plane search; an `Objective` with a
    non-quadratic loss (Huber) contributes hyper * (V vecs)^T diag(gr_coeffs(V x - data)) (V vecs).
    """
    if isinstance(objv_list, (QuadObjective, Objective)):
        objv_list = [objv_list]
    shape = np.shape(x0)
    vect = lambda op, v: np.reshape(op(np.reshape(v, shape)), (-1, 1))  # noqa: E731

    res = OptimizeResult()
    res["x"] = np.array(x0, dtype=np.float64).reshape((-1, 1)).copy()
    res["success"] = False
    res["status"] = 99
    res["nit"] = max_iter
    res["grad_norm"] = []
    res["time"] = [time.time()]
    move = np.zeros_like(res["x"])
    op_directions = [np.tile(vect(obj.operator, move), 2) for obj in objv_list]
    step = np.ones((2, 1))
    for iteration in range(max_iter):
        arr = res["x"].reshape(shape)
        grad = np.zeros_like(res["x"])
        for obj in objv_list:
            grad = grad + np.reshape(obj.gradient(arr), (-1, 1))
        res["grad_norm"].append(float(np.sum(grad ** 2)))
        if res["grad_norm"][-1] < res["x"].size * tol and iteration >= min_iter:
            res["success"] = True
            res["status"] = 1
            res["nit"] = iteration
            break
        directions = np.c_[-grad, move]
        op_directions = [np.c_[vect(obj.operator, directions[:, 0]), prev @ step]
                         for obj, prev in zip(objv_list, op_directions)]
        mat = sum(obj.norm_mat_major(od, arr) if isinstance(obj, Objective) else obj.hyper * (od.T @ od)
                  for obj, od in zip(objv_list, op_directions))
        step = -np.linalg.lstsq(mat, directions.T @ grad, rcond=None)[0]
        move = directions @ step
        res["x"] += move
        res["time"].append(time.time())
        if callback is not None:
            callback(res)
    res["x"] = res["x"].reshape(shape)
    return res


def laplacian2_circular(x):
    """udft.laplacian(2) = [[0,-1,0],[-1,4,-1],[0,-1,0]] applied the way
    surfh/Simulation/fusion_CT.py:45-63 (Difference_Operator_Joint.D) does: ir2fr of the CENTRED impulse
    response, product in Fourier space, inverse transform -- i.e. the circular convolution
    4 x - x[i-1] - x[i+1] - x[j-1] - x[j+1] on the last two axes (the kernel is symmetric, so D_t = D and
    DtD = D D).  Evaluated here through the same Fourier route, with ir2fr as restated above."""
    kernel = np.array([[0.0, -1.0, 0.0], [-1.0, 4.0, -1.0], [0.0, -1.0, 0.0]])
    shape = x.shape[-2:]
    d_freq = ir2fr(kernel, shape)
    return np.fft.irfftn(d_freq * np.fft.rfftn(x, axes=(-2, -1)), s=shape, axes=(-2, -1))

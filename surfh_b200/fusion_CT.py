"""Quadratic-regularised fusion solve, API of `surfh.Simulation.fusion_CT`, run on the device.

    J(x) = 1/2 ( mu_spectro * ||y - H x||^2  +  mu_reg * (||D_r x||^2 + ||D_c x||^2) )

solved by linear conjugate gradient on  Q x = b,  Q = mu_s H^T H + mu_r (D_r^T D_r + D_c^T D_c),
b = mu_s H^T y.  The whole iteration stays on the GPU: H^T H through `surfh_fwadj`, the
regulariser stencil + <d,Qd>, the x/r updates + <r,r> and the direction update are three fused
kernels whose scalars never visit the host; the host reads the gradient norm only when a callback
or the stopping test asks for it.

Reference interface mirrored (paths relative to /root/reference):
    NpDiff_r, NpDiff_c                       surfh/Simulation/fusion_CT.py:16-43
    QuadCriterion_MRS(...).run_method(...)   surfh/Simulation/fusion_CT.py:67-238
    QuadCriterion_MRS.get_crit_val           surfh/Simulation/fusion_CT.py:242-265
    Difference_Operator_Joint (gradient="joint")   surfh/Simulation/fusion_CT.py:45-63
    qmm.lcg, qmm.mmmg (third party, restated; recurrences and stopping rules as documented in
    the test oracle's thirdparty module -- parity with qmm itself is unpinned)

With gradient="joint" the prior is mu_reg * ||L x||^2, L = the circular 5-point Laplacian (the
udft.laplacian(2) impulse response the reference applies in Fourier space), i.e. Q_reg = mu_reg L L;
with gradient="separated" Q_reg = mu_reg (D_r^T D_r + D_c^T D_c) = mu_reg L.
"""
from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np

from . import _capi
from .linop import LinOp

REFRESH_PERIOD = 50  # lcg recomputes r = b - Q x exactly on iterations 0, 50, 100, ...


class NpDiff_r(LinOp):
    """Circular first difference along rows: (D x)[k,i,j] = x[k,i-1,j] - x[k,i,j]."""

    def __init__(self, maps_shape):
        super().__init__(ishape=maps_shape, oshape=maps_shape)

    def forward(self, x):
        return np.roll(x, 1, axis=1) - x

    def adjoint(self, y):
        return np.roll(y, -1, axis=1) - y


class NpDiff_c(LinOp):
    """Circular first difference along columns: (D x)[k,i,j] = x[k,i,j-1] - x[k,i,j]."""

    def __init__(self, maps_shape):
        super().__init__(ishape=maps_shape, oshape=maps_shape)

    def forward(self, x):
        return np.roll(x, 1, axis=2) - x

    def adjoint(self, y):
        return np.roll(y, -1, axis=2) - y


class OptimizeResult(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _torch():
    import torch
    return torch


def estimate_lambda_weights(model) -> np.ndarray:
    """Mean gain w_l of the detector sampling A^T A (A = Sig R . L . Sum . S, summed over bands and dithers) on
    smooth images at every cube wavelength, from the host tables alone.  For a locally constant plane the
    box-sum over srf rows gives srf per kept row, the spectral response spreads it with the LSF, and A^T brings
    srf * sum_l' W[l', l, b]^2 back onto each covered pixel and dither:
        w_l = sum_bands  P * srf * mean_b sum_l' W_band[l', l, b]^2        (0 where no band observes).
    Only the quality of the preconditioner depends on it, never the solution."""
    w = np.zeros(len(model.wavelength_axis))
    for it in model.local_bands:
        t = model.band_tables[it]
        if t.lsf is None:
            raise ValueError("the Fourier-domain preconditioner is defined for bands with a spectral response")
        gain = np.einsum("mlb,mlb->l", t.lsf, t.lsf) / t.nb
        w[t.wave_local] += t.n_pointing * t.srf * gain
    return w


class FourierPreconditioner:
    """P ~ (mu_s H^T H + mu_r R)^-1 with H^T H replaced by its shift-invariant part  T^T C^T (w_l) C T:
    per spatial frequency a K x K block  mu_s sum_l w_l |OTF_l|^2 T_l T_l^T + mu_r d(f)^p I, inverted once on
    the device; applying it costs two K-map FFTs and one pointwise K x K product (SURVEY section 8f-3;
    reference: surfh/Models/mixing.py:131-207, surfh/ToolsDir/fusion_mixing.py:401-438)."""

    def __init__(self, model, mu_spectro: float, mu_reg: float, gradient: str = "separated", weights=None,
                 scale: float = 1.0):
        if getattr(model, "partial", False):
            raise ValueError("the Fourier-domain preconditioner needs the unsharded operator on this GPU")
        if not model.lmm:
            raise ValueError("the Fourier-domain preconditioner needs templates (LMM model)")
        self.model = model
        w = estimate_lambda_weights(model) if weights is None else np.asarray(weights, dtype=np.float64)
        if w.shape != (len(model.wavelength_axis),):
            raise ValueError("weights must have one entry per cube wavelength")
        self.weights = np.ascontiguousarray(w * float(scale))
        _capi.check(model.handle, model._lib.surfh_precond_build(model.handle, _capi.ptr(self.weights), float(mu_spectro),
                                                                 float(mu_reg), 1 if gradient == "joint" else 0))

    def apply(self, r, z):
        """z = P r on flat device tensors of the model's input size."""
        torch = _torch()
        _capi.check(self.model.handle, self.model._lib.surfh_precond_apply(
            self.model.handle, r.data_ptr(), z.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return z


class DeviceCG:
    """Device-resident state of one lcg solve on a `spectroSigRLSCT` model."""

    def __init__(self, model, y, mu_spectro: float, mu_reg: float, comm=None, gradient: str = "separated"):
        torch = _torch()
        if gradient not in ("separated", "joint"):
            raise ValueError("gradient must be 'separated' or 'joint'")
        self.gradient = gradient
        self._lap = None
        self.model = model
        self.lib = model._lib
        self.h = model.handle
        self.mu_s, self.mu_r = float(mu_spectro), float(mu_reg)
        # the model owns the collectives (its comm sums the partial results of a sharded operator); a comm
        # given here must be that same communicator -- it is never injected into the model
        model_comm = getattr(model, "comm", None)
        if comm is not None and getattr(model, "partial", False) and model_comm is None:
            raise ValueError("a sharded model must be built with comm=...; passing comm to the solver only "
                             "would leave its partial results un-summed")
        self.comm = model_comm if model_comm is not None else comm
        self.tdtype = model._torch_dtype()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.y = self._to_dev(y).reshape(-1)
        if self.y.numel() != model.osize:
            raise ValueError(f"data has {self.y.numel()} samples, model produces {model.osize}")
        self.n = model.isize
        self.x = self.b = self.r = None
        self._y_sq = None       # |y|^2, for the criterion from the CG state
        self._state_evals = 0   # criterion evaluations served from the CG state (no forward pass)
        # track_hx: keep H x_k next to the iterate by linearity (H x_{k+1} = H x_k + alpha H d, H d being the
        # detector vector the fused H^T H call leaves behind) so that J(x_k) needs no forward pass whatever the
        # adjoint flavour.  One extra axpy over the detector vector per iteration; unsharded models only.
        self.track_hx = False
        self.hx = self.hd = None
        self.precond = None     # FourierPreconditioner: preconditioned CG (qmm.lcg's precond= argument)
        self.z = None
        # sharded model: H d is the model's exchanged detector vector; this rank tracks H x_k on the blocks it owns
        self._sharded_track = False
        self._blocks = None

    def _to_dev(self, a):
        torch = _torch()
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device=self.dev, dtype=self.tdtype).contiguous()

    def _stream(self):
        return _torch().cuda.current_stream().cuda_stream

    def _check(self, code):
        _capi.check(self.h, code)

    def laplacian(self, x, out, a: float = 0.0, b: float = 1.0):
        """out = a * out + b * L x (every map, circular)."""
        self._check(self.lib.surfh_laplacian_axpby(self.h, x.data_ptr(), out.data_ptr(), float(a), float(b),
                                                   self._stream()))
        return out

    def hessp(self, v, out):
        """out = mu_s H^T H v + mu_r R v ; s[1] = <v, out>, with R = L ('separated': D_r^T D_r + D_c^T D_c)
        or R = L L ('joint')."""
        self.model.fwadj_into(v, out, y_scratch=None if self._sharded_track else self.hd)
        if self.gradient == "separated":
            self._check(self.lib.surfh_cg_regularise_dot(self.h, v.data_ptr(), out.data_ptr(), self.mu_s, self.mu_r,
                                                         self.s.data_ptr(), self._stream()))
            return
        if self._lap is None:
            self._lap = _torch().empty_like(out)
        self.laplacian(v, self._lap)
        self.laplacian(self._lap, out, a=self.mu_s, b=self.mu_r)      # out = mu_s H^T H v + mu_r L L v
        self._check(self.lib.surfh_cg_regularise_dot(self.h, v.data_ptr(), out.data_ptr(), 1.0, 0.0,
                                                     self.s.data_ptr(), self._stream()))  # s[1] = <v, out>

    def start(self, x0, max_iter: int):
        torch = _torch()
        self.x = self._to_dev(x0).reshape(-1).clone()
        self.s = torch.zeros(_capi.CG_NSCALARS + max_iter + 2, dtype=torch.float64, device=self.dev)
        self.b = self.model.adjoint(self.y).reshape(-1)  # all-reduced over shards when the model has a comm
        if self.mu_s != 1.0:
            self.b.mul_(self.mu_s)
        self.q = torch.empty_like(self.b)
        self.r = torch.empty_like(self.b)
        self.d = torch.empty_like(self.b)
        self.hx = self.hd = None
        self._sharded_track, self._blocks = False, None
        partial = getattr(self.model, "partial", False)
        if self.track_hx and not partial:
            self.hd = torch.empty(self.model.osize, dtype=self.tdtype, device=self.dev)
        self.hessp(self.x, self.q)
        if self.track_hx and partial and self.model.lambda_range is not None and self.comm is not None:
            # wavelength shards: after the band exchange the model's detector vector holds the complete H v on
            # the blocks of the bands this rank touches; the rank owning a band tracks H x_k on that block
            self._sharded_track = True
            self.hd = self.model._y_shard
            self._blocks = self.model._band_exchange().owned_blocks()
        if self.hd is not None:
            self.hx = self.hd.clone()                     # H x_0
        self._check(self.lib.surfh_cg_start(self.h, self.b.data_ptr(), self.q.data_ptr(), self.r.data_ptr(),
                                            self.d.data_ptr(), self.s.data_ptr(), self._stream()))
        if self.precond is not None:                      # d = z = P r ; rho_z = <r, z>
            self.z = torch.empty_like(self.b)
            self.precond.apply(self.r, self.z)
            self._check(self.lib.surfh_pcg_direction(self.h, self.r.data_ptr(), self.z.data_ptr(), self.d.data_ptr(),
                                                     self.s.data_ptr(), 1, self._stream()))
        self.iteration = 0

    def _advance_hx(self):
        if self.hx is None:
            return
        item = self.hx.element_size()
        blocks = self._blocks if self._sharded_track else [(0, self.hx.numel())]
        for off, size in blocks:
            self._check(self.lib.surfh_axpy_device_scalar(self.h, self.hx.data_ptr() + off * item,
                                                          self.hd.data_ptr() + off * item, size, self.s.data_ptr(), 2,
                                                          self._stream()))

    def _pcg_step(self, refresh: bool):
        """One preconditioned iteration: alpha = <r,z>/<d,Qd>; x, r updates; z = P r; beta = <r,z>'/<r,z>;
        d = z + beta d.  <r,r> still feeds the gradient-norm history and the stopping rule."""
        lib, st = self.lib, self._stream()
        self.hessp(self.d, self.q)
        if refresh:
            self._check(lib.surfh_cg_refresh(self.h, 0, self.x.data_ptr(), None, self.d.data_ptr(), None, None,
                                             self.s.data_ptr(), st))           # x += alpha d
            self.hessp(self.x, self.q)
            if self.hx is not None:
                self.hx.copy_(self.hd)
            self._check(lib.surfh_pcg_update(self.h, 1, None, self.r.data_ptr(), None, self.q.data_ptr(),
                                             self.b.data_ptr(), self.s.data_ptr(), st))   # r = b - Q x
        else:
            self._check(lib.surfh_pcg_update(self.h, 0, self.x.data_ptr(), self.r.data_ptr(), self.d.data_ptr(),
                                             self.q.data_ptr(), None, self.s.data_ptr(), st))
            self._advance_hx()
        self.precond.apply(self.r, self.z)
        self._check(lib.surfh_pcg_direction(self.h, self.r.data_ptr(), self.z.data_ptr(), self.d.data_ptr(),
                                            self.s.data_ptr(), 0, st))
        self.iteration += 1

    def step(self, refresh: bool):
        if self.precond is not None:
            return self._pcg_step(refresh)
        self.hessp(self.d, self.q)
        if refresh:
            self._check(self.lib.surfh_cg_refresh(self.h, 0, self.x.data_ptr(), None, self.d.data_ptr(), None, None,
                                                  self.s.data_ptr(), self._stream()))
            self.hessp(self.x, self.q)
            if self.hx is not None:
                self.hx.copy_(self.hd)                    # exact H x_{k+1}, like the residual
            self._check(self.lib.surfh_cg_refresh(self.h, 1, None, self.r.data_ptr(), self.d.data_ptr(),
                                                  self.b.data_ptr(), self.q.data_ptr(), self.s.data_ptr(),
                                                  self._stream()))
        else:
            self._check(self.lib.surfh_cg_update(self.h, self.x.data_ptr(), self.r.data_ptr(), self.d.data_ptr(),
                                                 self.q.data_ptr(), self.s.data_ptr(), self._stream()))
            self._advance_hx()                            # H x_{k+1} = H x_k + alpha H d   (s[2] = alpha)
        self.iteration += 1

    def grad_norm_history(self):
        return self.s[_capi.CG_NSCALARS: _capi.CG_NSCALARS + self.iteration + 1].cpu().numpy()

    def is_current_iterate(self, x) -> bool:
        """True when `x` is this solver's own iterate (the tensor the lcg callback hands out as `res.x`,
        or a reshaped view of it), for which r = b - Q x is known."""
        if self.x is None or self.r is None or not hasattr(x, "data_ptr"):
            return False
        return x.data_ptr() == self.x.data_ptr() and x.numel() == self.x.numel()

    def state_criterion_available(self) -> bool:
        """The criterion of the current iterate can be formed without applying H when H x_k is tracked
        (`track_hx`), or -- exact adjoint only, where Q is symmetric and b = H^T y -- from the residual."""
        return self.hx is not None or (self.r is not None and self.model.adjoint_mode == "exact")

    def criterion_from_state(self) -> float:
        """J(x_k) of the current iterate without applying H.
          * H x_k tracked: J = (mu_s |y - H x_k|^2 + mu_r prior(x_k)) / 2, one fused reduction kernel -- valid
            for both adjoint flavours (H x_k is refreshed exactly with the residual, every REFRESH_PERIOD);
          * otherwise, exact adjoint: J(x) = 1/2 x^T Q x - b^T x + mu_s |y|^2 / 2 and Q x = b - r give
            J = mu_s |y|^2 / 2 - <x, b + r> / 2 (one dot-product kernel).  Not valid with the reference's
            `gridding_t` "adjoint", whose Q is not H^T H.
        Two doubles cross PCIe."""
        torch = _torch()
        if not self.state_criterion_available():
            raise ValueError("criterion_from_state needs track_hx or an exact-adjoint model")
        self._state_evals += 1
        if self.hx is not None and self._sharded_track:
            return self._criterion_sharded()
        if self.hx is not None:
            return self._criterion_given_hx(self.x, self.hx)
        if self._y_sq is None:
            self._y_sq = float(torch.dot(self.y.double(), self.y.double()))
        out = torch.empty(1, dtype=torch.float64, device=self.dev)
        self._check(self.lib.surfh_cg_dot_x_b_plus_r(self.h, self.x.data_ptr(), self.b.data_ptr(), self.r.data_ptr(),
                                                     out.data_ptr(), self._stream()))
        return 0.5 * self.mu_s * self._y_sq - 0.5 * float(out.item())

    def _criterion_sharded(self) -> float:
        """Data term summed over the detector blocks this rank owns, all-reduced (one double); the prior is
        evaluated redundantly on the identical maps."""
        torch = _torch()
        item = self.hx.element_size()
        data = torch.zeros(1, dtype=torch.float64, device=self.dev)
        out = torch.zeros(2, dtype=torch.float64, device=self.dev)
        for off, size in self._blocks:
            self._check(self.lib.surfh_criterion_terms(self.h, self.y.data_ptr() + off * item, self.hx.data_ptr() + off * item,
                                                       size, None, out.data_ptr(), self._stream()))
            data += out[:1]
        self.comm.allreduce_sum(data)
        self._check(self.lib.surfh_criterion_terms(self.h, None, None, 0, self.x.data_ptr(), out.data_ptr(), self._stream()))
        prior = out[1]
        if self.gradient == "joint":
            lx = self.laplacian(self.x, torch.empty_like(self.x))
            prior = torch.dot(lx.double(), lx.double())
        return float((self.mu_s * data[0] + self.mu_r * prior).item() / 2)

    def _criterion_given_hx(self, x, hx) -> float:
        torch = _torch()
        out = torch.zeros(2, dtype=torch.float64, device=self.dev)
        self._check(self.lib.surfh_criterion_terms(self.h, self.y.data_ptr(), hx.data_ptr(), int(hx.numel()),
                                                   x.data_ptr(), out.data_ptr(), self._stream()))
        if self.gradient == "joint":  # ||L x||^2 instead of ||D_r x||^2 + ||D_c x||^2
            lx = self.laplacian(x, torch.empty_like(x))
            out[1] = torch.dot(lx.double(), lx.double())
        t = out.cpu().numpy()
        return float((self.mu_s * t[0] + self.mu_r * t[1]) / 2)

    def criterion(self, x) -> float:
        """J(x), everything reduced on the device; the only host traffic is two doubles.  Applies H once
        (the reference's get_crit_val, fusion_CT.py:242-265); `criterion_from_state` avoids that for the
        solver's own iterate."""
        torch = _torch()
        x = self._to_dev(x).reshape(-1)
        hx = self.model.forward(x.reshape(self.model.ishape))
        # forward() returns the complete detector vector on every rank (all-reduced when sharded),
        # so the criterion is evaluated redundantly and needs no scalar collective
        if self.comm is None and getattr(self.model, "partial", False):
            raise ValueError("a sharded model needs a comm to evaluate the criterion")
        return self._criterion_given_hx(x, hx.reshape(-1))


def lcg(model, y, mu_spectro=1.0, mu_reg=1.0, x0=None, tol=1e-4, max_iter=500, min_iter=0,
        callback: Optional[Callable] = None, refresh: int = REFRESH_PERIOD, check_every: int = 1, comm=None,
        numpy_result: bool = True, gradient: str = "separated", solver: Optional[DeviceCG] = None,
        precond=None) -> OptimizeResult:
    """Linear CG on the device.  `res.x` is a flat device tensor while iterating (callbacks may
    `.reshape` it) and, at return, a numpy array of the model's input shape (`numpy_result`).

    The stopping rule is qmm.lcg's, tested after every iteration (`check_every=1`: one double read back
    per iteration).  `check_every=n` tests it every n-th iteration only -- the loop then runs up to n-1
    iterations past the tolerance but the host never waits on the device in between (opt-in).

    `precond`: None (qmm.lcg's default), a `FourierPreconditioner`, or True to build one for these
    hyper-parameters -- preconditioned CG; the iterates then differ from plain CG's (same minimiser)."""
    torch = _torch()
    cg = solver if solver is not None else DeviceCG(model, y, mu_spectro, mu_reg, comm=comm, gradient=gradient)
    if precond is True:
        precond = FourierPreconditioner(model, mu_spectro, mu_reg, gradient)
    cg.precond = precond
    if x0 is None:
        x0 = np.zeros(model.ishape)
    cg.start(x0, max_iter)
    res = OptimizeResult(x=cg.x, success=True, status=99, nit=max_iter, grad_norm=[], time=[time.time()],
                         message="maximum number of iterations reached")
    size_tol = cg.n * tol
    check_every = max(1, int(check_every))
    for iteration in range(max_iter):
        cg.step(refresh=bool(refresh) and iteration % refresh == 0)
        last = iteration == max_iter - 1
        if callback is not None or last or (iteration + 1) % check_every == 0:
            res["grad_norm"] = list(cg.grad_norm_history())
            res["time"].append(time.time())
            if callback is not None:
                callback(res)
            if np.sqrt(res["grad_norm"][-1]) < size_tol and iteration >= min_iter:
                res["status"], res["nit"] = 1, iteration + 1
                res["message"] = "gradient norm below tolerance"
                break
    torch.cuda.current_stream().synchronize()
    res["grad_norm"] = list(cg.grad_norm_history())
    res["time"].append(time.time())
    res["x"] = cg.x.reshape(model.ishape).cpu().numpy().astype(np.float64) if numpy_result \
        else cg.x.reshape(model.ishape)
    res["solver"] = cg
    return res


def mmmg(model, y, mu_spectro=1.0, mu_reg=1.0, x0=None, tol=1e-4, max_iter=500, min_iter=0,
         callback: Optional[Callable] = None, comm=None, numpy_result: bool = True,
         gradient: str = "separated") -> OptimizeResult:
    """qmm.mmmg (3MG: majorize-minimize memory gradient) for the quadratic objectives of
    `QuadCriterion_MRS.run_method(method != 'lcg')` (fusion_CT.py:194-197), on the device.

    Every iteration minimises J exactly over the plane spanned by -grad and the previous move (for
    quadratic objectives the quadratic majorant is J itself):
        g = Q x - b;  D = [-g, m];  M = sum_i hyper_i (V_i D)^T (V_i D) (2x2);  step = -lstsq(M, D^T g);
        m = D step;  x += m;  V_i m is carried by linearity (V_i D) step, as qmm does.
    One H^T H application (gradient) and one H application (H g) per iteration; the 2x2 system is solved
    on the host (five scalars cross PCIe per iteration).  Stops when |g|^2 < x.size * tol (qmm's rule).
    """
    torch = _torch()
    cg = DeviceCG(model, y, mu_spectro, mu_reg, comm=comm, gradient=gradient)
    if x0 is None:
        x0 = np.zeros(model.ishape)
    cg.start(x0, 1)                       # b = mu_s H^T y; x, q = Q x, r = b - Q x
    x, b = cg.x, cg.b
    n = x.numel()
    move = torch.zeros_like(x)
    h_move = torch.zeros(model.osize, dtype=x.dtype, device=x.device)   # H m
    q = torch.empty_like(x)
    lap_g, lap_m = torch.empty_like(x), torch.empty_like(x)
    res = OptimizeResult(x=x, success=False, status=99, nit=max_iter, grad_norm=[], time=[time.time()],
                         message="maximum number of iterations reached")

    def reg_gram(u, lu, v, lv):
        """<u, R v> with R = L (separated) or L L (joint), given lu = L u and lv = L v."""
        return torch.dot(u, lv) if gradient == "separated" else torch.dot(lu, lv)

    for iteration in range(max_iter):
        cg.hessp(x, q)                                   # q = Q x
        g = q.sub_(b)                                    # gradient, in place
        gn = float(torch.dot(g, g))
        res["grad_norm"].append(gn)
        if gn < n * tol and iteration >= min_iter:
            res["success"], res["status"], res["nit"] = True, 1, iteration
            res["message"] = "gradient norm below tolerance"
            break
        h_g = model.forward(g.reshape(model.ishape))     # H g  (the direction is -g: signs handled below)
        cg.laplacian(g, lap_g)
        cg.laplacian(move, lap_m)
        # M = mu_s [H d_i . H d_j] + mu_r [d_i . R d_j]  for d_0 = -g, d_1 = m
        m00 = cg.mu_s * torch.dot(h_g, h_g) + cg.mu_r * reg_gram(g, lap_g, g, lap_g)
        m01 = -(cg.mu_s * torch.dot(h_g, h_move) + cg.mu_r * reg_gram(g, lap_g, move, lap_m))
        m11 = cg.mu_s * torch.dot(h_move, h_move) + cg.mu_r * reg_gram(move, lap_m, move, lap_m)
        rhs = torch.stack([-torch.dot(g, g), torch.dot(move, g)])
        vals = torch.stack([m00, m01, m11]).double().cpu().numpy()
        rhs = rhs.double().cpu().numpy()
        mat = np.array([[vals[0], vals[1]], [vals[1], vals[2]]])
        step = -np.linalg.lstsq(mat, rhs, rcond=None)[0]
        # m = D step ; H m = (H D) step
        move.mul_(float(step[1])).add_(g, alpha=-float(step[0]))
        h_move.mul_(float(step[1])).add_(h_g, alpha=-float(step[0]))
        x.add_(move)
        res["time"].append(time.time())
        if callback is not None:
            callback(res)
    torch.cuda.current_stream().synchronize()
    res["x"] = x.reshape(model.ishape).cpu().numpy().astype(np.float64) if numpy_result else x.reshape(model.ishape)
    res["solver"] = cg
    return res


def huber_value(u, delta: float):
    """qmm.Huber: u^2 / 2 inside [-delta, delta], delta |u| - delta^2 / 2 outside (torch tensors)."""
    a = u.abs()
    return _torch().where(a <= delta, 0.5 * u * u, delta * a - 0.5 * delta * delta)


def mmmg_huber(model, y, mu_spectro=1.0, spat_reg=1.0, spat_th=1.0, x0=None, tol=1e-4, max_iter=500, min_iter=0,
               callback: Optional[Callable] = None, numpy_result: bool = True) -> OptimizeResult:
    """`lmm_reconstruction` (surfh/ToolsDir/algorithms.py:71-106) on the device: quadratic data term +
    Huber priors (threshold `spat_th`, weight `spat_reg`) on the row and column differences of every map,
    minimised by qmm.mmmg (3MG): each iteration minimises, over the plane spanned by -grad and the previous
    move, the quadratic majorant whose curvature for the priors is diag(min(1, delta / |D x|)).

        J(x) = mu_s / 2 |y - H x|^2 + spat_reg * sum_{D in (D_r, D_c)} sum huber_delta(D x)

    One H^T H application and one H application per iteration (the operator work, on the CUDA library); the
    prior terms are elementwise / circular-shift tensor operations on the K maps.  D_r, D_c are the circular
    differences of fusion_CT.py:16-43 (the reference routine's aljabr.Diff is not in the reference tree);
    qmm itself is restated: parity unpinned."""
    torch = _torch()
    cg = DeviceCG(model, y, mu_spectro, 0.0)
    if x0 is None:
        x0 = np.zeros(model.ishape)
    cg.start(x0, 1)                          # b = mu_s H^T y
    x, b = cg.x, cg.b
    shape = model.ishape
    n = x.numel()
    delta, lam = float(spat_th), float(spat_reg)
    d_r = lambda v: torch.roll(v, 1, dims=1) - v      # noqa: E731  NpDiff_r.forward
    d_c = lambda v: torch.roll(v, 1, dims=2) - v      # noqa: E731  NpDiff_c.forward
    d_r_t = lambda v: torch.roll(v, -1, dims=1) - v   # noqa: E731
    d_c_t = lambda v: torch.roll(v, -1, dims=2) - v   # noqa: E731
    move = torch.zeros_like(x)
    h_move = torch.zeros(model.osize, dtype=x.dtype, device=x.device)
    q = torch.empty_like(x)
    res = OptimizeResult(x=x, success=False, status=99, nit=max_iter, grad_norm=[], time=[time.time()],
                         message="maximum number of iterations reached")
    for iteration in range(max_iter):
        xm = x.reshape(shape)
        ur, uc = d_r(xm), d_c(xm)
        cg.hessp(x, q)                                   # mu_s H^T H x  (mu_reg = 0 in this solver object)
        g = q.sub_(b)
        g.add_((d_r_t(ur.clamp(-delta, delta)) + d_c_t(uc.clamp(-delta, delta))).reshape(-1), alpha=lam)
        gn = float(torch.dot(g, g))
        res["grad_norm"].append(gn)
        if gn < n * tol and iteration >= min_iter:
            res["success"], res["status"], res["nit"] = True, 1, iteration
            res["message"] = "gradient norm below tolerance"
            break
        h_g = model.forward(g.reshape(shape))            # H g; the direction is -g
        gm, mm = g.reshape(shape), move.reshape(shape)
        m00 = cg.mu_s * torch.dot(h_g, h_g)
        m01 = -cg.mu_s * torch.dot(h_g, h_move)
        m11 = cg.mu_s * torch.dot(h_move, h_move)
        for u, dg, dm in ((ur, d_r(gm), d_r(mm)), (uc, d_c(gm), d_c(mm))):
            w = torch.clamp(delta / u.abs().clamp_min(1e-300), max=1.0)       # gr_coeffs
            m00 = m00 + lam * (w * dg * dg).sum()
            m01 = m01 - lam * (w * dg * dm).sum()
            m11 = m11 + lam * (w * dm * dm).sum()
        rhs = torch.stack([-torch.dot(g, g), torch.dot(move, g)]).double().cpu().numpy()
        vals = torch.stack([m00, m01, m11]).double().cpu().numpy()
        mat = np.array([[vals[0], vals[1]], [vals[1], vals[2]]])
        step = -np.linalg.lstsq(mat, rhs, rcond=None)[0]
        move.mul_(float(step[1])).add_(g, alpha=-float(step[0]))
        h_move.mul_(float(step[1])).add_(h_g, alpha=-float(step[0]))
        x.add_(move)
        res["time"].append(time.time())
        if callback is not None:
            callback(res)
    torch.cuda.current_stream().synchronize()
    res["x"] = x.reshape(shape).cpu().numpy().astype(np.float64) if numpy_result else x.reshape(shape)
    res["solver"] = cg
    return res


def criterion_huber(model, y, x, mu_spectro, spat_reg, spat_th) -> float:
    """J(x) of `mmmg_huber` (one forward pass)."""
    torch = _torch()
    cg = DeviceCG(model, y, mu_spectro, 0.0)
    xd = cg._to_dev(x).reshape(model.ishape)
    hx = model.forward(xd).reshape(-1)
    data = 0.5 * float(mu_spectro) * float(torch.sum((cg.y - hx) ** 2))
    prior = float(torch.sum(huber_value(torch.roll(xd, 1, dims=1) - xd, float(spat_th))
                            + huber_value(torch.roll(xd, 1, dims=2) - xd, float(spat_th))))
    return data + float(spat_reg) * prior


class QuadCriterion_MRS:
    """Same constructor and `run_method` as the reference class, run on the device: gradient =
    'separated' (NpDiff_r / NpDiff_c) or 'joint' (Difference_Operator_Joint), method = 'lcg' or anything
    else -> mmmg, exactly like the reference's dispatch (fusion_CT.py:194-197)."""

    def __init__(self, mu_spectro, y_spectro, model_spectro, mu_reg, printing=False, gradient="separated",
                 comm=None):
        self.mu_spectro = mu_spectro
        self.y_spectro = y_spectro
        self.model_spectro = model_spectro
        self.n_spec = model_spectro.ishape[0]
        self.it = 1
        assert isinstance(mu_reg, (float, int, list, np.ndarray))
        self.mu_reg = mu_reg
        if isinstance(mu_reg, (list, np.ndarray)):
            assert len(mu_reg) == self.n_spec
        shape_target = model_spectro.ishape[1:]
        self.shape_of_output = (self.n_spec, shape_target[0], shape_target[1])
        if gradient not in ("separated", "joint"):
            raise ValueError("gradient must be 'separated' or 'joint'")
        if gradient == "separated":
            self.npdiff_r = NpDiff_r(self.shape_of_output)
            self.npdiff_c = NpDiff_c(self.shape_of_output)
        self.L_mu = np.copy(mu_reg) if isinstance(mu_reg, (list, np.ndarray)) else np.ones(self.n_spec) * mu_reg
        self.printing = printing
        self.gradient = gradient
        self.comm = comm
        self.L_crit_val = []
        self._cg: Optional[DeviceCG] = None
        self.criterion_from_state = True  # False: always evaluate J through a forward pass, like the reference

    def _solver(self) -> DeviceCG:
        if self._cg is None:
            self._cg = DeviceCG(self.model_spectro, self.y_spectro, self.mu_spectro, self.mu_reg, comm=self.comm,
                                gradient=self.gradient)
        return self._cg

    def run_method(self, method="lcg", maximum_iterations=10, tolerance=1e-12, calc_crit=False, perf_crit=None,
                   value_init=0.5):
        assert isinstance(self.mu_reg, (int, float))
        if isinstance(value_init, (float, int)):
            init = np.ones(self.shape_of_output) * value_init
        else:
            assert tuple(value_init.shape) == self.shape_of_output
            init = value_init
        self.L_crit_val = []

        def perf_crit_with_reshape(res):
            crit_val = self.get_crit_val(res.x.reshape(self.shape_of_output))
            self.L_crit_val.append(crit_val)
            if self.printing:
                print(f"Criterion value = {crit_val}\n")

        def print_last_grad_norm(res):
            if self.printing:
                print(f"Iteration n°{self.it}, Grad norm = {res.grad_norm[-1]}")
            self.it = self.it + 1

        def print_last_grad_norm_and_crit(res):
            print_last_grad_norm(res)
            if self.it % 5 == 2:
                perf_crit_with_reshape(res)

        if calc_crit and perf_crit is None:
            callback = lambda res: self.L_crit_val.append(self.get_crit_val(res.x.reshape(self.shape_of_output)))  # noqa: E731
        elif not calc_crit and perf_crit is not None:
            callback = print_last_grad_norm
        elif calc_crit and perf_crit is not None:
            callback = print_last_grad_norm_and_crit
        else:
            callback = None
        t1 = time.time()
        if method == "lcg":
            # the solver is shared with get_crit_val: the criterion of the current iterate then comes from
            # the CG state (no extra forward pass per evaluation)
            self._solver().track_hx = bool(calc_crit) and self.criterion_from_state
            res = lcg(self.model_spectro, self.y_spectro, self.mu_spectro, self.mu_reg, init, tol=tolerance,
                      max_iter=maximum_iterations, callback=callback, comm=self.comm, gradient=self.gradient,
                      solver=self._solver())
        else:
            res = mmmg(self.model_spectro, self.y_spectro, self.mu_spectro, self.mu_reg, init, tol=tolerance,
                       max_iter=maximum_iterations, callback=callback, comm=self.comm, gradient=self.gradient)
        if self.printing:
            print(f"Total time needed for {method} :", round(time.time() - t1, 3))
        return res

    def get_crit_val(self, x_hat) -> float:
        """J(x_hat) (fusion_CT.py:242-265).  When x_hat is the running lcg iterate (what the callbacks of
        `run_method` pass) the value comes from the CG state; any other argument applies H once."""
        cg = self._solver()
        if self.criterion_from_state and cg.is_current_iterate(x_hat) and cg.state_criterion_available():
            return cg.criterion_from_state()
        return cg.criterion(x_hat)

"""Operator protocol of the reference's `aljabr.LinOp` (third-party, not vendored), as far as
the hot path and its callers use it: `forward`/`adjoint` on shaped arrays, `matvec`/`rmatvec` on
flat ones, `fwadj`, shapes and sizes; and `dottest` (call sites: surfh/Models/spectroModel.py:39,116;
test/test_fw_ad.py:608; semantics as in test/sandbox_dottest.py:16-27)."""
from __future__ import annotations

import numpy as np


class LinOp:
    def __init__(self, ishape, oshape, name: str = "_", dtype=np.float64):
        self.ishape = tuple(int(v) for v in ishape)
        self.oshape = tuple(int(v) for v in oshape)
        self.name = name
        self.dtype = dtype

    @property
    def isize(self) -> int:
        return int(np.prod(self.ishape))

    @property
    def osize(self) -> int:
        return int(np.prod(self.oshape))

    @property
    def shape(self):
        return (self.osize, self.isize)

    def forward(self, point):
        raise NotImplementedError

    def adjoint(self, point):
        raise NotImplementedError

    def matvec(self, point):
        return self.forward(_reshape(point, self.ishape)).reshape(-1)

    def rmatvec(self, point):
        return self.adjoint(_reshape(point, self.oshape)).reshape(-1)

    def fwadj(self, point):
        return self.adjoint(self.forward(point))

    fwback = fwadj

    def __call__(self, point):
        return self.forward(point)


def _reshape(point, shape):
    return point.reshape(shape)


def dottest(linop: LinOp, num: int = 1, rtol: float = 1e-5, atol: float = 1e-8, echo: bool = False,
            seed=None) -> bool:
    """True when <A^T v, u> == <v, A u> within tolerance for `num` draws of randn vectors."""
    rng = np.random.default_rng(seed)
    ok = True
    for _ in range(num):
        u = rng.standard_normal(linop.isize)
        v = rng.standard_normal(linop.osize)
        left = float(np.vdot(np.asarray(linop.rmatvec(v)), u))
        right = float(np.vdot(v, np.asarray(linop.matvec(u))))
        if echo:
            print(f"(A^T v)^T u = {left} ~= {right} = v^T (A u)")
        ok = ok and bool(np.allclose(left, right, rtol=rtol, atol=atol))
    return ok

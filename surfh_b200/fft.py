"""Batched 2-D real FFT pair on the GPU through the library's hand-written chirp-z kernels.

`dft` / `idft` mirror `surfh.ToolsDir.jax_utils.dft / idft` (jax_utils.py:30-46; numpy twins
python_utils.py:41-71): `rfftn` / `irfftn` over the last two axes with norm="ortho".  They take and
return torch CUDA tensors; there is no CPU path."""
from __future__ import annotations

import math

from . import _capi


def _check(t, complex_ok: bool):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("surfh_b200.fft works on torch CUDA tensors only (no CPU fallback)")
    if complex_ok:
        if t.dtype not in (torch.complex64, torch.complex128):
            raise TypeError("expected a complex64 / complex128 tensor")
    elif t.dtype not in (torch.float32, torch.float64):
        raise TypeError("expected a float32 / float64 tensor")


def rfft2(x):
    """Un-normalised rfft2 over the last two axes (== numpy.fft.rfft2)."""
    import torch
    _check(x, False)
    x = x.contiguous()
    na, nb = x.shape[-2:]
    batch = x.numel() // (na * nb)
    f64 = x.dtype == torch.float64
    out = torch.empty(x.shape[:-1] + (nb // 2 + 1,), dtype=torch.complex128 if f64 else torch.complex64, device=x.device)
    lib = _capi.load()
    with torch.cuda.device(x.device):
        code = lib.surfh_rfft2(_capi.F64 if f64 else _capi.F32, na, nb, batch, 0, x.data_ptr(), out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _capi.check(None, code)
    return out


def irfft2(xf, shape):
    """Un-normalised inverse (== numpy.fft.irfft2(xf, shape) * shape[0] * shape[1])."""
    import torch
    _check(xf, True)
    xf = xf.contiguous()
    na, nb = int(shape[0]), int(shape[1])
    if xf.shape[-2] != na or xf.shape[-1] != nb // 2 + 1:
        raise ValueError(f"spectrum of shape {tuple(xf.shape[-2:])} does not match image shape {(na, nb)}")
    batch = xf.numel() // (na * (nb // 2 + 1))
    f64 = xf.dtype == torch.complex128
    out = torch.empty(xf.shape[:-2] + (na, nb), dtype=torch.float64 if f64 else torch.float32, device=xf.device)
    lib = _capi.load()
    with torch.cuda.device(xf.device):
        code = lib.surfh_rfft2(_capi.F64 if f64 else _capi.F32, na, nb, batch, 1, xf.data_ptr(), out.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    _capi.check(None, code)
    return out


def dft(inarray):
    """jax_utils.dft: rfftn over the last two axes, norm="ortho"."""
    na, nb = inarray.shape[-2:]
    return rfft2(inarray) * (1.0 / math.sqrt(na * nb))


def idft(inarray, im_shape):
    """jax_utils.idft: irfftn over the last two axes to `im_shape`, norm="ortho"."""
    na, nb = int(im_shape[0]), int(im_shape[1])
    return irfft2(inarray, (na, nb)) * (1.0 / math.sqrt(na * nb))

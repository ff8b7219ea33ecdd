"""The hand-written 2-D real FFT pair (csrc/kernels_fft.cuh) against numpy's pocketfft in fp64:
the arithmetic of jax_utils.dft / idft (surfh/ToolsDir/jax_utils.py:30-46).  Tolerances: relative
L2 <= 1e-13 in fp64 (well inside the 1e-10 operator budget), <= 2e-6 in fp32 (budget 1e-5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(251, 251), (501, 501), (301, 301), (2, 2), (3, 5), (16, 9), (64, 64), (128, 128), (129, 127), (256, 256),
          (257, 255), (300, 512), (512, 257), (5, 500), (512, 512), (40, 96), (96, 96)]
TOL = {"float64": 1e-13, "float32": 2e-6}


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


@pytest.fixture(scope="module")
def fft():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200 import fft as f
    return f


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_rfft2_and_irfft2_match_numpy(fft, shape, dtype):
    import torch
    rng = np.random.default_rng(shape[0] * 4099 + shape[1])
    batch = 7 if shape[0] * shape[1] < 200000 else 3
    x = rng.standard_normal((batch,) + shape)
    want = np.fft.rfft2(x)
    xt = torch.as_tensor(x, device="cuda", dtype=getattr(torch, dtype))
    got = fft.rfft2(xt).cpu().numpy()
    assert got.shape == want.shape
    assert rel(got, want) <= TOL[dtype]
    # inverse from an arbitrary half-spectrum (numpy ignores the imaginary part of the DC / Nyquist bins)
    s = rng.standard_normal(want.shape) + 1j * rng.standard_normal(want.shape)
    back = np.fft.irfft2(s, shape) * (shape[0] * shape[1])
    st = torch.as_tensor(s, device="cuda", dtype=torch.complex128 if dtype == "float64" else torch.complex64)
    got_back = fft.irfft2(st, shape).cpu().numpy()
    assert got_back.shape == back.shape
    assert rel(got_back, back) <= TOL[dtype]


def test_ortho_pair_round_trip_and_impulse(fft):
    import torch
    x = torch.zeros((2, 251, 251), dtype=torch.float64, device="cuda")
    x[0, 0, 0] = 1.0
    x[1, 17, 133] = -2.5
    f = fft.dft(x)
    assert torch.allclose(f[0], torch.full_like(f[0], 1.0 / 251))
    assert float((f[1].abs() - 2.5 / 251).abs().max()) < 1e-15
    back = fft.idft(f, (251, 251))
    assert float((back - x).abs().max()) < 1e-14


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_bitwise_reproducible_under_load(fft, dtype):
    """compute-sanitizer is not available on the pool: a race in the shared-memory exchanges, the named
    barriers or the cp.async staging would show up as run-to-run differences with every SM busy."""
    import torch
    x = torch.randn((300, 251, 251), dtype=getattr(torch, dtype), device="cuda")
    f0 = fft.rfft2(x)
    b0 = fft.irfft2(f0, (251, 251))
    for _ in range(10):
        assert torch.equal(fft.rfft2(x), f0)
        assert torch.equal(fft.irfft2(f0, (251, 251)), b0)


def test_rejects_cpu_tensors_and_long_axes(fft):
    import torch
    from surfh_b200 import _capi
    with pytest.raises(TypeError):
        fft.rfft2(torch.zeros((4, 4), dtype=torch.float64))
    with pytest.raises(_capi.SurfhError):
        fft.rfft2(torch.zeros((4, 513), dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-5)])
def test_operator_backends_agree(dtype, tol):
    """The operator gives the same result through the hand-written FFT and through cuFFT."""
    from cases import CASES
    from surfh_b200.model import spectroSigRLSCT
    cfg = CASES["mini_2band_4p"]()
    own = spectroSigRLSCT(**cfg.model_args(), dtype=dtype, fft_backend="own")
    lib = spectroSigRLSCT(**cfg.model_args(), dtype=dtype, fft_backend="cufft")
    y0, y1 = own.forward(cfg.maps), lib.forward(cfg.maps)
    assert rel(y0, y1) <= tol
    v = np.random.default_rng(3).standard_normal(own.osize)
    assert rel(own.adjoint(v), lib.adjoint(v)) <= tol

"""CUDA path vs the CPU oracle and vs the golden vectors made by the reference's own code.
Tolerances are BASELINE.json's: relative L2 <= 1e-10 in fp64, <= 1e-5 in fp32; dot-test 1e-6."""
import os

import numpy as np
import pytest

from cases import CASES, FULL, MINI

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-10, "float32": 1e-5}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))


@pytest.fixture(scope="module")
def built():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200.model import spectroSigRLSCT
    return spectroSigRLSCT


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MINI)
def test_mini_vs_oracle_and_golden(built, golden_dir, name, dtype):
    from surfh_oracle import model as om
    cfg = CASES[name]()
    args = cfg.model_args()
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    v = np.random.default_rng(1234).standard_normal(int(gold["idx"][-1]))
    for mode in ("reference", "exact"):
        gpu = built(**args, dtype=dtype, adjoint_mode=mode)
        cpu = om.SpectroLMM(**args, adjoint_mode=mode)
        assert gpu.ishape == cpu.ishape and gpu.oshape == cpu.oshape
        assert np.array_equal(gpu._idx, gold["idx"])
        y = gpu.forward(cfg.maps)
        assert y.shape == cpu.oshape and y.dtype == np.float64
        assert rel(y, cpu.forward(cfg.maps)) <= TOL[dtype]
        x = gpu.adjoint(v)
        assert x.shape == cpu.ishape
        assert rel(x, cpu.adjoint(v)) <= TOL[dtype]
        if mode == "reference":  # the reference's own outputs
            assert rel(y, gold["fwd"]) <= TOL[dtype]
            assert rel(x, gold["adj"]) <= TOL[dtype]


@pytest.mark.parametrize("name", FULL)
def test_full_size_vs_reference_golden(built, golden_dir, name):
    cfg = CASES[name]()
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    gpu = built(**cfg.model_args(), dtype="float64", adjoint_mode="reference")
    y = gpu.forward(cfg.maps)
    stride = int(gold["fwd_stride"])
    assert rel(y[::stride], gold["fwd_sample"]) <= 1e-10
    assert abs(np.linalg.norm(y) - float(gold["fwd_norm"])) <= 1e-10 * float(gold["fwd_norm"])
    v = np.random.default_rng(1234).standard_normal(gpu.osize)
    x = gpu.adjoint(v)
    assert rel(x[:, ::7, ::7], gold["adj_sample"]) <= 1e-10
    assert abs(np.linalg.norm(x) - float(gold["adj_norm"])) <= 1e-10 * float(gold["adj_norm"])


def test_c3_vs_reference_golden(built, golden_dir):
    """BASELINE.json configs[2] (C3: bands 1A, 2A, 3A, 4A, K = 4, N = 501, 3612 cube wavelengths, 4 dithers)
    against the vector the reference's own code produced for the same inputs (oracle/make_golden.py)."""
    import torch
    from surfh_b200 import synthetic
    cfg = CASES["c3"]()
    gold = np.load(os.path.join(golden_dir, "c3.npz"))
    dev = torch.device("cuda")
    sotf = lambda lo, hi: synthetic.ir2fr_device(cfg.psf[lo:hi], cfg.imshape, dev, torch.float64)  # noqa: E731
    gpu = built(sotf, cfg.templates, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, cfg.instrs, cfg.step_degree,
                cfg.pointings, dtype="float64", adjoint_mode="reference")
    assert np.array_equal(gpu._idx, gold["idx"]) and gpu.ishape == tuple(gold["ishape"])
    y = gpu.forward(cfg.maps)
    stride = int(gold["fwd_stride"])
    assert rel(y[::stride], gold["fwd_sample"]) <= 1e-10
    assert abs(np.linalg.norm(y) - float(gold["fwd_norm"])) <= 1e-10 * float(gold["fwd_norm"])
    # every band's block on its own (a band-local error cannot hide behind the others' norm)
    for c in range(4):
        lo, hi = int(gold["idx"][c]), int(gold["idx"][c + 1])
        first = -(-lo // stride) * stride
        assert rel(y[first:hi:stride], gold["fwd_sample"][first // stride: (hi - 1) // stride + 1]) <= 1e-10
    v = np.random.default_rng(1234).standard_normal(gpu.osize)
    x = gpu.adjoint(v)
    assert rel(x[:, ::7, ::7], gold["adj_sample"]) <= 1e-10
    assert abs(np.linalg.norm(x) - float(gold["adj_norm"])) <= 1e-10 * float(gold["adj_norm"])


@pytest.mark.parametrize("gemm", ["ozaki", "tf32", "simt"])
def test_fp32_full_size_vs_reference_golden(built, golden_dir, gemm, monkeypatch):
    """fp32 mode at full size (contraction length 3144) against the reference's golden vectors, 1e-5 budget,
    for the three spectral-response kernels: the int8-sliced product on tcgen05 (4 digits, the default), 3xTF32
    mma.sync (slab-wise accumulation, without which the truncating tensor-core adds bias the sum by 2e-5) and the
    FFMA kernel."""
    monkeypatch.setenv("SURFH_F32_GEMM", gemm)
    cfg = CASES["c1_band1a"]()
    gold = np.load(os.path.join(golden_dir, "c1_band1a.npz"))
    gpu = built(**cfg.model_args(), dtype="float32", adjoint_mode="reference")
    y = gpu.forward(cfg.maps)
    assert rel(y[::int(gold["fwd_stride"])], gold["fwd_sample"]) <= 1e-5
    v = np.random.default_rng(1234).standard_normal(gpu.osize)
    x = gpu.adjoint(v)
    assert rel(x[:, ::7, ::7], gold["adj_sample"]) <= 1e-5


@pytest.mark.parametrize("name", ["mini_2band_4p", "c1_band1a"])
def test_dot_test_exact_adjoint(built, name):
    """<Hx, y> = <x, H^T y> to 1e-6 (BASELINE.json) in exact mode; it holds to ~1e-13 in fp64."""
    cfg = CASES[name]()
    gpu = built(**cfg.model_args(), dtype="float64", adjoint_mode="exact")
    rng = np.random.default_rng(0)
    for _ in range(2):
        u, v = rng.standard_normal(gpu.isize), rng.standard_normal(gpu.osize)
        left = float(np.vdot(gpu.rmatvec(v), u))
        right = float(np.vdot(v, gpu.matvec(u)))
        assert abs(left - right) <= 1e-6 * abs(right)
        assert abs(left - right) <= 1e-11 * abs(right)


def test_device_tensors_and_fwadj(built):
    import torch
    cfg = CASES["mini_2band_4p"]()
    gpu = built(**cfg.model_args(), dtype="float64", adjoint_mode="exact")
    x = torch.as_tensor(cfg.maps, device="cuda")
    y = gpu.forward(x)
    assert y.is_cuda and y.shape == (gpu.osize,)
    assert rel(y.cpu().numpy(), gpu.forward(cfg.maps)) <= 1e-14
    q = gpu.fwadj(x)
    assert rel(q.cpu().numpy(), gpu.adjoint(gpu.forward(cfg.maps))) <= 1e-13
    # host path of the same LinOp call: numpy in -> numpy out, optionally into a caller-provided array
    q_np = gpu.fwadj(cfg.maps)
    assert isinstance(q_np, np.ndarray) and q_np.shape == gpu.ishape and q_np.dtype == np.float64
    assert rel(q_np, q.cpu().numpy()) <= 1e-15
    into = np.empty(gpu.ishape)
    assert gpu.fwadj(cfg.maps, out=into) is into and np.array_equal(into, q_np)
    # deterministic: bitwise-identical across runs (no atomics anywhere on the path)
    assert torch.equal(q, gpu.fwadj(x))
    assert torch.equal(gpu.adjoint(y), gpu.adjoint(y))


def test_wavelength_shards_sum_to_the_full_operator(built):
    """Two wavelength shards built in ONE process (no communicator): their partial forwards / adjoints add up to the
    unsharded operator's.  The cuts are chosen so that a band's local window has an odd number of wavelengths and
    an odd beta width: the K-fast slit-space rows then need their padded (even) pitch for the TMA tensor maps."""
    import torch
    cfg = CASES["mini_2band_4p"]()
    args = cfg.model_args()
    n_l = len(cfg.wavelength_axis)
    full = built(**args, adjoint_mode="exact")
    x = torch.as_tensor(cfg.maps, device="cuda")
    v = torch.as_tensor(np.random.default_rng(2).standard_normal(full.osize), device="cuda")
    y_full, a_full = full.forward(x), full.adjoint(v)
    for cut in (n_l // 2, n_l // 2 + 1, 7):
        lo = built(**args, adjoint_mode="exact", lambda_range=(0, cut))
        hi = built(**args, adjoint_mode="exact", lambda_range=(cut, n_l))
        assert lo.partial and hi.partial
        assert rel((lo.forward(x) + hi.forward(x)).cpu().numpy(), y_full.cpu().numpy()) <= 1e-13
        assert rel((lo.adjoint(v) + hi.adjoint(v)).cpu().numpy(), a_full.cpu().numpy()) <= 1e-13
        with pytest.raises(ValueError, match="needs comm"):
            lo.fwadj(x)


def test_partial_handle_host_forward_after_adjoint_has_zero_foreign_blocks(built):
    """A handle that owns only some bands shares one host staging buffer between `adjoint` (input) and `forward`
    (output): the blocks of the bands it does not own must read as zeros whatever was staged before
    (round-1 ADVICE: they used to return the previous adjoint's input)."""
    cfg = CASES["mini_2band_4p"]()
    full = built(**cfg.model_args())
    part = built(**cfg.model_args(), local_bands=[0])
    v = np.random.default_rng(6).standard_normal(full.osize) + 3.0
    part.adjoint(v)                       # stages v, non-zero everywhere
    y = part.forward(cfg.maps)
    cut = int(full._idx[1])
    assert np.all(y[cut:] == 0.0)
    assert rel(y[:cut], full.forward(cfg.maps)[:cut]) <= 1e-13


def test_contraction_backends_agree(built, monkeypatch):
    """The spectral response runs as an int8-sliced product on the tcgen05 tensor cores by default (8 digits in
    fp64, 4 in fp32: csrc/kernels_ozaki.cuh).  It has to agree with the FP64 DMMA kernels it replaces -- to the
    rounding of those kernels with 8 digits, to 2^-(7 digits) with fewer -- and with the oracle (reference:
    jax_utils.wblur_subSampling / wblur_t, surfh/ToolsDir/jax_utils.py:72-91)."""
    from surfh_oracle import model as om
    cfg = CASES["band2a_4p"]()
    args = cfg.model_args()
    v = np.random.default_rng(7).standard_normal(int(om.SpectroLMM(**args).osize))
    monkeypatch.delenv("SURFH_F64_GEMM", raising=False)
    monkeypatch.delenv("SURFH_OZAKI_DIGITS", raising=False)
    oz = built(**args, dtype="float64", adjoint_mode="exact")
    info = oz.contraction_info()
    assert (info["mode"], info["digits"]) == ("ozaki_i8", 8) and 0.3 < info["executed_fraction"] <= 1.0
    y_oz, x_oz = oz.forward(cfg.maps), oz.adjoint(v)
    # skipping the all-zero digit tiles of the line-spread function changes nothing: the sums are integers
    monkeypatch.setenv("SURFH_OZAKI_DENSE", "1")
    dense = built(**args, dtype="float64", adjoint_mode="exact")
    assert dense.contraction_info()["executed_fraction"] == 1.0
    assert np.array_equal(dense.forward(cfg.maps), y_oz) and np.array_equal(dense.adjoint(v), x_oz)
    monkeypatch.delenv("SURFH_OZAKI_DENSE")
    monkeypatch.setenv("SURFH_F64_GEMM", "tma")
    dm = built(**args, dtype="float64", adjoint_mode="exact")
    assert dm.contraction_info()["mode"] == "dmma_tma"
    y_dm, x_dm = dm.forward(cfg.maps), dm.adjoint(v)
    monkeypatch.setenv("SURFH_F64_GEMM", "mma")
    assert built(**args, dtype="float64").contraction_info()["mode"] == "mma_sync"
    assert rel(y_oz, y_dm) <= 1e-13 and rel(x_oz, x_dm) <= 1e-13
    monkeypatch.setenv("SURFH_F64_GEMM", "ozaki")
    for digits, tol in ((7, 1e-11), (6, 1e-9)):
        monkeypatch.setenv("SURFH_OZAKI_DIGITS", str(digits))
        m = built(**args, dtype="float64", adjoint_mode="exact")
        assert m.contraction_info()["digits"] == digits
        assert rel(m.forward(cfg.maps), y_dm) <= tol and rel(m.adjoint(v), x_dm) <= tol
    monkeypatch.delenv("SURFH_OZAKI_DIGITS")
    monkeypatch.delenv("SURFH_F64_GEMM")
    f32 = built(**args, dtype="float32", adjoint_mode="exact")
    assert (f32.contraction_info()["mode"], f32.contraction_info()["digits"]) == ("ozaki_i8", 4)
    assert rel(f32.forward(cfg.maps), y_dm) <= 1e-5 and rel(f32.adjoint(v), x_dm) <= 1e-5
    monkeypatch.setenv("SURFH_F32_GEMM", "tf32")
    t32 = built(**args, dtype="float32", adjoint_mode="exact")
    assert t32.contraction_info()["mode"] == "mma_sync"
    assert rel(t32.forward(cfg.maps), y_dm) <= 1e-5


def test_contraction_nonfinite_and_tiny_inputs(built):
    """The sliced contraction scales every row by its own power of two: a NaN in the input must still poison the
    output (the row maximum alone would drop it), and inputs of any magnitude keep their relative accuracy."""
    cfg = CASES["mini_2band_4p"]()
    gpu = built(**cfg.model_args(), dtype="float64", adjoint_mode="exact")
    y = gpu.forward(cfg.maps)
    tiny = gpu.forward(1e-290 * cfg.maps)
    assert rel(tiny * 1e290, y) <= 1e-13
    huge = gpu.forward(1e250 * cfg.maps)
    assert rel(huge * 1e-250, y) <= 1e-13
    x = cfg.maps.copy()
    x[0, x.shape[1] // 2, x.shape[2] // 2] = np.nan
    assert np.isnan(gpu.forward(x)).any()
    v = np.random.default_rng(3).standard_normal(gpu.osize)
    v[17] = np.inf
    assert not np.isfinite(gpu.adjoint(v)).all()

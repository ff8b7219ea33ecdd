"""Multi-GPU parity inside `pytest -m gpu`: launches tests/dist_check.py under torchrun on every GPU of the
box (2 at least; skipped on a single-GPU box) -- wavelength-sharded forward / fwadj / CG / criterion against
the unsharded model, on the mini configuration, on MRSBlurred and on BASELINE.json's full-size C4."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return torch.cuda.device_count()


def _launch(n, which, timeout):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "dist_check.py")] + which
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST OK " + " ".join(which) in res.stdout, res.stdout[-3000:]
    return res.stdout


def test_sharded_equals_unsharded_mini_and_blind():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    print(_launch(min(n, 8), ["mini", "blind"], 600))


def test_sharded_equals_unsharded_c4():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    print(_launch(min(n, 8), ["c4"], 1200))

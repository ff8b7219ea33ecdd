"""Iterations-to-tolerance of plain vs Fourier-block-preconditioned CG on a BASELINE
configuration, for a few scalings of the wavelength weights.  Run on the GPU box:
    python tools/precond_sweep.py c2 200 > gpurun_out/precond_c2.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from surfh_b200 import fusion_CT, synthetic
    from surfh_b200.model import spectroSigRLSCT
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    mu = float(sys.argv[3]) if len(sys.argv) > 3 else 5e3
    cfg = synthetic.baseline_config(name)
    dev = torch.device("cuda")
    sotf = lambda lo, hi: synthetic.ir2fr_device(cfg.psf[lo:hi], cfg.imshape, dev, torch.float64)  # noqa: E731
    model = spectroSigRLSCT(sotf, cfg.templates, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, cfg.instrs,
                            cfg.step_degree, cfg.pointings, adjoint_mode="exact")
    x_true = torch.as_tensor(cfg.maps, device=dev)
    y = model.forward(x_true)
    g = torch.Generator(device=dev).manual_seed(1)
    y = y + 0.01 * y.pow(2).mean().sqrt() * torch.randn(y.shape, dtype=y.dtype, device=dev, generator=g)
    crit = fusion_CT.QuadCriterion_MRS(1.0, y, model, mu)

    def its(gn, drop):
        gn = np.asarray(gn)
        hit = np.flatnonzero(gn <= drop * gn[0])
        return int(hit[0]) if len(hit) else -1

    def report(tag, res, seconds):
        gn = res.grad_norm
        print(f"{tag:>28s}: its to |r|^2/|r0|^2 <= 1e-4 / 1e-6 / 1e-8 / 1e-10: "
              f"{its(gn, 1e-4):4d} {its(gn, 1e-6):4d} {its(gn, 1e-8):4d} {its(gn, 1e-10):4d}   "
              f"J({len(gn) - 1}) = {crit.get_crit_val(res.x):.8e}   {seconds:.2f} s", flush=True)

    import time
    print(f"config {name}, mu_reg {mu:g}, {n_it} iterations, N = {len(cfg.alpha_axis)}, K = {cfg.templates.shape[0]}")
    t0 = time.time()
    plain = fusion_CT.lcg(model, y, 1.0, mu, np.zeros(model.ishape), tol=1e-30, max_iter=n_it, check_every=n_it)
    report("plain CG", plain, time.time() - t0)
    for scale in (1.0, 0.3, 0.1, 0.03, 0.01):
        pre = fusion_CT.FourierPreconditioner(model, 1.0, mu, scale=scale)
        t0 = time.time()
        res = fusion_CT.lcg(model, y, 1.0, mu, np.zeros(model.ishape), tol=1e-30, max_iter=n_it, check_every=n_it,
                            precond=pre)
        report(f"PCG, weights x {scale:g}", res, time.time() - t0)


if __name__ == "__main__":
    main()

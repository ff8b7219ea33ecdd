"""Multi-GPU parity check, run under torchrun on N GPUs (tests/test_gpu_dist.py launches it):
the wavelength-sharded operator / CG on N ranks must equal the single-GPU result.

    torchrun --nproc-per-node N tests/dist_check.py [mini] [c4] [blind]

  mini   mini_2band_4p: forward, fwadj, an 8-iteration CG with a refresh, the criterion (both ways)
  c4     BASELINE.json's full-size workload (12 bands, N = 501, K = 6): sharded forward / fwadj vs the
         unsharded model built on the same GPU, and 5 CG iterations
  blind  MRSBlurred (beta-sum band: every rank writes only the detector rows of its own wavelengths) sharded
         by wavelength, applied three times in a row (the staleness bug ADVICE r1 describes shows from the
         second application on)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    return float((a - b).norm() / b.norm())


def check_mini(torch, comm, dev, bench):
    from cases import CASES
    from surfh_b200 import fusion_CT
    from surfh_b200.model import spectroSigRLSCT
    cfg = CASES["mini_2band_4p"]()
    args = cfg.model_args()
    lam = bench.shard_for_rank(cfg, comm, 8)
    sharded = spectroSigRLSCT(**args, adjoint_mode="exact", lambda_range=lam, comm=comm)
    full = spectroSigRLSCT(**args, adjoint_mode="exact")
    x = torch.as_tensor(cfg.maps, device=dev)
    e_f = rel(sharded.forward(x), full.forward(x))
    e_q = max(rel(sharded.fwadj(x), full.fwadj(x)) for _ in range(2))
    y = full.forward(x).cpu().numpy()
    y = y + 0.01 * np.sqrt(np.mean(y ** 2)) * np.random.default_rng(9).standard_normal(y.shape)
    r_s = fusion_CT.lcg(sharded, y, 1.0, 5.0, max_iter=8, tol=1e-12, refresh=4)
    r_f = fusion_CT.lcg(full, y, 1.0, 5.0, max_iter=8, tol=1e-12, refresh=4)
    e_x = float(np.linalg.norm(r_s.x - r_f.x) / np.linalg.norm(r_f.x))
    j_s = fusion_CT.QuadCriterion_MRS(1, y, sharded, 5.0).get_crit_val(r_s.x)
    j_f = fusion_CT.QuadCriterion_MRS(1, y, full, 5.0).get_crit_val(r_f.x)
    j_state = r_s.solver.criterion_from_state()
    print(f"[mini] rank {comm.rank} range {lam}: forward {e_f:.2e} fwadj {e_q:.2e} cg {e_x:.2e} "
          f"crit {abs(j_s - j_f) / abs(j_f):.2e} crit(state) {abs(j_state - j_f) / abs(j_f):.2e}", flush=True)
    assert e_f < 1e-12 and e_q < 1e-12 and e_x < 1e-9 and abs(j_s - j_f) < 1e-10 * abs(j_f)
    assert abs(j_state - j_f) < 1e-10 * abs(j_f)
    # the reference driver's call on a sharded model, reference adjoint: the criterion trace comes from H x_k
    # tracked on the detector blocks each rank owns (one scalar all-reduce), no forward pass
    sh_ref = spectroSigRLSCT(**args, adjoint_mode="reference", lambda_range=lam, comm=comm)
    fu_ref = spectroSigRLSCT(**args, adjoint_mode="reference")
    qs = fusion_CT.QuadCriterion_MRS(1, y, sh_ref, 5.0)
    qf = fusion_CT.QuadCriterion_MRS(1, y, fu_ref, 5.0)
    rs = qs.run_method("lcg", 12, perf_crit=1, calc_crit=True, value_init=0)
    rf = qf.run_method("lcg", 12, perf_crit=1, calc_crit=True, value_init=0)
    e_c = max(abs(a - b) / abs(b) for a, b in zip(qs.L_crit_val, qf.L_crit_val))
    print(f"[mini] rank {comm.rank}: run_method criterion trace sharded vs unsharded {e_c:.2e} "
          f"({qs._solver()._state_evals} of {len(qs.L_crit_val)} from the CG state)", flush=True)
    assert len(qs.L_crit_val) == len(qf.L_crit_val) == 3 and e_c < 1e-10
    assert qs._solver()._state_evals == len(qs.L_crit_val)
    assert float(np.linalg.norm(rs.x - rf.x) / np.linalg.norm(rf.x)) < 1e-9
    # an unsharded model with a comm must not sum W identical copies
    dup = spectroSigRLSCT(**args, adjoint_mode="exact", comm=comm)
    assert rel(dup.fwadj(x), full.fwadj(x)) < 1e-14 and rel(dup.forward(x), full.forward(x)) < 1e-14


def check_c4(torch, comm, dev, bench):
    from surfh_b200 import fusion_CT, synthetic
    from surfh_b200.model import spectroSigRLSCT
    cfg = synthetic.baseline_config("c4")
    sotf = lambda lo, hi: synthetic.ir2fr_device(cfg.psf[lo:hi], cfg.imshape, dev, torch.float64)  # noqa: E731
    lam = bench.shard_for_rank(cfg, comm, 8)
    common = (cfg.templates, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, cfg.instrs, cfg.step_degree,
              cfg.pointings)
    sharded = spectroSigRLSCT(sotf, *common, adjoint_mode="reference", lambda_range=lam, comm=comm)
    full = spectroSigRLSCT(sotf, *common, adjoint_mode="reference")
    x = torch.as_tensor(cfg.maps, device=dev)
    y_full = full.forward(x)
    e_f = rel(sharded.forward(x), y_full)
    q_full = full.fwadj(x)
    e_q = max(rel(sharded.fwadj(x), q_full) for _ in range(2))
    y = y_full.cpu().numpy()
    y = y + 0.01 * np.sqrt(np.mean(y ** 2)) * np.random.default_rng(9).standard_normal(y.shape)
    r_s = fusion_CT.lcg(sharded, y, 1.0, 5e3, max_iter=5, tol=1e-12)
    r_f = fusion_CT.lcg(full, y, 1.0, 5e3, max_iter=5, tol=1e-12)
    e_x = float(np.linalg.norm(r_s.x - r_f.x) / np.linalg.norm(r_f.x))
    print(f"[c4] rank {comm.rank} range {lam}: forward {e_f:.2e} fwadj {e_q:.2e} cg(5) {e_x:.2e}", flush=True)
    assert e_f < 1e-12 and e_q < 1e-12 and e_x < 1e-10


def check_blind(torch, comm, dev, bench):
    from cases import BLIND
    from surfh_b200 import synthetic
    from surfh_b200.spectro_blind import MRSBlurred
    factory, _ = BLIND["blind_mini_2p"]
    cfg = factory()
    n_l = len(cfg.wavelength_axis)
    cut = [round(n_l * r / comm.world_size) for r in range(comm.world_size + 1)]
    lam = (cut[comm.rank], cut[comm.rank + 1])
    sotf = cfg.sotf()
    args = (cfg.alpha_axis, cfg.beta_axis, cfg.instrs[0], cfg.step_degree, cfg.pointings[0])
    sharded = MRSBlurred(sotf, *args, adjoint_mode="exact", lambda_range=lam, comm=comm)
    full = MRSBlurred(sotf, *args, adjoint_mode="exact")
    x = torch.as_tensor(cfg.maps, device=dev)
    e_f = max(rel(sharded.forward(x), full.forward(x)) for _ in range(2))
    want = full.fwadj(x)
    errs = [rel(sharded.fwadj(x), want) for _ in range(3)]   # stale sums would show at the 2nd call
    v = torch.randn(full.osize, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
    e_a = rel(sharded.adjoint(v), full.adjoint(v))
    print(f"[blind] rank {comm.rank} range {lam}: forward {e_f:.2e} fwadj x3 {max(errs):.2e} adjoint {e_a:.2e}", flush=True)
    assert e_f < 1e-12 and max(errs) < 1e-12 and e_a < 1e-12


def main():
    import torch
    from surfh_b200 import dist
    sys.path.insert(0, ROOT)
    import bench

    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["mini", "blind"]
    comm = dist.init_from_env("nccl")
    assert comm is not None, "run under torchrun with at least 2 ranks"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    for name in which:
        {"mini": check_mini, "c4": check_c4, "blind": check_blind}[name](torch, comm, dev, bench)
        comm.barrier()
    import torch.distributed as td
    td.destroy_process_group()
    if comm.rank == 0:
        print("DIST OK " + " ".join(which))


if __name__ == "__main__":
    main()

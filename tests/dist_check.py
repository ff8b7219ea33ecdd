"""Multi-GPU parity check, run under torchrun on N GPUs (not collected by pytest):
sharded fwadj / CG on N ranks must equal the single-GPU result computed on rank 0's device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from cases import CASES
    from surfh_b200 import dist, fusion_CT
    from surfh_b200.model import spectroSigRLSCT
    sys.path.insert(0, ROOT)
    import bench

    comm = dist.init_from_env("nccl")
    rank = comm.rank if comm else 0
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    cfg = CASES["mini_2band_4p"]()
    args = cfg.model_args()
    lam = bench.shard_for_rank(cfg, comm, 8)
    sharded = spectroSigRLSCT(**args, adjoint_mode="exact", lambda_range=lam, comm=comm)
    full = spectroSigRLSCT(**args, adjoint_mode="exact")
    x = torch.as_tensor(cfg.maps, device=dev)
    rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
    e_f = rel(sharded.forward(x), full.forward(x))
    e_q = rel(sharded.fwadj(x), full.fwadj(x))
    y = full.forward(x).cpu().numpy()
    y = y + 0.01 * np.sqrt(np.mean(y ** 2)) * np.random.default_rng(9).standard_normal(y.shape)
    r_s = fusion_CT.lcg(sharded, y, 1.0, 5.0, max_iter=8, tol=1e-12, refresh=4)
    r_f = fusion_CT.lcg(full, y, 1.0, 5.0, max_iter=8, tol=1e-12, refresh=4)
    e_x = float(np.linalg.norm(r_s.x - r_f.x) / np.linalg.norm(r_f.x))
    j_s = fusion_CT.QuadCriterion_MRS(1, y, sharded, 5.0).get_crit_val(r_s.x)
    j_f = fusion_CT.QuadCriterion_MRS(1, y, full, 5.0).get_crit_val(r_f.x)
    print(f"rank {rank} range {lam}: forward {e_f:.2e} fwadj {e_q:.2e} cg {e_x:.2e} crit {abs(j_s - j_f) / abs(j_f):.2e}",
          flush=True)
    assert e_f < 1e-12 and e_q < 1e-12 and e_x < 1e-9 and abs(j_s - j_f) < 1e-10 * abs(j_f)
    if comm:
        comm.barrier()
        import torch.distributed as td
        td.destroy_process_group()
    if rank == 0:
        print("DIST OK")


if __name__ == "__main__":
    main()

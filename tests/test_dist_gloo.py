"""Wavelength sharding across ranks, exercised with 2 gloo processes on the CPU: each rank builds
the tables of its own wavelength range, applies them (numpy emulation of the kernels), and the
all-reduce of the partial detector vectors / partial cubes must reproduce the unsharded oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as tdist

    import _emulate
    from cases import CASES
    from surfh_b200 import dist, geometry, instru
    from surfh_oracle import model as om

    tdist.init_process_group("gloo", rank=rank, world_size=world)
    comm = dist.Comm()
    cfg = CASES["mini_2band_4p"]()
    srfs = instru.get_srf([i.det_pix_size for i in cfg.instrs], cfg.step_degree * 3600)
    # cost model from band sizes only, then this rank's wavelength range
    light = [geometry.build_band(i, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, s, cfg.pointings[k],
                                 cfg.step_degree, with_adjoint=False, lambda_range=(0, 0))
             for k, (i, s) in enumerate(zip(cfg.instrs, srfs))]
    bands = [dict(wave_start=t.wslice.start, n_wave=t.wslice.stop - t.wslice.start, n_det=t.n_det, nb=t.nb,
                  n_pointing=t.n_pointing, n_slit=t.n_slit, na=t.na, local_a=t.local_shape[0],
                  local_b=t.local_shape[1]) for t in light]
    costs = dist.lambda_costs(len(cfg.wavelength_axis), bands, len(cfg.alpha_axis))
    ranges = dist.partition_lambda(costs, world)
    lo, hi = ranges[rank]
    tabs = [geometry.build_band(i, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, s, cfg.pointings[k],
                                cfg.step_degree, lambda_range=(lo, hi))
            for k, (i, s) in enumerate(zip(cfg.instrs, srfs))]
    oracle = om.SpectroLMM(**cfg.model_args(), adjoint_mode="exact")
    blurred = oracle.blurred_cube(cfg.maps)
    # forward: partial sums over this rank's wavelengths, bands it does not touch contribute zeros
    parts = []
    for tb in tabs:
        if tb.is_local:
            parts.append(_emulate.forward([tb], blurred, len(cfg.beta_axis)))
        else:
            parts.append(np.zeros(int(np.prod(tb.oshape))))
    y = torch.from_numpy(np.concatenate(parts))
    comm.allreduce_sum(y)
    y_ref = oracle.forward(cfg.maps)
    err_f = float(np.linalg.norm(y.numpy() - y_ref) / np.linalg.norm(y_ref))
    # adjoint: every rank scatters only its wavelengths; the cube is the disjoint union
    v = np.random.default_rng(3).standard_normal(oracle.osize)
    cube = torch.from_numpy(_emulate.adjoint_cube(tabs, v, oracle.cube_shape, "exact"))
    comm.allreduce_sum(cube)
    c_ref = oracle.adjoint_cube(v)
    err_a = float(np.linalg.norm(cube.numpy() - c_ref) / np.linalg.norm(c_ref))
    ret[rank] = (err_f, err_a, ranges)
    tdist.destroy_process_group()


def test_two_rank_lambda_sharding_reproduces_oracle():
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err_f, err_a, ranges = ret[rank]
        assert err_f < 1e-13 and err_a < 1e-13, (rank, err_f, err_a)
        assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0]


def test_partition_helpers():
    from surfh_b200 import dist
    parts = dist.partition_bands([5, 3, 8, 1, 7, 2, 6, 4, 9, 3, 2, 1], 8)
    assert sorted(i for p in parts for i in p) == list(range(12)) and all(parts)
    c = np.ones(100)
    c[40:60] = 3
    r = dist.partition_lambda(c, 4)
    assert r[0][0] == 0 and r[-1][1] == 100 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sums = [c[a:b].sum() for a, b in r]
    assert max(sums) - min(sums) <= 6
    assert dist.partition_lambda(np.ones(8), 8) == [(i, i + 1) for i in range(8)]
    with pytest.raises(ValueError):
        dist.partition_lambda(np.zeros(4), 2)


def _exchange_worker(rank, world, port, ret):
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as tdist
    from surfh_b200 import dist

    tdist.init_process_group("gloo", rank=rank, world_size=world)
    comm = dist.Comm()
    ranges = [(0, 10), (10, 20), (20, 30)]
    windows = [(0, 15), (12, 25), (26, 30), (0, 30)]      # shared by {0,1}, {1,2}, {2}, everyone
    blocks = [(0, 5), (5, 7), (12, 3), (15, 4)]
    ex = dist.BandExchange(comm, windows, blocks, ranges[rank])
    y = torch.zeros(19, dtype=torch.float64)
    for b, (off, size) in enumerate(blocks):                # partial sums: rank r contributes (r + 1) * (b + 1)
        if rank in ex.rank_sets[b]:
            y[off: off + size] = (rank + 1) * (b + 1)
    ex.reduce_shared(y)
    ret[rank] = (ex.ranges, ex.rank_sets, y.numpy().copy(), ex.bytes_per_application(8))
    tdist.destroy_process_group()


def test_band_exchange_sums_each_band_among_its_ranks_only():
    """3 gloo ranks: every shared band is summed on a sub-communicator of exactly the ranks that hold a
    share of it; a rank ends up with complete blocks for the bands it touches."""
    import torch.multiprocessing as mp
    world = 3
    port = 31500 + (os.getpid() % 2000)
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_exchange_worker, args=(world, port, ret), nprocs=world, join=True)
    blocks = [(0, 5), (5, 7), (12, 3), (15, 4)]
    sets = [[0, 1], [1, 2], [2], [0, 1, 2]]
    for rank in range(world):
        ranges, rank_sets, y, nbytes = ret[rank]
        assert ranges == [(0, 10), (10, 20), (20, 30)] and rank_sets == sets
        for b, (off, size) in enumerate(blocks):
            if rank in sets[b]:
                want = (b + 1) * sum(r + 1 for r in sets[b])
                assert np.all(y[off: off + size] == want), (rank, b, y)
            else:
                assert np.all(y[off: off + size] == 0)
        assert nbytes == 8 * sum(size for b, (off, size) in enumerate(blocks) if rank in sets[b] and len(sets[b]) > 1)


def test_band_rank_sets():
    from surfh_b200 import dist
    assert dist.band_rank_sets([(0, 10), (8, 20), (25, 30)], [(0, 9), (9, 18), (18, 40)]) == [[0, 1], [0, 1, 2], [2]]
    assert dist.band_rank_sets([(5, 6)], [(0, 5), (5, 6), (6, 9)]) == [[1]]


def test_c4_partition_is_contiguous_complete_and_balanced():
    """The wavelength partition bench.py uses for the full 12-band workload at 2, 4 and 8 ranks."""
    sys.path.insert(0, ROOT)
    import bench
    from surfh_b200 import dist, synthetic
    cfg = synthetic.baseline_config("c4")
    bands = bench.band_summaries(cfg)
    assert len(bands) == 12 and all(b["hull_rows"] <= 501 and b["srf"] in (7, 9, 10) for b in bands)
    costs = dist.lambda_costs(len(cfg.wavelength_axis), bands, len(cfg.alpha_axis), 8)
    assert costs.min() >= 0 and (costs > 0).mean() > 0.95  # a few wavelengths at the ends lie outside every band
    for world in (2, 4, 8):
        parts = dist.partition_lambda(costs, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(costs)
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(parts, parts[1:])) and parts[-1][0] < parts[-1][1]
        loads = [costs[a:b].sum() for a, b in parts]
        assert max(loads) / min(loads) < 1.02
        # every band is shared by a contiguous set of ranks, and no rank is left without a band
        sets = dist.band_rank_sets([(b["wave_start"], b["wave_start"] + b["n_wave"]) for b in bands], parts)
        assert all(s == list(range(s[0], s[-1] + 1)) for s in sets)
        assert set(r for s in sets for r in s) == set(range(world))

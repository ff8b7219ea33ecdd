#!/usr/bin/env python
"""Benchmark of the LMM instrument operator (forward + adjoint application) and of the CG loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4] [--dtype float64]
    python bench.py --impl reference ...      # the reference's CPU path (oracle) on the host cores

One JSON line on stdout (rank 0).  A "step" is one application of the normal operator
H^T H (one `forward` + one `adjoint`, the unit of work of one CG iteration) on the synthetic
configuration named in `config.workload`:

  value      applications/s, maps resident in HBM, detector vector never leaves the device
             (timed with CUDA events on the launching stream, max over ranks)
  e2e        the same application through the reference-facing LinOp calls with HOST buffers:
             numpy maps -> forward -> numpy y -> adjoint -> numpy maps (H2D/D2H inside)
  cg_iters_per_s   K iterations of the device-resident CG (fwadj + fused vector kernels
             (+ NCCL all-reduce of the map gradient when sharded))
  roofline   dominant hand-written kernel, achieved vs measured peak; `stages` lists every stage
  cpu_baseline     the CPU oracle (numpy/scipy restatement of the reference path) on a bounded
             sample of the same workload, same box, same run (rank 0, N=1)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

if "reference" in sys.argv[1:]:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must see every host core like a plain
    # `python bench.py --impl reference` does (all host cores) -- before numpy loads
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lmm_fwd_adj_applications_per_s"
UNIT = "applications/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", help="c1..c4 of BASELINE.md (default: c4, the 12-band workload)")
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--adjoint", default="reference", choices=["reference", "exact"])
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--solve-iters", type=int, default=100, help="iterations of the timed CG solve (0 = skip)")
    return ap.parse_args()


# --------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_fp_peak(torch, tdtype):
    """Arithmetic peak of the compute dtype on this GPU, measured here: cuBLAS GEMM 6144^3 (fp64 runs on the
    FP64 tensor pipe, fp32 with TF32 off on the FMA pipe), best of 3."""
    try:
        n = 6144
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        a = torch.randn((n, n), dtype=tdtype, device="cuda")
        b = torch.randn((n, n), dtype=tdtype, device="cuda")
        torch.matmul(a, b)
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        return {"tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "source": f"measured here: cuBLAS {tdtype} GEMM {n}^3, best of 3"}
    except Exception as exc:  # noqa: BLE001
        return {"tflops": None, "source": f"unavailable ({exc})"}


def measured_int8_peak():
    """Dense int8 tensor rate (TOP/s).  MEASURED_PEAKS.json holds no int8 figure (its dense bf16 GEMM runs power-capped
    at 1312 MHz; the int8 contraction holds 1965 MHz and exceeds twice that bf16 rate), so the denominator is the
    recipe's nominal dense 8-bit rate: 4.5 POP/s (B200_PROFILING.md; 16384 ops per clock and SM x 148 SMs x 1.965 GHz
    = 4.76 is what ncu's sm__ops_path_tensor_op_utcimma peak uses)."""
    return 4500.0, "nominal dense 8-bit tensor rate, 4.5 POP/s (B200_PROFILING.md: MEASURED_PEAKS.json has no int8 figure)"


def ncu_traffic():
    """DRAM bytes per launch of each stage's kernels from the committed ncu --set full capture of this same
    command (profiles/traffic.json, written by tools/ncu_traffic.py); {} when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f).get("stages", {})


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_config(name: str):
    from surfh_b200 import synthetic
    return synthetic.baseline_config(name)


def band_summaries(cfg):
    """Sizes of every band from the instrument table alone (no device tables built)."""
    from surfh_b200 import geometry, instru
    srfs = instru.get_srf([i.det_pix_size for i in cfg.instrs], cfg.step_degree * 3600)
    out = []
    for ifu, srf, pts in zip(cfg.instrs, srfs, cfg.pointings):
        band = ifu.pix(cfg.step_degree)
        la, lb = geometry.local_axes(band.fov, cfg.step_degree, geometry.N_MARGIN_PIX * cfg.step_degree)
        _, _, na, nbw, _ = geometry.slit_layout(band, cfg.beta_axis, la, lb, srf)
        wsl = band.wslice(cfg.wavelength_axis, geometry.WAVE_MARGIN_UM)
        ang = np.deg2rad(band.fov.angle)
        hull_rows = len(la) * abs(np.cos(ang)) + len(lb) * abs(np.sin(ang)) + 6  # rotated local grid + dithers
        out.append(dict(wave_start=wsl.start, n_wave=wsl.stop - wsl.start, n_det=band.n_wavel, nb=nbw,
                        n_pointing=len(pts), n_slit=band.n_slit, na=na, local_a=len(la), local_b=len(lb),
                        hull_rows=float(hull_rows), srf=int(srf)))
    return out


def shard_for_rank(cfg, comm, dtype_bytes):
    """Contiguous wavelength range of this rank (None when single-process)."""
    if comm is None:
        return None
    from surfh_b200 import dist
    costs = dist.lambda_costs(len(cfg.wavelength_axis), band_summaries(cfg), len(cfg.alpha_axis), dtype_bytes)
    return dist.partition_lambda(costs, comm.world_size)[comm.rank]


# ------------------------------------------------------------------- CPU reference legs
def _channel_representatives(cfg):
    """One band per MRS channel present in the configuration (bands of a channel share S, srf, the local
    grid, na and nb -- SURVEY section 8d -- so they cost the same per wavelength): [(band index, bands in
    that channel)]."""
    groups = {}
    for i, name in enumerate(cfg.band_names):
        groups.setdefault(name[0], []).append(i)
    return [(idx[0], len(idx)) for _, idx in sorted(groups.items())]


def cpu_sample(cfg, repeats: int, warmup: int):
    """Time the CPU oracle (numpy/scipy fp64 restatement of the reference path, all host threads) on a bounded
    sample of the workload and extrapolate one full application from the sample's OWN stage times.

    Sample = one band per MRS channel, ONE dither, on that band's wavelength window, same N and K.  For each
    sampled band the global stages (templates T and spatial blur C of the forward, C^T and T^T of the adjoint:
    per cube plane, independent of the dithers) and the per-dither stages (gridding, box-sum FFTs, slit
    slicing, spectral response and their adjoints) are timed separately; then
        t_application = sum_channels bands_in_channel * ( planes_scale * t_global + P * t_per_dither )
    where planes_scale = (planes of the configuration's union of windows) / (sum of the bands' windows)
    accounts for the reference transforming every cube plane once however many bands read it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import band_subconfig
    from surfh_oracle import model as om  # executed here only as the timed CPU baseline

    if cfg.templates is None:
        return cpu_sample_blind(cfg, repeats, warmup)
    reps = _channel_representatives(cfg)
    n_point = len(cfg.pointings[0])
    windows = [i.wslice(cfg.wavelength_axis, 0.1) for i in cfg.instrs]
    covered = np.zeros(len(cfg.wavelength_axis), dtype=bool)
    for w in windows:
        covered[w] = True
    planes_scale = float(covered.sum()) / float(sum(w.stop - w.start for w in windows))
    models = [(om.SpectroLMM(**band_subconfig(cfg, b, pointing=0), adjoint_mode="reference"), nb) for b, nb in reps]
    rng = np.random.default_rng(0)
    probes = [rng.standard_normal(m.osize) for m, _ in models]
    samples, estimates = [], []
    for it in range(warmup + repeats):
        t_sample, t_app = 0.0, 0.0
        for (m, n_bands), v in zip(models, probes):
            ch = m.channels[0]
            t0 = time.perf_counter()
            blurred = m.blurred_cube(cfg.maps)                     # T, C (global)
            t1 = time.perf_counter()
            ch.forward(blurred)                                    # S, Sum, L, Sig R (one dither)
            t2 = time.perf_counter()
            cube = m.adjoint_cube(v)                               # R^T Sig^T, L^T, Sum^T, gridding_t (one dither)
            t3 = time.perf_counter()
            om.lmm_cube2maps(om.idft(om.dft(cube) * m.sotf.conj(), m.imshape), m.templates)  # C^T, T^T (global)
            t4 = time.perf_counter()
            t_global, t_dither = (t1 - t0) + (t4 - t3), (t2 - t1) + (t3 - t2)
            t_sample += t4 - t0
            t_app += n_bands * (planes_scale * t_global + n_point * t_dither)
        if it >= warmup:
            samples.append(t_sample)
            estimates.append(t_app)
    t_sample, t_app = float(np.median(samples)), float(np.median(estimates))
    names = ",".join(cfg.band_names[b].upper() for b, _ in reps)
    sample = (f"bands {names} of {cfg.name} (one per MRS channel), one of the {n_point} dithers each, on each band's "
              f"own wavelength window, K={cfg.templates.shape[0]}, N={len(cfg.alpha_axis)}, fp64, forward + adjoint "
              f"(reference gridding_t): {t_sample:.2f} s per sample (median of {repeats}, {warmup} warm-up); one "
              f"application extrapolated from the sample's own stage times as sum_channels n_bands * "
              f"({planes_scale:.3f} * t_global + {n_point} * t_per_dither) = {t_app:.1f} s")
    return {"value": 1.0 / t_app, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
            "seconds_per_sample": t_sample, "sample_fraction": t_sample / t_app, "extrapolated_application_s": t_app,
            "repeats": repeats, "warmup": warmup}


def cpu_sample_blind(cfg, repeats: int, warmup: int, n_wave: int = 16):
    """Config 5: the oracle's MRSBlurred on the first `n_wave` wavelengths, one after the other like the
    reference's per-wavelength script (deconvolution_mrs_single_wavelength.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from surfh_b200 import synthetic
    from surfh_oracle import blind as ob  # executed here only as the timed CPU baseline
    sotf = synthetic.ir2fr(cfg.psf[:n_wave], cfg.imshape)
    models = [ob.MRSBlurred(sotf[l], cfg.alpha_axis, cfg.beta_axis, cfg.instrs[0], cfg.step_degree, cfg.pointings[0])
              for l in range(n_wave)]
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        for l, m in enumerate(models):
            m.adjoint(m.forward(cfg.maps[l]))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.median(times))
    frac = n_wave / len(cfg.wavelength_axis)
    sample = (f"{n_wave} of {len(cfg.wavelength_axis)} wavelengths of {cfg.name} (band {cfg.band_names[0]}, "
              f"N={len(cfg.alpha_axis)}, {len(cfg.pointings[0])} pointings, fp64) = {frac:.4f} of one application in "
              f"{t:.2f} s (median of {repeats}); value = fraction / seconds")
    return {"value": frac / t, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
            "seconds_per_sample": t, "sample_fraction": frac}


def full_output_size(cfg) -> int:
    from surfh_b200 import geometry, instru
    srfs = instru.get_srf([i.det_pix_size for i in cfg.instrs], cfg.step_degree * 3600)
    total = 0
    for ifu, srf, pts in zip(cfg.instrs, srfs, cfg.pointings):
        band = ifu.pix(cfg.step_degree)
        la, lb = geometry.local_axes(band.fov, cfg.step_degree, geometry.N_MARGIN_PIX * cfg.step_degree)
        _, _, na, _, _ = geometry.slit_layout(band, cfg.beta_axis, la, lb, srf)
        total += len(pts) * band.n_slit * band.n_wavel * na
    return total


def run_reference(args):
    """The CPU arm: one step = one bounded sample (see cpu_sample); `ms_per_step` is the measured sample time,
    `value` the applications/s extrapolated from it (= sample_fraction / seconds_per_sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = build_config(args.config)
    steps, warm = max(1, min(args.steps, 3)), min(max(args.warmup, 0), 1)
    base = cpu_sample(cfg, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * base["seconds_per_sample"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_description(cfg, "float64", args),
        "step_definition": "one step = one bounded CPU sample of the workload (cpu_baseline.sample); value = "
                           "sample_fraction / seconds_per_sample, i.e. whole applications per second",
        "sample_fraction": base["sample_fraction"],
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
    }
    emit(line)


def workload_description(cfg, dtype, args):
    if cfg.templates is None:
        return {"workload": f"{cfg.name}: MRSBlurred (non-LMM, per-wavelength C-S-Sum-L-beta-sum chain of "
                            f"spectro_blind.py) batched over {len(cfg.wavelength_axis)} cube wavelengths, band "
                            f"{cfg.band_names[0]}, {len(cfg.alpha_axis)}x{len(cfg.beta_axis)} planes, "
                            f"{len(cfg.pointings[0])} pointings; one step = one forward + one adjoint of the batch",
                "adjoint_mode": args.adjoint, "compute_dtype": dtype,
                "l2_policy": "inputs larger than L2: every step streams the cube, the OTF and the FFT intermediates",
                "parallelism": "single GPU"}
    return {"workload": f"{cfg.name}: {len(cfg.instrs)} MRS band(s) {','.join(cfg.band_names)}, "
                        f"K={cfg.templates.shape[0]} templates, {len(cfg.alpha_axis)}x{len(cfg.beta_axis)} maps, "
                        f"{len(cfg.wavelength_axis)} cube wavelengths, {len(cfg.pointings[0])} pointings; "
                        f"one step = one forward + one adjoint (H^T H) application",
            "adjoint_mode": args.adjoint, "compute_dtype": dtype,
            "l2_policy": "inputs larger than L2: every step streams the OTF and cube chunks (GBs) through HBM",
            "parallelism": f"cube wavelength axis sharded over {args.gpus} GPU(s) (contiguous, cost-balanced); per "
                           f"application: the detector blocks of bands shared by several ranks are summed among those "
                           f"ranks (NCCL sub-communicators), then one all-reduce of the [K,N,N] maps"}


# ------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    from surfh_b200 import dist as sdist
    from surfh_b200 import fusion_CT, synthetic
    from surfh_b200.model import spectroSigRLSCT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; surfh_b200 has no CPU fallback")
    comm = sdist.init_from_env("nccl")
    rank = comm.rank if comm else 0
    world = comm.world_size if comm else 1
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    tdtype = torch.float64 if args.dtype == "float64" else torch.float32
    esz = 8 if args.dtype == "float64" else 4

    cfg = build_config(args.config)
    lam_range = shard_for_rank(cfg, comm, esz) if cfg.templates is not None else None
    shape = cfg.imshape
    t_setup = time.time()
    sotf = lambda lo, hi: synthetic.ir2fr_device(cfg.psf[lo:hi], shape, dev, torch.float64)  # noqa: E731
    if cfg.templates is None:  # configuration 5: the batched non-LMM operator, one GPU
        if comm is not None:
            raise SystemExit("bench.py: config c5 is a single-GPU stress of the FFT + slit path")
        from surfh_b200.spectro_blind import MRSBlurred
        model = MRSBlurred(sotf, cfg.alpha_axis, cfg.beta_axis, cfg.instrs[0], cfg.step_degree, cfg.pointings[0],
                           n_lambda=len(cfg.wavelength_axis), dtype=args.dtype, adjoint_mode=args.adjoint,
                           chunk=args.chunk, device=local_rank)
    else:
        model = spectroSigRLSCT(sotf, cfg.templates, cfg.alpha_axis, cfg.beta_axis, cfg.wavelength_axis, cfg.instrs,
                                cfg.step_degree, cfg.pointings, dtype=args.dtype, adjoint_mode=args.adjoint,
                                lambda_range=lam_range, comm=comm, chunk=args.chunk, device=local_rank)
    setup_s = time.time() - t_setup

    x = torch.as_tensor(cfg.maps, device=dev, dtype=tdtype)
    q = torch.empty_like(x)
    lib, h = model._lib, model.handle
    from surfh_b200 import _capi

    def application():
        model.fwadj_into(x, q)  # sharded: forward | all-reduce(y) | adjoint | all-reduce(maps)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
        torch.cuda.synchronize()
        if profile:
            model.profile(True)
        launches0 = model.own_launch_count()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if comm:
            comm.allreduce_max(ms)
        stages = model.profile_read() if profile else None
        if profile:
            model.profile(False)
        return float(ms.item()), model.own_launch_count() - launches0, clocks, stages

    # ---- device-resident applications (the headline `value`), stage events inside the timed region
    ms_total, launches, clocks, stages = timed(application, args.steps, args.warmup, profile=True)
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # per-rank kernel time of one application (sum of the stage events): load balance of the wavelength shards
    rank_ms = torch.zeros(world, dtype=torch.float64, device=dev)
    rank_ms[rank] = sum(s["ms"] for s in stages) / args.steps
    if comm:
        comm.allreduce_sum(rank_ms)
    rank_ms = [float(v) for v in rank_ms.cpu()]

    # ---- CG iterations (stage events on: the fused vector kernels show up as the `cg_fused` stage row)
    y = model.forward(x)
    g = torch.Generator(device=dev).manual_seed(1)
    y = y + 0.01 * y.pow(2).mean().sqrt() * torch.randn(y.shape, dtype=y.dtype, device=dev, generator=g)
    cg = fusion_CT.DeviceCG(model, y, 1.0, 5e3, comm=comm)
    cg.start(np.zeros(model.ishape), args.steps + args.warmup + 8)
    ms_cg, _, _, cg_stages = timed(lambda: cg.step(False), args.steps, args.warmup, profile=True)
    cg_iters = 1e3 * args.steps / ms_cg
    cg_row = next((st for st in cg_stages if st["stage"] == "cg_fused"), None)
    del cg

    # ---- the north-star solve: `--solve-iters` CG iterations from x0 = 0 (mu_reg = 5e3, refresh every 50,
    # criterion every 5th iteration from the CG state), wall clock around the whole call, result on the host
    solve = None
    if args.solve_iters > 0:
        try:
            quad = fusion_CT.QuadCriterion_MRS(1.0, y, model, 5e3, comm=comm)
            torch.cuda.synchronize()
            if comm:
                comm.barrier()
            t0 = time.perf_counter()
            res = quad.run_method("lcg", args.solve_iters, tolerance=1e-12, perf_crit=1, calc_crit=True, value_init=0)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if comm:
                comm.allreduce_max(dt)
            solve = {"iterations": int(res.nit), "seconds": float(dt.item()),
                     "iters_per_s": int(res.nit) / float(dt.item()),
                     "criterion_first": float(quad.L_crit_val[0]), "criterion_last": float(quad.L_crit_val[-1]),
                     "criterion_evaluations": len(quad.L_crit_val),
                     "criterion_forward_passes": len(quad.L_crit_val) - quad._solver()._state_evals,
                     "grad_norm_first": float(res.grad_norm[0]), "grad_norm_last": float(res.grad_norm[-1]),
                     "mu_reg": 5e3, "what": "QuadCriterion_MRS.run_method('lcg', n, perf_crit=1, calc_crit=True, "
                                            "value_init=0) as scripts/main_fusion.py:179-190 calls it; data = H x_true + 1 % noise"}
            del quad, res
        except Exception as exc:  # noqa: BLE001  the solve is an extra: never lose the bench line over it
            if comm is None:
                solve = {"error": f"{type(exc).__name__}: {exc}"}
            else:
                raise
    del y

    # ---- end to end through the LinOp API with host buffers: numpy maps in pinned memory ->
    # spectroSigRLSCT.fwadj (H2D, forward, [all-reduce], adjoint, [all-reduce], D2H) -> numpy maps
    e2e = None
    if not args.no_e2e:
        maps_pin = torch.from_numpy(cfg.maps.copy()).pin_memory()
        out_pin = torch.empty(model.ishape, dtype=torch.float64).pin_memory()
        maps_np, out_np = maps_pin.numpy(), out_pin.numpy()

        def host_timed(fn, steps):
            fn()
            torch.cuda.synchronize()
            if comm:
                comm.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if comm:
                comm.allreduce_max(dt)
            return steps / float(dt.item())

        steps_h = max(1, min(args.steps, 10))
        rate = host_timed(lambda: model.fwadj(maps_np, out=out_np), steps_h)
        assert np.isfinite(out_np).all()
        e2e = {"value": rate, "unit": UNIT, "h2d_bytes_per_step": int(8 * model.isize),
               "d2h_bytes_per_step": int(8 * model.isize), "steps": steps_h,
               "cg_iters_per_s": cg_iters, "cg_solve": solve,
               "path": "spectroSigRLSCT.fwadj (aljabr LinOp.fwadj): float64 numpy maps in pinned host memory -> "
                       "H2D -> forward -> adjoint -> D2H -> numpy maps; wall clock around the calls, max over ranks"}
        if world == 1:
            # the stricter two-call path of qmm (forward and adjoint as separate LinOp calls): the detector
            # vector crosses PCIe twice per application
            y_pin = torch.zeros(model.osize, dtype=torch.float64).pin_memory()
            n_out = int(lib.surfh_output_size(h))

            def two_calls():
                _capi.check(h, lib.surfh_forward_host(h, maps_pin.data_ptr(), y_pin.data_ptr()))
                _capi.check(h, lib.surfh_adjoint_host(h, y_pin.data_ptr(), out_pin.data_ptr(), model.mode_code))

            e2e["forward_then_adjoint_host_calls"] = {
                "value": host_timed(two_calls, max(1, min(args.steps, 5))), "unit": UNIT,
                "h2d_bytes_per_step": int(8 * (model.isize + n_out)), "d2h_bytes_per_step": int(8 * (n_out + model.isize))}

    if rank != 0:
        return

    # ---- roofline from the stage events of the timed region
    hbm_peak, peak_src = measured_peaks()
    fp_peak = measured_fp_peak(torch, tdtype)  # cuBLAS GEMM of the compute dtype, this GPU, this run
    traffic = ncu_traffic()
    contraction = model.contraction_info()
    stage_rows = []
    tot_ms = sum(s["ms"] for s in stages) or 1.0
    for s in stages:
        name = s["stage"]
        per_step_ms = s["ms"] / args.steps
        own = not (name.startswith("cufft_") or name == "memset")
        row = {"stage": name, "ms_per_step": per_step_ms, "share": s["ms"] / tot_ms, "own": own,
               "launches_per_step": s["launches"] / args.steps}
        sec = s["ms"] * 1e-3
        if name.startswith("spectral_gemm") and contraction["mode"] == "ozaki_i8":
            # int8-sliced error-free product on tcgen05: S digits per operand = S (S + 1) / 2 int8 products of the
            # contraction's shape; measured against the int8 tensor rate.  The stage time includes cutting the
            # per-call operand into digits.
            digits = contraction["digits"]
            products = digits * (digits + 1) // 2
            i8_peak, i8_src = measured_int8_peak()
            eq = s["flops"] / sec / 1e12 if sec > 0 else None
            executed = contraction.get("executed_fraction", 1.0)
            row.update(bound="tensor", unit="TFLOP/s", achieved=None if eq is None else eq * products * executed, peak=i8_peak,
                       peak_source=i8_src, digits=digits, int8_products=products, executed_fraction=executed,
                       fp64_equivalent_tflops=eq,
                       note="achieved / peak are int8 tensor operations actually executed (TOP/s; the all-zero digit "
                            "tiles of the line-spread function are skipped: executed_fraction); fp64_equivalent_tflops = "
                            "the contraction's own 2 M N K flops over the same time")
            if eq is not None:
                row["frac"] = row["achieved"] / i8_peak
                if fp_peak["tflops"]:
                    row["vs_cublas_gemm_of_dtype"] = eq / fp_peak["tflops"]
        elif name.startswith("spectral_gemm"):
            # dense contraction on the FP64 tensor pipe (DMMA) / FP32 FMA: flop-bound, not HBM-bound
            row.update(bound="tensor" if esz == 8 else "fma_fp32", unit="TFLOP/s",
                       achieved=s["flops"] / sec / 1e12 if sec > 0 else None, peak=fp_peak["tflops"])
            if row["achieved"] is not None and fp_peak["tflops"]:
                row["frac"] = row["achieved"] / fp_peak["tflops"]
        else:
            row.update(bound="hbm", unit="GB/s", achieved=s["bytes"] / sec / 1e9 if sec > 0 else None, peak=hbm_peak)
            if row["achieved"] is not None:
                row["frac"] = row["achieved"] / hbm_peak
            if name.startswith("chirpz_") and sec > 0:
                # the chirp-z FFT is bounded by the FP64 pipe and the shared-memory crossbar, not by HBM:
                # report its arithmetic rate beside the (low, by construction) HBM fraction
                row["tflops"] = s["flops"] / sec / 1e12
                if fp_peak["tflops"]:
                    row["frac_of_fp_peak"] = row["tflops"] / fp_peak["tflops"]
                row["limiter"] = "fp64 pipe + shared-memory crossbar (see profiles/)"
        if s["launches"]:
            row["algorithmic_bytes_per_launch"] = s["bytes"] / s["launches"]
        if name in traffic:
            row["traffic"] = traffic[name]
        stage_rows.append(row)
    if cg_row is not None and cg_row["ms"] > 0:
        sec = cg_row["ms"] * 1e-3
        stage_rows.append({"stage": "cg_fused", "ms_per_step": cg_row["ms"] / args.steps, "share": None, "own": True,
                           "launches_per_step": cg_row["launches"] / args.steps, "bound": "hbm", "unit": "GB/s",
                           "achieved": cg_row["bytes"] / sec / 1e9, "peak": hbm_peak,
                           "frac": cg_row["bytes"] / sec / 1e9 / hbm_peak,
                           "algorithmic_bytes_per_launch": cg_row["bytes"] / max(1, cg_row["launches"]),
                           "note": "per CG iteration, outside the application's timed region; the vectors "
                                   "(K*N^2 reals) are L2-resident, the cost is launch latency"})
    own_rows = [r for r in stage_rows if r["own"] and r.get("achieved") and r.get("share") is not None]
    top = max(own_rows, key=lambda r: r["ms_per_step"]) if own_rows else None
    roofline = None
    if top:
        roofline = {"kernel": top["stage"], "bound": top["bound"] if top["bound"] in ("hbm", "tensor") else "hbm",
                    "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top.get("frac"),
                    "traffic": top.get("traffic"),
                    "peak_source": top.get("peak_source") or (peak_src if top["unit"] == "GB/s" else fp_peak["source"]),
                    "share_of_step": top["share"],
                    "fp_peak": {"tflops": fp_peak["tflops"], "source": fp_peak["source"]},
                    "hbm_peak": {"gbs": hbm_peak, "source": peak_src}}
        for k in ("tflops", "frac_of_fp_peak", "limiter", "algorithmic_bytes_per_launch"):
            if k in top:
                roofline[k] = top[k]
        # every stage in one compact list, inside the object the driver keeps
        roofline["contraction"] = contraction
        roofline["stages"] = [{"stage": r["stage"], "ms": round(r["ms_per_step"], 4), "bound": r.get("bound"),
                               "frac": None if r.get("frac") is None else round(r["frac"], 4),
                               "frac_of_fp_peak": None if r.get("frac_of_fp_peak") is None else round(r["frac_of_fp_peak"], 4)}
                              for r in stage_rows]

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_sample(cfg, 1, 0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if esz == 8 else "f32", "data": "synthetic", "config": workload_description(cfg, args.dtype, args),
        "cg_iters_per_s": cg_iters, "cg_solve": solve, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "stages": stage_rows, "cpu_baseline": cpu,
        "per_rank_kernel_ms": rank_ms, "exchange_and_idle_ms": ms_step - max(rank_ms),
        "setup_s": setup_s, "lambda_range_rank0": lam_range, "workspace_gb": model.workspace_bytes() / 1e9,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, warnings) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()

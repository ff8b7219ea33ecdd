#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the REFERENCE'S OWN CODE.

Run here (the container that has /root/reference):   python oracle/make_golden.py

What it does:
  1. copies /root/reference/surfh to a scratch directory (the reference tree is read-only and
     nothing of it is copied into this repository);
  2. in that scratch copy only, replaces the dataclass default `origin: Coord = Coord(0, 0)`
     (surfh/Models/instru.py:280), which Python >= 3.11 rejects, by a default_factory -- no
     arithmetic is touched;
  3. builds the reference's Cython kernels (surfh/ToolsDir/cythons_files.pyx) with gcc;
  4. registers stand-ins for the packages that are not installed (udft, aljabr, qmm restated
     in oracle/surfh_oracle/thirdparty.py; jax.numpy -> numpy, so the "JAX" kernels run in
     fp64; astropy / matplotlib / xarray / progressbar as empty shells) and the module aliases
     of the reference's unfinished rename (surfh.DottestModels.*, surfh.Models.slicer_new);
  5. replaces the two broadcast-sum spectral blurs by the identical einsum (a [L',L,a,b]
     temporary of several GB otherwise);
  6. runs `spectroSigRLSCT.forward / .adjoint`, the geometry helpers and the
     `QuadCriterion_MRS` criterion on the seeded inputs of surfh_b200.synthetic and stores the
     outputs (full arrays for the mini cases, strided samples + norms for full-size C1).

The fixtures are what pins the oracle (tests/test_oracle_golden.py) and, through it, the
CUDA path.  /root/reference is never read at test / smoke / bench time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = "/root/reference"
GOLDEN = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from surfh_oracle import thirdparty  # noqa: E402
from surfh_b200 import synthetic  # noqa: E402


def prepare_scratch() -> str:
    scratch = os.environ.get("SURFH_REF_SCRATCH") or os.path.join(tempfile.gettempdir(), "surfh_ref_scratch")
    pkg = os.path.join(scratch, "pkg")
    if not os.path.exists(os.path.join(pkg, "surfh", "ToolsDir", "BUILT")):
        shutil.rmtree(scratch, ignore_errors=True)
        os.makedirs(pkg)
        shutil.copytree(os.path.join(REFERENCE, "surfh"), os.path.join(pkg, "surfh"))
        path = os.path.join(pkg, "surfh", "Models", "instru.py")
        src = open(path).read()
        src = src.replace("from dataclasses import dataclass", "from dataclasses import dataclass, field")
        src = src.replace("origin: Coord = Coord(0, 0)",
                          "origin: Coord = field(default_factory=lambda: Coord(0, 0))")
        open(path, "w").write(src)
        setup = (
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\nimport numpy\n"
            "ext = Extension('surfh.ToolsDir.cythons_files', ['surfh/ToolsDir/cythons_files.pyx'],\n"
            "    extra_compile_args=['-O3', '-fopenmp'], extra_link_args=['-fopenmp'],\n"
            "    include_dirs=[numpy.get_include()])\n"
            "setup(ext_modules=cythonize([ext], compiler_directives={'binding': True, 'language_level': 3}))\n"
        )
        open(os.path.join(pkg, "setup_cy.py"), "w").write(setup)
        env = dict(os.environ, CC="/usr/bin/gcc", LDSHARED="/usr/bin/gcc -shared")
        subprocess.check_call([sys.executable, "setup_cy.py", "build_ext", "--inplace"], cwd=pkg, env=env,
                              stdout=subprocess.DEVNULL)
        open(os.path.join(pkg, "surfh", "ToolsDir", "BUILT"), "w").write("ok\n")
    if not os.path.exists(os.path.join(pkg, "surfh", "ToolsDir", "BUILT_SHEPARD")):
        setup = (
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\nimport numpy\n"
            "ext = Extension('surfh.ToolsDir.shepard_interpolation', ['surfh/ToolsDir/shepard_interpolation.pyx'],\n"
            "    extra_compile_args=['-O2'], include_dirs=[numpy.get_include()])\n"
            "setup(ext_modules=cythonize([ext], compiler_directives={'binding': True, 'language_level': 3}))\n"
        )
        open(os.path.join(pkg, "setup_shepard.py"), "w").write(setup)
        env = dict(os.environ, CC="/usr/bin/gcc", LDSHARED="/usr/bin/gcc -shared")
        subprocess.check_call([sys.executable, "setup_shepard.py", "build_ext", "--inplace"], cwd=pkg, env=env,
                              stdout=subprocess.DEVNULL)
        open(os.path.join(pkg, "surfh", "ToolsDir", "BUILT_SHEPARD"), "w").write("ok\n")
    return pkg


def shepard_inputs(seed=3, n_lambda=64, n_alpha_det=26, na=19):
    """One synthetic slit: detector samples on a sheared, slightly curved (alpha, lambda) lattice with a few
    NaN-removed holes, to be interpolated onto the model's regular [n_lambda, na] grid (the geometry of
    distorsion_correction.py:150-176: alpha_res, lambda_res from the grid extents, p = 2, alpha = 2, cutoff = 2)."""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(n_lambda * 2), np.arange(n_alpha_det), indexing="ij")
    lam = 5.0 + 0.0004 * ii + 0.00003 * jj + 2e-7 * jj ** 2
    alpha = -1.6 + 0.13 * jj + 0.0008 * ii + 0.01 * rng.standard_normal(ii.shape)
    val = np.sin(3.0 * alpha) + 0.2 * np.cos(900.0 * (lam - 5.0)) + 0.05 * rng.standard_normal(ii.shape)
    keep = rng.random(ii.shape) > 0.03
    alpha, lam, val = alpha[keep], lam[keep], val[keep]
    chan_wavelength = np.linspace(lam.min(), lam.max(), n_lambda)
    grid_alpha = np.linspace(alpha.min(), alpha.max(), na)
    alpha_mesh, lambda_mesh = np.meshgrid(grid_alpha, chan_wavelength)
    alpha_res = (grid_alpha.max() - grid_alpha.min()) / alpha_mesh.shape[1]
    lambda_res = (chan_wavelength.max() - chan_wavelength.min()) / lambda_mesh.shape[0]
    return dict(alpha=alpha, lam=lam, val=val, alpha_mesh=alpha_mesh, lambda_mesh=lambda_mesh,
                alpha_res=alpha_res, lambda_res=lambda_res)


def run_shepard():
    """The reference's compiled shepard_interpolation.pyx on the synthetic slit, through the same float32 casts
    as perform_shepard_interpolation (distorsion_correction.py:89-94)."""
    from surfh.ToolsDir import shepard_interpolation
    d = shepard_inputs()
    f = np.float32
    out = {}
    for tag, (p, a_exp, cut) in {"p2": (2, 2.0, 2), "p15": (1.5, 1.0, 3)}.items():
        out[tag] = np.asarray(shepard_interpolation.exponential_modified_shepard(
            d["alpha"].astype(f), d["lam"].astype(f), d["val"].astype(f), d["alpha_mesh"].astype(f),
            d["lambda_mesh"].astype(f), p=p, alpha=a_exp, pixel_cutoff=cut, alpha_res=d["alpha_res"],
            lambda_res=d["lambda_res"]))
    path = os.path.join(GOLDEN, "shepard.npz")
    np.savez_compressed(path, **out)
    print(f"shepard: {out['p2'].shape} grid from {len(d['val'])} samples, |out|={np.linalg.norm(out['p2']):.6e} "
          f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def _shell(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package for "from x.y import z"
    sys.modules[name] = mod
    return mod


def install_stubs():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    _shell("udft", ir2fr=thirdparty.ir2fr, rdft2=thirdparty.rdft2, irdftn=thirdparty.irdftn,
           laplacian=_Anything(), diff_ir=_Anything(), dft2=None, idft2=None)
    _shell("aljabr", LinOp=thirdparty.LinOp, dottest=thirdparty.dottest)
    _shell("qmm", QuadObjective=thirdparty.QuadObjective, lcg=thirdparty.lcg, mmmg=None,
           Objective=object, Huber=_Anything, HebertLeahy=_Anything)

    def jit(fn=None, **kwargs):
        return fn if fn is not None else (lambda f: f)

    jnp = _shell("jax.numpy", **{k: getattr(np, k) for k in
                                 ("sum", "expand_dims", "newaxis", "concatenate", "multiply", "zeros",
                                  "ones", "fft", "array", "asarray")})
    _shell("jax", jit=jit, numpy=jnp, lax=_shell("jax.lax"))
    for name in ("xarray", "matplotlib", "matplotlib.pyplot", "astropy", "astropy.units", "astropy.io",
                 "astropy.io.fits", "astropy.coordinates", "progressbar", "pytest", "einops",
                 "sklearn", "sklearn.neighbors"):
        if name not in sys.modules:
            _shell(name, Angle=_Anything, fits=_Anything(), rearrange=_Anything(), KDTree=_Anything)
    try:
        import loguru  # noqa: F401
    except ImportError:
        _shell("loguru", logger=_Anything())
    import scipy
    if not hasattr(scipy, "misc"):
        scipy.misc = _shell("scipy.misc")


def import_reference(pkg):
    sys.path.insert(0, pkg)
    install_stubs()
    import surfh.Models.slicer as slicer
    sys.modules["surfh.Models.slicer_new"] = slicer
    import surfh.Models as models_pkg
    models_pkg.slicer_new = slicer
    from surfh.ToolsDir import jax_utils
    jax_utils.wblur_subSampling = lambda arr, wpsf: np.einsum("lab,mlb->ma", arr, wpsf, optimize=True)
    jax_utils.wblur_t = lambda arr, wpsf: np.einsum("mab,mlb->lab", arr, wpsf, optimize=True)
    import surfh.Models.spectroModelChannel as chan_mod
    dott = _shell("surfh.DottestModels", MCMO_SigRLSCT_Channel_Model=chan_mod)
    import surfh
    surfh.DottestModels = dott
    import surfh.Models.spectroModel as model_mod
    dott.MCMO_SigRLSCT_Model = model_mod
    from surfh.Models import instru as ref_instru
    return model_mod, ref_instru


def to_reference_objects(ref_instru, cfg):
    instrs = [ref_instru.IFU(
        fov=ref_instru.FOV(i.fov.alpha_width, i.fov.beta_width,
                           origin=ref_instru.Coord(i.fov.origin.alpha, i.fov.origin.beta), angle=i.fov.angle),
        det_pix_size=i.det_pix_size, n_slit=i.n_slit,
        w_blur=ref_instru.SpectralBlur(i.w_blur.grating_resolution), pce=None,
        wavel_axis=i.wavel_axis, name=i.name) for i in cfg.instrs]
    pointings = [ref_instru.CoordList([ref_instru.Coord(c.alpha, c.beta) for c in pl]) for pl in cfg.pointings]
    return instrs, pointings


def build_reference_model(model_mod, ref_instru, cfg):
    instrs, pointings = to_reference_objects(ref_instru, cfg)
    return model_mod.spectroSigRLSCT(cfg.sotf(), cfg.templates, cfg.alpha_axis, cfg.beta_axis,
                                     cfg.wavelength_axis, instrs, cfg.step_degree, pointings)


def geometry_record(model):
    rec = {}
    for c, ch in enumerate(model.channels):
        sl = [ch.slicer.get_slit_slices(s) for s in range(ch.instr.n_slit)]
        rec[f"b{c}_slices"] = np.array([[a.start, a.stop, b.start, b.stop] for a, b in sl])
        w = [ch.slicer.get_slit_weights(s, sl[s])[0, 0, :] for s in range(ch.instr.n_slit)]
        wmax = max(len(v) for v in w)
        rec[f"b{c}_weights"] = np.array([np.pad(v, (0, wmax - len(v)), constant_values=-1) for v in w])
        rec[f"b{c}_oshape"] = np.array(ch.oshape)
        rec[f"b{c}_wslice"] = np.array([ch.wslice.start, ch.wslice.stop])
        rec[f"b{c}_local_shape"] = np.array(ch.local_im_shape)
        rec[f"b{c}_local_alpha"] = np.asarray(ch.local_alpha_axis)
        rec[f"b{c}_local_beta"] = np.asarray(ch.local_beta_axis)
        rec[f"b{c}_srf"] = np.array(ch.srf)
        rec[f"b{c}_nbw"] = np.array(ch.slicer.npix_slit_beta_width)
        rec[f"b{c}_wpsf_sample"] = np.asarray(ch.wpsf)[::7, ::5, :].copy()
        rec[f"b{c}_wpsf_sum"] = np.asarray(ch.wpsf).sum(axis=(1, 2))
    rec["idx"] = np.asarray(model._idx)
    return rec


def run_cg_case(model_mod, ref_instru, cfg, n_iter: int, mu_reg: float):
    """BASELINE.json's configs[1]: the reference's own QuadCriterion_MRS.run_method('lcg', n_iter,
    perf_crit=1, calc_crit=True, value_init=0) (scripts/main_fusion.py:179-190) on the reference operator,
    driven by the restated qmm.lcg.  Stores the gradient-norm history, the criterion trace and a strided
    sample of the final maps."""
    import contextlib
    import io
    from surfh.Simulation import fusion_CT
    model = build_reference_model(model_mod, ref_instru, cfg)
    fwd = np.asarray(model.forward(cfg.maps))
    y = fwd + 0.01 * np.sqrt(np.mean(fwd ** 2)) * np.random.default_rng(99).standard_normal(fwd.shape)
    crit = fusion_CT.QuadCriterion_MRS(1, np.copy(y), model, mu_reg, printing=False, gradient="separated")
    rec = {"idx": np.asarray(model._idx), "mu_reg": np.array(mu_reg), "n_iter": np.array(n_iter),
           "fwd_norm": np.array(np.linalg.norm(fwd)), "crit_at_zero": np.array(crit.get_crit_val(np.zeros(model.ishape)))}
    with contextlib.redirect_stdout(io.StringIO()):
        res = crit.run_method("lcg", n_iter, perf_crit=1, calc_crit=True, value_init=0)
    x = np.asarray(res.x).reshape(model.ishape)
    rec["cg_x_sample"] = x[:, ::3, ::3].copy()
    rec["cg_x_norm"] = np.array(np.linalg.norm(x))
    rec["cg_grad_norm"] = np.asarray(res.grad_norm)
    rec["cg_crit"] = np.asarray(crit.L_crit_val)
    rec["crit_final"] = np.array(crit.get_crit_val(x))
    return rec


CG_CASES = {
    # name: (config factory, iterations, mu_reg)
    "c2_cg50": (lambda: synthetic.baseline_config("c2"), 50, 5e3),
}


def run_case(model_mod, ref_instru, cfg, full: bool, with_cg: bool = False):
    model = build_reference_model(model_mod, ref_instru, cfg)
    rng = np.random.default_rng(1234)
    y_probe = rng.standard_normal(model.oshape[0])
    fwd = np.asarray(model.forward(cfg.maps))
    adj = np.asarray(model.adjoint(y_probe))
    rec = geometry_record(model)
    rec["ishape"] = np.array(model.ishape)
    rec["fwd_norm"] = np.array(np.linalg.norm(fwd))
    rec["adj_norm"] = np.array(np.linalg.norm(adj))
    if full:
        rec["fwd"] = fwd
        rec["adj"] = adj
        ch0 = model.channels[0]
        # the blurred cube (T then C) and one gridded plane, for stage-level checks
        from surfh.ToolsDir import jax_utils
        cube = jax_utils.lmm_maps2cube(cfg.maps, cfg.templates).reshape(model.cube_shape) if model.lmm else cfg.maps
        blurred = np.asarray(jax_utils.idft(jax_utils.dft(cube) * model.sotf, (model.ishape[1], model.ishape[2])))
        rec["blurred_sample"] = blurred[::5, ::3, ::3].copy()
        rec["gridded0"] = np.asarray(ch0.gridding(blurred[ch0.wslice], ch0.pointings[0]))[::4].copy()
    else:
        rec["fwd_stride"] = np.array(997)
        rec["fwd_sample"] = fwd[::997].copy()
        rec["adj_sample"] = adj[:, ::7, ::7].copy()
    if with_cg:
        from surfh.Simulation import fusion_CT
        y = fwd + 0.01 * np.sqrt(np.mean(fwd ** 2)) * np.random.default_rng(99).standard_normal(fwd.shape)
        crit = fusion_CT.QuadCriterion_MRS(1, y, model, 5.0, printing=False, gradient="separated")
        rec["crit_at_maps"] = np.array(crit.get_crit_val(cfg.maps))
        rec["diff_r"] = crit.npdiff_r.forward(cfg.maps)[:, ::9, ::9].copy()
        rec["diff_c_t"] = crit.npdiff_c.adjoint(cfg.maps)[:, ::9, ::9].copy()
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            res = crit.run_method("lcg", 6, perf_crit=1, calc_crit=True, value_init=0)
        rec["cg_x"] = np.asarray(res.x)
        rec["cg_grad_norm"] = np.asarray(res.grad_norm)
        rec["cg_crit"] = np.asarray(crit.L_crit_val)
    return rec


CASES = {
    # name: (config factory, store full arrays?, run the CG criterion?)
    "mini_1band_1p": (lambda: synthetic.mini_config(1, 1), True, True),
    "mini_2band_4p": (lambda: synthetic.mini_config(2, 4), True, False),
    "mini_2band_2p_cube": (lambda: synthetic.mini_config(2, 2, lmm=False, n_pix=96), True, False),
    "c1_band1a": (lambda: synthetic.baseline_config("c1"), False, False),
    "band2a_4p": (lambda: synthetic.mrs_config(["2a"], 251, 4, 4, seed=3, name="band2a_4p"), False, False),
    # channels 3 and 4 at the north-star map size (srf 9 / 10, nb 16 / 26; the 275x319 local grid of
    # channel 4 is what forces N = 501), 4 dithers, and BASELINE.json's configs[2] (C3) as a whole
    "band3a_n501_4p": (lambda: synthetic.mrs_config(["3a"], 501, 4, 4, seed=11, name="band3a_n501_4p"), False, False),
    "band4a_n501_4p": (lambda: synthetic.mrs_config(["4a"], 501, 4, 4, seed=12, name="band4a_n501_4p"), False, False),
    "c3": (lambda: synthetic.baseline_config("c3"), False, False),
}


def blind_cases():
    """Single-wavelength MRSBlurred inputs (surfh/Models/spectro_blind.py): (name, cfg, wavelength index)."""
    mini = synthetic.mini_config(1, 2, lmm=False, n_pix=96, n_slit_a=11)  # MRSBlurred needs n_slit >= nb
    full = synthetic.mrs_config(["1c"], 301, 0, 4, seed=5, name="blind_1c", lmm=False,
                                wavel=np.array([6.9, 7.0, 7.1]))
    return [("blind_mini_2p", mini, 7), ("blind_1c_4p", full, 1)]


def run_blind(ref_instru, name, cfg, l_idx):
    import surfh.Models.spectro_blind as blind_mod
    instrs, pointings = to_reference_objects(ref_instru, cfg)
    sotf = cfg.sotf()[l_idx]
    model = blind_mod.MRSBlurred(sotf, cfg.alpha_axis, cfg.beta_axis, instrs[0], cfg.step_degree, pointings[0])
    x = cfg.maps[l_idx]
    fwd = np.asarray(model.forward(x))
    v = np.random.default_rng(1234).standard_normal(fwd.shape[0])
    adj = np.asarray(model.adjoint(v))
    sl = [model.get_slit_slices(s) for s in range(instrs[0].n_slit)]
    rec = {"fwd": fwd, "adj": adj, "l_idx": np.array(l_idx), "slices_shape": np.array(model.slices_shape),
           "slices": np.array([[a.start, a.stop, b.start, b.stop] for a, b in sl]),
           "weights": np.array([model.get_slit_weights(s, sl[s])[0, 0, :] for s in range(instrs[0].n_slit)]),
           "jansky": np.asarray(model.real_data_janskySR_to_jansky(fwd.copy()))}
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: oshape={fwd.shape[0]} |fwd|={np.linalg.norm(fwd):.6e} |adj|={np.linalg.norm(adj):.6e} "
          f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def main(names=None):
    pkg = prepare_scratch()
    model_mod, ref_instru = import_reference(pkg)
    os.makedirs(GOLDEN, exist_ok=True)
    if not names or "shepard" in names:
        run_shepard()
    for name, cfg, l_idx in blind_cases():
        if names and name not in names:
            continue
        run_blind(ref_instru, name, cfg, l_idx)
    for name, (factory, n_iter, mu_reg) in CG_CASES.items():
        if names and name not in names:
            continue
        rec = run_cg_case(model_mod, ref_instru, factory(), n_iter, mu_reg)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: {n_iter} iterations, J0={float(rec['crit_at_zero']):.6e} J={float(rec['crit_final']):.6e} "
              f"-> {os.path.getsize(path) / 1e3:.0f} kB")
    for name, (factory, full, with_cg) in CASES.items():
        if names and name not in names:
            continue
        cfg = factory()
        rec = run_case(model_mod, ref_instru, cfg, full, with_cg)
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: oshape={rec['idx'][-1]} |fwd|={float(rec['fwd_norm']):.6e} "
              f"|adj|={float(rec['adj_norm']):.6e} -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main(sys.argv[1:])

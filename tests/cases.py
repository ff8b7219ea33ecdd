"""Seeded configurations shared by the oracle-vs-golden and CUDA-vs-oracle tests.  The names match
the fixtures written by oracle/make_golden.py."""
from surfh_b200 import synthetic

CASES = {
    "mini_1band_1p": lambda: synthetic.mini_config(1, 1),
    "mini_2band_4p": lambda: synthetic.mini_config(2, 4),
    "mini_2band_2p_cube": lambda: synthetic.mini_config(2, 2, lmm=False, n_pix=96),
    "c1_band1a": lambda: synthetic.baseline_config("c1"),
    "band2a_4p": lambda: synthetic.mrs_config(["2a"], 251, 4, 4, seed=3, name="band2a_4p"),
    # channels 3 and 4 at the north-star map size (srf 9 / 10, the 275 x 319 local grid that forces N = 501)
    "band3a_n501_4p": lambda: synthetic.mrs_config(["3a"], 501, 4, 4, seed=11, name="band3a_n501_4p"),
    "band4a_n501_4p": lambda: synthetic.mrs_config(["4a"], 501, 4, 4, seed=12, name="band4a_n501_4p"),
    "c3": lambda: synthetic.baseline_config("c3"),   # BASELINE.json configs[2]: 1A, 2A, 3A, 4A at N = 501
}
MINI = ["mini_1band_1p", "mini_2band_4p", "mini_2band_2p_cube"]
FULL = ["c1_band1a", "band2a_4p", "band3a_n501_4p", "band4a_n501_4p"]


def band_subconfig(cfg, band: int, pointing=None, margin: int = 3):
    """Arguments of a ONE-band model that produces exactly band `band`'s block of the multi-band model `cfg`
    describes: the cube axis is cut to the band's wavelength window (+ `margin` planes each side, so that
    `wslice` selects the same planes), templates and PSF stamps are cut alike, the maps are shared.
    `pointing`: keep only that dither (the block of pointing p is [p] of the band's [P, S, L', na] block).
    Used to put the CPU oracle beside a full-size multi-band CUDA model at a cost of seconds per band."""
    import numpy as np
    from surfh_b200 import instru
    ifu = cfg.instrs[band]
    ws = ifu.wslice(cfg.wavelength_axis, 0.1)
    lo, hi = max(0, ws.start - margin), min(len(cfg.wavelength_axis), ws.stop + margin)
    axis = cfg.wavelength_axis[lo:hi]
    sub = ifu.wslice(axis, 0.1)
    assert (sub.start + lo, sub.stop + lo) == (ws.start, ws.stop), "sub-axis selects different planes"
    pts = cfg.pointings[band] if pointing is None else instru.CoordList([cfg.pointings[band][pointing]])
    return dict(sotf=synthetic.ir2fr(cfg.psf[lo:hi], cfg.imshape),
                templates=None if cfg.templates is None else np.ascontiguousarray(cfg.templates[:, lo:hi]),
                alpha_axis=cfg.alpha_axis, beta_axis=cfg.beta_axis, wavelength_axis=axis, instrs=[ifu],
                step_degree=cfg.step_degree, pointings=[pts])

# MRSBlurred (spectro_blind.py) single-wavelength fixtures: name -> (config factory, wavelength index)
BLIND = {
    "blind_mini_2p": (lambda: synthetic.mini_config(1, 2, lmm=False, n_pix=96, n_slit_a=11), 7),
    "blind_1c_4p": (lambda: synthetic.mrs_config(["1c"], 301, 0, 4, seed=5, name="blind_1c", lmm=False,
                                                 wavel=__import__("numpy").array([6.9, 7.0, 7.1])), 1),
}

"""Seeded configurations shared by the oracle-vs-golden and CUDA-vs-oracle tests.  The names match
the fixtures written by oracle/make_golden.py."""
from surfh_b200 import synthetic

CASES = {
    "mini_1band_1p": lambda: synthetic.mini_config(1, 1),
    "mini_2band_4p": lambda: synthetic.mini_config(2, 4),
    "mini_2band_2p_cube": lambda: synthetic.mini_config(2, 2, lmm=False, n_pix=96),
    "c1_band1a": lambda: synthetic.baseline_config("c1"),
    "band2a_4p": lambda: synthetic.mrs_config(["2a"], 251, 4, 4, seed=3, name="band2a_4p"),
}
MINI = ["mini_1band_1p", "mini_2band_4p", "mini_2band_2p_cube"]
FULL = ["c1_band1a", "band2a_4p"]

# MRSBlurred (spectro_blind.py) single-wavelength fixtures: name -> (config factory, wavelength index)
BLIND = {
    "blind_mini_2p": (lambda: synthetic.mini_config(1, 2, lmm=False, n_pix=96, n_slit_a=11), 7),
    "blind_1c_4p": (lambda: synthetic.mrs_config(["1c"], 301, 0, 4, seed=5, name="blind_1c", lmm=False,
                                                 wavel=__import__("numpy").array([6.9, 7.0, 7.1])), 1),
}

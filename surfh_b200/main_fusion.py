"""The reference's fusion driver flow (`scripts/main_fusion.py`) on top of the CUDA operator.

Same functions, same argument meaning, same files read and written, with the three imports the
reference script makes for this path swapped:

    surfh.DottestModels.MCMO_SigRLSCT_Model.spectroSigRLSCT  ->  surfh_b200.model.spectroSigRLSCT
    surfh.Models.instru                                      ->  surfh_b200.instru
    surfh.Simulation.fusion_CT.QuadCriterion_MRS             ->  surfh_b200.fusion_CT.QuadCriterion_MRS

Reference interface mirrored (paths relative to /root/reference):
    load_data               scripts/main_fusion.py:30-63     FITS slices [L', S, na] -> [S, L', na] per exposure
    initialize_parameters   scripts/main_fusion.py:65-77     Fusion/ directory layout
    load_simulation_data    scripts/main_fusion.py:79-104    Templates/*.npy, PSF/*.npy -> axes, templates, OTF
    create_instruments      scripts/main_fusion.py:106-137   the 12-band table
    create_model            scripts/main_fusion.py:139-160
    reconstruction_method   scripts/main_fusion.py:164-204   run_method -> res_x.npy, res_cube.npy, criterion.npy

Everything here is host-side glue; the arithmetic is `spectroSigRLSCT` / `QuadCriterion_MRS`.  The
reference reads its FITS files with astropy, which this image does not have: `read_fits_primary` is a
reader for the one thing the driver needs (the primary HDU's header cards and image).
"""
from __future__ import annotations

import os
import pathlib
from typing import Dict, List, Sequence

import numpy as np

from . import instru, synthetic
from .fusion_CT import QuadCriterion_MRS
from .model import spectroSigRLSCT

LIST_CHAN = ["1a", "1b", "1c", "2a", "2b", "2c", "3a", "3b", "3c", "4a", "4b", "4c"]

# (n_slit, n_det, na) of every band's exposure files, scripts/main_fusion.py:34-39
DATASHAPE = {
    "1a": (21, 1050, 19), "1b": (21, 1213, 19), "1c": (21, 1400, 19),
    "2a": (17, 970, 24), "2b": (17, 1124, 24), "2c": (17, 1300, 24),
    "3a": (16, 769, 24), "3b": (16, 892, 24), "3c": (16, 1028, 24),
    "4a": (12, 542, 27), "4b": (12, 632, 27), "4c": (12, 717, 27),
}

PSF_FILE = "psfs_pixscale0.025_npix_501_fov12.525_chan_1ABC_2ABC_3ABC_4ABC_SS4.npy"
_FITS_BLOCK = 2880
_BITPIX = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


# ----------------------------------------------------------------------------- minimal FITS I/O
def _parse_card_value(text: str):
    text = text.split("/")[0].strip() if not text.strip().startswith("'") else text.strip()
    if text.startswith("'"):
        end = text.find("'", 1)
        return text[1:end].rstrip()
    if text in ("T", "F"):
        return text == "T"
    try:
        return int(text)
    except ValueError:
        return float(text.replace("D", "E"))


def read_fits_primary(path: str):
    """(header dict, data array) of the primary HDU of a FITS file (what `fits.open(path)[0]` gives)."""
    with open(path, "rb") as f:
        raw = f.read()
    header: Dict[str, object] = {}
    pos, done = 0, False
    while not done:
        block = raw[pos: pos + _FITS_BLOCK]
        if len(block) < _FITS_BLOCK:
            raise ValueError(f"{path}: truncated FITS header")
        pos += _FITS_BLOCK
        for i in range(0, _FITS_BLOCK, 80):
            card = block[i: i + 80].decode("ascii", errors="replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= " and key:
                header[key] = _parse_card_value(card[10:])
    if not header.get("SIMPLE", False):
        raise ValueError(f"{path}: not a standard FITS file")
    naxis = int(header.get("NAXIS", 0))
    if naxis == 0:
        return header, None
    shape = tuple(int(header[f"NAXIS{k}"]) for k in range(naxis, 0, -1))
    dt = np.dtype(_BITPIX[int(header["BITPIX"])])
    count = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=count, offset=pos).reshape(shape)
    data = data.astype(dt.newbyteorder("="))
    if "BSCALE" in header or "BZERO" in header:
        data = data * float(header.get("BSCALE", 1.0)) + float(header.get("BZERO", 0.0))
    return header, data


def write_fits_primary(path: str, data: np.ndarray, cards: Dict[str, object]) -> None:
    """Write a single-HDU FITS image (tests and synthetic Fusion/ trees)."""
    data = np.asarray(data)
    bitpix = {np.dtype("float64"): -64, np.dtype("float32"): -32, np.dtype("int32"): 32, np.dtype("int16"): 16}[data.dtype]

    def card(key, value, quote=False):
        if isinstance(value, bool):
            v = "T" if value else "F"
            return f"{key:<8}= {v:>20}".ljust(80)
        if quote or isinstance(value, str):
            return f"{key:<8}= '{value}'".ljust(80)
        return f"{key:<8}= {value!r:>20}".ljust(80)

    lines = [card("SIMPLE", True), card("BITPIX", bitpix), card("NAXIS", data.ndim)]
    for k, n in enumerate(reversed(data.shape), start=1):
        lines.append(card(f"NAXIS{k}", int(n)))
    for key, value in cards.items():
        lines.append(card(key, value))
    lines.append("END".ljust(80))
    head = "".join(lines).encode("ascii")
    head += b" " * (-len(head) % _FITS_BLOCK)
    body = data.astype(data.dtype.newbyteorder(">")).tobytes()
    body += b"\0" * (-len(body) % _FITS_BLOCK)
    with open(path, "wb") as f:
        f.write(head + body)


# ------------------------------------------------------------------------------- the driver flow
def load_data(list_chan: Sequence[str], save_filter_corrected_dir: str):
    """Exposure files of every channel, `[L', S, na]` on disk -> `[S, L', na]` (main_fusion.py:30-63)."""
    data_dict = {"data": {}, "target": {}, "rotation": {}}
    for chan in list_chan:
        data_dict["data"][chan] = []
        data_dict["target"][chan] = []
        data_dict["rotation"][chan] = 0.0
    for file in sorted(os.listdir(save_filter_corrected_dir)):
        for chan in list_chan:
            if chan in file:
                n_slit, n_det, na = DATASHAPE[chan]
                header, data = read_fits_primary(os.path.join(save_filter_corrected_dir, file))
                ndata = np.asarray(data, dtype=np.float64).reshape(n_det, n_slit, na).transpose(1, 0, 2)
                data_dict["data"][chan].append(ndata)
                data_dict["target"][chan].append((header["TARG_RA"], header["TARG_DEC"]))
                data_dict["rotation"][chan] = header["PA_V3"]
    return data_dict


def initialize_parameters(fusion_dir_path: str):
    paths = {
        "psf_dir": os.path.join(fusion_dir_path, "PSF/"),
        "template_dir": os.path.join(fusion_dir_path, "Templates/"),
        "save_filter_corrected_dir": os.path.join(fusion_dir_path, "Filtered_slices/"),
        "result_path": os.path.join(fusion_dir_path, "Results/"),
        "mask_path": os.path.join(fusion_dir_path, "Masks/"),
    }
    step = 0.025  # arcsec
    return paths, step, step / 3600.0  # Angle(step, u.arcsec).degree


def load_simulation_data(paths, step, step_angle, Npix, nTemplates):
    imshape = (Npix, Npix)
    ax = np.arange(imshape[0]) * step_angle
    origin_alpha_axis = ax - np.mean(ax)
    origin_beta_axis = np.arange(imshape[1]) * step_angle - np.mean(np.arange(imshape[1]) * step_angle)
    if nTemplates not in (4, 6):
        raise NameError("No corresponding Templates name")
    wavel_file = f"wavel_axis_orion_1ABC_2ABC_3ABC_4ABC_{nTemplates}_templates_SS4.npy"
    templates_file = f"nmf_orion_1ABC_2ABC_3ABC_4ABC_{nTemplates}_templates_SS4.npy"
    wavel_axis = np.load(os.path.join(paths["template_dir"], wavel_file))
    templates = np.load(os.path.join(paths["template_dir"], templates_file))
    spsf = np.load(os.path.join(paths["psf_dir"], PSF_FILE))
    sotf = synthetic.ir2fr(spsf, imshape)  # udft.ir2fr(spsf, imshape)
    templates = templates / 10e3
    return origin_alpha_axis, origin_beta_axis, wavel_axis, templates, sotf


def create_instruments(data_dict, list_chan: Sequence[str] = LIST_CHAN):
    instruments = {}
    for chan in list_chan:
        n_slit, r_min, r_max, det_pix_size, fov_x, fov_y = synthetic.MRS_BANDS[chan]
        instruments[chan] = instru.IFU(
            fov=instru.FOV(fov_x / 3600, fov_y / 3600, origin=instru.Coord(0, 0), angle=-data_dict["rotation"][chan]),
            det_pix_size=det_pix_size, n_slit=n_slit, w_blur=instru.SpectralBlur(float(np.mean([r_min, r_max]))),
            pce=None, wavel_axis=synthetic.mrs_detector_axis(chan), name=chan.upper())
    return instruments


def create_model(sotf, templates, origin_alpha_axis, origin_beta_axis, wavel_axis, instruments, step_angle, data_dict,
                 **model_kwargs):
    main_pointing = instru.Coord(0, 0)
    pointings = []
    for chan in instruments.keys():
        pointing_chan = [main_pointing + instru.Coord(ra, dec) for ra, dec in data_dict["target"][chan]]
        pointings.append(instru.CoordList(pointing_chan).pix(step_angle))
    # the reference centres the cube on the third exposure of band 2A (main_fusion.py:148-149)
    anchor = "2a" if "2a" in data_dict["target"] and len(data_dict["target"]["2a"]) > 2 else next(iter(instruments))
    centre = data_dict["target"][anchor][min(2, len(data_dict["target"][anchor]) - 1)]
    alpha_axis = origin_alpha_axis + centre[0]
    beta_axis = origin_beta_axis + centre[1]
    return spectroSigRLSCT(sotf=sotf, templates=templates, alpha_axis=alpha_axis, beta_axis=beta_axis,
                           wavelength_axis=wavel_axis, instrs=list(instruments.values()), step_degree=step_angle,
                           pointings=pointings, **model_kwargs)


def assemble_data(data_dict, list_chan: Sequence[str]) -> np.ndarray:
    """`ndata` of main_fusion.py:257-260: every channel's exposures, raveled and concatenated."""
    return np.concatenate([np.array(data_dict["data"][chan]).ravel() for chan in list_chan])


def reconstruction_method(spectroModel, ndata, templates, result_path, hyperParameter, niter, method, scale_data,
                          printing: bool = False):
    """main_fusion.py:164-204; returns (path of the result directory, OptimizeResult, criterion object)."""
    value_init = 0
    result_dir = (f"{method}_MC_{len(spectroModel.instrs)}_MO_4_Temp_{templates.shape[0]}_nit_{str(niter)}"
                  f"_mu_{str('{:.2e}'.format(hyperParameter))}_SD_{scale_data}/")
    path = pathlib.Path(str(result_path) + result_dir)
    path.mkdir(parents=True, exist_ok=True)
    quadCrit_fusion = QuadCriterion_MRS(mu_spectro=1, y_spectro=np.copy(ndata), model_spectro=spectroModel,
                                        mu_reg=hyperParameter, printing=printing, gradient="separated")
    res_fusion = quadCrit_fusion.run_method(method, niter, perf_crit=1, calc_crit=True, value_init=value_init)
    y_cube = spectroModel.mapsToCube(res_fusion.x)
    np.save(path / "res_x.npy", res_fusion.x)
    np.save(path / "res_cube.npy", y_cube)
    np.save(path / "criterion.npy", quadCrit_fusion.L_crit_val)
    return path, res_fusion, quadCrit_fusion


def run(fusion_dir: str, npix: int = 501, hyper_parameter: float = 1.0, niter: int = 5, n_templates: int = 4,
        scale_data: bool = False, method: str = "lcg", list_chan: Sequence[str] = LIST_CHAN, **model_kwargs):
    """`parse_options` of main_fusion.py:211-270 as a function."""
    paths, step, step_angle = initialize_parameters(fusion_dir)
    origin_alpha_axis, origin_beta_axis, wavel_axis, templates, sotf = load_simulation_data(
        paths, step, step_angle, npix, n_templates)
    data_dict = load_data(list_chan, paths["save_filter_corrected_dir"])
    instruments = create_instruments(data_dict, list_chan)
    model = create_model(sotf, templates, origin_alpha_axis, origin_beta_axis, wavel_axis, instruments, step_angle,
                         data_dict, **model_kwargs)
    ndata = assemble_data(data_dict, list_chan)
    if scale_data:
        ndata = model.real_data_janskySR_to_jansky(ndata)
    return reconstruction_method(model, ndata, templates, paths["result_path"], hyper_parameter, niter, method,
                                 scale_data) + (model,)

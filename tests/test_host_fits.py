"""Host-side ingest of the reference driver (scripts/main_fusion.py:30-63): the minimal FITS reader and the
[L', S, na] -> [S, L', na] exposure layout.  No GPU."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def mf():
    try:
        from surfh_b200 import main_fusion
    except ImportError as exc:  # the CUDA library is not built: the module cannot bind
        pytest.skip(str(exc))
    return main_fusion


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fits_round_trip(tmp_path, mf, dtype):
    data = np.random.default_rng(0).standard_normal((5, 7, 3)).astype(dtype)
    path = str(tmp_path / "x.fits")
    mf.write_fits_primary(path, data, {"PA_V3": 251.25, "TARG_RA": -1.5e-4, "TARG_DEC": 3.0e-5, "NOTE": "abc"})
    assert os.path.getsize(path) % 2880 == 0
    hdr, back = mf.read_fits_primary(path)
    assert back.dtype == dtype and np.array_equal(back, data)
    assert hdr["PA_V3"] == 251.25 and hdr["TARG_RA"] == -1.5e-4 and hdr["TARG_DEC"] == 3.0e-5 and hdr["NOTE"] == "abc"
    assert hdr["NAXIS"] == 3 and (hdr["NAXIS1"], hdr["NAXIS2"], hdr["NAXIS3"]) == (3, 7, 5)


def test_load_data_layout(tmp_path, mf):
    """Files hold [L', S, na]; the driver hands [S, L', na] per exposure, exposures in file-name order."""
    n_slit, n_det, na = mf.DATASHAPE["4a"]
    rng = np.random.default_rng(1)
    exposures = [rng.standard_normal((n_det, n_slit, na)) for _ in range(2)]
    for i, e in enumerate(exposures):
        mf.write_fits_primary(str(tmp_path / f"ch4a_exp{i}.fits"), e,
                              {"PA_V3": 8.2, "TARG_RA": 1e-5 * i, "TARG_DEC": -2e-5 * i})
    d = mf.load_data(["4a"], str(tmp_path))
    assert len(d["data"]["4a"]) == 2 and d["rotation"]["4a"] == 8.2
    for i, e in enumerate(exposures):
        assert d["data"]["4a"][i].shape == (n_slit, n_det, na)
        assert np.array_equal(d["data"]["4a"][i], e.transpose(1, 0, 2))
        assert d["target"]["4a"][i] == (1e-5 * i, -2e-5 * i)
    flat = mf.assemble_data(d, ["4a"])
    assert flat.shape == (2 * n_slit * n_det * na,)
    assert np.array_equal(flat[: n_slit * n_det * na], exposures[0].transpose(1, 0, 2).ravel())


def test_instrument_table_matches_reference_script(mf):
    inst = mf.create_instruments({"rotation": {c: 10.0 for c in mf.LIST_CHAN}})
    assert list(inst) == mf.LIST_CHAN
    for chan, ifu in inst.items():
        n_slit, n_det, _ = mf.DATASHAPE[chan]
        assert ifu.n_slit == n_slit and ifu.n_wavel == n_det and ifu.fov.angle == -10.0 and ifu.name == chan.upper()

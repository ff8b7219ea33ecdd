"""The reference driver's sequence (scripts/main_fusion.py:136-204, 211-270) with the three imports swapped,
on a synthetic `Fusion/`-layout tree: files in, res_x.npy / res_cube.npy / criterion.npy out."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a.ravel() - b.ravel()) / np.linalg.norm(b.ravel()))


@pytest.fixture(scope="module")
def fusion_tree(tmp_path_factory):
    """Templates/, PSF/ and Filtered_slices/ of a 2-exposure band-1A observation; the exposure files are the
    operator applied to seeded maps plus 1 % noise, written as FITS [L', S, na] like the reference's inputs."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    from surfh_b200 import instru, main_fusion as mf, synthetic
    root = tmp_path_factory.mktemp("Fusion")
    for d in ("PSF", "Templates", "Filtered_slices"):
        os.makedirs(root / d)
    npix, k = 251, 4
    cfg = synthetic.mrs_config(["1a"], npix, k, 2, seed=21, name="flow")
    np.save(root / "Templates" / f"wavel_axis_orion_1ABC_2ABC_3ABC_4ABC_{k}_templates_SS4.npy", cfg.wavelength_axis)
    np.save(root / "Templates" / f"nmf_orion_1ABC_2ABC_3ABC_4ABC_{k}_templates_SS4.npy", cfg.templates * 10e3)
    np.save(root / "PSF" / mf.PSF_FILE, cfg.psf)
    # exposures: the data the instrument would record at the two dither targets
    targets = [(c.alpha, c.beta) for c in cfg.pointings[0]]
    rotation = -synthetic.FOV_ANGLE   # create_instruments uses angle = -PA_V3
    paths, step, step_angle = mf.initialize_parameters(str(root) + "/")
    data_dict = {"data": {"1a": []}, "target": {"1a": targets}, "rotation": {"1a": rotation}}
    ax = mf.load_simulation_data(paths, step, step_angle, npix, k)
    model = mf.create_model(ax[4], ax[3], ax[0], ax[1], ax[2], mf.create_instruments(data_dict, ["1a"]), step_angle,
                            data_dict)
    y = model.forward(cfg.maps)
    y = y + 0.01 * np.sqrt(np.mean(y ** 2)) * np.random.default_rng(5).standard_normal(y.shape)
    n_slit, n_det, na = mf.DATASHAPE["1a"]
    for p, (ra, dec) in enumerate(targets):
        block = y.reshape(2, n_slit, n_det, na)[p]
        mf.write_fits_primary(str(root / "Filtered_slices" / f"ch1a_exposure{p}.fits"),
                              np.ascontiguousarray(block.transpose(1, 0, 2)),
                              {"PA_V3": rotation, "TARG_RA": ra, "TARG_DEC": dec})
    return str(root) + "/", cfg, y


def test_driver_flow_writes_reference_result_files(fusion_tree):
    from surfh_b200 import fusion_CT, main_fusion as mf
    fusion_dir, cfg, y = fusion_tree
    niter, mu = 6, 5e3
    path, res, quad, model = mf.run(fusion_dir, npix=251, hyper_parameter=mu, niter=niter, n_templates=4,
                                    scale_data=False, method="lcg", list_chan=["1a"])
    assert os.path.isdir(path) and os.path.basename(os.path.normpath(path)) == "lcg_MC_1_MO_4_Temp_4_nit_6_mu_5.00e+03_SD_False"
    res_x, res_cube, crit = (np.load(os.path.join(path, f)) for f in ("res_x.npy", "res_cube.npy", "criterion.npy"))
    # shapes and dtypes of the reference's files
    assert res_x.shape == (4, 251, 251) and res_x.dtype == np.float64
    assert res_cube.shape == (len(cfg.wavelength_axis), 251, 251) and res_cube.dtype == np.float32
    # perf_crit=1, calc_crit=True: the criterion is appended when self.it % 5 == 2 after the increment
    assert crit.shape == (len([i for i in range(2, niter + 2) if i % 5 == 2]),) and crit.dtype == np.float64
    # the data the driver assembled from the FITS files is the vector the exposures were cut from
    assert np.array_equal(mf.assemble_data(mf.load_data(["1a"], fusion_dir + "Filtered_slices/"), ["1a"]), y)
    # values: the same solve called directly, criterion through the explicit forward pass
    direct = fusion_CT.lcg(model, y, 1.0, mu, np.zeros(model.ishape), tol=1e-12, max_iter=niter)
    assert rel(res_x, direct.x) <= 1e-13
    explicit = fusion_CT.QuadCriterion_MRS(1, y, model, mu)
    explicit.criterion_from_state = False
    trace = []
    fusion_CT.lcg(model, y, 1.0, mu, np.zeros(model.ishape), tol=1e-12, max_iter=niter,
                  callback=lambda r: trace.append(explicit.get_crit_val(r.x.reshape(model.ishape))))
    want = [trace[i - 2] for i in range(2, niter + 2) if i % 5 == 2]
    assert np.allclose(crit, want, rtol=1e-10)
    assert quad._solver()._state_evals == len(crit)      # ... and no forward pass was spent on them
    expect_cube = np.einsum("kij,kl->lij", res_x.astype(np.float32), (cfg.templates).astype(np.float32))
    assert rel(res_cube, expect_cube) <= 1e-6
    # scale_data=True goes through real_data_janskySR_to_jansky (main_fusion.py:262-265)
    path2, res2, _, _ = mf.run(fusion_dir, npix=251, hyper_parameter=mu, niter=2, n_templates=4, scale_data=True,
                               method="lcg", list_chan=["1a"])
    assert path2 != path and np.load(os.path.join(path2, "res_x.npy")).shape == (4, 251, 251)


def test_visualisation_helpers_raise_clearly(fusion_tree):
    from surfh_b200 import main_fusion as mf
    fusion_dir, cfg, y = fusion_tree
    from surfh_b200.model import spectroSigRLSCT
    model = spectroSigRLSCT(**cfg.model_args())
    for name in ("make_mask", "plot_slice", "project_FOV"):
        with pytest.raises(NotImplementedError, match="visualisation helper"):
            getattr(model, name)(None)
    with pytest.raises(NotImplementedError, match="sliceToCube"):
        model.channels[0].sliceToCube(None)

"""Distortion-correction pre-processing of the MRS detector slices, API of
`surfh.Preprocessing.distorsion_correction` (surfh/Preprocessing/distorsion_correction.py), with the
Shepard interpolation on the GPU.

The detector samples of a slit sit on a distorted (alpha, lambda) lattice; the model wants them on its regular
`[L', na]` grid.  The reference interpolates every slit with a single-threaded float32 Cython loop
(surfh/ToolsDir/shepard_interpolation.pyx:77-141: ~5e8 pair tests per slit); here that loop is one CUDA kernel
(`surfh_shepard`).  Labelling of the slits' connected components is host-side glue (scipy.ndimage instead of
skimage, which is not in this image).

Reference interface mirrored (paths relative to /root/reference):
    generate_label_image                 surfh/Preprocessing/distorsion_correction.py:27-35
    sort_labels_by_centroid              :38-53
    perform_shepard_interpolation        :55-98
    mrs_slices_distrorsion_correction    :106-181
"""
from __future__ import annotations

import numpy as np

from . import _capi


def generate_label_image(binary_grid):
    """Connected components of the binary slit mask (skimage.measure.label's default: full connectivity)."""
    from scipy import ndimage
    labels, _ = ndimage.label(np.asarray(binary_grid), structure=np.ones((3, 3), dtype=int))
    return labels


def sort_labels_by_centroid(label_image):
    """Relabel the components in increasing order of their centroid's column."""
    from scipy.ndimage import center_of_mass
    num_labels = int(label_image.max())
    centroids = center_of_mass(label_image, label_image, range(1, num_labels + 1))
    sorted_labels = np.argsort([c[1] for c in centroids]) + 1
    out = np.zeros_like(label_image)
    for new_label, old_label in enumerate(sorted_labels, start=1):
        out[label_image == old_label] = new_label
    return out


def perform_shepard_interpolation(alpha_valid, lambda_valid, intensity_valid, alpha_mesh, lambda_mesh, p, alpha_exp,
                                  pixel_cutoff, alpha_res, lambda_res, epsilon: float = 1e-6):
    """Exponential modified-Shepard interpolation onto the (alpha_mesh, lambda_mesh) grid; float32 like the
    reference (it casts every argument to float32 before the Cython call).  numpy in -> numpy out."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("surfh_b200 has no CPU fallback: the Shepard interpolation needs a CUDA device")
    lib = _capi.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    as32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev).reshape(-1)  # noqa: E731
    a_in, l_in, v_in = as32(alpha_valid), as32(lambda_valid), as32(intensity_valid)
    if not (a_in.numel() == l_in.numel() == v_in.numel()):
        raise ValueError("alpha, lambda and intensity of the samples must have the same length")
    a_mesh, l_mesh = as32(alpha_mesh), as32(lambda_mesh)
    if a_mesh.numel() != l_mesh.numel():
        raise ValueError("alpha_mesh and lambda_mesh must have the same shape")
    out = torch.empty_like(a_mesh)
    code = lib.surfh_shepard(a_in.data_ptr(), l_in.data_ptr(), v_in.data_ptr(), a_in.numel(), a_mesh.data_ptr(),
                             l_mesh.data_ptr(), a_mesh.numel(), float(p), float(alpha_exp), float(pixel_cutoff),
                             float(alpha_res), float(lambda_res), float(epsilon), out.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    _capi.check(None, code)
    return out.cpu().numpy().reshape(np.shape(alpha_mesh))


def mrs_slices_distrorsion_correction(model_channel, sorted_labeled_image, detector2world, data, chan_wavelength, mode):
    """Every labelled slit of a detector image interpolated onto the channel's `[S, L', na]` grid
    (distorsion_correction.py:106-181; same skipping rules, p = 2, alpha = 2, cutoff = 2 pixels)."""
    corrected_slices = np.zeros(model_channel.oshape[1:])
    i = 0
    for slit in range(len(np.unique(sorted_labeled_image))):
        if slit == 0:
            continue
        pixel_set = np.where(sorted_labeled_image == slit)
        alpha, beta, lam = detector2world(pixel_set[1], pixel_set[0])
        if mode == 0 and np.any(lam > np.max(chan_wavelength) + 1):
            continue
        if mode == 1 and np.any(lam < np.min(chan_wavelength) - 1):
            continue
        intensity = data[pixel_set]
        valid = ~np.isnan(intensity)
        grid_alpha = np.linspace(np.min(alpha), np.max(alpha), model_channel.oshape[-1])
        alpha_mesh, lambda_mesh = np.meshgrid(grid_alpha, chan_wavelength)
        alpha_res = (np.max(grid_alpha) - np.min(grid_alpha)) / alpha_mesh.shape[1]
        lambda_res = (np.max(chan_wavelength) - np.min(chan_wavelength)) / lambda_mesh.shape[0]
        corrected_slices[i] = perform_shepard_interpolation(alpha[valid], lam[valid], intensity[valid], alpha_mesh,
                                                            lambda_mesh, 2, 2.0, 2, alpha_res, lambda_res)
        i += 1
    return corrected_slices

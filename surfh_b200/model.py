"""`spectroSigRLSCT`: the reference's multi-band, multi-pointing LMM instrument operator, with
the same constructor, shapes, ordering and LinOp surface, computed by the CUDA library.

    y = sum_bands  Sig R . L . Sum . S . C . T  x

Reference interface mirrored (paths relative to /root/reference):
    class spectroSigRLSCT      surfh/Models/spectroModel.py:39-198
    class Channel              surfh/Models/spectroModelChannel.py:26-264 (as `.channels[i]` views)
Call sites kept working: scripts/main_fusion.py:136-156, 160-204, 257-265;
surfh/Simulation/fusion_CT.py:130-136, 242-265.

numpy in -> numpy out (host buffers; H2D/D2H inside the call, like the reference's LinOp);
torch CUDA tensor in -> torch CUDA tensor out on the current stream (no synchronisation).
"""
from __future__ import annotations

import ctypes as C
from math import ceil
from typing import Callable, List, Optional, Sequence, Union

import numpy as np

from . import _capi, geometry, instru
from .linop import LinOp

_DTYPES = {"float64": _capi.F64, "fp64": _capi.F64, "f64": _capi.F64, np.float64: _capi.F64,
           "float32": _capi.F32, "fp32": _capi.F32, "f32": _capi.F32, np.float32: _capi.F32}


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class SlicerView:
    """What callers read from `Channel.slicer` (surfh/Models/slicer.py): slit slices/weights."""

    def __init__(self, tables: geometry.BandTables):
        self._t = tables
        self.srf = tables.srf
        self.npix_slit_beta_width = tables.nb
        self.npix_slit_alpha_width = tables.npix_slit_alpha_width
        self.slices_shape = (tables.n_slit, tables.na)

    def get_slit_slices(self, slit_idx: int):
        return self._t.slices[slit_idx]

    def get_slit_weights(self, slit_idx: int, slices=None):
        n_alpha = self._t.slices[slit_idx][0].stop - self._t.slices[slit_idx][0].start
        return np.broadcast_to(self._t.weights[slit_idx][None, None, :], (1, n_alpha, self._t.nb)).copy()

    def get_slit_shape_t(self):
        sl = self._t.slices[0]
        return (self._t.n_wave, sl[0].stop - sl[0].start, sl[1].stop - sl[1].start)

    get_slit_shape = get_slit_shape_t


class ChannelView:
    """Read-only view of one band with the attribute names of the reference's `Channel`."""

    def __init__(self, tables: geometry.BandTables, n_pointing: int):
        self.tables = tables
        self.instr = tables.instr
        self.pointings = tables.pointings
        self.srf = tables.srf
        self.wslice = tables.wslice
        self.local_alpha_axis = tables.local_alpha_axis
        self.local_beta_axis = tables.local_beta_axis
        self.local_im_shape = tables.local_shape
        self.oshape = (n_pointing,) + tables.oshape[1:]
        self.slices_shape = (n_pointing, tables.n_slit, tables.na)
        self.slicer = SlicerView(tables)
        self.wpsf = tables.lsf

    @property
    def name(self):
        return self.instr.name

    def sliceToCube(self, *args, **kwargs):
        """`Channel.sliceToCube` (spectroModelChannel.py:266) re-projects detector slices onto the cube for
        plots: a visualisation helper outside the operator path, not provided."""
        raise NotImplementedError(
            "surfh_b200: Channel.sliceToCube is a visualisation helper of the reference "
            "(surfh/Models/spectroModelChannel.py:266) and is outside the forward/adjoint/CG path this "
            "library replaces; use surfh.Models.spectroModelChannel.Channel for plots")


def _viz_stub(name: str, where: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(
            f"surfh_b200: spectroSigRLSCT.{name} is a host-side visualisation helper of the reference ({where}) "
            f"and is outside the forward/adjoint/CG path this library replaces; call it on the reference class")
    method.__name__ = name
    method.__doc__ = f"Not provided: visualisation helper of the reference ({where})."
    return method


class spectroSigRLSCT(LinOp):
    """Drop-in for `surfh.Models.spectroModel.spectroSigRLSCT`.

    Extra keyword arguments (all optional, defaults reproduce the reference's behaviour):
      dtype          "float64" (default; parity <= 1e-10 vs the reference numpy path) or "float32"
      adjoint_mode   "reference" (default: bug-for-bug `gridding_t` interpolation) or "exact"
                     (true transpose; <Hx,y> = <x,H^T y> to rounding)
      lambda_range   (l0, l1): this process computes only cube wavelengths [l0, l1) (wavelength
                     sharding across GPUs, see surfh_b200.dist): `forward` then yields this shard's
                     partial sum of y and `adjoint` its partial maps
      local_bands    indices of the bands this process computes (coarser band sharding)
      comm           surfh_b200.dist.Comm: when given, device-tensor forward/adjoint/fwadj all-reduce
                     their partial results so every rank returns the full operator's output
      chunk          wavelengths per pipeline chunk (0 = library default)
      fft_backend    "auto" (default: hand-written chirp-z FFT kernels for maps up to 1024 pixels a side,
                     cuFFT beyond), "own" or "cufft"
      device         CUDA device index (default: current device)
      sotf           may also be a torch CUDA complex tensor, or a callable (l0, l1) -> complex
                     array/tensor of planes [l0, l1), so a multi-GB OTF never sits on the host
    """

    _rules = "channel"  # geometry rule set (geometry.build_band); MRSBlurred overrides it

    def __init__(self, sotf, templates, alpha_axis, beta_axis, wavelength_axis,
                 instrs: List[instru.IFU], step_degree: float, pointings: Sequence[instru.CoordList],
                 dtype="float64", adjoint_mode: str = "reference", local_bands: Optional[Sequence[int]] = None,
                 chunk: int = 0, device: Optional[int] = None, lambda_range=None, comm=None,
                 fft_backend: str = "auto"):
        self._h = None
        self._lib = _capi.load()
        if adjoint_mode not in _capi.ADJOINT_MODES:
            raise ValueError("adjoint_mode must be 'reference' or 'exact'")
        if fft_backend not in _capi.FFT_BACKENDS:
            raise ValueError("fft_backend must be 'auto', 'own' or 'cufft'")
        self.fft_backend = fft_backend
        self.adjoint_mode = adjoint_mode
        self._dtype_code = _DTYPES[dtype]
        self.np_dtype = np.float64 if self._dtype_code == _capi.F64 else np.float32
        self.alpha_axis = np.asarray(alpha_axis, dtype=np.float64)
        self.beta_axis = np.asarray(beta_axis, dtype=np.float64)
        self.wavelength_axis = np.asarray(wavelength_axis, dtype=np.float64)
        self.step_degree = step_degree
        self.instrs = [i.pix(step_degree) for i in instrs]
        self.templates = None if templates is None else np.ascontiguousarray(templates, dtype=np.float64)
        self.lmm = self.templates is not None
        self.pointings = pointings
        self.sotf = sotf if isinstance(sotf, np.ndarray) else None
        self.srfs = instru.get_srf([i.det_pix_size for i in instrs], step_degree * 3600)
        n_bands = len(instrs)
        self.local_bands = list(range(n_bands)) if local_bands is None else sorted(int(b) for b in local_bands)
        self.lambda_range = None if lambda_range is None else (int(lambda_range[0]), int(lambda_range[1]))
        self.comm = comm
        if device is not None:
            import torch
            torch.cuda.set_device(device)

        self.band_tables: List[geometry.BandTables] = [
            geometry.build_band(instr, self.alpha_axis, self.beta_axis, self.wavelength_axis, srf, pointings[it],
                                step_degree, with_adjoint=(it in self.local_bands),
                                lambda_range=self.lambda_range if it in self.local_bands else (0, 0),
                                rules=self._rules)
            for it, (srf, instr) in enumerate(zip(self.srfs, instrs))]
        self.local_bands = [it for it in self.local_bands if self.band_tables[it].is_local]
        self.partial = self.lambda_range is not None or len(self.local_bands) < n_bands
        n_point = len(pointings[0])
        self.instrs_oshape = [(n_point, t.n_slit, t.n_det, t.na) for t in self.band_tables]
        self._idx = np.cumsum([0] + [int(np.prod(s)) for s in self.instrs_oshape])
        for t, off in zip(self.band_tables, self._idx[:-1]):
            t.out_offset = int(off)
            if t.n_pointing != n_point:
                raise ValueError("every band must have the same number of pointings")
        self.channels = [ChannelView(t, n_point) for t in self.band_tables]
        self.list_wslice = [t.wslice for t in self.band_tables]
        self.list_local_alpha_axis = [t.local_alpha_axis for t in self.band_tables]
        self.list_local_beta_axis = [t.local_beta_axis for t in self.band_tables]
        self.list_local_im_shape = [t.local_shape for t in self.band_tables]
        self.cube_shape = (len(self.wavelength_axis), len(self.alpha_axis), len(self.beta_axis))
        self.imshape = self.cube_shape[1:]
        ishape = ((self.templates.shape[0],) + self.imshape) if self.lmm else self.cube_shape
        super().__init__(ishape=ishape, oshape=(int(self._idx[-1]),))
        if self.lmm and self.templates.shape[1] != len(self.wavelength_axis):
            raise ValueError("templates must be [K, n_lambda]")

        desc = _capi.ModelDesc(self._dtype_code, self.templates.shape[0] if self.lmm else 0,
                               len(self.alpha_axis), len(self.beta_axis), len(self.wavelength_axis), int(chunk),
                               _capi.ptr(self.templates) if self.lmm else None,
                               _capi.FFT_BACKENDS[fft_backend])
        handle = C.c_void_p()
        code = self._lib.surfh_create(C.byref(desc), C.byref(handle))
        _capi.check(None, code)
        self._h = handle
        if not self.local_bands:
            raise ValueError("this process has no band / wavelength to compute (empty shard)")
        self._upload_otf(sotf)
        for it in self.local_bands:
            self._add_band(self.band_tables[it])
        _capi.check(self._h, self._lib.surfh_finalize(self._h))
        # the big host tables are on the device now
        for it in self.local_bands:
            self.band_tables[it].adj_exact = None
            self.band_tables[it].adj_reference = None

    # ------------------------------------------------------------------ construction helpers
    def _needed_planes(self):
        need = np.zeros(len(self.wavelength_axis), dtype=bool)
        for it in self.local_bands:
            need[self.band_tables[it].wave_local] = True
        return need

    def _upload_otf(self, sotf, planes_per_call: int = 128):
        n_l, n_a, n_b = self.cube_shape
        shape = (n_a, n_b // 2 + 1)
        need = self._needed_planes()
        idx = np.flatnonzero(need)
        if len(idx) == 0:
            return
        # contiguous runs of needed planes, cut into pieces
        runs = np.split(idx, np.flatnonzero(np.diff(idx) > 1) + 1)
        for run in runs:
            for lo in range(int(run[0]), int(run[-1]) + 1, planes_per_call):
                hi = min(int(run[-1]) + 1, lo + planes_per_call)
                block = sotf(lo, hi) if callable(sotf) else sotf[lo:hi]
                if tuple(block.shape) != (hi - lo,) + shape:
                    raise ValueError(f"sotf planes must have shape {shape}, got {tuple(block.shape[1:])}")
                if _is_torch(block):
                    import torch
                    block = block.to(torch.complex128).contiguous()
                    if block.is_cuda:
                        torch.cuda.current_stream().synchronize()
                    p = block.data_ptr()
                else:
                    block = np.ascontiguousarray(block, dtype=np.complex128)
                    p = _capi.ptr(block)
                _capi.check(self._h, self._lib.surfh_set_otf(self._h, lo, hi - lo, p))

    def _add_band(self, t: geometry.BandTables):
        keep: list = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return _capi.ptr(a)

        d = _capi.BandDesc(
            t.n_pointing, t.n_slit, t.na, t.nb, t.srf, t.local_shape[0], t.local_shape[1], t.wave_local.start,
            t.n_wave, t.n_det,
            _capi.SPECTRAL_BETA_SUM if t.lsf is None else _capi.SPECTRAL_LSF, t.wave_local.start - t.wslice.start,
            t.out_offset,
            arr(t.slit_a0, np.int32), arr(t.slit_b0, np.int32), arr(t.weights, np.float64),
            None if t.lsf is None else arr(t.lsf, np.float64), arr(t.grid_base, np.int32),
            arr(t.grid_frac, np.float64),
            _capi.csr_desc(t.adj_exact, keep), _capi.csr_desc(t.adj_reference, keep))
        _capi.check(self._h, self._lib.surfh_add_band(self._h, C.byref(d)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and getattr(self, "_lib", None) is not None:
            self._lib.surfh_destroy(h)

    # ------------------------------------------------------------------------------ helpers
    @property
    def alpha_step(self) -> float:
        return self.alpha_axis[1] - self.alpha_axis[0]

    @property
    def beta_step(self) -> float:
        return self.beta_axis[1] - self.beta_axis[0]

    @property
    def mode_code(self) -> int:
        return _capi.ADJOINT_MODES[self.adjoint_mode]

    @property
    def handle(self):
        return self._h

    def _torch_dtype(self):
        import torch
        return torch.float64 if self._dtype_code == _capi.F64 else torch.float32

    def _stream(self) -> int:
        import torch
        return torch.cuda.current_stream().cuda_stream

    def _dev_in(self, t, size: int):
        t = t.to(self._torch_dtype()).contiguous().reshape(-1)
        if not t.is_cuda:
            raise ValueError("torch inputs must be CUDA tensors (pass numpy arrays for host data)")
        if t.numel() != size:
            raise ValueError(f"expected {size} elements, got {t.numel()}")
        return t

    # ------------------------------------------------------------------------------ LinOp
    def forward(self, maps):
        """y = H maps.  maps: [K, N, N] (or the [n_lambda, N, N] cube when templates is None)."""
        if _is_torch(maps):
            import torch
            x = self._dev_in(maps, self.isize)
            # a partial model (band / wavelength shard) does not write every element of y: bands it does not
            # own, and the detector rows of beta-sum bands outside its wavelength window, must read as zeros
            alloc = torch.zeros if self.partial else torch.empty
            y = alloc(self.osize, dtype=x.dtype, device=x.device)
            _capi.check(self._h, self._lib.surfh_forward(self._h, x.data_ptr(), y.data_ptr(), self._stream()))
            self._reduce(y)
            return y
        x = np.ascontiguousarray(np.asarray(maps, dtype=np.float64).reshape(self.ishape))
        y = np.zeros(self.oshape, dtype=np.float64)
        _capi.check(self._h, self._lib.surfh_forward_host(self._h, _capi.ptr(x), _capi.ptr(y)))
        return y

    def adjoint(self, inarray):
        """x = H^T y in the model's adjoint_mode."""
        if _is_torch(inarray):
            import torch
            y = self._dev_in(inarray, self.osize)
            x = torch.empty(self.ishape, dtype=y.dtype, device=y.device)
            _capi.check(self._h, self._lib.surfh_adjoint(self._h, y.data_ptr(), x.data_ptr(), self.mode_code,
                                                         self._stream()))
            self._reduce(x)
            return x
        y = np.ascontiguousarray(np.asarray(inarray, dtype=np.float64).reshape(-1))
        if y.size != self.osize:
            raise ValueError(f"expected {self.osize} samples, got {y.size}")
        x = np.empty(self.ishape, dtype=np.float64)
        _capi.check(self._h, self._lib.surfh_adjoint_host(self._h, _capi.ptr(y), _capi.ptr(x), self.mode_code))
        return x

    def fwadj(self, maps, out=None):
        """H^T H maps without the detector vector leaving the device (aljabr.LinOp.fwadj as inherited by
        the reference class).  numpy in -> numpy out (H2D of the maps, D2H of the result inside the call;
        `out`: optional float64 host array to fill, e.g. a view of pinned memory); CUDA tensor in -> CUDA
        tensor out."""
        import torch
        if _is_torch(maps):
            x = self._dev_in(maps, self.isize)
            return self.fwadj_into(x, torch.empty(self.ishape, dtype=x.dtype, device=x.device))
        host = torch.from_numpy(np.ascontiguousarray(np.asarray(maps, dtype=np.float64).reshape(self.ishape)))
        dt = self._torch_dtype()
        if getattr(self, "_fwadj_io", None) is None:
            self._fwadj_io = (torch.empty(self.ishape, dtype=torch.float64, device="cuda"),
                              torch.empty(self.ishape, dtype=dt, device="cuda"))
        stage, q = self._fwadj_io
        stage.copy_(host, non_blocking=True)
        x = stage if dt == torch.float64 else stage.to(dt)
        self.fwadj_into(x.reshape(-1), q)
        res = torch.from_numpy(out).reshape(self.ishape) if out is not None else torch.empty(self.ishape, dtype=torch.float64)
        res.copy_(q)  # device -> host (converts fp32 results), synchronises
        return out if out is not None else res.numpy()

    def fwadj_into(self, x, out, y_scratch=None):
        """out = H^T H x on device tensors.  Sharded (partial) models exchange the detector vector
        (all-reduce of the partial sums over wavelength shards) between the two halves and the
        [K, N, N] result at the end; unsharded models run the fused library call.  `y_scratch`: optional
        device tensor of `osize` elements that receives H x (unsharded models only)."""
        if not self.partial:
            # every rank holds the whole operator: nothing to sum (a comm on an unsharded model is ignored --
            # summing W identical copies would scale H^T H by W)
            _capi.check(self._h, self._lib.surfh_fwadj(self._h, x.data_ptr(), out.data_ptr(), self.mode_code,
                                                       None if y_scratch is None else y_scratch.data_ptr(),
                                                       self._stream()))
            return out
        if y_scratch is not None:
            raise ValueError("y_scratch is only available on an unsharded model")
        if self.comm is None:
            raise ValueError("a sharded model (lambda_range / local_bands) needs comm= to apply H^T H: "
                             "the detector vector has to be summed over the shards between H and H^T")
        import torch
        if getattr(self, "_y_shard", None) is None or self._y_shard.dtype != x.dtype:
            self._y_shard = torch.zeros(self.osize, dtype=x.dtype, device=x.device)
        y = self._y_shard
        if self._y_needs_zeroing:
            # elements this shard does not write (bands it does not touch, detector rows of beta-sum bands
            # outside its wavelength window) still hold the previous application's exchanged SUM
            y.zero_()
        _capi.check(self._h, self._lib.surfh_forward(self._h, x.data_ptr(), y.data_ptr(), self._stream()))
        # the adjoint of this shard reads only the detector blocks of the bands it touches: sum each shared
        # band among the ranks that hold a share of it (sub-communicators) instead of all-reducing all of y
        if self.lambda_range is not None and len(self.local_bands) > 0:
            self._band_exchange().reduce_shared(y)
        else:
            self.comm.allreduce_sum(y)
        _capi.check(self._h, self._lib.surfh_adjoint(self._h, y.data_ptr(), out.data_ptr(), self.mode_code,
                                                     self._stream()))
        self.comm.allreduce_sum(out)
        return out

    def _reduce(self, tensor):
        """Sum partial results over the shards.  Only a partial model has anything to sum."""
        if self.partial and self.comm is not None:
            self.comm.allreduce_sum(tensor)
        return tensor

    @property
    def _y_needs_zeroing(self) -> bool:
        if len(self.local_bands) < len(self.band_tables):
            return True
        return any(self.band_tables[it].lsf is None for it in self.local_bands)

    def _band_exchange(self):
        """Collective on first use: every rank of the communicator must reach it (fwadj does)."""
        if getattr(self, "_exchange", None) is None:
            from .dist import BandExchange
            windows = [(t.wslice.start, t.wslice.stop) for t in self.band_tables]
            blocks = [(int(self._idx[c]), int(self._idx[c + 1] - self._idx[c])) for c in range(len(self.band_tables))]
            self._exchange = BandExchange(self.comm, windows, blocks, self.lambda_range)
        return self._exchange

    fwback = fwadj
    project_FOV = _viz_stub("project_FOV", "surfh/Models/spectroModel.py:201")
    plot_slice = _viz_stub("plot_slice", "surfh/Models/spectroModel.py:242")
    make_mask = _viz_stub("make_mask", "surfh/Models/spectroModel.py:289")

    def matvec(self, point):
        return self.forward(point.reshape(self.ishape)).reshape(-1)

    def rmatvec(self, point):
        return self.adjoint(point.reshape(-1)).reshape(-1)

    # ---------------------------------------------------------------- result export / scaling
    def mapsToCube(self, maps):
        """float32 cube = T maps (spectroModel.py:190-192)."""
        import torch
        if not self.lmm:
            raise ValueError("mapsToCube needs templates")
        was_numpy = not _is_torch(maps)
        x = torch.as_tensor(np.ascontiguousarray(maps, dtype=np.float64), device="cuda") if was_numpy else maps
        x = self._dev_in(x, self.isize)
        cube = torch.empty(self.cube_shape, dtype=torch.float32, device=x.device)
        _capi.check(self._h, self._lib.surfh_maps_to_cube(self._h, x.data_ptr(), cube.data_ptr(), self._stream()))
        return cube.cpu().numpy() if was_numpy else cube

    def cubeTomaps(self, cube):
        """maps[k] = sum_l cube[l] * templates[k, l] (spectroModel.py:187-188); host-side export helper."""
        return np.einsum("lij,kl->kij", np.asarray(cube), self.templates)

    def real_data_janskySR_to_jansky(self, data: np.ndarray) -> np.ndarray:
        """Jy/sr -> Jy: every slit scaled by (sum of its beta weights) * srf (spectroModel.py:225-239)."""
        out = np.zeros_like(data)
        for c, t in enumerate(self.band_tables):
            block = np.array(data[self._idx[c]: self._idx[c + 1]], dtype=np.float64).reshape(self.instrs_oshape[c])
            scale = t.weights.sum(axis=1) * t.srf
            out[self._idx[c]: self._idx[c + 1]] = (block * scale[None, :, None, None]).ravel()
        return out

    # -------------------------------------------------------------------------- diagnostics
    def launch_count(self) -> int:
        return int(self._lib.surfh_launch_count(self._h))

    def own_launch_count(self) -> int:
        return int(self._lib.surfh_own_launch_count(self._h))

    def contraction_info(self) -> dict:
        """How the spectral response is evaluated: {'mode': 'ozaki_i8' | 'dmma_tma' | 'mma_sync' | 'simt' | 'none', 'digits': n,
        'executed_fraction': share of the digit products that is run (all-zero digit tiles of the LSF are skipped)}."""
        import ctypes
        mode, digits, frac = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_double(1.0)
        _capi.check(self._h, self._lib.surfh_contraction_info(self._h, ctypes.byref(mode), ctypes.byref(digits),
                                                              ctypes.byref(frac)))
        return {"mode": {-1: "none", 0: "mma_sync", 1: "dmma_tma", 2: "ozaki_i8", 3: "simt"}[mode.value],
                "digits": int(digits.value), "executed_fraction": float(frac.value)}

    def workspace_bytes(self) -> int:
        return int(self._lib.surfh_workspace_bytes(self._h))

    def profile(self, enable: bool) -> None:
        _capi.check(self._h, self._lib.surfh_profile_enable(self._h, 1 if enable else 0))

    def profile_read(self):
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        nbytes = (C.c_double * cap)()
        flops = (C.c_double * cap)()
        launches = (C.c_int32 * cap)()
        n = self._lib.surfh_profile_read(self._h, cap, names, ms, nbytes, flops, launches)
        if n < 0:
            raise _capi.SurfhError(n, "profile_read failed")
        return [dict(stage=names[i].decode(), ms=float(ms[i]), bytes=float(nbytes[i]), flops=float(flops[i]),
                     launches=int(launches[i])) for i in range(n)]


# the reference's callers import the class under this module alias too
SpectroLMM = spectroSigRLSCT

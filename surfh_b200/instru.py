"""Host-side instrument description, API-compatible with the subset of
`surfh.Models.instru` that the LMM hot path and `scripts/main_fusion.py` touch.

Only *description* lives here (coordinates, field of view, spectral resolution, IFU band).
Everything the GPU needs is derived from these objects once, in `surfh_b200.geometry`.

Reference interface mirrored (paths relative to /root/reference):
  Coord, CoordList        surfh/Models/instru.py:88-152, 155-255
  FOV                     surfh/Models/instru.py:257-410
  SpectralBlur            surfh/Models/instru.py:484-572
  IFU                     surfh/Models/instru.py:576-697
  get_srf                 surfh/Models/instru.py:67-84
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence

import numpy as np

__all__ = ["Coord", "CoordList", "FOV", "SpectralBlur", "IFU", "get_srf", "rotmatrix"]


def rotmatrix(degree: float) -> np.ndarray:
    """Counter-clockwise 2x2 rotation, angle in degrees."""
    c, s = math.cos(math.radians(degree)), math.sin(math.radians(degree))
    return np.array([[c, -s], [s, c]], dtype=np.float64)


def get_srf(det_pix_size_list: Sequence[float], step: float) -> List[int]:
    """Super-resolution factor of each band: detector pixel size // cube step (arcsec)."""
    return [int(size // step) for size in det_pix_size_list]


class Coord:
    """An (alpha, beta) coordinate in degrees; supports + and - and grid rounding."""

    __slots__ = ("alpha", "beta")

    def __init__(self, alpha: float, beta: float):
        self.alpha = alpha
        self.beta = beta

    @classmethod
    def from_array(cls, arr) -> "Coord":
        return cls(arr[0], arr[1])

    def _other(self, other) -> "Coord":
        if not isinstance(other, Coord):
            raise ValueError("`coord` must be a `Coord`")
        return other

    def __add__(self, other) -> "Coord":
        o = self._other(other)
        return Coord(self.alpha + o.alpha, self.beta + o.beta)

    def __sub__(self, other) -> "Coord":
        o = self._other(other)
        return Coord(self.alpha - o.alpha, self.beta - o.beta)

    def rotate(self, degree: float) -> "Coord":
        v = rotmatrix(-degree) @ np.array([self.alpha, self.beta], dtype=np.float64)
        return Coord(float(v[0]), float(v[1]))

    def pix(self, step: float) -> "Coord":
        """Round both coordinates to the `step` grid (python round(), ties to even)."""
        return Coord(round(self.alpha / step) * step, round(self.beta / step) * step)

    def __iter__(self):
        yield self.alpha
        yield self.beta

    def __eq__(self, other) -> bool:
        return isinstance(other, Coord) and self.alpha == other.alpha and self.beta == other.beta

    def __repr__(self) -> str:
        return f"Coord(alpha={self.alpha!r}, beta={self.beta!r})"


class CoordList(list):
    """A list of `Coord` (one per dither pointing) with bounding-box helpers."""

    @classmethod
    def from_array(cls, arr) -> "CoordList":
        return cls(Coord.from_array(a) for a in arr)

    def pix(self, step: float) -> "CoordList":
        return CoordList(c.pix(step) for c in self)

    alpha_min = property(lambda self: min(c.alpha for c in self))
    alpha_max = property(lambda self: max(c.alpha for c in self))
    beta_min = property(lambda self: min(c.beta for c in self))
    beta_max = property(lambda self: max(c.beta for c in self))
    alpha_mean = property(lambda self: (self.alpha_max + self.alpha_min) / 2)
    beta_mean = property(lambda self: (self.beta_max + self.beta_min) / 2)
    alpha_box = property(lambda self: self.alpha_max - self.alpha_min)
    beta_box = property(lambda self: self.beta_max - self.beta_min)
    box = property(lambda self: (self.alpha_box, self.beta_box))


class FOV:
    """A rectangular field of view: widths in degrees, centre `origin`, rotation `angle` (deg)."""

    def __init__(self, alpha_width: float, beta_width: float,
                 origin: Optional[Coord] = None, angle: float = 0):
        self.alpha_width = alpha_width
        self.beta_width = beta_width
        self.origin = Coord(0, 0) if origin is None else origin
        self.angle = angle

    def __add__(self, coord: Coord) -> "FOV":
        return FOV(self.alpha_width, self.beta_width, self.origin + coord, self.angle)

    def __sub__(self, coord: Coord) -> "FOV":
        return FOV(self.alpha_width, self.beta_width, self.origin - coord, self.angle)

    def rotate(self, degree: float) -> None:
        self.angle += degree

    def shift(self, coord: Coord) -> None:
        self.origin = self.origin + coord

    def _corner(self, sa: float, sb: float) -> Coord:
        return Coord(sa * self.alpha_width / 2, sb * self.beta_width / 2).rotate(self.angle) + self.origin

    @property
    def vertices(self):
        """Corners, counter-clockwise from the lower left."""
        return (self._corner(-1, -1), self._corner(1, -1), self._corner(1, 1), self._corner(-1, 1))

    @property
    def bbox(self):
        v = self.vertices
        return (Coord(min(p.alpha for p in v), min(p.beta for p in v)),
                Coord(max(p.alpha for p in v), max(p.beta for p in v)))

    def __repr__(self) -> str:
        return (f"FOV(alpha_width={self.alpha_width!r}, beta_width={self.beta_width!r}, "
                f"origin={self.origin!r}, angle={self.angle!r})")


class SpectralBlur:
    """Spectral response of a grating of resolution R = lambda / delta-lambda."""

    def __init__(self, grating_resolution: float):
        self.grating_resolution = grating_resolution

    @property
    def grating_len(self) -> float:
        return 2 * 0.44245 / np.pi * self.grating_resolution


class IFU:
    """One MRS band: field of view, detector pixel size (arcsec), number of slits, spectral
    blur and detector wavelength axis (um)."""

    def __init__(self, fov: FOV, det_pix_size: float, n_slit: int, w_blur: SpectralBlur,
                 pce=None, wavel_axis: Iterable[float] = (), name: str = "_"):
        self.fov = fov
        self.det_pix_size = det_pix_size
        self.n_slit = n_slit
        self.w_blur = w_blur
        self.pce = pce
        self.wavel_axis = np.asarray(wavel_axis, dtype=np.float64)
        self.name = name

    wavel_min = property(lambda self: self.wavel_axis[0])
    wavel_max = property(lambda self: self.wavel_axis[-1])
    wavel_step = property(lambda self: self.wavel_axis[1] - self.wavel_axis[0])
    n_wavel = property(lambda self: len(self.wavel_axis))
    slit_beta_width = property(lambda self: self.fov.beta_width / self.n_slit)

    def wslice(self, wavel_input_axis, margin: float = 0) -> slice:
        """Cube wavelengths this band observes, widened by `margin` um on each side."""
        wavel_input_axis = np.asarray(wavel_input_axis)
        lo = max(self.wavel_min - margin, wavel_input_axis.min())
        hi = min(self.wavel_max + margin, wavel_input_axis.max())
        return slice(int(np.flatnonzero(wavel_input_axis <= lo)[-1]),
                     int(np.flatnonzero(wavel_input_axis >= hi)[0]))

    def pix(self, step: float) -> "IFU":
        """Same band with the FoV origin rounded onto the `step` grid."""
        name = self.name if self.name.endswith("_pix") else self.name + "_pix"
        return IFU(FOV(self.fov.alpha_width, self.fov.beta_width, self.fov.origin.pix(step),
                       self.fov.angle),
                   self.det_pix_size, self.n_slit, self.w_blur, self.pce, self.wavel_axis, name)

    def __repr__(self) -> str:
        return f"IFU(name={self.name!r}, n_slit={self.n_slit}, det_pix_size={self.det_pix_size})"

"""surfh_b200: B200-native (sm_100a CUDA + cuFFT) implementation of surfh's LMM MIRI-MRS
instrument operator and of the CG fusion loop that drives it, behind the reference's own
LinOp API.  See DESIGN.md and INTEGRATION.md."""
from . import instru  # noqa: F401
from .linop import LinOp, dottest  # noqa: F401

__all__ = ["instru", "LinOp", "dottest", "spectroSigRLSCT", "SpectroLMM", "QuadCriterion_MRS", "lcg"]


def __getattr__(name):
    # heavy modules (ctypes library, solver) are imported on first use
    if name in ("spectroSigRLSCT", "SpectroLMM"):
        from .model import spectroSigRLSCT
        return spectroSigRLSCT
    if name in ("QuadCriterion_MRS", "lcg", "NpDiff_r", "NpDiff_c"):
        from . import fusion_CT
        return getattr(fusion_CT, name)
    raise AttributeError(name)

"""TEST INFRASTRUCTURE (CPU restatement, never on the product path): the integer-sliced arithmetic of the tcgen05
contraction (surfh_b200/csrc/kernels_ozaki.cuh) in numpy.

The product evaluated is the reference's spectral response, `wblur_subSampling` / `wblur_t`
(surfh/ToolsDir/jax_utils.py:72-91): y[m, n] = sum_k W[m, k] G[n, k].  The CUDA kernel writes every row of both operands
as 2^(e-6) * sum_p d_p 2^(-7p) with int8 digits, multiplies the digit planes as exact integer matrices (levels
L_t = sum_{p+q=t} dA_p dB_q^T, t < S) and sums the levels in fp64.  This module states that arithmetic step by step so that a
CPU test can bound its error against an extended-precision product; it is not a port of any reference file (the reference
multiplies in fp64 through numpy / JAX einsum)."""
import numpy as np


def cut_rows(x: np.ndarray, digits: int):
    """Digit planes [S, rows, K] (int8 values, |d| <= 64) and the per-row scale 2^(e-6), e = the row's exponent."""
    x = np.asarray(x, dtype=np.float64)
    amax = np.abs(x).max(axis=1)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))).astype(np.int64) + 1, 0)
    e = np.maximum(e, -1000)
    r = x * np.ldexp(1.0, (6 - e))[:, None]          # |r| < 64, exact (power of two)
    planes = np.empty((digits,) + x.shape, dtype=np.int8)
    for p in range(digits):
        d = np.rint(r)                                # round to nearest even, like the +- 1.5 * 2^52 trick
        planes[p] = d.astype(np.int8)
        r = (r - d) * 128.0                           # exact
    return planes, np.ldexp(1.0, e - 6)


def product(a: np.ndarray, b: np.ndarray, digits: int = 8) -> np.ndarray:
    """a [M, K] @ b[N, K]^T through `digits` int8 digits per operand: S (S + 1) / 2 exact integer products."""
    da, sa = cut_rows(a, digits)
    db, sb = cut_rows(b, digits)
    levels = np.zeros((digits, a.shape[0], b.shape[0]), dtype=np.int64)
    for p in range(digits):
        for q in range(digits - p):
            levels[p + q] += da[p].astype(np.int64) @ db[q].astype(np.int64).T
    assert np.abs(levels).max() < 2 ** 31, "a level does not fit the int32 accumulators of the tensor memory"
    acc = levels[digits - 1].astype(np.float64)
    for t in range(digits - 2, -1, -1):
        acc = acc * 0.0078125 + levels[t]             # Horner in 2^-7, like the epilogue
    return acc * sa[:, None] * sb[None, :]


def tile_mask(planes: np.ndarray, tile_rows: int = 128, tile_k: int = 64) -> np.ndarray:
    """[row tiles, k-blocks] bit mask: bit p set when digit p of that tile is not all zero (k-block 0: always dense)."""
    s, rows, k = planes.shape
    mt, nk = -(-rows // tile_rows), -(-k // tile_k)
    mask = np.zeros((mt, nk), dtype=np.uint8)
    for i in range(mt):
        for j in range(nk):
            blk = planes[:, i * tile_rows:(i + 1) * tile_rows, j * tile_k:(j + 1) * tile_k]
            bits = 0
            for p in range(s):
                if np.any(blk[p] != 0):
                    bits |= 1 << p
            mask[i, j] = bits if j else (1 << s) - 1
    return mask

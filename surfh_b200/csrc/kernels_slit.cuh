// k3 / k3^T: IFU slit sampling and its two "adjoints".
//
// Forward (gather), fused S . Sum . L . alpha-decimation for all pointings of a band:
//   G[l, (p,a,s,b)] = w[s,b] * sum_{m<srf} bilinear(cube[l]; grid[p][(a0[s] + a*srf + m) mod A, b0[s] + b])
// Replaces (paths relative to the reference tree)
//   Channel.gridding            surfh/Models/spectroModelChannel.py:158-177
//     -> cythons_files.find_indices / solve_2D_hypercube   surfh/ToolsDir/cythons_files.pyx:109-193
//   the FFT "Sum" stage          spectroModelChannel.py:220-223  (== circular box-sum of srf rows)
//   Slicer.slicing               surfh/Models/slicer.py:64-68
//   the `[:, : na*srf : srf]` decimation                  spectroModelChannel.py:229
// Only the rows the detector keeps are ever computed (the reference computes and discards
// (srf-1)/srf of them), and the two FFTs per pointing of the Sum stage disappear.
//
// Adjoint (scatter as gather): the composition of the slit placement, Sum^T and either the true
// transpose of the bilinear gridding (exact) or the reference's `gridding_t` interpolation
// (spectroModelChannel.py:180-199, 234-264) is precomputed on the host as ONE sparse table
// cube pixel -> (slit-space column, weight), shared by every wavelength.  Each cube pixel is owned
// by one thread, so the reduction is a fixed-order segmented sum: deterministic, no atomics.
//
// Both kernels are HBM/L2-bound index streams; a thread carries `LB` wavelengths so that one
// table lookup serves LB planes.
#pragma once
#include "common.cuh"

#ifndef SURFH_GATHER_MERGE
#define SURFH_GATHER_MERGE 1
#endif

namespace surfh {

template <typename T> struct SlitTables {
    const int32_t* slit_a0;    // [S]
    const int32_t* slit_b0;    // [S]
    const T* slit_w;           // [S, nb]
    const int32_t* grid_base;  // [P, A*B]
    const T* grid_frac;        // [P, A*B, 2]
    int32_t P, S, na, nb, srf, A, B;
    int32_t ncol;  // P*S*na*nb
    // slit-space vector G: element (wavelength l, detector column n, beta b) lives at n * g_col + l * g_l + b.
    // Bands with a spectral response store it K-fast per detector column, columns in the detector's own order
    // n = (p*S + s)*na + a (g_col = Lambda_b * nb, g_l = nb, g_psa = 1): the contraction's operand is then a plain
    // [n][k = l*nb + b] matrix, a 2-D TMA tensor map.  Beta-sum bands keep [l][n'][b] with n' = (p*na + a)*S + s
    // (g_col = nb, g_l = ncol, g_psa = 0).
    int32_t g_col, g_l, g_psa;
};

template <typename T, int LB>
__global__ void __launch_bounds__(128)
slit_gather_kernel(const T* __restrict__ cube, size_t plane /* elements per cube plane */, int n_beta,
                   int n_l, SlitTables<T> t, T* __restrict__ G) {
    // Thread -> output mapping: the warps of a CTA take the SAME 32 outputs (s, a, b) of DIFFERENT pointings.
    // The dithers shift the field of view by a couple of pixels only, so the warps of a CTA read almost the
    // same cube sectors at almost the same time and share them in L1 (the kernel is bound by the L2 -> L1
    // sector stream of its scattered taps, not by HBM).
    const int n_per_p = t.S * t.na * t.nb;         // outputs per pointing
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int pp = t.P < n_warps ? t.P : n_warps;  // pointings side by side in a CTA
    const int chunks = n_warps / pp;               // 32-output chunks per CTA
    if (warp >= pp * chunks) return;
    const int o = (blockIdx.x * chunks + warp / pp) * 32 + lane;
    if (o >= n_per_p) return;
    const int l0 = blockIdx.y * LB;
    // o enumerates (a, s, b) with (s, b) fastest: a warp's 32 lanes walk ~32 consecutive local columns of ONE
    // local row, i.e. ~5 runs of ~7 cube pixels, instead of 4 local rows of 8 columns
    const int b = o % t.nb;
    int r = o / t.nb;
    const int s = r % t.S;
    const int a = r / t.S;
  for (int p = warp % pp; p < t.P; p += pp) {
    const int n_col = t.g_psa ? (p * t.S + s) * t.na + a : (p * t.na + a) * t.S + s;
    const size_t c = (size_t)n_col * t.g_col + b;
    const int j = t.slit_b0[s] + b;
    const int i_first = t.slit_a0[s] + a * t.srf;
    const int32_t* gb = t.grid_base + (size_t)p * t.A * t.B;
    const T* gf = t.grid_frac + (size_t)p * t.A * t.B * 2;
    T acc[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) acc[u] = T(0);
    const T* base_l = cube + (size_t)l0 * plane;
#if SURFH_GATHER_MERGE
    // The srf bilinear samples of the box-sum walk down (almost) one cube column: the lower row of sample m is the
    // upper row of sample m + 1 (the rotated local row step is 0.99 pixel), usually in the same two columns.  The
    // lower-row weights are therefore carried to the next sample and merged with its upper-row weights: 2 loads per
    // sample and plane (+ 2 to flush) instead of 4 -- 16 instead of 28 for srf = 7.  Whenever the carried pair does
    // not sit where the next sample's upper row is (the column index stepped, or the row index did not), it is
    // flushed on its own: every case stays exact, only the order of the additions differs from 4 taps per sample.
    auto emit = [&](int32_t o, T w0, T w1) {
#pragma unroll
        for (int u = 0; u < LB; ++u) {
            if (l0 + u < n_l) {
                const T* pl = base_l + (size_t)u * plane + o;
                acc[u] = fma(__ldg(pl), w0, acc[u]);
                acc[u] = fma(__ldg(pl + 1), w1, acc[u]);
            }
        }
    };
    int32_t poff = -1;
    T p0 = T(0), p1 = T(0);
    for (int m = 0; m < t.srf; ++m) {
        int i = i_first + m;
        i = i >= t.A ? i - t.A : i;  // circular wrap of the FFT box-sum
        const int q = i * t.B + j;
        const int32_t off = __ldg(gb + q);
        const T y0 = __ldg(gf + 2 * q), y1 = __ldg(gf + 2 * q + 1);
        T u0 = (T(1) - y0) * (T(1) - y1), u1 = (T(1) - y0) * y1;
        if (poff == off) {
            u0 += p0;
            u1 += p1;
        } else if (poff >= 0) {
            emit(poff, p0, p1);
        }
        emit(off, u0, u1);
        poff = off + n_beta;
        p0 = y0 * (T(1) - y1);
        p1 = y0 * y1;
    }
    emit(poff, p0, p1);
#else   // 4 taps per sample (round 1; a software-pipelined table read was measured: 3.11 vs 3.07 ms, removed)
    for (int m = 0; m < t.srf; ++m) {
        int i = i_first + m;
        i = i >= t.A ? i - t.A : i;  // circular wrap of the FFT box-sum
        const int q = i * t.B + j;
        const int32_t off = __ldg(gb + q);
        const T y0 = __ldg(gf + 2 * q), y1 = __ldg(gf + 2 * q + 1);
        const T w00 = (T(1) - y0) * (T(1) - y1), w01 = (T(1) - y0) * y1;
        const T w10 = y0 * (T(1) - y1), w11 = y0 * y1;
#pragma unroll
        for (int u = 0; u < LB; ++u) {
            if (l0 + u < n_l) {
                const T* pl = base_l + (size_t)u * plane + off;
                T v = __ldg(pl) * w00;
                v = fma(__ldg(pl + 1), w01, v);
                v = fma(__ldg(pl + n_beta), w10, v);
                v = fma(__ldg(pl + n_beta + 1), w11, v);
                acc[u] += v;
            }
        }
    }
#endif
    const T w = t.slit_w[s * t.nb + b];
#pragma unroll
    for (int u = 0; u < LB; ++u)
        if (l0 + u < n_l) G[c + (size_t)(l0 + u) * t.g_l] = w * acc[u];
  }
}

// (A shared-memory-tiled variant -- per-tile footprint analysis, parallelogram-shaped staging area filled by
// double-buffered cp.async, taps read from shared memory -- was built and measured in round 1: bitwise-equal
// results, 152 us against this kernel's 137 us on config C2: the 4 taps per sample cost as many LSU
// wavefronts from shared memory (bank conflicts across the rotated rows) as they do from L1.  Removed.)

// The host hands the adjoint tables over as CSR (include/surfh_b200.h); on the device they are stored
// "sliced ELL": rows in slices of 32 (one warp), the k-th entries of a slice's 32 rows adjacent in memory,
// every slice padded to its longest row with (column 0, weight 0).  A warp then reads its k-th entries as one
// contiguous 128 / 256-byte run; in row-major CSR the same read touched 32 different sectors (rows hold
// 16-32 entries), and those two table reads per entry were the bulk of the kernel's LSU wavefronts.
template <typename T> struct CsrTable {
    const int32_t* row_pixel;   // [n_rows]
    const int64_t* slice_ptr;   // [n_slices + 1] entry offset of every slice (multiples of 32)
    const int32_t* col;         // [slice_ptr[n_slices]]  entry k of row r at slice_ptr[r/32] + 32 k + r % 32;
                                // the value is the element offset n' * g_col + b of the entry in plane 0 of G
    const T* val;               // same layout
    int32_t n_rows;
};

// cube[l, pixel] += sum_e val[e] * Gt[l, col[e]]     (the rows of the cube are zeroed by the caller)
template <typename T, int LB>
__global__ void __launch_bounds__(128)
slit_scatter_kernel(const T* __restrict__ Gt, int g_l /* elements between wavelengths of G */, int n_l, CsrTable<T> t,
                    T* __restrict__ cube, size_t plane) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int slice = r >> 5, lane = r & 31;             // warp-uniform slice: no divergence on the loop bound
    if ((slice << 5) >= t.n_rows) return;
    const int l0 = blockIdx.y * LB;
    const int64_t base = t.slice_ptr[slice];
    const int width = (int)((t.slice_ptr[slice + 1] - base) >> 5);
    T acc[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) acc[u] = T(0);
    const T* g = Gt + (size_t)l0 * g_l;
    const int32_t* col = t.col + base + lane;
    const T* val = t.val + base + lane;
#pragma unroll 4  // 4 entries = 16 dependent G loads in flight (unroll 1: 3.28 ms, 4: 2.59 ms, 8: 2.79 ms on C4)
    for (int k = 0; k < width; ++k) {
        const int c = __ldg(col + 32 * k);
        const T v = __ldg(val + 32 * k);
#pragma unroll
        for (int u = 0; u < LB; ++u)
            if (l0 + u < n_l) acc[u] = fma(v, __ldg(g + (size_t)u * g_l + c), acc[u]);
    }
    if (r >= t.n_rows) return;
    T* dst = cube + (size_t)l0 * plane + t.row_pixel[r];
#pragma unroll
    for (int u = 0; u < LB; ++u)
        if (l0 + u < n_l) dst[(size_t)u * plane] += acc[u];
}

// Zero, in every plane of a chunk, only the row pairs [lo, lo + cnt) that the pruned R2C transform will read
// (the rows outside are never looked at): replaces a memset of the whole working cube.
template <typename T>
__global__ void __launch_bounds__(256)
zero_row_hull_kernel(T* __restrict__ cube, size_t plane, int n_alpha, int n_beta, const int2* __restrict__ pair_range) {
    const int2 pr = pair_range[blockIdx.y];
    const int r0 = 2 * pr.x, r1 = min(n_alpha, r0 + 2 * pr.y);
    const size_t n = (size_t)(r1 - r0) * n_beta;
    T* dst = cube + (size_t)blockIdx.y * plane + (size_t)r0 * n_beta;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = T(0);
}

// No spectral response (MRSBlurred, surfh/Models/spectro_blind.py:191-235): the detector value is the
// beta-sum of the weighted slit, one output row per cube wavelength; the transpose replicates.
//   y[(row0 + l) * Nn + n] = sum_b G[l, n*nb + b]          Gt[l, n*nb + b] = y[(row0 + l) * Nn + n]
// Detector sample n = (p*S + s)*na + a lives at slit-space column ((p*na + a)*S + s)*nb (+ b).
__device__ __forceinline__ size_t slit_col_of_sample(size_t n, int S, int na, int nb) {
    const size_t a = n % na, ps = n / na, s = ps % S, p = ps / S;
    return ((p * na + a) * S + s) * nb;
}

template <typename T>
__global__ void __launch_bounds__(256)
beta_sum_fwd_kernel(const T* __restrict__ G, int n_l, int Nn, int nb, int S, int na, int row0, T* __restrict__ y) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_l * Nn) return;
    const size_t l = idx / Nn, n = idx % Nn;
    const T* g = G + l * ((size_t)Nn * nb) + slit_col_of_sample(n, S, na, nb);
    T s = T(0);
    for (int b = 0; b < nb; ++b) s += g[b];
    y[(size_t)row0 * Nn + idx] = s;
}

template <typename T>
__global__ void __launch_bounds__(256)
beta_sum_adj_kernel(const T* __restrict__ y, int n_l, int Nn, int nb, int S, int na, int row0, T* __restrict__ Gt) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_l * Nn * nb) return;
    const size_t b = idx % nb, ln = idx / nb, l = ln / Nn, n = ln % Nn;
    Gt[l * ((size_t)Nn * nb) + slit_col_of_sample(n, S, na, nb) + b] = y[(size_t)row0 * Nn + ln];
}

}  // namespace surfh

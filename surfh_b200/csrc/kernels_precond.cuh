// kP: Fourier-domain block preconditioner of the normal operator.
//
// The reference's Fourier-domain mixing model (surfh/Models/mixing.py:131-272, `Model_WCT`) builds, per spatial
// frequency f, the K x K Hessian of  C T :  sum_l |OTF_l(f)|^2 T_l T_l^T  (`hess_spec_freq`, :175-207) and
// `Inv_Regul_Fusion_Model3` (surfh/ToolsDir/fusion_mixing.py:401-438, algorithms.make_iHtH_spectro) inverts
// "Hessian + regulariser" bin by bin to solve that model in closed form.  For the MRS operator H = A C T the
// detector sampling A is not shift-invariant, so the same per-frequency inverse is not the solution -- it is a
// preconditioner for the CG of fusion_CT.py:194-232 (what `qmm.lcg(precond=...)` takes):
//     P(f) = ( mu_s * sum_l w_l |OTF_l(f)|^2 T_l T_l^T  +  mu_r * d(f)^p I )^-1 ,
// w_l = the mean gain of A^T A at wavelength l (host-side estimate), d(f) = 4 - 2 cos(2 pi i / Na) - 2 cos(2 pi j / Nb)
// the eigenvalue of the circular 5-point Laplacian D_r^T D_r + D_c^T D_c (p = 1; p = 2 for the joint prior).
//
//   precond_gram_kernel     one pass over the OTF (HBM-bound, like k1^T): the upper triangle of the K x K Gram
//                           blocks, reduced over wavelength in a fixed order (deterministic)
//   precond_factor_kernel   per bin: add the regulariser, Cholesky in registers, store the inverse
//   precond_apply_kernel    zhat[k, f] = scale * sum_k' P[k, k'](f) rhat[k', f]   between two K-map FFTs
#pragma once
#include "common.cuh"

namespace surfh {

constexpr int kGramLanes = 8;

// gram[(k, k') upper][f] = sum_l w[l] |otf[l, f]|^2 tpl[k, l] tpl[k', l]      (double, structure of arrays)
template <typename T, int K>
__global__ void __launch_bounds__(32 * kGramLanes)
precond_gram_kernel(const cplx_t<T>* __restrict__ otf, const T* __restrict__ tpl, int tpl_ld, const double* __restrict__ w,
                    int n_l, size_t nfp, double* __restrict__ gram) {
    using C = cplx_t<T>;
    constexpr int KK = K * (K + 1) / 2;
    __shared__ double red[kGramLanes][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const size_t f = (size_t)blockIdx.x * 32 + tx;
    double a[KK];
#pragma unroll
    for (int e = 0; e < KK; ++e) a[e] = 0.0;
    if (f < nfp) {
        for (int l = ty; l < n_l; l += kGramLanes) {
            const double wl = w[l];
            if (wl == 0.0) continue;
            const C o = ld_stream(otf + (size_t)l * nfp + f);
            const double p = wl * ((double)o.x * (double)o.x + (double)o.y * (double)o.y);
            double t[K];
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = (double)__ldg(tpl + (size_t)k * tpl_ld + l);
            int e = 0;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int k2 = k; k2 < K; ++k2) {
                    a[e] = fma(p * t[k], t[k2], a[e]);
                    ++e;
                }
        }
    }
    // fixed-order reduction over the wavelength lanes, one Gram entry at a time
#pragma unroll
    for (int e = 0; e < KK; ++e) {
        red[ty][tx] = a[e];
        __syncthreads();
        if (ty == 0 && f < nfp) {
            double s = red[0][tx];
#pragma unroll
            for (int j = 1; j < kGramLanes; ++j) s += red[j][tx];
            gram[(size_t)e * nfp + f] = s;
        }
        __syncthreads();
    }
}

// P[(k, k') full][f] = ( mu_s * gram(f) + mu_r * d(f)^power * I )^-1  as T.  `transposed`: bin f = j*na + i
// (the hand-written FFT's spectrum layout), else f = i*nh + j.  Bins >= nf (plane padding) get the identity.
template <typename T, int K>
__global__ void __launch_bounds__(128)
precond_factor_kernel(const double* __restrict__ gram, size_t nf, size_t nfp, int na, int nb, int nh, int transposed,
                      double mu_s, double mu_r, int power, T* __restrict__ pinv) {
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfp) return;
    double a[K][K];
    bool ok = f < nf;
    if (ok) {
        const int i = transposed ? (int)(f % (size_t)na) : (int)(f / (size_t)nh);
        const int j = transposed ? (int)(f / (size_t)na) : (int)(f % (size_t)nh);
        double d = 4.0 - 2.0 * cospi(2.0 * (double)i / (double)na) - 2.0 * cospi(2.0 * (double)j / (double)nb);
        if (power == 2) d *= d;
        int e = 0;
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int k2 = k; k2 < K; ++k2) {
                const double v = mu_s * gram[(size_t)e * nfp + f];
                a[k][k2] = a[k2][k] = v;
                ++e;
            }
#pragma unroll
        for (int k = 0; k < K; ++k) a[k][k] += mu_r * d;
        // in-place Cholesky A = L L^T (lower triangle)
#pragma unroll
        for (int c = 0; c < K; ++c) {
#pragma unroll
            for (int m = 0; m < c; ++m) a[c][c] -= a[c][m] * a[c][m];
            if (!(a[c][c] > 0.0)) ok = false;
            const double dinv = ok ? rsqrt(a[c][c]) : 0.0;
            a[c][c] = ok ? a[c][c] * dinv : 1.0;  // sqrt
#pragma unroll
            for (int r = c + 1; r < K; ++r) {
#pragma unroll
                for (int m = 0; m < c; ++m) a[r][c] -= a[r][m] * a[c][m];
                a[r][c] *= dinv;
            }
        }
    }
    double inv[K][K];
    if (ok) {
        // columns of the inverse by forward / backward substitution
#pragma unroll
        for (int c = 0; c < K; ++c) {
            double y[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                double s = r == c ? 1.0 : 0.0;
#pragma unroll
                for (int m = 0; m < r; ++m) s -= a[r][m] * y[m];
                y[r] = s / a[r][r];
            }
#pragma unroll
            for (int r = K - 1; r >= 0; --r) {
                double s = y[r];
#pragma unroll
                for (int m = r + 1; m < K; ++m) s -= a[m][r] * inv[m][c];
                inv[r][c] = s / a[r][r];
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) inv[r][c] = r == c ? 1.0 : 0.0;
    }
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
        for (int c = 0; c < K; ++c) pinv[(size_t)(r * K + c) * nfp + f] = (T)(0.5 * (inv[r][c] + inv[c][r]));
}

// zhat[k, f] = scale * sum_k' P[k, k'](f) * rhat[k', f]  (in place allowed: every bin is read before it is written)
template <typename T, int K>
__global__ void __launch_bounds__(256)
precond_apply_kernel(const T* __restrict__ pinv, cplx_t<T>* __restrict__ spec, size_t nfp, T scale) {
    using C = cplx_t<T>;
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfp) return;
    C r[K];
#pragma unroll
    for (int k = 0; k < K; ++k) r[k] = spec[(size_t)k * nfp + f];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        C s = make_c<T>(T(0), T(0));
#pragma unroll
        for (int k2 = 0; k2 < K; ++k2) {
            const T p = pinv[(size_t)(k * K + k2) * nfp + f];
            s.x = fma(p, r[k2].x, s.x);
            s.y = fma(p, r[k2].y, s.y);
        }
        spec[(size_t)k * nfp + f] = make_c<T>(s.x * scale, s.y * scale);
    }
}

}  // namespace surfh

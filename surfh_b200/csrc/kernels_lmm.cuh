// k1 / k1^T: fused Fourier-domain template contraction with the wavelength-dependent OTF.
//
//   forward   spec[l, f] = otf[l, f] * sum_k tpl[k, l] * xhat[k, f]
//   adjoint   acc[k, f] (+)= sum_l tpl[k, l] * conj(otf[l, f]) * spec[l, f]
//
// Replaces, fused into one pass each (paths relative to the reference tree):
//   jax_utils.lmm_maps2cube  (surfh/ToolsDir/jax_utils.py:10-15)  + `dft(cube) * self.sotf`
//                            (surfh/Models/spectroModel.py:160-166)
//   `dft_mult(global_cube, sotf.conj())` + jax_utils.lmm_cube2maps
//                            (spectroModel.py:178-181, jax_utils.py:17-26)
// The K maps are Fourier-transformed ONCE (K small FFTs) instead of FFT-ing the L-plane cube,
// which is valid because the LMM is linear and wavelength-separable; the 1/N^2 of the two
// ortho-normalised transforms is folded into `tpl`.
//
// Both kernels are HBM-bound streams over the OTF (and, for the adjoint, the cube spectrum):
// 16 bytes per complex bin per plane in fp64, arithmetic intensity < 1 flop/byte, so no tensor
// cores.  Planes are padded to `nfp` bins (multiple of 16) so every row start is 256-byte aligned.
#pragma once
#include "common.cuh"

namespace surfh {

constexpr int kLmmLsub = 16;  // wavelengths handled by one CTA of the forward kernel

template <typename T, int K>
__global__ void __launch_bounds__(256)
lmm_otf_fwd_kernel(const cplx_t<T>* __restrict__ xhat, const cplx_t<T>* __restrict__ otf,
                   const T* __restrict__ tpl, int tpl_ld, int l_first, int n_l, size_t nfp,
                   cplx_t<T>* __restrict__ spec) {
    using C = cplx_t<T>;
    __shared__ T st[K][kLmmLsub];
    const int l_begin = blockIdx.y * kLmmLsub;
    const int nl = min(kLmmLsub, n_l - l_begin);
    if (threadIdx.x < K * kLmmLsub) {
        const int k = threadIdx.x / kLmmLsub, l = threadIdx.x % kLmmLsub;
        st[k][l] = l < nl ? tpl[(size_t)k * tpl_ld + l_first + l_begin + l] : T(0);
    }
    __syncthreads();
    const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfp) return;
    C xk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xk[k] = xhat[(size_t)k * nfp + f];
    const C* op = otf + (size_t)l_begin * nfp + f;
    C* sp = spec + (size_t)l_begin * nfp + f;
#pragma unroll 4
    for (int l = 0; l < nl; ++l) {
        const C o = ld_stream(op + (size_t)l * nfp);
        C s = make_c<T>(T(0), T(0));
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s.x = fma(st[k][l], xk[k].x, s.x);
            s.y = fma(st[k][l], xk[k].y, s.y);
        }
        sp[(size_t)l * nfp] = cmul(o, s);
    }
}

constexpr int kAdjLanes = 8;  // wavelength lanes of one CTA of the adjoint kernel

template <typename T, int K>
__global__ void __launch_bounds__(32 * kAdjLanes)
lmm_otf_adj_kernel(const cplx_t<T>* __restrict__ spec, const cplx_t<T>* __restrict__ otf,
                   const T* __restrict__ tpl, int tpl_ld, int l_first, int n_l, size_t nfp,
                   cplx_t<T>* __restrict__ acc, int accumulate) {
    using C = cplx_t<T>;
    __shared__ C red[kAdjLanes][K][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const size_t f = (size_t)blockIdx.x * 32 + tx;
    C a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = make_c<T>(T(0), T(0));
    if (f < nfp) {
#pragma unroll 4  // 8 loads in flight per thread: 0.92 of the HBM rate (unroll 2: 0.87, unroll 8: 0.80)
        for (int l = ty; l < n_l; l += kAdjLanes) {
            const C o = ld_stream(otf + (size_t)l * nfp + f);
            const C s = ld_stream(spec + (size_t)l * nfp + f);
            const C v = cmul_conj(o, s);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const T t = __ldg(tpl + (size_t)k * tpl_ld + l_first + l);
                a[k].x = fma(t, v.x, a[k].x);
                a[k].y = fma(t, v.y, a[k].y);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) red[ty][k][tx] = a[k];
    __syncthreads();
    // fixed-order reduction over the wavelength lanes: deterministic
    for (int k = ty; k < K; k += kAdjLanes) {
        if (f < nfp) {
            C s = red[0][k][tx];
#pragma unroll
            for (int j = 1; j < kAdjLanes; ++j) {
                s.x += red[j][k][tx].x;
                s.y += red[j][k][tx].y;
            }
            C* dst = acc + (size_t)k * nfp + f;
            if (accumulate) {
                const C old = *dst;
                s.x += old.x;
                s.y += old.y;
            }
            *dst = s;
        }
    }
}

// No-LMM path (templates=None): spec[l, f] = scale * (CONJ ? conj(otf) : otf)[l, f] * spec[l, f]
template <typename T, bool CONJ>
__global__ void __launch_bounds__(256)
otf_mul_kernel(cplx_t<T>* __restrict__ spec, const cplx_t<T>* __restrict__ otf, size_t n, T scale) {
    using C = cplx_t<T>;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const C o = ld_stream(otf + i);
    const C s = spec[i];
    C r = CONJ ? cmul_conj(o, s) : cmul(o, s);
    r.x *= scale;
    r.y *= scale;
    spec[i] = r;
}

// complex128 -> handle dtype with plane padding (used by surfh_set_otf)
template <typename T>
__global__ void otf_convert_kernel(const double2* __restrict__ src, cplx_t<T>* __restrict__ dst, size_t nf,
                                   size_t nfp, int n_planes, int na, int nh, int transposed) {
    // dst element i of a plane: natural [na][nh] order, or -- the layout of the hand-written FFT's spectra --
    // transposed [nh][na]
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (i >= nfp || l >= n_planes) return;
    cplx_t<T> v = make_c<T>(T(0), T(0));
    if (i < nf) {
        size_t from = i;
        if (transposed) {
            const size_t col = i / (size_t)na, row = i - col * (size_t)na;
            from = row * (size_t)nh + col;
        }
        const double2 s = src[(size_t)l * nf + from];
        v = make_c<T>((T)s.x, (T)s.y);
    }
    dst[(size_t)l * nfp + i] = v;
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ src, TO* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (TO)src[i];
}

// maps -> float32 cube (result export): cube[l, p] = sum_k maps[k, p] * tpl[k, l]
template <typename T, int K>
__global__ void __launch_bounds__(256)
maps_to_cube_kernel(const T* __restrict__ maps, const T* __restrict__ tpl, int tpl_ld, int n_l, size_t npix,
                    float* __restrict__ cube) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float m[K];
#pragma unroll
    for (int k = 0; k < K; ++k) m[k] = (float)maps[(size_t)k * npix + p];
    const int l0 = blockIdx.y * 32;
    for (int l = l0; l < min(l0 + 32, n_l); ++l) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) s += m[k] * (float)__ldg(tpl + (size_t)k * tpl_ld + l);
        cube[(size_t)l * npix + p] = s;
    }
}

}  // namespace surfh

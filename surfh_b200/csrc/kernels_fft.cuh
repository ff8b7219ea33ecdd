// kF: hand-written batched 2-D real FFT pair for the cube planes (sm_100a).
//
// Replaces, for the L-plane cube, the `rfftn` / `irfftn(norm="ortho")` pair of
//   jax_utils.dft / idft          surfh/ToolsDir/jax_utils.py:30-41   (python_utils.py:41-71)
// called at surfh/Models/spectroModel.py:166 (forward) and :178 (adjoint).  The transforms here are
// un-normalised, like cuFFT's; the two ortho factors 1/sqrt(Na*Nb) are folded into the templates.
//
// Why hand-written: the reference's map sizes are 251 (prime) and 501 = 3 x 167, for which cuFFT
// falls back to a multi-kernel Bluestein that runs at ~6 % of the HBM roofline and was 72 % of one
// operator application.  Here each 1-D transform of length N is ONE chirp-z (Bluestein) evaluation
// kept entirely in registers + shared memory:
//     X[k] = a[k] * sum_n (x[n] a[n]) conj(a)[k-n],   a[n] = exp(-i pi n^2 / N)
// i.e. chirp multiply on load -> radix-16 x radix-16 x radix-R3 FFT of length M = 256*R3 >= 2N-1
// (decimation in frequency, output left in digit-reversed order) -> pointwise multiply with the
// precomputed, identically permuted spectrum of the chirp filter -> the mirrored inverse FFT
// (decimation in time, natural order out) -> chirp multiply on store.  No bit-reversal pass exists.
//
// A CTA of 256 threads runs G = 4096/M transforms side by side; a transform is spread over
// TT = M/16 threads that hold 16 points each.  Threads of different transforms are interleaved
// lane-wise (transform = tid % G) so that the G adjacent columns of a column pass are read and
// written as one contiguous segment per row.  Of the 4 register<->shared exchanges of one chirp-z
// only 2 need a CTA barrier; the other 2 stay inside a group of R3 lanes of one warp.
//
// 2-D real transforms use the two-for-one trick: a pair of real rows is transformed as one complex
// row; the pass along the other axis separates / re-assembles the two Hermitian spectra on load, so
// no extra pass or exchange is spent on it.
//   R2C:  rows_r2c  (real [Na][Nb] -> pair spectra Z [ceil(Na/2)][Nb])  ->  cols_r2c (-> spec [Na][Nh])
//   C2R:  cols_c2r  (spec [Na][Nh] -> column-transformed [Na][Nh])      ->  rows_c2r (-> real [Na][Nb])
#pragma once
#include "common.cuh"

namespace surfh {

template <int M> struct FftGeom {
    static_assert(M == 256 || M == 512 || M == 1024 || M == 2048, "chirp-z length must be 256..2048");
    static constexpr int R3 = M / 256;          // last radix: 1, 2, 4, 8
    static constexpr int TT = M / 16;           // threads per transform
    static constexpr int G = 256 / TT;          // transforms per CTA
    static constexpr int TP = 16 * (R3 + 1);    // padded length of one of the 16 sub-sequences
    // per-transform buffer; the tail pad staggers the G buffers over the shared-memory banks
    static constexpr int BUF = 16 * TP + (G == 4 ? 2 : (G == 2 ? 4 : 1));
    static constexpr int HALF = M / 2;          // the input / output length N must be <= HALF
};

template <typename T, int M> constexpr size_t fft_smem_bytes() {
    return (size_t)FftGeom<M>::G * FftGeom<M>::BUF * sizeof(cplx_t<T>);
}

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
// a * (-i) for the forward transform, a * (+i) for the inverse
template <bool INV, typename C> __device__ __forceinline__ C rot90(C a) {
    C r;
    if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; }
    return r;
}
// a * w (forward) or a * conj(w) (inverse)
template <bool INV, typename C> __device__ __forceinline__ C twmul(C a, C w) { return INV ? cmul_conj(w, a) : cmul(w, a); }

template <bool INV, typename C> __device__ __forceinline__ void dft2(C& a0, C& a1) {
    const C s = cadd(a0, a1);
    a1 = csub(a0, a1);
    a0 = s;
}

// 4-point DFT, natural order in and out
template <bool INV, typename C> __device__ __forceinline__ void dft4(C& a0, C& a1, C& a2, C& a3) {
    const C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot90<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// a * exp(-+ i pi e / 8) for the few exponents the 8- and 16-point kernels need
template <bool INV, int E, typename C> __device__ __forceinline__ C mul_w16(C a) {
    using R = decltype(a.x);
    constexpr double Cc = 0.92387953251128673848, Ss = 0.38268343236508978178, Hh = 0.70710678118654752440;
    C r;
    if (E == 0) return a;
    if (E == 4) return rot90<INV>(a);
    if (E == 2) {  // (H, -H)
        if (INV) { r.x = R(Hh) * (a.x - a.y); r.y = R(Hh) * (a.x + a.y); }
        else { r.x = R(Hh) * (a.x + a.y); r.y = R(Hh) * (a.y - a.x); }
        return r;
    }
    if (E == 6) {  // (-H, -H)
        if (INV) { r.x = -R(Hh) * (a.x + a.y); r.y = R(Hh) * (a.x - a.y); }
        else { r.x = R(Hh) * (a.y - a.x); r.y = -R(Hh) * (a.x + a.y); }
        return r;
    }
    C w;
    if (E == 1) { w.x = R(Cc); w.y = R(-Ss); }
    if (E == 3) { w.x = R(Ss); w.y = R(-Cc); }
    if (E == 9) { w.x = R(-Cc); w.y = R(Ss); }
    return twmul<INV>(a, w);
}

// 8-point DFT, natural order in and out
template <bool INV, typename C> __device__ __forceinline__ void dft8(C* v) {
    dft4<INV>(v[0], v[2], v[4], v[6]);  // n = 0: y_q[0] in v[2q]
    dft4<INV>(v[1], v[3], v[5], v[7]);  // n = 1: y_q[1] in v[2q+1]
    v[3] = mul_w16<INV, 2>(v[3]);
    v[5] = mul_w16<INV, 4>(v[5]);
    v[7] = mul_w16<INV, 6>(v[7]);
    dft2<INV>(v[0], v[1]);
    dft2<INV>(v[2], v[3]);
    dft2<INV>(v[4], v[5]);
    dft2<INV>(v[6], v[7]);
    // X[4k + q] sits in v[2q + k]
    const C x1 = v[2], x2 = v[4], x3 = v[6], x4 = v[1], x5 = v[3], x6 = v[5];
    v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6;
}

// 16-point DFT, natural order in and out
template <bool INV, typename C> __device__ __forceinline__ void dft16(C* v) {
#pragma unroll
    for (int n = 0; n < 4; ++n) dft4<INV>(v[n], v[n + 4], v[n + 8], v[n + 12]);  // y_q[n] in v[n + 4q]
    v[5] = mul_w16<INV, 1>(v[5]);
    v[9] = mul_w16<INV, 2>(v[9]);
    v[13] = mul_w16<INV, 3>(v[13]);
    v[6] = mul_w16<INV, 2>(v[6]);
    v[10] = mul_w16<INV, 4>(v[10]);
    v[14] = mul_w16<INV, 6>(v[14]);
    v[7] = mul_w16<INV, 3>(v[7]);
    v[11] = mul_w16<INV, 6>(v[11]);
    v[15] = mul_w16<INV, 9>(v[15]);
#pragma unroll
    for (int q = 0; q < 4; ++q) dft4<INV>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    // X[4k + q] sits in v[4q + k]: transpose the 4 x 4 register tile
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = q + 1; k < 4; ++k) {
            const C tmp = v[4 * q + k];
            v[4 * q + k] = v[4 * k + q];
            v[4 * k + q] = tmp;
        }
}

template <bool INV, int R, typename C> __device__ __forceinline__ void dft_r(C* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    if (R == 8) dft8<INV>(v);
}

// Length-M forward FFT of the sequence held as v[m] = x[t + TT*m]; the result stays in registers in
// a digit-reversed order that only fft_inv() (and the filter table built by the same code) needs to
// know.  `buf` is this transform's shared buffer; every thread of the CTA must call.
template <typename T, int M>
__device__ __forceinline__ void fft_fwd(cplx_t<T>* v, cplx_t<T>* buf, int t, const cplx_t<T>* __restrict__ tw) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    constexpr int R3 = Gm::R3, TP = Gm::TP;
    dft16<false>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        C x = v[q];
        if (q) x = cmul(x, __ldg(tw + t * q));
        buf[q * TP + t] = x;
    }
    __syncthreads();
    const int q = t / R3, n2 = t % R3;
    C* blk = buf + q * TP;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = blk[n2 + R3 * m];
    dft16<false>(v);
    if (R3 > 1) {
        __syncwarp();
#pragma unroll
        for (int q2 = 0; q2 < 16; ++q2) {
            C x = v[q2];
            if (q2) x = cmul(x, __ldg(tw + 16 * n2 * q2));
            blk[(q2 / R3) * (R3 * (R3 + 1)) + n2 * (R3 + 1) + (q2 % R3)] = x;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16 / R3; ++c) {
#pragma unroll
            for (int n = 0; n < R3; ++n) v[c * R3 + n] = blk[c * (R3 * (R3 + 1)) + n * (R3 + 1) + n2];
            dft_r<false, R3>(v + c * R3);
        }
    }
}

// Mirror of fft_fwd: takes the digit-reversed spectrum in registers, returns M * x[t + TT*m] in v[m].
template <typename T, int M>
__device__ __forceinline__ void fft_inv(cplx_t<T>* v, cplx_t<T>* buf, int t, const cplx_t<T>* __restrict__ tw) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    constexpr int R3 = Gm::R3, TP = Gm::TP;
    const int q = t / R3, n2 = t % R3;
    C* blk = buf + q * TP;
    if (R3 > 1) {
        __syncwarp();  // the group's last reads of blk in fft_fwd precede these writes
#pragma unroll
        for (int c = 0; c < 16 / R3; ++c) {
            dft_r<true, R3>(v + c * R3);
#pragma unroll
            for (int n = 0; n < R3; ++n) blk[c * (R3 * (R3 + 1)) + n * (R3 + 1) + n2] = v[c * R3 + n];
        }
        __syncwarp();
#pragma unroll
        for (int q2 = 0; q2 < 16; ++q2) {
            C x = blk[(q2 / R3) * (R3 * (R3 + 1)) + n2 * (R3 + 1) + (q2 % R3)];
            if (q2) x = cmul_conj(__ldg(tw + 16 * n2 * q2), x);
            v[q2] = x;
        }
    }
    dft16<true>(v);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) blk[n2 + R3 * m] = v[m];
    __syncthreads();
#pragma unroll
    for (int qq = 0; qq < 16; ++qq) {
        C x = buf[qq * TP + t];
        if (qq) x = cmul_conj(__ldg(tw + t * qq), x);
        v[qq] = x;
    }
    dft16<true>(v);
}

// Device tables of one 1-D chirp-z plan (length N through M-point FFTs).
template <typename T> struct FftPlan1d {
    const cplx_t<T>* chirp;  // [N]   a[n] = exp(-i pi n^2 / N)
    const cplx_t<T>* filt;   // [M]   FFT_M(conj(a) wrapped) / M, in fft_fwd's register order [j*TT + t]
    const cplx_t<T>* tw;     // [M]   exp(-2 pi i j / M)
    int n;
};

// Circular convolution with the chirp filter: v[m] = (u * conj(a))[t + TT*m], u given the same way.
template <typename T, int M>
__device__ __forceinline__ void chirp_convolve(cplx_t<T>* v, cplx_t<T>* buf, int t, const FftPlan1d<T>& p) {
    constexpr int TT = FftGeom<M>::TT;
    fft_fwd<T, M>(v, buf, t, p.tw);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], __ldg(p.filt + j * TT + t));
    fft_inv<T, M>(v, buf, t, p.tw);
}

// Builds FftPlan1d::filt from the natural-order filter `b` (already scaled by 1/M) with the very code
// that consumes it, so the digit-reversed order never has to be spelled out.  One CTA.
template <typename T, int M>
__global__ void __launch_bounds__(256)
fft_filter_kernel(const cplx_t<T>* __restrict__ b, const cplx_t<T>* __restrict__ tw, cplx_t<T>* __restrict__ filt) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = b[t + Gm::TT * m];
    fft_fwd<T, M>(v, buf, t, tw);
    if (g == 0)
#pragma unroll
        for (int j = 0; j < 16; ++j) filt[j * Gm::TT + t] = v[j];
}

// Plain length-M FFT round trip / forward transform in natural order, for the self-test only:
// out[f] for the register order is recovered by transforming unit impulses on the host side.
template <typename T, int M>
__global__ void __launch_bounds__(256)
fft_selftest_kernel(const cplx_t<T>* __restrict__ in, const cplx_t<T>* __restrict__ tw, cplx_t<T>* __restrict__ fwd_regs,
                    cplx_t<T>* __restrict__ roundtrip) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = in[t + Gm::TT * m];
    fft_fwd<T, M>(v, buf, t, tw);
    if (g == 0)
#pragma unroll
        for (int j = 0; j < 16; ++j) fwd_regs[j * Gm::TT + t] = v[j];
    fft_inv<T, M>(v, buf, t, tw);
    if (g == 0)
#pragma unroll
        for (int m = 0; m < 16; ++m) roundtrip[t + Gm::TT * m] = v[m];
}

struct FftShape {
    int na, nb, nh;          // rows, columns, nb/2+1
    int npair;               // ceil(na / 2)
    int zpitch;              // row pitch (complex) of the pair-spectrum buffer Z, >= nb
    size_t real_plane;       // elements between real planes
    size_t spec_plane;       // complex elements between spectrum planes ([na][nh], row pitch nh)
    size_t z_plane;          // complex elements between planes of Z / of the column-transformed buffer
    int batch;
};

// ---- R2C pass 1: pairs of real rows -> full complex spectrum of (row_even + i row_odd)
template <typename T, int M>
__global__ void __launch_bounds__(256, 2)
fft_rows_r2c_kernel(const T* __restrict__ in, cplx_t<T>* __restrict__ z, FftShape s, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    const long long item = (long long)blockIdx.x * Gm::G + g;
    const bool live = item < (long long)s.batch * s.npair;
    const int plane = live ? (int)(item / s.npair) : 0, pair = live ? (int)(item % s.npair) : 0;
    const int r0 = 2 * pair, r1 = r0 + 1;
    const T* ra = in + (size_t)plane * s.real_plane + (size_t)r0 * s.nb;
    const T* rb = ra + s.nb;
    const bool has_b = r1 < s.na;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + Gm::TT * m;
        C u = make_c<T>(T(0), T(0));
        if (m < 8 && live && n < s.nb) {
            u.x = ra[n];
            u.y = has_b ? rb[n] : T(0);
            u = cmul(u, __ldg(p.chirp + n));
        }
        v[m] = u;
    }
    chirp_convolve<T, M>(v, buf, t, p);
    if (!live) return;
    C* dst = z + (size_t)plane * s.z_plane + (size_t)pair * s.zpitch;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int n = t + Gm::TT * m;
        if (n < s.nb) dst[n] = cmul(v[m], __ldg(p.chirp + n));
    }
}

// ---- R2C pass 2: column transforms; the two Hermitian row spectra are separated on load
template <typename T, int M>
__global__ void __launch_bounds__(256, 2)
fft_cols_r2c_kernel(const cplx_t<T>* __restrict__ z, cplx_t<T>* __restrict__ spec, FftShape s, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    const int tiles = (s.nh + Gm::G - 1) / Gm::G;
    const int plane = blockIdx.x / tiles, j = (blockIdx.x % tiles) * Gm::G + g;
    const bool live = j < s.nh;
    const int jm = live ? (j == 0 ? 0 : s.nb - j) : 0;
    const C* zp = z + (size_t)plane * s.z_plane;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = t + Gm::TT * m;
        C u = make_c<T>(T(0), T(0));
        if (m < 8 && live && i < s.na) {
            const C* row = zp + (size_t)(i >> 1) * s.zpitch;
            const C a = row[j], b = row[jm];
            if (i & 1) { u.x = T(0.5) * (a.y + b.y); u.y = T(0.5) * (b.x - a.x); }
            else { u.x = T(0.5) * (a.x + b.x); u.y = T(0.5) * (a.y - b.y); }
            u = cmul(u, __ldg(p.chirp + i));
        }
        v[m] = u;
    }
    chirp_convolve<T, M>(v, buf, t, p);
    if (!live) return;
    C* dst = spec + (size_t)plane * s.spec_plane + j;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int i = t + Gm::TT * m;
        if (i < s.na) dst[(size_t)i * s.nh] = cmul(v[m], __ldg(p.chirp + i));
    }
}

// ---- C2R pass 1: inverse column transforms (conj in, conj out around the forward chirp-z)
template <typename T, int M>
__global__ void __launch_bounds__(256, 2)
fft_cols_c2r_kernel(const cplx_t<T>* __restrict__ spec, cplx_t<T>* __restrict__ w, FftShape s, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    const int tiles = (s.nh + Gm::G - 1) / Gm::G;
    const int plane = blockIdx.x / tiles, j = (blockIdx.x % tiles) * Gm::G + g;
    const bool live = j < s.nh;
    const C* src = spec + (size_t)plane * s.spec_plane + j;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int i = t + Gm::TT * m;
        C u = make_c<T>(T(0), T(0));
        if (m < 8 && live && i < s.na) {
            u = src[(size_t)i * s.nh];
            u.y = -u.y;
            u = cmul(u, __ldg(p.chirp + i));
        }
        v[m] = u;
    }
    chirp_convolve<T, M>(v, buf, t, p);
    if (!live) return;
    C* dst = w + (size_t)plane * s.z_plane + j;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int i = t + Gm::TT * m;
        if (i < s.na) {
            C r = cmul(v[m], __ldg(p.chirp + i));
            r.y = -r.y;
            dst[(size_t)i * s.nh] = r;
        }
    }
}

// ---- C2R pass 2: pairs of Hermitian half-rows -> pairs of real rows
template <typename T, int M>
__global__ void __launch_bounds__(256, 2)
fft_rows_c2r_kernel(const cplx_t<T>* __restrict__ w, T* __restrict__ out, FftShape s, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using Gm = FftGeom<M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int g = threadIdx.x % Gm::G, t = threadIdx.x / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    const long long item = (long long)blockIdx.x * Gm::G + g;
    const bool live = item < (long long)s.batch * s.npair;
    const int plane = live ? (int)(item / s.npair) : 0, pair = live ? (int)(item % s.npair) : 0;
    const int r0 = 2 * pair;
    const bool has_b = r0 + 1 < s.na;
    const C* wa = w + (size_t)plane * s.z_plane + (size_t)r0 * s.nh;
    const C* wb = wa + s.nh;
    const int nyq = (s.nb & 1) ? -1 : s.nb / 2;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + Gm::TT * m;
        C u = make_c<T>(T(0), T(0));
        if (m < 8 && live && n < s.nb) {
            // z[n] = A[n] + i B[n] with A, B extended by Hermitian symmetry; conj(z) feeds the forward chirp-z
            const bool mir = n >= s.nh;
            const int k = mir ? s.nb - n : n;
            C a = wa[k];
            C b = has_b ? wb[k] : make_c<T>(T(0), T(0));
            if (k == 0 || k == nyq) { a.y = T(0); b.y = T(0); }
            if (mir) { a.y = -a.y; b.y = -b.y; }
            u.x = a.x - b.y;
            u.y = -(a.y + b.x);
            u = cmul(u, __ldg(p.chirp + n));
        }
        v[m] = u;
    }
    chirp_convolve<T, M>(v, buf, t, p);
    if (!live) return;
    T* oa = out + (size_t)plane * s.real_plane + (size_t)r0 * s.nb;
    T* ob = oa + s.nb;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int n = t + Gm::TT * m;
        if (n < s.nb) {
            const C r = cmul(v[m], __ldg(p.chirp + n));
            oa[n] = r.x;
            if (has_b) ob[n] = -r.y;
        }
    }
}

}  // namespace surfh

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

#ifndef SURFH_FFT_SHFL
#define SURFH_FFT_SHFL 1
#endif

namespace surfh {

template <int M, int NT, int W> struct FftGeom {
    static_assert(M == 256 || M == 512 || M == 1024 || M == 2048, "chirp-z length must be 256..2048");
    static constexpr int R3 = M / 256;          // last radix: 1, 2, 4, 8
    static constexpr int TT = M / 16;           // threads per transform
    static_assert(NT % TT == 0 && NT >= TT, "CTA size must be a multiple of the threads per transform");
    static constexpr int G = NT / TT;           // transforms per CTA
    // length of one of the 16 sub-sequences in shared memory; the shared-memory flavour of the
    // radix-R3 exchange needs a padded transposed staging area, the shuffled one does not
    static constexpr bool SHFL = SURFH_FFT_SHFL && R3 * G <= 32;
    static constexpr int TP = SHFL ? 16 * R3 : 16 * (R3 + 1);
    // per-transform buffer; the tail pad staggers the G buffers over the shared-memory banks: a wavefront
    // serves W lanes (8 complex doubles or 16 complex floats = 128 bytes), i.e. W/G consecutive t of G
    // transforms, which must fall into W distinct element slots modulo W
    static constexpr int PAD = G >= W ? 1 : (W / G) % W;
    static constexpr int BUF = 16 * TP + PAD;
    static constexpr int HALF = M / 2;          // the input / output length N must be <= HALF
};

// CTA shape per arithmetic type: 16 complex doubles per thread need ~160 registers to stay out of
// local memory, so fp64 runs 3 CTAs of 128 threads per SM; fp32 fits 2 CTAs of 256 threads.
template <typename T> struct FftCta;
template <> struct FftCta<double> { static constexpr int NT = 128, MINB = 3; };
template <> struct FftCta<float> { static constexpr int NT = 256, MINB = 2; };
template <int M, int NT> constexpr int fft_cta_threads() { return NT < M / 16 ? M / 16 : NT; }
template <typename T, int M> struct FftK {
    static constexpr int NT = fft_cta_threads<M, FftCta<T>::NT>();
    static constexpr int MINB = FftCta<T>::MINB;
    using Gm = FftGeom<M, NT, 128 / (2 * (int)sizeof(T))>;
};

template <typename T, int M> __host__ __device__ constexpr size_t fft_smem_bytes() {
    return ((size_t)FftK<T, M>::Gm::G * FftK<T, M>::Gm::BUF * sizeof(cplx_t<T>) + 15) / 16 * 16;
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

// 16-point DFT whose inputs v[8..15] are known to be zero (the zero padding of the chirp-z input)
template <bool INV, typename C> __device__ __forceinline__ void dft16_in8(C* v) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const C a0 = v[n], a1 = v[n + 4], r = rot90<INV>(a1);
        v[n] = cadd(a0, a1);
        v[n + 4] = cadd(a0, r);
        v[n + 8] = csub(a0, a1);
        v[n + 12] = csub(a0, r);
    }
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
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = q + 1; k < 4; ++k) {
            const C tmp = v[4 * q + k];
            v[4 * q + k] = v[4 * k + q];
            v[4 * k + q] = tmp;
        }
}

// 16-point DFT of which only the outputs X[0..7] are wanted (left in v[0..7]; v[8..15] are garbage)
template <bool INV, typename C> __device__ __forceinline__ void dft16_out8(C* v) {
#pragma unroll
    for (int n = 0; n < 4; ++n) dft4<INV>(v[n], v[n + 4], v[n + 8], v[n + 12]);
    v[5] = mul_w16<INV, 1>(v[5]);
    v[9] = mul_w16<INV, 2>(v[9]);
    v[13] = mul_w16<INV, 3>(v[13]);
    v[6] = mul_w16<INV, 2>(v[6]);
    v[10] = mul_w16<INV, 4>(v[10]);
    v[14] = mul_w16<INV, 6>(v[14]);
    v[7] = mul_w16<INV, 3>(v[7]);
    v[11] = mul_w16<INV, 6>(v[11]);
    v[15] = mul_w16<INV, 9>(v[15]);
    C o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // X[q + 4k] for k = 0, 1
        const C t0 = cadd(v[4 * q], v[4 * q + 2]), t1 = csub(v[4 * q], v[4 * q + 2]);
        const C t2 = cadd(v[4 * q + 1], v[4 * q + 3]), t3 = rot90<INV>(csub(v[4 * q + 1], v[4 * q + 3]));
        o[q] = cadd(t0, t2);
        o[q + 4] = cadd(t1, t3);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = o[k];
}

template <bool INV, int R, typename C> __device__ __forceinline__ void dft_r(C* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    if (R == 8) dft8<INV>(v);
}

template <typename C> __device__ __forceinline__ C shfl_xor_c(C a, int lane_mask) {
    a.x = __shfl_xor_sync(0xffffffffu, a.x, lane_mask);
    a.y = __shfl_xor_sync(0xffffffffu, a.y, lane_mask);
    return a;
}

// In-register transpose among the R3 threads n2 = 0..R3-1 of one group (lane stride G): for every block
// c of R3 registers, thread n2 ends up with v[c*R3 + n] = (thread n's v[c*R3 + n2]).  Self-inverse.
// Replaces a shared-memory round trip (16 stores + 16 loads of 16 bytes) by 8 log2(R3) shuffled values.
template <int R3, int G, typename C> __device__ __forceinline__ void group_transpose(C* v, int n2) {
#pragma unroll
    for (int s = R3 / 2; s >= 1; s >>= 1) {
        const bool up = (n2 & s) != 0;
#pragma unroll
        for (int c = 0; c < 16 / R3; ++c)
#pragma unroll
            for (int j = 0; j < R3; ++j) {
                if (j & s) continue;
                C& lo = v[c * R3 + j];
                C& hi = v[c * R3 + (j | s)];
                const C send = up ? lo : hi;
                const C recv = shfl_xor_c(send, s * G);
                if (up) lo = recv; else hi = recv;
            }
    }
}

// Length-M forward FFT of the sequence held as v[m] = x[t + TT*m]; the result stays in registers in
// a digit-reversed order that only fft_inv() (and the filter table built by the same code) needs to
// know.  `buf` is this transform's shared buffer; every thread of the CTA must call.
template <typename T, int M, bool HALF_IN = false>
__device__ __forceinline__ void fft_fwd(cplx_t<T>* v, cplx_t<T>* buf, int t, const cplx_t<T>* __restrict__ tw) {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    constexpr int R3 = Gm::R3, TP = Gm::TP;
    if (HALF_IN) dft16_in8<false>(v);  // v[8..15] are the zero padding
    else dft16<false>(v);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        C x = v[q];
        if (q) x = cmul(x, __ldg(tw + q * Gm::TT + t));
        buf[q * TP + t] = x;
    }
    __syncthreads();
    const int q = t / R3, n2 = t % R3;
    C* blk = buf + q * TP;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = blk[n2 + R3 * m];
    dft16<false>(v);
    if (R3 > 1) {
        if (Gm::SHFL) {
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) v[q2] = cmul(v[q2], __ldg(tw + M + q2 * R3 + n2));
            group_transpose<R3, Gm::G>(v, n2);
#pragma unroll
            for (int c = 0; c < 16 / R3; ++c) dft_r<false, R3>(v + c * R3);
        } else {
            __syncwarp();
#pragma unroll
            for (int q2 = 0; q2 < 16; ++q2) {
                C x = v[q2];
                if (q2) x = cmul(x, __ldg(tw + M + q2 * R3 + n2));
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
}

// Mirror of fft_fwd: takes the digit-reversed spectrum in registers, returns M * x[t + TT*m] in v[m].
template <typename T, int M, bool HALF_OUT = false>
__device__ __forceinline__ void fft_inv(cplx_t<T>* v, cplx_t<T>* buf, int t, const cplx_t<T>* __restrict__ tw) {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    constexpr int R3 = Gm::R3, TP = Gm::TP;
    const int q = t / R3, n2 = t % R3;
    C* blk = buf + q * TP;
    if (R3 > 1) {
        if (Gm::SHFL) {
#pragma unroll
            for (int c = 0; c < 16 / R3; ++c) dft_r<true, R3>(v + c * R3);
            group_transpose<R3, Gm::G>(v, n2);
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) v[q2] = cmul_conj(__ldg(tw + M + q2 * R3 + n2), v[q2]);
        } else {
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
                if (q2) x = cmul_conj(__ldg(tw + M + q2 * R3 + n2), x);
                v[q2] = x;
            }
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
        if (qq) x = cmul_conj(__ldg(tw + qq * Gm::TT + t), x);
        v[qq] = x;
    }
    if (HALF_OUT) dft16_out8<true>(v);  // only x[t + TT*m], m < 8, is wanted
    else dft16<true>(v);
}

// Device tables of one 1-D chirp-z plan (length N through M-point FFTs).
template <typename T> struct FftPlan1d {
    const cplx_t<T>* chirp;  // [N]   a[n] = exp(-i pi n^2 / N)
    const cplx_t<T>* filt;   // [M]   FFT_M(conj(a) wrapped) / M, in fft_fwd's register order [j*TT + t]
    const cplx_t<T>* tw;     // [M + 16*R3] twiddles laid out for coalesced reads:
                             //   tw[q*TT + t] = exp(-2 pi i t q / M), then tw[M + q2*R3 + n2] = exp(-2 pi i n2 q2 / TT)
    int n;
};

// Circular convolution with the chirp filter: v[m] = (u * conj(a))[t + TT*m] for m < 8, with u given
// the same way and u[t + TT*m] = 0 for m >= 8 (v[8..15] are ignored on entry, garbage on exit).
template <typename T, int M>
__device__ __forceinline__ void chirp_convolve(cplx_t<T>* v, cplx_t<T>* buf, int t, const FftPlan1d<T>& p) {
    constexpr int TT = FftK<T, M>::Gm::TT;
    fft_fwd<T, M, true>(v, buf, t, p.tw);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], __ldg(p.filt + j * TT + t));
    fft_inv<T, M, true>(v, buf, t, p.tw);
}

// Builds FftPlan1d::filt from the natural-order filter `b` (already scaled by 1/M) with the very code
// that consumes it, so the digit-reversed order never has to be spelled out.  One CTA.
template <typename T, int M>
__global__ void __launch_bounds__(FftK<T, M>::NT)
fft_filter_kernel(const cplx_t<T>* __restrict__ b, const cplx_t<T>* __restrict__ tw, cplx_t<T>* __restrict__ filt) {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
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

struct FftShape {
    int na, nb, nh;          // rows, columns, nb/2+1
    int npair;               // ceil(na / 2)
    int zpitch;              // row pitch (complex) of the pair-spectrum buffer Z, >= nb
    size_t real_plane;       // elements between real planes
    size_t spec_plane;       // complex elements between spectrum planes ([na][nh], row pitch nh)
    size_t z_plane;          // complex elements between planes of Z / of the column-transformed buffer
    int batch;
};

// ---- asynchronous staging of the next item's inputs -------------------------------------------
// Every thread copies exactly the elements it will itself consume into thread-private shared-memory
// slots, so a cp.async.wait_group is all the synchronisation the staging needs.
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async moves 4, 8 or 16 bytes");
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Decomposition of a work item into (plane, index within the plane) for the two kinds of pass.
struct FftItem {
    bool live;
    int plane, idx;
};

// ---- R2C pass 1: pairs of real rows -> full complex spectrum of (row_even + i row_odd)
template <typename T, int M> struct RowsR2C {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    using Elem = T;                       // staged element
    static constexpr int SLOTS = 16;      // staged elements per thread: 8 points x 2 rows
    const T* in;
    C* z;
    FftShape s;
    __device__ long long items() const { return ((long long)s.batch * s.npair + Gm::G - 1) / Gm::G; }  // < 2^31 (checked on the host)
    __device__ FftItem item(long long blk, int g) const {
        const long long it = blk * Gm::G + g;
        FftItem r;
        r.live = it < (long long)s.batch * s.npair;
        r.plane = r.live ? (int)(it / s.npair) : 0;
        r.idx = r.live ? (int)(it % s.npair) : 0;
        return r;
    }
    __device__ void prefetch(const FftItem& it, int t, Elem* stage, int nt, int tid) const {
        if (!it.live) return;
        const int r0 = 2 * it.idx;
        const T* ra = in + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        const bool has_b = r0 + 1 < s.na;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            if (n < s.nb) {
                cp_async<sizeof(T)>(stage + (2 * m) * nt + tid, ra + n);
                if (has_b) cp_async<sizeof(T)>(stage + (2 * m + 1) * nt + tid, ra + s.nb + n);
            }
        }
    }
    __device__ void load(const FftItem& it, int t, const Elem* stage, int nt, int tid, C* v, const C* chirp) const {
        const bool has_b = 2 * it.idx + 1 < s.na;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && n < s.nb) {
                u.x = stage[(2 * m) * nt + tid];
                u.y = has_b ? stage[(2 * m + 1) * nt + tid] : T(0);
                u = cmul(u, __ldg(chirp + n));
            }
            v[m] = u;
        }
    }
    __device__ void store(const FftItem& it, int t, const C* v, const C* chirp) const {
        if (!it.live) return;
        C* dst = z + (size_t)it.plane * s.z_plane + (size_t)it.idx * s.zpitch;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            if (n < s.nb) dst[n] = cmul(v[m], __ldg(chirp + n));
        }
    }
};

// ---- R2C pass 2: column transforms; the two Hermitian row spectra are separated on load
template <typename T, int M> struct ColsR2C {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    using Elem = C;
    static constexpr int SLOTS = 16;      // 8 points x (Z[j], Z[nb - j])
    const C* z;
    C* spec;
    FftShape s;
    __device__ int tiles() const { return (s.nh + Gm::G - 1) / Gm::G; }
    __device__ long long items() const { return (long long)s.batch * tiles(); }
    __device__ FftItem item(long long blk, int g) const {
        const int tl = tiles();
        FftItem r;
        r.plane = (int)(blk / tl);
        r.idx = (int)(blk % tl) * Gm::G + g;
        r.live = r.idx < s.nh;
        return r;
    }
    __device__ void prefetch(const FftItem& it, int t, Elem* stage, int nt, int tid) const {
        if (!it.live) return;
        const int j = it.idx, jm = j == 0 ? 0 : s.nb - j;
        const C* zp = z + (size_t)it.plane * s.z_plane;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            if (i < s.na) {
                const C* row = zp + (size_t)(i >> 1) * s.zpitch;
                cp_async<sizeof(C)>(stage + (2 * m) * nt + tid, row + j);
                cp_async<sizeof(C)>(stage + (2 * m + 1) * nt + tid, row + jm);
            }
        }
    }
    __device__ void load(const FftItem& it, int t, const Elem* stage, int nt, int tid, C* v, const C* chirp) const {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && i < s.na) {
                const C a = stage[(2 * m) * nt + tid], b = stage[(2 * m + 1) * nt + tid];
                if (i & 1) { u.x = T(0.5) * (a.y + b.y); u.y = T(0.5) * (b.x - a.x); }
                else { u.x = T(0.5) * (a.x + b.x); u.y = T(0.5) * (a.y - b.y); }
                u = cmul(u, __ldg(chirp + i));
            }
            v[m] = u;
        }
    }
    __device__ void store(const FftItem& it, int t, const C* v, const C* chirp) const {
        if (!it.live) return;
        C* dst = spec + (size_t)it.plane * s.spec_plane + it.idx;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            if (i < s.na) dst[(size_t)i * s.nh] = cmul(v[m], __ldg(chirp + i));
        }
    }
};

// ---- C2R pass 1: inverse column transforms (conj in, conj out around the forward chirp-z)
template <typename T, int M> struct ColsC2R {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    using Elem = C;
    static constexpr int SLOTS = 8;
    const C* spec;
    C* w;
    FftShape s;
    __device__ int tiles() const { return (s.nh + Gm::G - 1) / Gm::G; }
    __device__ long long items() const { return (long long)s.batch * tiles(); }
    __device__ FftItem item(long long blk, int g) const {
        const int tl = tiles();
        FftItem r;
        r.plane = (int)(blk / tl);
        r.idx = (int)(blk % tl) * Gm::G + g;
        r.live = r.idx < s.nh;
        return r;
    }
    __device__ void prefetch(const FftItem& it, int t, Elem* stage, int nt, int tid) const {
        if (!it.live) return;
        const C* src = spec + (size_t)it.plane * s.spec_plane + it.idx;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            if (i < s.na) cp_async<sizeof(C)>(stage + m * nt + tid, src + (size_t)i * s.nh);
        }
    }
    __device__ void load(const FftItem& it, int t, const Elem* stage, int nt, int tid, C* v, const C* chirp) const {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && i < s.na) {
                u = stage[m * nt + tid];
                u.y = -u.y;
                u = cmul(u, __ldg(chirp + i));
            }
            v[m] = u;
        }
    }
    __device__ void store(const FftItem& it, int t, const C* v, const C* chirp) const {
        if (!it.live) return;
        C* dst = w + (size_t)it.plane * s.z_plane + it.idx;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + Gm::TT * m;
            if (i < s.na) {
                C r = cmul(v[m], __ldg(chirp + i));
                r.y = -r.y;
                dst[(size_t)i * s.nh] = r;
            }
        }
    }
};

// ---- C2R pass 2: pairs of Hermitian half-rows -> pairs of real rows
template <typename T, int M> struct RowsC2R {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    using Elem = C;
    static constexpr int SLOTS = 16;      // 8 points x 2 rows
    const C* w;
    T* out;
    FftShape s;
    __device__ long long items() const { return ((long long)s.batch * s.npair + Gm::G - 1) / Gm::G; }  // < 2^31 (checked on the host)
    __device__ FftItem item(long long blk, int g) const {
        const long long it = blk * Gm::G + g;
        FftItem r;
        r.live = it < (long long)s.batch * s.npair;
        r.plane = r.live ? (int)(it / s.npair) : 0;
        r.idx = r.live ? (int)(it % s.npair) : 0;
        return r;
    }
    __device__ void prefetch(const FftItem& it, int t, Elem* stage, int nt, int tid) const {
        if (!it.live) return;
        const int r0 = 2 * it.idx;
        const bool has_b = r0 + 1 < s.na;
        const C* wa = w + (size_t)it.plane * s.z_plane + (size_t)r0 * s.nh;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            if (n < s.nb) {
                const int k = n >= s.nh ? s.nb - n : n;
                cp_async<sizeof(C)>(stage + (2 * m) * nt + tid, wa + k);
                if (has_b) cp_async<sizeof(C)>(stage + (2 * m + 1) * nt + tid, wa + s.nh + k);
            }
        }
    }
    __device__ void load(const FftItem& it, int t, const Elem* stage, int nt, int tid, C* v, const C* chirp) const {
        const bool has_b = 2 * it.idx + 1 < s.na;
        const int nyq = (s.nb & 1) ? -1 : s.nb / 2;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && n < s.nb) {
                // z[n] = A[n] + i B[n] with A, B extended by Hermitian symmetry; conj(z) feeds the forward chirp-z
                const bool mir = n >= s.nh;
                const int k = mir ? s.nb - n : n;
                C a = stage[(2 * m) * nt + tid];
                C b = has_b ? stage[(2 * m + 1) * nt + tid] : make_c<T>(T(0), T(0));
                if (k == 0 || k == nyq) { a.y = T(0); b.y = T(0); }
                if (mir) { a.y = -a.y; b.y = -b.y; }
                u.x = a.x - b.y;
                u.y = -(a.y + b.x);
                u = cmul(u, __ldg(chirp + n));
            }
            v[m] = u;
        }
    }
    __device__ void store(const FftItem& it, int t, const C* v, const C* chirp) const {
        if (!it.live) return;
        const int r0 = 2 * it.idx;
        const bool has_b = r0 + 1 < s.na;
        T* oa = out + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        T* ob = oa + s.nb;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + Gm::TT * m;
            if (n < s.nb) {
                const C r = cmul(v[m], __ldg(chirp + n));
                oa[n] = r.x;
                if (has_b) ob[n] = -r.y;
            }
        }
    }
};

template <typename T, int M, typename Pass> __host__ __device__ constexpr size_t fft_pass_smem_bytes() {
    return fft_smem_bytes<T, M>() + (size_t)Pass::SLOTS * FftK<T, M>::NT * sizeof(typename Pass::Elem);
}

// Persistent kernel shared by the four passes: a CTA walks the work items with a grid stride; the inputs
// of item i+1 are in flight (cp.async into thread-private slots) while item i is transformed, which is
// what hides the HBM latency at 12 warps per SM.
template <typename T, int M, typename Pass>
__global__ void __launch_bounds__(FftK<T, M>::NT, FftK<T, M>::MINB) fft_pass_kernel(Pass pass, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using Gm = typename FftK<T, M>::Gm;
    using Elem = typename Pass::Elem;
    constexpr int NT = FftK<T, M>::NT;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    const int tid = threadIdx.x, g = tid % Gm::G, t = tid / Gm::G;
    C* buf = reinterpret_cast<C*>(fft_smem) + g * Gm::BUF;
    Elem* stage = reinterpret_cast<Elem*>(fft_smem + fft_smem_bytes<T, M>());
    const int n_items = (int)pass.items();
    int blk = blockIdx.x;
    if (blk >= n_items) return;
    pass.prefetch(pass.item(blk, g), t, stage, NT, tid);
    cp_async_commit();
    for (; blk < n_items; blk += gridDim.x) {
        C v[16];
        const FftItem cur = pass.item(blk, g);
        cp_async_wait_all();
        pass.load(cur, t, stage, NT, tid, v, p.chirp);
        if (blk + (int)gridDim.x < n_items) pass.prefetch(pass.item(blk + gridDim.x, g), t, stage, NT, tid);
        cp_async_commit();
        chirp_convolve<T, M>(v, buf, t, p);
        pass.store(pass.item(blk, g), t, v, p.chirp);
    }
}

}  // namespace surfh

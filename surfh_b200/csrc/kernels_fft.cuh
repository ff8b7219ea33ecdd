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
// Execution model: ONE persistent CTA per SM walks the work items with a grid stride.
//   * All tables (twiddles, filter spectrum, chirp) live in shared memory for the CTA's lifetime, so
//     every table read has shared-memory latency; the only global traffic is the data itself.
//   * The inputs of item i+1 are copied global -> shared by cp.async into thread-private slots while
//     item i is transformed: HBM latency is off the critical path at 12 warps per SM.
//   * A transform is spread over TT = M/16 threads that hold 16 points each; a CTA runs G transforms
//     side by side, split into independent groups that synchronise with their own named barrier
//     so that one group's shared-memory phases overlap the others' FP64 phases.
//   * Of the 4 register<->register exchanges of one chirp-z, 2 go through shared memory (one barrier
//     each), 2 are transposes among R3 adjacent lanes done with warp shuffles.
//
// 2-D real transforms use the two-for-one trick: a pair of real rows is transformed as one complex row.
//   R2C:  rows_r2c  real [Na][Nb] -> A/B-separated half spectra Y [Na][Nh]   (separation through the
//                   transform's own shared buffer)            ->  cols  (-> spec [Na][Nh])
//   C2R:  cols_c2r  spec [Na][Nh] -> Z [ceil(Na/2)][Nb], Z[p] = W[2p] + i W[2p+1] Hermitian-extended
//                   (rows 2p, 2p+1 sit in adjacent lanes: pairing by shuffle) ->  rows_c2r (-> real [Na][Nb])
// In the operator the rows passes only visit the row pairs some band's field of view touches (FftRanges).
#pragma once
#include "common.cuh"

// build-time switches for A/B measurements (python -c "build.build(out=..., defines=[...])")
#ifndef SURFH_FFT_TWIDDLE_CHAIN
#define SURFH_FFT_TWIDDLE_CHAIN 1
#endif
#ifndef SURFH_FFT_GROUPS
#define SURFH_FFT_GROUPS 3
#endif

namespace surfh {

// Geometry of one chirp-z length M for arithmetic type T.
template <typename T, int M> struct FftK {
    static_assert(M == 256 || M == 512 || M == 1024 || M == 2048, "chirp-z length must be 256..2048");
    using C = cplx_t<T>;
    static constexpr int R3 = M / 256;          // last radix: 1, 2, 4, 8
    static constexpr int TT = M / 16;           // threads per transform
    // 16 complex doubles per thread need ~168 registers to stay out of local memory: 384 threads per SM
    // in fp64 (256 for M = 2048, whose tables are twice as large); fp32 runs 512 threads at 128 registers
    static constexpr int NT = sizeof(T) == 8 ? (M == 2048 ? 256 : 384) : 512;
    static constexpr int G = NT / TT;           // transforms per CTA step
    // independently synchronised groups (named barriers 1..NH): one group's shared-memory phases overlap
    // the others' FP64 phases.  Measured at N = 501 fp64 (ms per 512 planes, R2C / C2R): 2 groups 2.39 / 2.40,
    // 3 groups 2.17 / 2.36, 6 groups (one transform each) 2.96 / 3.08 -- the transforms interleaved lane-wise
    // inside a group keep the column accesses in contiguous runs, so fewer, fatter groups win.
    static constexpr int pick_groups() {
        int best = 1;
        for (int nh = 1; nh <= SURFH_FFT_GROUPS && nh <= G; ++nh)
            if (G % nh == 0 && (NT / nh) % 32 == 0) best = nh;
        return best;
    }
    static constexpr int NH = pick_groups();
    static constexpr int HT = NT / NH;          // threads per group
    static constexpr int GH = G / NH;           // transforms per group
    static_assert(GH * NH * TT == NT && HT % 32 == 0 && NH <= 15, "CTA shape");
    // Exchange buffer of one transform: 16 blocks of TT elements (block q = the q-th sub-sequence) at pitch
    // TP, buffers at pitch BUF.  A 128-byte wavefront serves 8/R3 (or 16/R3 in fp32) consecutive
    // (block, transform) pairs of R3 elements each: with several transforms per group the pad of BUF
    // staggers them over the banks, with one transform per group the pad of TP staggers the blocks.
    static constexpr int TP = TT + (GH == 1 ? R3 : 0);
    static constexpr int BUF = 16 * TP + R3;
    // natural index n of the transform <-> its slot in the buffer (block n / TT, element n % TT)
    __host__ __device__ static constexpr int slot(int n) { return n + (n / TT) * (TP - TT); }
    static constexpr int N_TW = M + 16 * R3;    // tw1[q*TT + t] then tw2[q2*R3 + n2]
    static constexpr int HALF = M / 2;          // the transform length N must be <= HALF
    static constexpr int SLOTS = 8;             // staged complex-sized elements per thread
    // The kernels are bound by shared-memory (LSU) wavefronts while the FP64 pipe has slack, so in fp64 the
    // inter-stage twiddles w^q, q = 1..15, are generated from the one loaded value w by a multiplication
    // chain (+14 complex products, -14 table reads of 16 bytes per stage; error <= 15 ulp on |w^q| = 1).
    // fp32 keeps the table: its FMA pipe is not idle and its tolerance is tighter relative to eps.
    static constexpr bool CHAIN = SURFH_FFT_TWIDDLE_CHAIN && sizeof(T) == 8;
    // shared-memory layout, in units of C
    static constexpr int OFF_TW = 0;
    static constexpr int OFF_FILT = OFF_TW + N_TW;
    static constexpr int OFF_CHIRP = OFF_FILT + M;
    static constexpr int OFF_BUF = OFF_CHIRP + HALF;
    static constexpr int OFF_STAGE = OFF_BUF + G * BUF;
    static constexpr int N_SMEM = OFF_STAGE + SLOTS * NT;
    // after the complex area: per-plane row-pair ranges of a pruned launch (start[MAX_PLANES+1], lo[MAX_PLANES])
    static constexpr int MAX_PLANES = 512;
    static constexpr size_t OFF_RANGES_BYTES = (size_t)N_SMEM * sizeof(C);
    static constexpr size_t SMEM_BYTES = OFF_RANGES_BYTES + (2 * MAX_PLANES + 2) * sizeof(int);
    static constexpr size_t SMEM_BYTES_FILTER = (size_t)OFF_STAGE * sizeof(C);
};

// Who am I: transform g (of G), thread t (of TT) = q*R3 + n2.  The R3 threads that exchange registers
// in the last radix stage are adjacent lanes; the GH transforms of a group are interleaved next, so that a
// warp touches GH adjacent columns in a column pass.  `half` = index of the barrier group.
template <typename T, int M> struct FftThread {
    using K = FftK<T, M>;
    int half, g, t, q, n2;
    __device__ __forceinline__ FftThread(int tid) {
        half = tid / K::HT;
        const int l = tid % K::HT;
        n2 = l % K::R3;
        const int c = l / K::R3;
        g = half * K::GH + c % K::GH;
        q = c / K::GH;
        t = q * K::R3 + n2;
    }
    // barrier among the threads of this group only
    __device__ __forceinline__ void sync() const {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(K::HT) : "memory");
    }
};

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

// In-register transpose among the R3 adjacent lanes n2 = 0..R3-1 of one group: for every block c of R3
// registers, thread n2 ends up with v[c*R3 + n] = (thread n's v[c*R3 + n2]).  Self-inverse.
template <int R3, typename C> __device__ __forceinline__ void group_transpose(C* v, int n2) {
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
                const C recv = shfl_xor_c(send, s);
                if (up) lo = recv; else hi = recv;
            }
    }
}

// Length-M forward FFT of the sequence held as v[m] = x[t + TT*m]; the result stays in registers in a
// digit-reversed order that only fft_inv() (and the filter table built by the same code) needs to know.
// `buf` is this transform's shared buffer, `tw` the shared twiddle table.  PRE_SYNC: other threads may
// still be reading this buffer from the previous step (the passes that post-process through it).
template <typename T, int M, bool HALF_IN, bool PRE_SYNC>
__device__ __forceinline__ void fft_fwd(cplx_t<T>* v, cplx_t<T>* buf, const FftThread<T, M>& th, const cplx_t<T>* tw) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    constexpr int R3 = K::R3, TT = K::TT;
    if (HALF_IN) dft16_in8<false>(v);  // v[8..15] are the zero padding
    else dft16<false>(v);
    if (PRE_SYNC) th.sync();
    if (K::CHAIN) {
        const C w = tw[TT + th.t];
        C pw = w;
        buf[th.t] = v[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) {
            buf[q * K::TP + th.t] = cmul(v[q], pw);
            if (q < 15) pw = cmul(pw, w);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            C x = v[q];
            if (q) x = cmul(x, tw[q * TT + th.t]);
            buf[q * K::TP + th.t] = x;
        }
    }
    th.sync();
    C* blk = buf + th.q * K::TP;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = blk[th.n2 + R3 * m];
    dft16<false>(v);
    if (R3 > 1) {
        if (K::CHAIN) {
            const C w = tw[M + R3 + th.n2];
            C pw = w;
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) {
                v[q2] = cmul(v[q2], pw);
                if (q2 < 15) pw = cmul(pw, w);
            }
        } else {
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) v[q2] = cmul(v[q2], tw[M + q2 * R3 + th.n2]);
        }
        group_transpose<R3>(v, th.n2);
#pragma unroll
        for (int c = 0; c < 16 / R3; ++c) dft_r<false, R3>(v + c * R3);
    }
}

// Mirror of fft_fwd: takes the digit-reversed spectrum in registers, returns M * x[t + TT*m] in v[m].
template <typename T, int M, bool HALF_OUT>
__device__ __forceinline__ void fft_inv(cplx_t<T>* v, cplx_t<T>* buf, const FftThread<T, M>& th, const cplx_t<T>* tw) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    constexpr int R3 = K::R3, TT = K::TT;
    if (R3 > 1) {
#pragma unroll
        for (int c = 0; c < 16 / R3; ++c) dft_r<true, R3>(v + c * R3);
        group_transpose<R3>(v, th.n2);
        if (K::CHAIN) {
            const C w = tw[M + R3 + th.n2];
            C pw = w;
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) {
                v[q2] = cmul_conj(pw, v[q2]);
                if (q2 < 15) pw = cmul(pw, w);
            }
        } else {
#pragma unroll
            for (int q2 = 1; q2 < 16; ++q2) v[q2] = cmul_conj(tw[M + q2 * R3 + th.n2], v[q2]);
        }
    }
    dft16<true>(v);
    // these are the very locations this thread read in fft_fwd: no barrier needed before the writes
    C* blk = buf + th.q * K::TP;
#pragma unroll
    for (int m = 0; m < 16; ++m) blk[th.n2 + R3 * m] = v[m];
    th.sync();
    if (K::CHAIN) {
        const C w = tw[TT + th.t];
        C pw = w;
        v[0] = buf[th.t];
#pragma unroll
        for (int qq = 1; qq < 16; ++qq) {
            v[qq] = cmul_conj(pw, buf[qq * K::TP + th.t]);
            if (qq < 15) pw = cmul(pw, w);
        }
    } else {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            C x = buf[qq * K::TP + th.t];
            if (qq) x = cmul_conj(tw[qq * TT + th.t], x);
            v[qq] = x;
        }
    }
    if (HALF_OUT) dft16_out8<true>(v);  // only x[t + TT*m], m < 8, is wanted
    else dft16<true>(v);
}

// Device tables of one 1-D chirp-z plan (length n through M-point FFTs), in global memory; the kernels
// copy them to shared memory once per CTA.
template <typename T> struct FftPlan1d {
    const cplx_t<T>* chirp;  // [n]   a[j] = exp(-i pi j^2 / n)
    const cplx_t<T>* filt;   // [M]   FFT_M(conj(a) wrapped) / M, in fft_fwd's register order [j*TT + t]
    const cplx_t<T>* tw;     // [M + 16*R3]  tw[q*TT + t] = exp(-2 pi i t q / M), then
                             //              tw[M + q2*R3 + n2] = exp(-2 pi i n2 q2 / TT)
    int n;
};

template <typename T, int M>
__device__ __forceinline__ void fft_load_tables(cplx_t<T>* smem, const FftPlan1d<T>& p, bool with_filter) {
    using K = FftK<T, M>;
    for (int i = threadIdx.x; i < K::N_TW; i += blockDim.x) smem[K::OFF_TW + i] = p.tw[i];
    if (with_filter) {
        for (int i = threadIdx.x; i < M; i += blockDim.x) smem[K::OFF_FILT + i] = p.filt[i];
        for (int i = threadIdx.x; i < p.n; i += blockDim.x) smem[K::OFF_CHIRP + i] = p.chirp[i];
    }
    __syncthreads();
}

// Circular convolution with the chirp filter: v[m] = (u * conj(a))[t + TT*m] for m < 8, with u given
// the same way and u[t + TT*m] = 0 for m >= 8 (v[8..15] are ignored on entry, garbage on exit).
template <typename T, int M, bool PRE_SYNC>
__device__ __forceinline__ void chirp_convolve(cplx_t<T>* v, cplx_t<T>* smem, cplx_t<T>* buf, const FftThread<T, M>& th) {
    using K = FftK<T, M>;
    fft_fwd<T, M, true, PRE_SYNC>(v, buf, th, smem + K::OFF_TW);
    const cplx_t<T>* filt = smem + K::OFF_FILT;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], filt[j * K::TT + th.t]);
    fft_inv<T, M, true>(v, buf, th, smem + K::OFF_TW);
}

// Builds FftPlan1d::filt from the natural-order filter `b` (already scaled by 1/M) with the very code
// that consumes it, so the digit-reversed order never has to be spelled out.  One CTA.
template <typename T, int M>
__global__ void __launch_bounds__(FftK<T, M>::NT, 1)
fft_filter_kernel(const cplx_t<T>* __restrict__ b, FftPlan1d<T> p, cplx_t<T>* __restrict__ filt) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    C* smem = reinterpret_cast<C*>(fft_smem);
    fft_load_tables<T, M>(smem, p, false);
    const FftThread<T, M> th(threadIdx.x);
    C* buf = smem + K::OFF_BUF + th.g * K::BUF;
    C v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = b[th.t + K::TT * m];
    fft_fwd<T, M, false, false>(v, buf, th, smem + K::OFF_TW);
    if (th.g == 0)
#pragma unroll
        for (int j = 0; j < 16; ++j) filt[j * K::TT + th.t] = v[j];
}

struct FftShape {
    int na, nb, nh;          // rows, columns, nb/2+1
    int npair;               // ceil(na / 2)
    size_t real_plane;       // elements between real planes
    size_t spec_plane;       // complex elements between spectrum planes ([na][nh], row pitch nh)
    size_t z_plane;          // complex elements between planes of the intermediate buffer
    int batch;
    // Pruned transforms: per plane, only the row pairs [lo, lo + cnt) of the real image matter (C2R: the
    // others are not produced; R2C: the others are known to be zero).  NULL = all rows.  [batch] (lo, cnt)
    const int2* pair_range;
};

// Shared-memory view of the row-pair ranges of one launch (built once per CTA).
struct FftRanges {
    const int* start;  // [batch + 1] exclusive prefix sum of cnt; NULL when the launch is not pruned
    const int* lo;     // [batch]
    int batch;
    __device__ __forceinline__ int total() const { return start[batch]; }
    // plane holding work item `it` (0 <= it < total): last p with start[p] <= it
    __device__ __forceinline__ int plane_of(int it) const {
        int a = 0, b = batch;
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (start[mid] <= it) a = mid; else b = mid;
        }
        return a;
    }
    __device__ __forceinline__ int cnt(int p) const { return start[p + 1] - start[p]; }
};

// Builds the ranges in shared memory: thread-serial chunks + one warp scan (batch <= MAX_PLANES).
template <int MAX_PLANES>
__device__ __forceinline__ FftRanges fft_build_ranges(int* smem_i, const int2* pair_range, int batch) {
    FftRanges r;
    r.batch = batch;
    r.start = nullptr;
    r.lo = nullptr;
    if (pair_range == nullptr) return r;
    int* start = smem_i;
    int* lo = smem_i + MAX_PLANES + 1;
    if (threadIdx.x < 32) {
        constexpr int PER = MAX_PLANES / 32;
        const int base = threadIdx.x * PER;
        int sum = 0;
        for (int k = 0; k < PER; ++k) {
            const int p = base + k;
            if (p < batch) {
                const int2 pr = pair_range[p];
                lo[p] = pr.x;
                sum += pr.y;
            }
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += v;
        }
        int run = incl - sum;
        for (int k = 0; k < PER; ++k) {
            const int p = base + k;
            if (p < batch) {
                start[p] = run;
                run += pair_range[p].y;
            }
        }
        if (threadIdx.x == 31) start[batch] = incl;
    }
    r.start = start;
    r.lo = lo;
    return r;  // visibility: the caller's __syncthreads (fft_load_tables) follows
}

// ---- asynchronous staging of the next item's inputs -------------------------------------------
// Every thread copies exactly the elements it will itself consume into thread-private shared-memory
// slots, so a cp.async.wait_group is all the synchronisation the staging needs.
// A work item of a CTA step, seen from one thread: which plane and which row pair / column.
struct FftItem {
    bool live;
    int plane, idx;
};

// ---- R2C pass 1: pairs of real rows -> the two Hermitian half spectra, rows 2p and 2p+1 of Y [Na][Nh]
template <typename T, int M> struct RowsR2C {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    static constexpr bool POST = true;    // post-processes through the shared buffer
    const T* in;
    C* y;
    FftShape s;
    __device__ int steps(const FftRanges& rg) const {
        const long long total = rg.start ? rg.total() : (long long)s.batch * s.npair;
        return (int)((total + K::G - 1) / K::G);
    }
    __device__ FftItem item(int step, int g, const FftRanges& rg) const {
        const long long it = (long long)step * K::G + g;
        FftItem r;
        if (rg.start) {
            r.live = it < rg.total();
            r.plane = r.live ? rg.plane_of((int)it) : 0;
            r.idx = r.live ? rg.lo[r.plane] + ((int)it - rg.start[r.plane]) : 0;
        } else {
            r.live = it < (long long)s.batch * s.npair;
            r.plane = r.live ? (int)(it / s.npair) : 0;
            r.idx = r.live ? (int)(it % s.npair) : 0;
        }
        return r;
    }
    // staged: 16 reals per thread = 8 complex-sized slots; slot (m, k) at stage[(2m + k) * NT + tid] in reals
    __device__ void prefetch(const FftItem& it, int t, C* stage_c, int tid, const FftRanges&) const {
        if (!it.live) return;
        T* stage = reinterpret_cast<T*>(stage_c);
        const int r0 = 2 * it.idx;
        const T* ra = in + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        const bool has_b = r0 + 1 < s.na;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            if (n < s.nb) {
                cp_async<sizeof(T)>(stage + (2 * m) * K::NT + tid, ra + n);
                if (has_b) cp_async<sizeof(T)>(stage + (2 * m + 1) * K::NT + tid, ra + s.nb + n);
            }
        }
    }
    __device__ void load(const FftItem& it, int t, const C* stage_c, int tid, C* v, const C* chirp, const FftRanges&) const {
        const T* stage = reinterpret_cast<const T*>(stage_c);
        const bool has_b = 2 * it.idx + 1 < s.na;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && n < s.nb) {
                u.x = stage[(2 * m) * K::NT + tid];
                u.y = has_b ? stage[(2 * m + 1) * K::NT + tid] : T(0);
                u = cmul(u, chirp[n]);
            }
            v[m] = u;
        }
    }
    // Z[n] = FFT(row_even + i row_odd)[n] is written to the transform's buffer in natural order (these
    // are the thread's own final-stage locations), then every thread separates A[j] = (Z[j] + conj
    // Z[nb-j]) / 2 and B[j] = (Z[j] - conj Z[nb-j]) / 2i for its share of j < nh.
    __device__ void finish(const FftItem& it, const FftThread<T, M>& th, C* v, C* buf, const C* chirp, const FftRanges& rg) const {
        const int t = th.t;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            if (n < s.nb) buf[K::slot(n)] = cmul(v[m], chirp[n]);
        }
        th.sync();
        if (!it.live) return;
        const bool has_b = 2 * it.idx + 1 < s.na;
        C* ya = y + (size_t)it.plane * s.z_plane + (size_t)(2 * it.idx) * s.nh;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const int j = t + K::TT * m;
            if (j < s.nh) {
                const C a = buf[K::slot(j)], b = buf[K::slot(j == 0 ? 0 : s.nb - j)];
                ya[j] = make_c<T>(T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y));
                if (has_b) ya[s.nh + j] = make_c<T>(T(0.5) * (a.y + b.y), T(0.5) * (b.x - a.x));
            }
        }
    }
};

// ---- column transforms of a half-complex plane, forward (R2C pass 2) or inverse (C2R pass 1)
// INVERSE = false:  Y [Na][Nh] -> spec [Na][Nh]
// INVERSE = true:   spec [Na][Nh] -> Z [npair][Nb], Z[p][j] = W[2p][j] + i W[2p+1][j] and its Hermitian
//                   extension Z[p][nb-j] = conj(W[2p][j]) + i conj(W[2p+1][j]), W = inverse column
//                   transform (conj in, conj out around the forward chirp-z)
template <typename T, int M, bool INVERSE> struct ColsPass {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    // The C2R pairing is a lane-pair shuffle when the transform's adjacent threads are adjacent lanes
    // (R3 >= 2), else a round trip through the shared buffer (which then needs the PRE_SYNC of fft_fwd).
    static constexpr bool PAIR_BY_SHUFFLE = INVERSE && K::R3 >= 2;
    static constexpr bool POST = INVERSE && !PAIR_BY_SHUFFLE;
    const C* src;
    C* dst;
    FftShape s;
    __device__ int tiles() const { return (s.nh + K::G - 1) / K::G; }
    __device__ int steps(const FftRanges&) const { return s.batch * tiles(); }
    __device__ FftItem item(int step, int g, const FftRanges&) const {
        const int tl = tiles();
        FftItem r;
        r.plane = step / tl;
        r.idx = (step % tl) * K::G + g;
        r.live = r.idx < s.nh;
        return r;
    }
    __device__ size_t src_plane() const { return INVERSE ? s.spec_plane : s.z_plane; }
    // forward direction of a pruned launch: rows outside the plane's range are zero and are not read
    __device__ void row_window(const FftItem& it, const FftRanges& rg, int& r0, int& r1) const {
        r0 = 0;
        r1 = s.na;
        if (!INVERSE && rg.start) {
            r0 = 2 * rg.lo[it.plane];
            r1 = min(s.na, r0 + 2 * rg.cnt(it.plane));
        }
    }
    __device__ void prefetch(const FftItem& it, int t, C* stage, int tid, const FftRanges& rg) const {
        if (!it.live) return;
        int r0, r1;
        row_window(it, rg, r0, r1);
        const C* col = src + (size_t)it.plane * src_plane() + it.idx;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + K::TT * m;
            if (i >= r0 && i < r1) cp_async<sizeof(C)>(stage + m * K::NT + tid, col + (size_t)i * s.nh);
        }
    }
    __device__ void load(const FftItem& it, int t, const C* stage, int tid, C* v, const C* chirp, const FftRanges& rg) const {
        int r0, r1;
        row_window(it, rg, r0, r1);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + K::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && i >= r0 && i < r1) {
                u = stage[m * K::NT + tid];
                if (INVERSE) u.y = -u.y;
                u = cmul(u, chirp[i]);
            }
            v[m] = u;
        }
    }
    __device__ void finish(const FftItem& it, const FftThread<T, M>& th, C* v, C* buf, const C* chirp, const FftRanges& rg) const {
        const int t = th.t;
        if (!INVERSE) {
            if (!it.live) return;
            C* col = dst + (size_t)it.plane * s.spec_plane + it.idx;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = t + K::TT * m;
                if (i < s.na) col[(size_t)i * s.nh] = cmul(v[m], chirp[i]);
            }
            return;
        }
        const int j = it.idx;
        const int nyq = (s.nb & 1) ? -1 : s.nb / 2;
        const bool self_mirror = j == 0 || j == nyq;  // numpy's irfft ignores the imaginary part there
        const int p0 = rg.start ? rg.lo[it.plane] : 0;
        const int p1 = rg.start ? p0 + rg.cnt(it.plane) : s.npair;
        C* zp = dst + (size_t)it.plane * s.z_plane;
        if (PAIR_BY_SHUFFLE) {
            // rows 2p and 2p+1 sit in the adjacent lanes t and t^1 (t = ... + n2, n2 the fastest lane index):
            // the even lane assembles Z[p][j], the odd lane its Hermitian mirror Z[p][nb-j]
            const bool odd = (t & 1) != 0;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int i = t + K::TT * m;
                C mine = make_c<T>(T(0), T(0));
                if (i < s.na) {
                    mine = cmul(v[m], chirp[i]);
                    mine.y = self_mirror ? T(0) : -mine.y;
                }
                const C other = shfl_xor_c(mine, 1);
                const C a = odd ? other : mine, b = odd ? mine : other;
                const int p = i >> 1;
                if (it.live && p >= p0 && p < p1) {
                    C* row = zp + (size_t)p * s.nb;
                    if (!odd) row[j] = make_c<T>(a.x - b.y, a.y + b.x);
                    else if (!self_mirror) row[s.nb - j] = make_c<T>(a.x + b.y, b.x - a.y);
                }
            }
            return;
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = t + K::TT * m;
            if (i < s.na) {
                C r = cmul(v[m], chirp[i]);
                r.y = self_mirror ? T(0) : -r.y;
                buf[K::slot(i)] = r;
            }
        }
        th.sync();
        if (!it.live) return;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int p = t + K::TT * m;
            if (p >= p0 && p < p1) {
                const C a = buf[K::slot(2 * p)];
                const C b = 2 * p + 1 < s.na ? buf[K::slot(2 * p + 1)] : make_c<T>(T(0), T(0));
                C* row = zp + (size_t)p * s.nb;
                row[j] = make_c<T>(a.x - b.y, a.y + b.x);
                if (!self_mirror) row[s.nb - j] = make_c<T>(a.x + b.y, b.x - a.y);
            }
        }
    }
};

// ---- C2R pass 2: Z [npair][Nb] -> pairs of real rows
template <typename T, int M> struct RowsC2R {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    static constexpr bool POST = false;
    const C* z;
    T* out;
    FftShape s;
    __device__ int steps(const FftRanges& rg) const {
        const long long total = rg.start ? rg.total() : (long long)s.batch * s.npair;
        return (int)((total + K::G - 1) / K::G);
    }
    __device__ FftItem item(int step, int g, const FftRanges& rg) const {
        const long long it = (long long)step * K::G + g;
        FftItem r;
        if (rg.start) {
            r.live = it < rg.total();
            r.plane = r.live ? rg.plane_of((int)it) : 0;
            r.idx = r.live ? rg.lo[r.plane] + ((int)it - rg.start[r.plane]) : 0;
        } else {
            r.live = it < (long long)s.batch * s.npair;
            r.plane = r.live ? (int)(it / s.npair) : 0;
            r.idx = r.live ? (int)(it % s.npair) : 0;
        }
        return r;
    }
    __device__ void prefetch(const FftItem& it, int t, C* stage, int tid, const FftRanges&) const {
        if (!it.live) return;
        const C* row = z + (size_t)it.plane * s.z_plane + (size_t)it.idx * s.nb;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            if (n < s.nb) cp_async<sizeof(C)>(stage + m * K::NT + tid, row + n);
        }
    }
    // IFFT(z) = conj(FFT(conj z)) = row_even + i row_odd
    __device__ void load(const FftItem& it, int t, const C* stage, int tid, C* v, const C* chirp, const FftRanges&) const {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            C u = make_c<T>(T(0), T(0));
            if (it.live && n < s.nb) {
                u = stage[m * K::NT + tid];
                u.y = -u.y;
                u = cmul(u, chirp[n]);
            }
            v[m] = u;
        }
    }
    __device__ void finish(const FftItem& it, const FftThread<T, M>& th, C* v, C*, const C* chirp, const FftRanges&) const {
        if (!it.live) return;
        const int t = th.t, r0 = 2 * it.idx;
        const bool has_b = r0 + 1 < s.na;
        T* oa = out + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        T* ob = oa + s.nb;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int n = t + K::TT * m;
            if (n < s.nb) {
                const C r = cmul(v[m], chirp[n]);
                oa[n] = r.x;
                if (has_b) ob[n] = -r.y;
            }
        }
    }
};

// Persistent kernel shared by the four passes.
template <typename T, int M, typename Pass>
__global__ void __launch_bounds__(FftK<T, M>::NT, 1) fft_pass_kernel(Pass pass, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    C* smem = reinterpret_cast<C*>(fft_smem);
    const int tid = threadIdx.x;
    const FftThread<T, M> th(tid);
    C* buf = smem + K::OFF_BUF + th.g * K::BUF;
    C* stage = smem + K::OFF_STAGE;
    const C* chirp = smem + K::OFF_CHIRP;
    const FftRanges rg = fft_build_ranges<K::MAX_PLANES>(reinterpret_cast<int*>(fft_smem + K::OFF_RANGES_BYTES),
                                                          pass.s.pair_range, pass.s.batch);
    fft_load_tables<T, M>(smem, p, true);
    const int n_steps = pass.steps(rg);
    int step = blockIdx.x;
    if (step < n_steps) pass.prefetch(pass.item(step, th.g, rg), th.t, stage, tid, rg);
    cp_async_commit();
    for (; step < n_steps; step += gridDim.x) {
        C v[16];
        const FftItem cur = pass.item(step, th.g, rg);
        cp_async_wait_all();
        pass.load(cur, th.t, stage, tid, v, chirp, rg);
        if (step + (int)gridDim.x < n_steps)
            pass.prefetch(pass.item(step + (int)gridDim.x, th.g, rg), th.t, stage, tid, rg);
        cp_async_commit();
        chirp_convolve<T, M, Pass::POST>(v, smem, buf, th);
        pass.finish(cur, th, v, buf, chirp, rg);
    }
}

}  // namespace surfh

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
// i.e. chirp multiply on load -> FFT of length M >= 2N-1 (decimation in frequency, output left in a
// digit-scrambled order) -> pointwise multiply with the precomputed, identically scrambled spectrum of
// the chirp filter -> the mirrored inverse FFT -> chirp multiply on store.  No bit-reversal pass exists.
//
// Execution model (round 2: ONE WARP = ONE TRANSFORM).
//   * A transform of length M = 32 * PTS lives in the 32 lanes of one warp, PTS = 8 / 16 / 32 points per
//     lane (M = 256 / 512 / 1024): FFT = radix-PTS in registers x radix-PTS in registers x radix-(32/PTS)
//     among adjacent lanes by warp shuffles.  For M = 1024 (the 501-pixel maps) that is 32 x 32: two register
//     stages and ONE transpose through shared memory, nothing else.  The only synchronisation inside a
//     transform is __syncwarp(): no CTA or named barrier exists in the steady state, so the G warps of a CTA
//     drift apart freely and one warp's shared-memory phase overlaps the others' FP64 phases.
//   * ONE persistent CTA per SM; every warp walks its own stream of work items with a grid stride.  The
//     chirp and the filter spectrum live in shared memory for the CTA's lifetime.  In fp64 the inter-stage
//     twiddles w^k are generated from the lane's single root w (held in registers) by a multiplication chain:
//     the kernels are co-limited by the FP64 pipe and the shared-memory crossbar, and a chain step costs 4
//     FP64 instructions where a table read costs 4 shared-memory wavefronts per warp.
//   * The inputs of a warp's next item are copied global -> shared by cp.async into that warp's staging
//     buffer (natural order; every lane copies exactly the elements it will consume, so cp.async.wait_group
//     is all the synchronisation the staging needs) while the current item is transformed.
//
// 2-D real transforms use the two-for-one trick: a pair of real rows is transformed as one complex row.
//   R2C:  rows_r2c  real [Na][Nb] -> A/B-separated half spectra Y, stored transposed [Nh][Na] (separation
//                   through the warp's own shared buffer)      ->  cols  (-> spectrum)
//   C2R:  cols_c2r  spectrum -> Z [ceil(Na/2)][Nb], Z[p] = W[2p] + i W[2p+1] Hermitian-extended
//                   (rows 2p, 2p+1 sit in adjacent lanes: pairing by shuffle) ->  rows_c2r (-> real [Na][Nb])
// Spectrum layout: the operator keeps every half spectrum (OTF, map spectra, the working cube's spectrum)
// TRANSPOSED, [Nh][Na]: a column of the 2-D transform is then one contiguous run of Na complex numbers, read
// and written by the column passes as 512-byte warp accesses; the pointwise OTF / template kernels do not
// care about the order.  The stand-alone surfh_rfft2 keeps numpy's [Na][Nh] (SPEC_T = false).
// In the operator the rows passes only visit the row pairs some band's field of view touches (FftRanges).
#pragma once
#include "common.cuh"

// build-time switches for A/B measurements (python -c "build.build(out=..., defines=[...])")
#ifndef SURFH_FFT_TWIDDLE_CHAIN
#define SURFH_FFT_TWIDDLE_CHAIN 1
#endif
#ifndef SURFH_FFT_WARPS_F64_1024
#define SURFH_FFT_WARPS_F64_1024 8
#endif
#ifndef SURFH_FFT_WARPS_F64_512
#define SURFH_FFT_WARPS_F64_512 8
#endif
#ifndef SURFH_FFT_BULK
#define SURFH_FFT_BULK 1
#endif
#ifndef SURFH_FFT_STAGGER
#define SURFH_FFT_STAGGER 0      // ns by which the second warp of every scheduler starts late (measured: no effect)
#endif

namespace surfh {

// Geometry of one chirp-z length M for arithmetic type T.
template <typename T, int M> struct FftK {
    static_assert(M == 256 || M == 512 || M == 1024, "chirp-z length must be 256, 512 or 1024");
    using C = cplx_t<T>;
    static constexpr int TT = 32;               // lanes per transform: one warp
    static constexpr int PTS = M / 32;          // points per lane: 8, 16, 32
    static constexpr int R3 = 32 / PTS;         // last radix, among adjacent lanes: 4, 2, 1
    static constexpr int HP = PTS / 2;          // live points per lane on the zero-padded side
    static constexpr int HALF = M / 2;          // the transform length N must be <= HALF
    // warps (= transforms in flight) per CTA: 32 complex doubles per lane need ~250 registers (8 warps fill the
    // register file), 16 need ~200 (10 warps); fp32 halves both
    static constexpr int G = sizeof(T) == 8 ? (M == 1024 ? SURFH_FFT_WARPS_F64_1024 : SURFH_FFT_WARPS_F64_512) : 16;
    static constexpr int NT = 32 * G;
    // Exchange buffer of one transform: PTS rows (row k1 = the k1-th output of every lane's first-stage DFT)
    // of 32 elements at pitch TP.  The transposed read of lane (q, n2) walks row q from element n2 in steps of
    // R3: the pad of R3 elements staggers the rows over the banks (conflict-free per 128-byte wavefront).
    static constexpr int TP = 32 + R3;
    static constexpr int BUF = PTS * TP;
    // natural index n of the transform <-> its slot in the buffer (row n / 32, element n % 32)
    __host__ __device__ static constexpr int slot(int n) { return n + (n / 32) * (TP - 32); }
    static constexpr int N_TW = M + PTS * R3;   // tw1[k1*32 + t] then tw2[q2*R3 + n2]
    // fp64: twiddles by multiplication chains from the lane's root (error <= PTS ulp on |w^k| = 1); fp32 keeps
    // the tables (its tolerance is tighter relative to eps, and its FMA pipe is not idle)
    static constexpr bool CHAIN = SURFH_FFT_TWIDDLE_CHAIN && sizeof(T) == 8;
    // shared-memory layout, in units of C
    static constexpr int OFF_FILT = 0;
    static constexpr int OFF_CHIRP = OFF_FILT + M;
    static constexpr int OFF_TW = OFF_CHIRP + HALF;
    static constexpr int N_ROOTS = 32 + R3;     // chains: w1[lane], then w2[n2]
    static constexpr int OFF_BUF = OFF_TW + (CHAIN ? N_ROOTS : N_TW);
    static constexpr int OFF_STAGE = OFF_BUF + G * BUF;
    static constexpr int N_SMEM = OFF_STAGE + G * HALF;
    // after the complex area: per-plane row-pair ranges of a pruned launch (start[MAX_PLANES+1], lo[MAX_PLANES])
    static constexpr int MAX_PLANES = 512;
    static constexpr size_t OFF_RANGES_BYTES = (size_t)N_SMEM * sizeof(C);
    // then one mbarrier per warp (bulk-copy staging, fp64)
    static constexpr size_t OFF_MBAR_BYTES = (OFF_RANGES_BYTES + (2 * MAX_PLANES + 2) * sizeof(int) + 7) / 8 * 8;
    static constexpr size_t SMEM_BYTES = OFF_MBAR_BYTES + G * 8;
    // A pass whose input of one item is ONE contiguous, 16-byte-aligned run can stage it with a single bulk
    // asynchronous copy (TMA) per warp instead of PTS/2 cp.async per lane: complex doubles only (16-byte
    // elements keep every run aligned whatever the odd map size)
    static constexpr bool BULK_OK = SURFH_FFT_BULK && sizeof(C) == 16;
    static_assert(SMEM_BYTES <= 232448, "shared-memory budget of one CTA (227 KB)");
};

// Who am I inside the warp: lane = q*R3 + n2.  The R3 lanes that exchange registers in the last radix stage
// are adjacent.  The lane's twiddle roots (chains) w1 = exp(-2 pi i lane / M) and w2 = exp(-2 pi i n2 / 32)
// are re-read from shared memory where a chain starts (32 + R3 table entries) rather than held in registers
// across the whole transform: the 32-point kernels are register-bound.
template <typename T, int M> struct FftLane {
    using K = FftK<T, M>;
    int lane, q, n2;
    __device__ __forceinline__ FftLane(int lane_) {
        lane = lane_;
        n2 = lane % K::R3;
        q = lane / K::R3;
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
// registers (PTS registers in all), thread n2 ends up with v[c*R3 + n] = (thread n's v[c*R3 + n2]).  Self-inverse.
template <int R3, int PTS, typename C> __device__ __forceinline__ void group_transpose(C* v, int n2) {
#pragma unroll
    for (int s = R3 / 2; s >= 1; s >>= 1) {
        const bool up = (n2 & s) != 0;
#pragma unroll
        for (int c = 0; c < PTS / R3; ++c)
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


// cos / sin of 2 pi E / 32
constexpr double kCos32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                               0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                               0.19509032201612826785, 0.0, -0.19509032201612826785, -0.38268343236508977173,
                               -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                               -0.92387953251128675613, -0.98078528040323044913};
constexpr double kSin32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                               0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613,
                               0.98078528040323044913, 1.0, 0.98078528040323044913, 0.92387953251128675613,
                               0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                               0.38268343236508977173, 0.19509032201612826785};

// a * exp(-+ 2 pi i E / 32) (forward / inverse), E in [0, 16)
template <bool INV, int E, typename C> __device__ __forceinline__ C mul_w32(C a) {
    using R = decltype(a.x);
    if constexpr (E == 0) return a;
    else if constexpr (E == 8) return rot90<INV>(a);
    else if constexpr (E == 4) return mul_w16<INV, 2>(a);
    else if constexpr (E == 12) return mul_w16<INV, 6>(a);
    else {
        constexpr double c = kCos32[E], sn = kSin32[E];
        C w;
        w.x = R(c);
        w.y = R(-sn);
        return twmul<INV>(a, w);
    }
}

// first level of the 32-point DFT for the pair (m, m + 16), m = I .. 15
template <bool INV, int MODE, int I, typename C> __device__ __forceinline__ void dft32_level1(C* v) {
    if constexpr (I < 16) {
        if (MODE == 1) {
            v[I + 16] = mul_w32<INV, I>(v[I]);
        } else {
            const C a = v[I], b = v[I + 16];
            v[I] = cadd(a, b);
            v[I + 16] = mul_w32<INV, I>(csub(a, b));
        }
        dft32_level1<INV, MODE, I + 1>(v);
    }
}

// 32-point DFT, natural order in and out: one radix-2 decimation-in-frequency level, then two 16-point DFTs.
// MODE 0: full.  MODE 1: inputs v[16..31] are zero (the zero padding of the chirp-z input; they are not read).
// MODE 2: only the outputs X[0..15] are wanted (left in v[0..15]; v[16..31] are garbage).
template <bool INV, int MODE, typename C> __device__ __forceinline__ void dft32(C* v) {
    dft32_level1<INV, MODE, 0>(v);
    C o[32];
    if (MODE == 2) {
        dft16_out8<INV>(v);        // X[2k]   = E[k]
        dft16_out8<INV>(v + 16);   // X[2k+1] = O[k]
#pragma unroll
        for (int k = 0; k < 8; ++k) { o[2 * k] = v[k]; o[2 * k + 1] = v[16 + k]; }
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = o[k];
    } else {
        dft16<INV>(v);
        dft16<INV>(v + 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) { o[2 * k] = v[k]; o[2 * k + 1] = v[16 + k]; }
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = o[k];
    }
}

// PTS-point DFT over the lane's registers, natural order in and out.  MODE as for dft32 (the 8-point version
// ignores it: its callers zero-fill / discard the unused half).
template <bool INV, int PTS, int MODE, typename C> __device__ __forceinline__ void dft_pts(C* v) {
    if (PTS == 32) dft32<INV, MODE>(v);
    if (PTS == 16) {
        if (MODE == 1) dft16_in8<INV>(v);
        else if (MODE == 2) dft16_out8<INV>(v);
        else dft16<INV>(v);
    }
    if (PTS == 8) dft8<INV>(v);
}

// Length-M forward FFT of the sequence held as v[m] = x[lane + 32*m]; the result stays in registers in a
// digit-scrambled order that only fft_inv() (and the filter table built by the same code) needs to know
// (natural -- X[lane + 32*j] in v[j] -- when R3 == 1).  `buf` is this warp's shared buffer, `tw` the shared
// twiddle tables (unused with chains).  HALF_IN: v[PTS/2..] are the zero padding.
template <typename T, int M, bool HALF_IN>
__device__ __forceinline__ void fft_fwd(cplx_t<T>* v, cplx_t<T>* buf, const FftLane<T, M>& th, const cplx_t<T>* tw) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    constexpr int R3 = K::R3, PTS = K::PTS;
    dft_pts<false, PTS, HALF_IN ? 1 : 0>(v);
    __syncwarp();  // the previous item's post-processing may still be reading this buffer
    if (K::CHAIN) {
        const C w1 = tw[th.lane];
        C pw = w1;
        buf[th.lane] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < PTS; ++k1) {
            buf[k1 * K::TP + th.lane] = cmul(v[k1], pw);
            if (k1 < PTS - 1) pw = cmul(pw, w1);
        }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < PTS; ++k1) {
            C x = v[k1];
            if (k1) x = cmul(x, tw[k1 * 32 + th.lane]);
            buf[k1 * K::TP + th.lane] = x;
        }
    }
    __syncwarp();
    C* blk = buf + th.q * K::TP;
#pragma unroll
    for (int m = 0; m < PTS; ++m) v[m] = blk[th.n2 + R3 * m];
    dft_pts<false, PTS, 0>(v);
    if (R3 > 1) {
        if (K::CHAIN) {
            const C w2 = tw[32 + th.n2];
            C pw = w2;
#pragma unroll
            for (int q2 = 1; q2 < PTS; ++q2) {
                v[q2] = cmul(v[q2], pw);
                if (q2 < PTS - 1) pw = cmul(pw, w2);
            }
        } else {
#pragma unroll
            for (int q2 = 1; q2 < PTS; ++q2) v[q2] = cmul(v[q2], tw[M + q2 * R3 + th.n2]);
        }
        group_transpose<R3, PTS>(v, th.n2);
#pragma unroll
        for (int c = 0; c < PTS / R3; ++c) dft_r<false, R3>(v + c * R3);
    }
}

// Mirror of fft_fwd: takes the scrambled spectrum in registers, returns M * x[lane + 32*m] in v[m].
// HALF_OUT: only m < PTS/2 is wanted.
template <typename T, int M, bool HALF_OUT>
__device__ __forceinline__ void fft_inv(cplx_t<T>* v, cplx_t<T>* buf, const FftLane<T, M>& th, const cplx_t<T>* tw) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    constexpr int R3 = K::R3, PTS = K::PTS;
    if (R3 > 1) {
#pragma unroll
        for (int c = 0; c < PTS / R3; ++c) dft_r<true, R3>(v + c * R3);
        group_transpose<R3, PTS>(v, th.n2);
        if (K::CHAIN) {
            const C w2 = tw[32 + th.n2];
            C pw = w2;
#pragma unroll
            for (int q2 = 1; q2 < PTS; ++q2) {
                v[q2] = cmul_conj(pw, v[q2]);
                if (q2 < PTS - 1) pw = cmul(pw, w2);
            }
        } else {
#pragma unroll
            for (int q2 = 1; q2 < PTS; ++q2) v[q2] = cmul_conj(tw[M + q2 * R3 + th.n2], v[q2]);
        }
    }
    dft_pts<true, PTS, 0>(v);
    // these are the very locations this lane read in fft_fwd: no barrier needed before the writes
    C* blk = buf + th.q * K::TP;
#pragma unroll
    for (int m = 0; m < PTS; ++m) blk[th.n2 + R3 * m] = v[m];
    __syncwarp();
    if (K::CHAIN) {
        const C w1 = tw[th.lane];
        C pw = w1;
        v[0] = buf[th.lane];
#pragma unroll
        for (int k1 = 1; k1 < PTS; ++k1) {
            v[k1] = cmul_conj(pw, buf[k1 * K::TP + th.lane]);
            if (k1 < PTS - 1) pw = cmul(pw, w1);
        }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < PTS; ++k1) {
            C x = buf[k1 * K::TP + th.lane];
            if (k1) x = cmul_conj(tw[k1 * 32 + th.lane], x);
            v[k1] = x;
        }
    }
    dft_pts<true, PTS, HALF_OUT ? 2 : 0>(v);
}

// Device tables of one 1-D chirp-z plan (length n through M-point FFTs), in global memory; the kernels
// copy chirp and filter (and, without chains, the twiddles) to shared memory once per CTA.
template <typename T> struct FftPlan1d {
    const cplx_t<T>* chirp;  // [n]   a[j] = exp(-i pi j^2 / n)
    const cplx_t<T>* filt;   // [M]   FFT_M(conj(a) wrapped) / M, in fft_fwd's register order [j*32 + lane]
    const cplx_t<T>* tw;     // [M + PTS*R3]  tw[k1*32 + t] = exp(-2 pi i t k1 / M), then
                             //               tw[M + q2*R3 + n2] = exp(-2 pi i n2 q2 / 32)
    int n;
};

template <typename T, int M>
__device__ __forceinline__ void fft_load_tables(cplx_t<T>* smem, const FftPlan1d<T>& p, bool with_filter) {
    using K = FftK<T, M>;
    if (K::CHAIN) {  // the chain roots only: tw1[1][lane], lane < 32, then tw2[1][n2], n2 < R3
        for (int i = threadIdx.x; i < K::N_ROOTS; i += blockDim.x)
            smem[K::OFF_TW + i] = i < 32 ? p.tw[32 + i] : p.tw[M + K::R3 + (i - 32)];
    } else {
        for (int i = threadIdx.x; i < K::N_TW; i += blockDim.x) smem[K::OFF_TW + i] = p.tw[i];
    }
    if (with_filter) {
        for (int i = threadIdx.x; i < M; i += blockDim.x) smem[K::OFF_FILT + i] = p.filt[i];
        for (int i = threadIdx.x; i < p.n; i += blockDim.x) smem[K::OFF_CHIRP + i] = p.chirp[i];
    }
    __syncthreads();
}

// Circular convolution with the chirp filter: v[m] = (u * conj(a))[lane + 32*m] for m < PTS/2, with u given
// the same way and u[lane + 32*m] = 0 for m >= PTS/2 (v[PTS/2..] are ignored on entry -- except for PTS == 8,
// whose callers zero them -- and garbage on exit).
template <typename T, int M>
__device__ __forceinline__ void chirp_convolve(cplx_t<T>* v, cplx_t<T>* smem, cplx_t<T>* buf, const FftLane<T, M>& th) {
    using K = FftK<T, M>;
    fft_fwd<T, M, true>(v, buf, th, smem + K::OFF_TW);
    const cplx_t<T>* filt = smem + K::OFF_FILT;
#pragma unroll
    for (int j = 0; j < K::PTS; ++j) v[j] = cmul(v[j], filt[j * 32 + th.lane]);
    fft_inv<T, M, true>(v, buf, th, smem + K::OFF_TW);
}

// Builds FftPlan1d::filt from the natural-order filter `b` (already scaled by 1/M) with the very code
// that consumes it, so the scrambled order never has to be spelled out.  One warp.
template <typename T, int M>
__global__ void __launch_bounds__(32, 1)
fft_filter_kernel(const cplx_t<T>* __restrict__ b, FftPlan1d<T> p, cplx_t<T>* __restrict__ filt) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    C* smem = reinterpret_cast<C*>(fft_smem);
    fft_load_tables<T, M>(smem, p, false);
    const FftLane<T, M> th(threadIdx.x);
    C* buf = smem + K::OFF_BUF;
    C v[K::PTS];
#pragma unroll
    for (int m = 0; m < K::PTS; ++m) v[m] = b[th.lane + 32 * m];
    fft_fwd<T, M, false>(v, buf, th, smem + K::OFF_TW);
#pragma unroll
    for (int j = 0; j < K::PTS; ++j) filt[j * 32 + th.lane] = v[j];
}

struct FftShape {
    int na, nb, nh;          // rows, columns, nb/2+1
    int npair;               // ceil(na / 2)
    size_t real_plane;       // elements between real planes
    size_t spec_plane;       // complex elements between spectrum planes
    size_t z_plane;          // complex elements between planes of the intermediate buffer
    int ypitch;              // R2C intermediate Y^T [nh][ypitch]: na rounded up to even (rows 2p, 2p+1 share a sector)
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

// A work item of one warp: which plane and which row pair / column.  Every pass exposes
//   items(rg)                      number of work items of the launch
//   item(it, rg)                   decode
//   prefetch(item, lane, stage)    cp.async the item's inputs into the warp's staging buffer (natural order;
//                                  lane copies the elements lane + 32 r it will itself consume)
//   load(item, lane, stage, v, chirp)        staged inputs x chirp -> registers v[0 .. PTS/2)
//   finish(item, th, v, buf, chirp, rg)      chirp multiply, pass-specific post-processing, store
struct FftItem {
    int plane, idx;
};

template <typename T, int M> struct RowItems {
    // row pairs of the (possibly pruned) launch
    __device__ static long long count(const FftShape& s, const FftRanges& rg) {
        return rg.start ? (long long)rg.total() : (long long)s.batch * s.npair;
    }
    __device__ static FftItem decode(int it, const FftShape& s, const FftRanges& rg) {
        FftItem r;
        if (rg.start) {
            r.plane = rg.plane_of(it);
            r.idx = rg.lo[r.plane] + (it - rg.start[r.plane]);
        } else {
            r.plane = it / s.npair;
            r.idx = it - r.plane * s.npair;
        }
        return r;
    }
};

// ---- R2C pass 1: pairs of real rows -> the two Hermitian half spectra, rows 2p and 2p+1 of Y (stored as Y^T)
// ALIGNED: the caller guarantees that every row pair starts on a 16-byte boundary and spans a multiple of 16
// bytes (the operator's working cube: fp64, plane stride padded to even) -> one TMA bulk copy per item; else
// (caller-owned maps, fp32) every lane copies its own elements with cp.async.
template <typename T, int M, bool ALIGNED> struct RowsR2C {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    static constexpr bool BULK = ALIGNED && K::BULK_OK;
    const T* in;
    C* y;
    FftShape s;
    __device__ long long items(const FftRanges& rg) const { return RowItems<T, M>::count(s, rg); }
    __device__ FftItem item(int it, const FftRanges& rg) const { return RowItems<T, M>::decode(it, s, rg); }
    // staged as reals: row 2p at stage[n], row 2p+1 at stage[nb + n] (the pair is contiguous in memory)
    __device__ const C* bulk_src(const FftItem& it, const FftRanges&, int& first, int& count) const {
        first = 0;
        count = s.nb;   // in units of C = two reals: 2 nb reals (a plane's last, unpaired row drags nb stray reals
                        // along: the buffer has a row of slack and `load` ignores them)
        return reinterpret_cast<const C*>(in + (size_t)it.plane * s.real_plane + (size_t)(2 * it.idx) * s.nb);
    }
    __device__ void prefetch(const FftItem& it, int lane, C* stage_c, const FftRanges&) const {
        T* stage = reinterpret_cast<T*>(stage_c);
        const int r0 = 2 * it.idx;
        const T* ra = in + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        const bool has_b = r0 + 1 < s.na;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            if (n < s.nb) {
                cp_async<sizeof(T)>(stage + n, ra + n);
                if (has_b) cp_async<sizeof(T)>(stage + s.nb + n, ra + s.nb + n);
            }
        }
    }
    __device__ void load(const FftItem& it, int lane, const C* stage_c, C* v, const C* chirp, const FftRanges&) const {
        const T* stage = reinterpret_cast<const T*>(stage_c);
        const bool has_b = 2 * it.idx + 1 < s.na;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            C u = make_c<T>(T(0), T(0));
            if (n < s.nb) {
                u.x = stage[n];
                u.y = has_b ? stage[s.nb + n] : T(0);
                u = cmul(u, chirp[n]);
            }
            v[m] = u;
        }
    }
    // Z[n] = FFT(row_even + i row_odd)[n] is written to the warp's buffer in natural order, then every lane
    // separates A[j] = (Z[j] + conj Z[nb-j]) / 2 and B[j] = (Z[j] - conj Z[nb-j]) / 2i for its share of j < nh.
    __device__ void finish(const FftItem& it, const FftLane<T, M>& th, C* v, C* buf, const C* chirp, const FftRanges&) const {
        const int lane = th.lane;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            if (n < s.nb) buf[K::slot(n)] = cmul(v[m], chirp[n]);
        }
        __syncwarp();
        const bool has_b = 2 * it.idx + 1 < s.na;
        // Y is stored TRANSPOSED, [nh][ypitch]: the two half spectra of a row pair land side by side (32 bytes,
        // one full sector per column j) and the column pass that follows reads whole columns as contiguous runs
        C* ya = y + (size_t)it.plane * s.z_plane + (size_t)(2 * it.idx);
#pragma unroll
        for (int m = 0; m < K::HP / 2 + 1; ++m) {
            const int j = lane + 32 * m;
            if (j < s.nh) {
                const C a = buf[K::slot(j)], b = buf[K::slot(j == 0 ? 0 : s.nb - j)];
                C* dst = ya + (size_t)j * s.ypitch;
                dst[0] = make_c<T>(T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y));
                if (has_b) dst[1] = make_c<T>(T(0.5) * (a.y + b.y), T(0.5) * (b.x - a.x));
            }
        }
    }
};

// ---- column transforms of a half-complex plane, forward (R2C pass 2) or inverse (C2R pass 1)
// INVERSE = false:  Y^T [Nh][ypitch] -> spectrum
// INVERSE = true:   spectrum -> Z [npair][Nb], Z[p][j] = W[2p][j] + i W[2p+1][j] and its Hermitian
//                   extension Z[p][nb-j] = conj(W[2p][j]) + i conj(W[2p+1][j]), W = inverse column
//                   transform (conj in, conj out around the forward chirp-z)
// SPEC_T: the spectrum is stored transposed, [Nh][Na] (the operator's layout: column j is contiguous);
//         otherwise numpy's [Na][Nh].
template <typename T, int M, bool INVERSE, bool SPEC_T> struct ColsPass {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    // forward: Y^T columns are contiguous; inverse: the spectrum's columns are, in the transposed layout
    static constexpr bool BULK = K::BULK_OK && (!INVERSE || SPEC_T);
    const C* src;
    C* dst;
    FftShape s;
    // the item's input as one contiguous run: first element index within the column, count, source
    __device__ const C* bulk_src(const FftItem& it, const FftRanges& rg, int& first, int& count) const {
        int r0, r1;
        row_window(it, rg, r0, r1);
        first = r0;
        count = r1 - r0;
        const C* base = src + (size_t)it.plane * (INVERSE ? s.spec_plane : s.z_plane);
        return base + (INVERSE ? (size_t)it.idx * s.na : (size_t)it.idx * s.ypitch) + r0;
    }
    __device__ long long items(const FftRanges&) const { return (long long)s.batch * s.nh; }
    __device__ FftItem item(int it, const FftRanges&) const {
        FftItem r;
        r.plane = it / s.nh;
        r.idx = it - r.plane * s.nh;
        return r;
    }
    // spectrum element (row i, column j) of a plane
    __device__ __forceinline__ size_t spec_at(int i, int j) const {
        return SPEC_T ? (size_t)j * s.na + i : (size_t)i * s.nh + j;
    }
    // forward direction of a pruned launch: rows outside the plane's range are zero and are not read
    __device__ void row_window(const FftItem& it, const FftRanges& rg, int& r0, int& r1) const {
        r0 = 0;
        r1 = s.na;
        if (!INVERSE && rg.start) {
            r0 = 2 * rg.lo[it.plane];
            r1 = min(s.na, r0 + 2 * rg.cnt(it.plane));
        }
    }
    __device__ void prefetch(const FftItem& it, int lane, C* stage, const FftRanges& rg) const {
        int r0, r1;
        row_window(it, rg, r0, r1);
        const C* base = src + (size_t)it.plane * (INVERSE ? s.spec_plane : s.z_plane);
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int i = lane + 32 * m;
            if (i >= r0 && i < r1)
                cp_async<sizeof(C)>(stage + i, base + (INVERSE ? spec_at(i, it.idx) : (size_t)it.idx * s.ypitch + i));
        }
    }
    __device__ void load(const FftItem& it, int lane, const C* stage, C* v, const C* chirp, const FftRanges& rg) const {
        int r0, r1;
        row_window(it, rg, r0, r1);
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int i = lane + 32 * m;
            C u = make_c<T>(T(0), T(0));
            if (i >= r0 && i < r1) {
                u = stage[i];
                if (INVERSE) u.y = -u.y;
                u = cmul(u, chirp[i]);
            }
            v[m] = u;
        }
    }
    __device__ void finish(const FftItem& it, const FftLane<T, M>& th, C* v, C*, const C* chirp, const FftRanges& rg) const {
        const int lane = th.lane;
        if (!INVERSE) {
            C* plane = dst + (size_t)it.plane * s.spec_plane;
#pragma unroll
            for (int m = 0; m < K::HP; ++m) {
                const int i = lane + 32 * m;
                if (i < s.na) plane[spec_at(i, it.idx)] = cmul(v[m], chirp[i]);
            }
            return;
        }
        const int j = it.idx;
        const int nyq = (s.nb & 1) ? -1 : s.nb / 2;
        const bool self_mirror = j == 0 || j == nyq;  // numpy's irfft ignores the imaginary part there
        const int p0 = rg.start ? rg.lo[it.plane] : 0;
        const int p1 = min(s.npair, rg.start ? p0 + rg.cnt(it.plane) : s.npair);
        C* zp = dst + (size_t)it.plane * s.z_plane;
        // rows 2p and 2p+1 sit in the adjacent lanes (lane even / odd, same register): the even lane assembles
        // Z[p][j], the odd lane its Hermitian mirror Z[p][nb-j].  Three separate sweeps (chirp products, lane-pair
        // exchange, stores) so that the HP shared-memory reads / shuffles of a sweep are in flight together
        // instead of one element's latency chain after the other (v[HP..] are dead here: registers are free).
        const bool odd = (lane & 1) != 0;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int i = lane + 32 * m;
            C mine = make_c<T>(T(0), T(0));
            if (i < s.na) {
                mine = cmul(v[m], chirp[i]);
                mine.y = self_mirror ? T(0) : -mine.y;
            }
            v[m] = mine;
        }
#pragma unroll
        for (int m = 0; m < K::HP; ++m) v[K::HP + m] = shfl_xor_c(v[m], 1);
        C* mycol = zp + (odd ? (size_t)(s.nb - j) : (size_t)j);
        const bool writes = !odd || !self_mirror;
        // with mine = this lane's row and other = its partner's: the even lane (rows a = mine, b = other) stores
        // Z[p][j] = (a.x - b.y, a.y + b.x), the odd lane (a = other, b = mine) Z[p][nb-j] = (a.x + b.y, b.x - a.y):
        // the same two numbers, swapped
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int p = (lane + 32 * m) >> 1;
            const T re = v[m].x - v[K::HP + m].y, im = v[m].y + v[K::HP + m].x;
            if (writes && p >= p0 && p < p1) mycol[(size_t)p * s.nb] = odd ? make_c<T>(im, re) : make_c<T>(re, im);
        }
    }
};

// ---- C2R pass 2: Z [npair][Nb] -> pairs of real rows
template <typename T, int M> struct RowsC2R {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    static constexpr bool BULK = K::BULK_OK;   // a row of Z is one contiguous run
    const C* z;
    T* out;
    FftShape s;
    __device__ const C* bulk_src(const FftItem& it, const FftRanges&, int& first, int& count) const {
        first = 0;
        count = s.nb;
        return z + (size_t)it.plane * s.z_plane + (size_t)it.idx * s.nb;
    }
    __device__ long long items(const FftRanges& rg) const { return RowItems<T, M>::count(s, rg); }
    __device__ FftItem item(int it, const FftRanges& rg) const { return RowItems<T, M>::decode(it, s, rg); }
    __device__ void prefetch(const FftItem& it, int lane, C* stage, const FftRanges&) const {
        const C* row = z + (size_t)it.plane * s.z_plane + (size_t)it.idx * s.nb;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            if (n < s.nb) cp_async<sizeof(C)>(stage + n, row + n);
        }
    }
    // IFFT(z) = conj(FFT(conj z)) = row_even + i row_odd
    __device__ void load(const FftItem&, int lane, const C* stage, C* v, const C* chirp, const FftRanges&) const {
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            C u = make_c<T>(T(0), T(0));
            if (n < s.nb) {
                u = stage[n];
                u.y = -u.y;
                u = cmul(u, chirp[n]);
            }
            v[m] = u;
        }
    }
    __device__ void finish(const FftItem& it, const FftLane<T, M>& th, C* v, C*, const C* chirp, const FftRanges&) const {
        const int lane = th.lane, r0 = 2 * it.idx;
        const bool has_b = r0 + 1 < s.na;
        T* oa = out + (size_t)it.plane * s.real_plane + (size_t)r0 * s.nb;
        T* ob = oa + s.nb;
#pragma unroll
        for (int m = 0; m < K::HP; ++m) {
            const int n = lane + 32 * m;
            if (n < s.nb) {
                const C r = cmul(v[m], chirp[n]);
                oa[n] = r.x;
                if (has_b) ob[n] = -r.y;
            }
        }
    }
};

// Persistent kernel shared by the four passes: every warp walks its own items.
template <typename T, int M, typename Pass>
__global__ void __launch_bounds__(FftK<T, M>::NT, 1) fft_pass_kernel(Pass pass, FftPlan1d<T> p) {
    using C = cplx_t<T>;
    using K = FftK<T, M>;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    C* smem = reinterpret_cast<C*>(fft_smem);
    const int warp = threadIdx.x >> 5;
    const FftLane<T, M> th(threadIdx.x & 31);
    C* buf = smem + K::OFF_BUF + warp * K::BUF;
    C* stage = smem + K::OFF_STAGE + warp * K::HALF;
    const C* chirp = smem + K::OFF_CHIRP;
    const FftRanges rg = fft_build_ranges<K::MAX_PLANES>(reinterpret_cast<int*>(fft_smem + K::OFF_RANGES_BYTES),
                                                          pass.s.pair_range, pass.s.batch);
    if constexpr (Pass::BULK) {
        if ((threadIdx.x & 31) == 0) mbar_init(fft_smem + K::OFF_MBAR_BYTES + 8 * warp, 1);
        mbar_init_fence();
    }
    fft_load_tables<T, M>(smem, p, true);   // ends with __syncthreads
    const int n_items = (int)pass.items(rg);   // < 2^31 (checked by the host)
    const int stride = (int)gridDim.x * K::G;
    int it = (int)blockIdx.x * K::G + warp;
    // staging of an item's inputs: one bulk asynchronous copy per warp (TMA; completion on the warp's own
    // mbarrier) where the input is a single contiguous run, else cp.async by every lane for its own elements
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(fft_smem + K::OFF_MBAR_BYTES) + warp;
    unsigned parity = 0;
    auto stage_in = [&](int item_index) {
        const FftItem nxt = pass.item(item_index, rg);
        if constexpr (Pass::BULK) {
            __syncwarp();                        // every lane has consumed the previous contents of `stage`
            if (th.lane == 0) {
                int first, count;
                const C* src = pass.bulk_src(nxt, rg, first, count);
                fence_proxy_async_smem();
                mbar_expect_tx(bar, (unsigned)count * (unsigned)sizeof(C));
                bulk_copy_g2s(stage + first, src, (unsigned)count * (unsigned)sizeof(C), bar);
            }
        } else {
            pass.prefetch(nxt, th.lane, stage, rg);
            cp_async_commit();
        }
    };
    if (it < n_items) stage_in(it);
#if SURFH_FFT_STAGGER
    // (Experiment, off: starting the second warp of every scheduler half a transform late, so that the two
    // alternate between FP64-dense and latency-bound phases, changed nothing measurable.)
    if ((warp >> 2) & 1) __nanosleep(SURFH_FFT_STAGGER);
#endif
    for (; it < n_items; it += stride) {
        C v[K::PTS];
        const FftItem cur = pass.item(it, rg);
        if constexpr (Pass::BULK) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            cp_async_wait_all();  // every lane consumes only what it copied itself
        }
        pass.load(cur, th.lane, stage, v, chirp, rg);
        if (K::PTS == 8) {
#pragma unroll
            for (int m = K::HP; m < K::PTS; ++m) v[m] = make_c<T>(T(0), T(0));
        }
        if (it + stride < n_items) stage_in(it + stride);
        chirp_convolve<T, M>(v, smem, buf, th);
        pass.finish(pass.item(it, rg), th, v, buf, chirp, rg);
    }
}

}  // namespace surfh

// k4 / k4^T: the spectral response as ONE dense contraction per band over all pointings,
// slits and kept detector rows:
//
//   forward   y[(p,s), l', a]      = sum_{(l,b)} W[l', (l,b)] * G[l, (p,s,a), b]
//   adjoint   Gt[l, (p,s,a), b]    = sum_{l'}    W[l', (l,b)] * y[(p,s), l', a]
//
// Replaces jax_utils.wblur_subSampling + the alpha decimation and jax_utils.wblur_t + np.repeat
// (surfh/ToolsDir/jax_utils.py:72-91; surfh/Models/spectroModelChannel.py:229, 242-252), which the
// reference evaluates slit by slit on all a1-a0 oversampled rows.  The LSF is a sinc^2 whose
// tails are not negligible at 1e-10, so the contraction is genuinely dense: it is bound by FP64
// (or FP32) FMA throughput, not HBM and cannot use reduced-precision
// tensor cores within the parity tolerance.
//
// The three operands live in layouts dictated by their neighbours (the reference's [P,S,L',na]
// detector order, the gather kernels' [L][(p,s,a,b)] slit space), so the kernel addresses every
// operand through two offset tables:  A(m,k) = A[aM[m] + aK[k]], etc.  Offsets along the non-K
// dimension are hoisted into registers; offsets along K cost one (warp-uniform or per-thread)
// table read per K-tile.
//
// Classic register-tiled SIMT GEMM: BMxBN CTA tile, BK slab double-buffered through shared
// memory in k-major order (padded), TMxTN accumulators per thread.
#pragma once
#include "common.cuh"

namespace surfh {

template <typename T> struct GemmArgs {
    int M, N, K;
    const T* A;
    const int32_t* aM;
    const int32_t* aK;
    const T* B;
    const int32_t* bK;
    const int32_t* bN;
    T* C;
    const int32_t* cM;
    const int32_t* cN;
};

// A_KFAST: consecutive k are close in memory for A (else consecutive m are); same for B (k vs n).
template <typename T, int BM, int BN, int BK, int TM, int TN, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
otgemm_kernel(GemmArgs<T> g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 2;
    constexpr int EA = BM * BK / NT;  // A elements loaded per thread per slab
    constexpr int EB = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
    static_assert(NT % BK == 0 && NT % BM == 0 && NT % BN == 0, "loader mapping needs divisibility");
    extern __shared__ __align__(16) unsigned char gemm_smem[];
    T (*As)[BK][BM + PAD] = reinterpret_cast<T (*)[BK][BM + PAD]>(gemm_smem);
    T (*Bs)[BK][BN + PAD] = reinterpret_cast<T (*)[BK][BN + PAD]>(gemm_smem + sizeof(T) * 2 * BK * (BM + PAD));

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // ---- loader mapping -------------------------------------------------------------------
    // KFAST: k_loc = tid % BK (same for all of a thread's elements), m_loc = tid / BK + i * (NT / BK)
    // MFAST: m_loc = tid % BM (same for all elements),               k_loc = tid / BM + i * (NT / BM)
    int32_t a_fix[A_KFAST ? EA : 1];
    int32_t b_fix[B_KFAST ? EB : 1];
    bool a_ok[A_KFAST ? EA : 1], b_ok[B_KFAST ? EB : 1];
    if (A_KFAST) {
#pragma unroll
        for (int i = 0; i < EA; ++i) {
            const int m = m0 + tid / BK + i * (NT / BK);
            a_ok[i] = m < g.M;
            a_fix[i] = a_ok[i] ? g.aM[m] : 0;
        }
    } else {
        const int m = m0 + tid % BM;
        a_ok[0] = m < g.M;
        a_fix[0] = a_ok[0] ? g.aM[m] : 0;
    }
    if (B_KFAST) {
#pragma unroll
        for (int i = 0; i < EB; ++i) {
            const int n = n0 + tid / BK + i * (NT / BK);
            b_ok[i] = n < g.N;
            b_fix[i] = b_ok[i] ? g.bN[n] : 0;
        }
    } else {
        const int n = n0 + tid % BN;
        b_ok[0] = n < g.N;
        b_fix[0] = b_ok[0] ? g.bN[n] : 0;
    }

    T ra[EA], rb[EB];
    auto load_slab = [&](int k0) {
        if (A_KFAST) {
            const int k = k0 + tid % BK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.aK + k) : 0;
#pragma unroll
            for (int i = 0; i < EA; ++i) ra[i] = (kok && a_ok[i]) ? __ldg(g.A + a_fix[i] + ko) : T(0);
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int k = k0 + tid / BM + i * (NT / BM);
                ra[i] = (k < g.K && a_ok[0]) ? __ldg(g.A + a_fix[0] + __ldg(g.aK + k)) : T(0);
            }
        }
        if (B_KFAST) {
            const int k = k0 + tid % BK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.bK + k) : 0;
#pragma unroll
            for (int i = 0; i < EB; ++i) rb[i] = (kok && b_ok[i]) ? __ldg(g.B + b_fix[i] + ko) : T(0);
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int k = k0 + tid / BN + i * (NT / BN);
                rb[i] = (k < g.K && b_ok[0]) ? __ldg(g.B + b_fix[0] + __ldg(g.bK + k)) : T(0);
            }
        }
    };
    auto store_slab = [&](int buf) {
        if (A_KFAST) {
#pragma unroll
            for (int i = 0; i < EA; ++i) As[buf][tid % BK][tid / BK + i * (NT / BK)] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) As[buf][tid / BM + i * (NT / BM)][tid % BM] = ra[i];
        }
        if (B_KFAST) {
#pragma unroll
            for (int i = 0; i < EB; ++i) Bs[buf][tid % BK][tid / BK + i * (NT / BK)] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) Bs[buf][tid / BN + i * (NT / BN)][tid % BN] = rb[i];
        }
    };

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

    load_slab(0);
    store_slab(0);
    __syncthreads();
    const int n_slab = (g.K + BK - 1) / BK;
    for (int sidx = 0; sidx < n_slab; ++sidx) {
        const int buf = sidx & 1;
        if (sidx + 1 < n_slab) load_slab((sidx + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T fa[TM], fb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) fa[i] = As[buf][kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) fb[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(fa[i], fb[j], acc[i][j]);
        }
        if (sidx + 1 < n_slab) {
            store_slab(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: scattered store through the C offset tables -----------------------------
    int32_t cn[TN];
    bool nok[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        nok[j] = n < g.N;
        cn[j] = nok[j] ? g.cN[n] : 0;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m < g.M) {
            const int32_t cm = g.cM[m];
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (nok[j]) g.C[cm + cn[j]] = acc[i][j];
        }
    }
}

template <typename T, int BM, int BN, int BK>
constexpr size_t otgemm_smem_bytes() {
    return sizeof(T) * 2 * BK * ((BM + 2) + (BN + 2));
}

}  // namespace surfh

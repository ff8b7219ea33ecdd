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

// ------------------------------------------------------------------------------------------------
// FP64 path: the same offset-table contraction on the FP64 tensor pipe (DMMA,
// mma.sync.aligned.m8n8k4.f64), grouped over all bands of the model in ONE launch so that the
// 128x64 tiles of every band fill the 148 SMs in many waves instead of 1.5 per band.
//
//   CTA tile 128x64, BK = 16, 8 warps as 4 (m) x 2 (n), warp tile 32x32 = 4x4 m8n8 tiles,
//   two CTAs per SM (<= 128 registers, 2 x 60 KB shared memory).
//   Shared layouts follow the operand's contiguous direction so global->shared needs no transpose:
//     K-fast operand  -> [row][k], leading dimension BK + 4
//     row-fast operand-> [k][row], leading dimension rows + 4
//   Both leading dimensions are = 4 (mod 16) doubles, which makes the m8n8k4 fragment reads
//   (4 k-lanes x 8 row-lanes, 64-bit each) bank-conflict-free.
constexpr int kMaxGemmGroup = 16;

struct GemmBatch {
    int count;
    int tile_start[kMaxGemmGroup + 1];  // prefix sum of CTA tiles per problem
    GemmArgs<double> p[kMaxGemmGroup];
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int kDBM = 128, kDBN = 64, kDBK = 16;
constexpr int kDLdK = kDBK + 4;    // [row][k] layouts
constexpr int kDLdM = kDBM + 4;    // A as [k][m]
constexpr int kDLdN = kDBN + 4;    // B as [k][n]

template <bool A_KFAST, bool B_KFAST>
constexpr size_t dgemm_smem_bytes() {
    return sizeof(double) * 2 * ((A_KFAST ? kDBM * kDLdK : kDBK * kDLdM) + (B_KFAST ? kDBN * kDLdK : kDBK * kDLdN));
}

template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256, 2)
dgemm_mma_kernel(const __grid_constant__ GemmBatch batch) {
    constexpr int NT = 256;
    constexpr int EA = kDBM * kDBK / NT;  // 8
    constexpr int EB = kDBN * kDBK / NT;  // 4
    constexpr int A_STAGE = A_KFAST ? kDBM * kDLdK : kDBK * kDLdM;
    constexpr int B_STAGE = B_KFAST ? kDBN * kDLdK : kDBK * kDLdN;
    extern __shared__ __align__(16) unsigned char gemm_smem[];
    double* As = reinterpret_cast<double*>(gemm_smem);
    double* Bs = As + 2 * A_STAGE;

    // ---- which problem, which tile --------------------------------------------------------
    int pi = 0;
    while (pi + 1 < batch.count && (int)blockIdx.x >= batch.tile_start[pi + 1]) ++pi;
    const GemmArgs<double>& g = batch.p[pi];
    const int local_tile = blockIdx.x - batch.tile_start[pi];
    const int tiles_n = (g.N + kDBN - 1) / kDBN;
    const int m0 = (local_tile / tiles_n) * kDBM, n0 = (local_tile % tiles_n) * kDBN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps
    const int gq = lane >> 2, tq = lane & 3;

    // ---- loader mapping (see otgemm_kernel) ----------------------------------------------
    // K-fast operands: a thread loads EA (EB) rows at one k per slab; the row offsets live in shared memory
    // (one table per CTA) instead of EA + EB registers per thread, which kept the forward variant spilling.
    __shared__ int32_t s_afix[A_KFAST ? kDBM : 1], s_bfix[B_KFAST ? kDBN : 1];
    int32_t a_fix0 = 0, b_fix0 = 0;
    bool a_ok0 = false, b_ok0 = false;
    if (A_KFAST) {
        for (int r = tid; r < kDBM; r += NT) s_afix[r] = m0 + r < g.M ? __ldg(g.aM + m0 + r) : -1;
    } else {
        const int m = m0 + tid % kDBM;
        a_ok0 = m < g.M;
        a_fix0 = a_ok0 ? __ldg(g.aM + m) : 0;
    }
    if (B_KFAST) {
        for (int r = tid; r < kDBN; r += NT) s_bfix[r] = n0 + r < g.N ? __ldg(g.bN + n0 + r) : -1;
    } else {
        const int n = n0 + tid % kDBN;
        b_ok0 = n < g.N;
        b_fix0 = b_ok0 ? __ldg(g.bN + n) : 0;
    }
    __syncthreads();
    double ra[EA], rb[EB];
    auto load_slab = [&](int k0) {
        if (A_KFAST) {
            const int k = k0 + tid % kDBK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.aK + k) : 0;
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int32_t fix = s_afix[tid / kDBK + i * (NT / kDBK)];
                ra[i] = (kok && fix >= 0) ? __ldg(g.A + fix + ko) : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int k = k0 + tid / kDBM + i * (NT / kDBM);
                ra[i] = (k < g.K && a_ok0) ? __ldg(g.A + a_fix0 + __ldg(g.aK + k)) : 0.0;
            }
        }
        if (B_KFAST) {
            const int k = k0 + tid % kDBK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.bK + k) : 0;
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int32_t fix = s_bfix[tid / kDBK + i * (NT / kDBK)];
                rb[i] = (kok && fix >= 0) ? __ldg(g.B + fix + ko) : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int k = k0 + tid / kDBN + i * (NT / kDBN);
                rb[i] = (k < g.K && b_ok0) ? __ldg(g.B + b_fix0 + __ldg(g.bK + k)) : 0.0;
            }
        }
    };
    auto store_slab = [&](int buf) {
        double* a = As + buf * A_STAGE;
        double* b = Bs + buf * B_STAGE;
        if (A_KFAST) {
#pragma unroll
            for (int i = 0; i < EA; ++i) a[(tid / kDBK + i * (NT / kDBK)) * kDLdK + tid % kDBK] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) a[(tid / kDBM + i * (NT / kDBM)) * kDLdM + tid % kDBM] = ra[i];
        }
        if (B_KFAST) {
#pragma unroll
            for (int i = 0; i < EB; ++i) b[(tid / kDBK + i * (NT / kDBK)) * kDLdK + tid % kDBK] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) b[(tid / kDBN + i * (NT / kDBN)) * kDLdN + tid % kDBN] = rb[i];
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    load_slab(0);
    store_slab(0);
    __syncthreads();
    const int n_slab = (g.K + kDBK - 1) / kDBK;
    for (int sidx = 0; sidx < n_slab; ++sidx) {
        const int buf = sidx & 1;
        if (sidx + 1 < n_slab) load_slab((sidx + 1) * kDBK);
        const double* a = As + buf * A_STAGE;
        const double* b = Bs + buf * B_STAGE;
#pragma unroll
        for (int kk = 0; kk < kDBK; kk += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = wm * 32 + i * 8 + gq;
                fa[i] = A_KFAST ? a[row * kDLdK + kk + tq] : a[(kk + tq) * kDLdM + row];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wn * 32 + j * 8 + gq;
                fb[j] = B_KFAST ? b[col * kDLdK + kk + tq] : b[(kk + tq) * kDLdN + col];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
        if (sidx + 1 < n_slab) {
            store_slab(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn * 32 + j * 8 + tq * 2;
        const bool ok0 = n < g.N, ok1 = n + 1 < g.N;
        const int32_t c0 = ok0 ? __ldg(g.cN + n) : 0, c1 = ok1 ? __ldg(g.cN + n + 1) : 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + wm * 32 + i * 8 + gq;
            if (m < g.M) {
                const int32_t cm = __ldg(g.cM + m);
                if (ok0) g.C[cm + c0] = acc[i][j][0];
                if (ok1) g.C[cm + c1] = acc[i][j][1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FP32 path: the same offset-table contraction on the tensor cores with the 3xTF32 split.
// TF32 alone (10-bit mantissa) misses the 1e-5 parity budget of the fp32 mode, so every operand is split
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and each product is accumulated in fp32 as
//     a_lo b_hi + a_hi b_lo + a_hi b_hi          (the dropped a_lo b_lo term is below 2^-22 |a b|),
// three mma.sync.m16n8k8.tf32 per tile instead of 128 FFMA per thread-tile: fp32-level accuracy at several
// times the SIMT rate.  Same CTA shape as the FP64 kernel: 128x64 tile, BK = 16, 8 warps as 4 (m) x 2 (n),
// warp tile 32x32 = 2 (m16) x 4 (n8) mma tiles, grouped over the bands of the model in one launch.
// Shared layouts: K-fast operand [row][k] with leading dimension BK + 4, row-fast operand [k][row] with
// leading dimension rows + 8; both make the fragment reads (8 row-lanes x 4 k-lanes) conflict-free.
struct GemmBatchF {
    int count;
    int tile_start[kMaxGemmGroup + 1];
    GemmArgs<float> p[kMaxGemmGroup];
};

constexpr int kFBM = 128, kFBN = 64, kFBK = 16;
constexpr int kFLdK = kFBK + 4;
constexpr int kFLdM = kFBM + 8;
constexpr int kFLdN = kFBN + 8;

template <bool A_KFAST, bool B_KFAST>
constexpr size_t sgemm_smem_bytes() {
    return sizeof(float) * 2 * ((A_KFAST ? kFBM * kFLdK : kFBK * kFLdM) + (B_KFAST ? kFBN * kFLdK : kFBK * kFLdN));
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float rest = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rest));
}

__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256, 2)
sgemm_tf32x3_kernel(const __grid_constant__ GemmBatchF batch) {
    constexpr int NT = 256;
    constexpr int EA = kFBM * kFBK / NT;  // 8
    constexpr int EB = kFBN * kFBK / NT;  // 4
    constexpr int A_STAGE = A_KFAST ? kFBM * kFLdK : kFBK * kFLdM;
    constexpr int B_STAGE = B_KFAST ? kFBN * kFLdK : kFBK * kFLdN;
    extern __shared__ __align__(16) unsigned char gemm_smem[];
    float* As = reinterpret_cast<float*>(gemm_smem);
    float* Bs = As + 2 * A_STAGE;
    __shared__ int32_t s_afix[A_KFAST ? kFBM : 1], s_bfix[B_KFAST ? kFBN : 1];

    int pi = 0;
    while (pi + 1 < batch.count && (int)blockIdx.x >= batch.tile_start[pi + 1]) ++pi;
    const GemmArgs<float>& g = batch.p[pi];
    const int local_tile = blockIdx.x - batch.tile_start[pi];
    const int tiles_n = (g.N + kFBN - 1) / kFBN;
    const int m0 = (local_tile / tiles_n) * kFBM, n0 = (local_tile % tiles_n) * kFBN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps
    const int gq = lane >> 2, tq = lane & 3;

    int32_t a_fix0 = 0, b_fix0 = 0;
    bool a_ok0 = false, b_ok0 = false;
    if (A_KFAST) {
        for (int r = tid; r < kFBM; r += NT) s_afix[r] = m0 + r < g.M ? __ldg(g.aM + m0 + r) : -1;
    } else {
        const int m = m0 + tid % kFBM;
        a_ok0 = m < g.M;
        a_fix0 = a_ok0 ? __ldg(g.aM + m) : 0;
    }
    if (B_KFAST) {
        for (int r = tid; r < kFBN; r += NT) s_bfix[r] = n0 + r < g.N ? __ldg(g.bN + n0 + r) : -1;
    } else {
        const int n = n0 + tid % kFBN;
        b_ok0 = n < g.N;
        b_fix0 = b_ok0 ? __ldg(g.bN + n) : 0;
    }
    __syncthreads();

    float ra[EA], rb[EB];
    auto load_slab = [&](int k0) {
        if (A_KFAST) {
            const int k = k0 + tid % kFBK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.aK + k) : 0;
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int32_t fix = s_afix[tid / kFBK + i * (NT / kFBK)];
                ra[i] = (kok && fix >= 0) ? __ldg(g.A + fix + ko) : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int k = k0 + tid / kFBM + i * (NT / kFBM);
                ra[i] = (k < g.K && a_ok0) ? __ldg(g.A + a_fix0 + __ldg(g.aK + k)) : 0.f;
            }
        }
        if (B_KFAST) {
            const int k = k0 + tid % kFBK;
            const bool kok = k < g.K;
            const int32_t ko = kok ? __ldg(g.bK + k) : 0;
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int32_t fix = s_bfix[tid / kFBK + i * (NT / kFBK)];
                rb[i] = (kok && fix >= 0) ? __ldg(g.B + fix + ko) : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int k = k0 + tid / kFBN + i * (NT / kFBN);
                rb[i] = (k < g.K && b_ok0) ? __ldg(g.B + b_fix0 + __ldg(g.bK + k)) : 0.f;
            }
        }
    };
    auto store_slab = [&](int buf) {
        float* a = As + buf * A_STAGE;
        float* b = Bs + buf * B_STAGE;
        if (A_KFAST) {
#pragma unroll
            for (int i = 0; i < EA; ++i) a[(tid / kFBK + i * (NT / kFBK)) * kFLdK + tid % kFBK] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) a[(tid / kFBM + i * (NT / kFBM)) * kFLdM + tid % kFBM] = ra[i];
        }
        if (B_KFAST) {
#pragma unroll
            for (int i = 0; i < EB; ++i) b[(tid / kFBK + i * (NT / kFBK)) * kFLdK + tid % kFBK] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) b[(tid / kFBN + i * (NT / kFBN)) * kFLdN + tid % kFBN] = rb[i];
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

    load_slab(0);
    store_slab(0);
    __syncthreads();
    const int n_slab = (g.K + kFBK - 1) / kFBK;
    for (int sidx = 0; sidx < n_slab; ++sidx) {
        const int buf = sidx & 1;
        if (sidx + 1 < n_slab) load_slab((sidx + 1) * kFBK);
        const float* a = As + buf * A_STAGE;
        const float* b = Bs + buf * B_STAGE;
        // The tensor core adds the products to the accumulator with truncation; chained over the whole K
        // dimension that is a bias of ~K/8 * 3 * 2^-24 on sums of same-sign terms (measured 2e-5 at K = 3144).
        // So every slab is accumulated from zero in `part` and added to `acc` by a rounded fp32 addition.
        float part[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int r = 0; r < 4; ++r) part[i][j][r] = 0.f;
#pragma unroll
        for (int kk = 0; kk < kFBK; kk += 8) {
            // A fragment of an m16 tile: (row gq, k tq), (row gq+8, k tq), (row gq, k tq+4), (row gq+8, k tq+4)
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = wm * 32 + i * 16 + gq;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int rr = row + (r & 1) * 8, kc = kk + tq + (r >> 1) * 4;
                    const float v = A_KFAST ? a[rr * kFLdK + kc] : a[kc * kFLdM + rr];
                    split_tf32(v, ah[i][r], al[i][r]);
                }
            }
            // B fragment of an n8 tile: (k tq, col gq), (k tq+4, col gq)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = wn * 32 + j * 8 + gq;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int kc = kk + tq + r * 4;
                    const float v = B_KFAST ? b[col * kFLdK + kc] : b[kc * kFLdN + col];
                    split_tf32(v, bh[j][r], bl[j][r]);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_tf32_m16n8k8(part[i][j], al[i], bh[j]);
                    mma_tf32_m16n8k8(part[i][j], ah[i], bl[j]);
                    mma_tf32_m16n8k8(part[i][j], ah[i], bh[j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[i][j][r] += part[i][j][r];
        if (sidx + 1 < n_slab) {
            store_slab(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: c0 (row gq, col 2 tq), c1 (gq, 2 tq + 1), c2 (gq + 8, 2 tq), c3 (gq + 8, 2 tq + 1)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn * 32 + j * 8 + tq * 2;
        const bool ok0 = n < g.N, ok1 = n + 1 < g.N;
        const int32_t c0 = ok0 ? __ldg(g.cN + n) : 0, c1 = ok1 ? __ldg(g.cN + n + 1) : 0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + wm * 32 + i * 16 + h * 8 + gq;
                if (m < g.M) {
                    const int32_t cm = __ldg(g.cM + m);
                    if (ok0) g.C[cm + c0] = acc[i][j][2 * h];
                    if (ok1) g.C[cm + c1] = acc[i][j][2 * h + 1];
                }
            }
    }
}

}  // namespace surfh

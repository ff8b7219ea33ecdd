// k4 / k4^T on TMA (fp64): the spectral response as plain "TN" matrix products whose operands are 2-D tensor maps.
//
//   forward   y[m, n]  = sum_k W [m, k]  * G [n, k]     m = detector wavelength l', k = (l, b), n = (p, s, a)
//   adjoint   Gt[n, k] = sum_m Wt[k, m]  * Yk[n, m]     Wt = W^T (a second copy, uploaded once), Yk = y permuted
//
// Replaces jax_utils.wblur_subSampling + the alpha decimation and jax_utils.wblur_t + np.repeat
// (surfh/ToolsDir/jax_utils.py:72-91; surfh/Models/spectroModelChannel.py:229, 242-252), like kernels_gemm.cuh.
// Round 1 gathered both operands element-wise through offset tables (__ldg -> registers -> st.shared, two stages,
// one __syncthreads per 16-deep slab) and reached 0.79-0.81 of the cuBLAS DGEMM rate.  Here the slit-space vector
// is stored K-fast per detector column (kernels_slit.cuh), so both operands of both products are row-major
// matrices with the contraction index contiguous:
//   * one elected thread issues two `cp.async.bulk.tensor.2d` (TMA) per 16-deep slab into a 4-stage ring of
//     shared-memory tiles (A: 128 rows x 128 bytes, B: 64 rows x 128 bytes), completion by transaction bytes on
//     the stage's `full` mbarrier; a stage is refilled (3 slabs ahead) once its `empty` mbarrier has collected
//     the arrivals of all eight warps -- one iteration after they consumed it, so the wait is normally over;
//   * eight warps (4 x 2, warp tile 32 x 32): wait on `full`, 16 DMMA (mma.sync m8n8k4 f64 -- the only FP64
//     tensor path: tcgen05 has no f64 kind) per 4-deep step, one arrive per warp on `empty`; no __syncthreads
//     in the main loop;
//   * the tiles use the hardware 128-byte swizzle (16-byte chunk index XOR row mod 8), which makes the DMMA
//     fragment reads -- 8 rows x 4 consecutive doubles per warp -- bank-conflict-free without padding;
//   * rows / slabs past the matrix edge are zero-filled by TMA: no predication in the main loop.
// The epilogue scatters the accumulators through two offset tables (the detector layout [P,S,L',na] is imposed by
// the reference; the slit-space layout by the gather / scatter kernels).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels_gemm.cuh"

namespace surfh {

#ifndef SURFH_GEMM_STAGES
#define SURFH_GEMM_STAGES 4
#endif
constexpr int kTBM = 128, kTBN = 64, kTBK = 16, kTStages = SURFH_GEMM_STAGES;
constexpr int kTConsumerWarps = 8;
constexpr int kTThreads = 32 * kTConsumerWarps;
constexpr int kTStageBytes = (kTBM + kTBN) * kTBK * 8;   // 24 KB
constexpr size_t kTSmemBytes = 1024 /* alignment slack */ + (size_t)kTStages * kTStageBytes + 2 * kTStages * 8;

struct GemmTmaProblem {
    CUtensorMap a;   // A [M][K], K contiguous; box {16, 128}, 128-byte swizzle
    CUtensorMap b;   // B [N][K], K contiguous; box {16, 64}
    int M, N, K;
    double* C;               // C(m, n) at C[cM[m] + cN[n]]
    const int32_t* cM;
    const int32_t* cN;
};

struct GemmTmaBatch {
    int count;
    int tile_start[kMaxGemmGroup + 1];  // prefix sum of CTA tiles per problem
    GemmTmaProblem p[kMaxGemmGroup];
};

__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, void* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// double at (row r, column k) of a [rows][16] tile written by TMA with the 128-byte swizzle (tile base 1024-aligned)
__device__ __forceinline__ double swz_ld(const unsigned char* tile, int r, int k) {
    const unsigned off = (unsigned)r * 128u + ((((unsigned)k >> 1) ^ ((unsigned)r & 7u)) << 4) + (((unsigned)k & 1u) << 3);
    return *reinterpret_cast<const double*>(tile + off);
}

__global__ void __launch_bounds__(kTThreads, 2)
dgemm_tma_kernel(const __grid_constant__ GemmTmaBatch batch) {
    extern __shared__ unsigned char gemm_tma_smem[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_tma_smem) + 1023) & ~(uintptr_t)1023);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)kTStages * kTStageBytes);
    unsigned long long* empty = full + kTStages;

    int pi = 0;
    while (pi + 1 < batch.count && (int)blockIdx.x >= batch.tile_start[pi + 1]) ++pi;
    const GemmTmaProblem& g = batch.p[pi];
    const int local_tile = blockIdx.x - batch.tile_start[pi];
    const int tiles_n = (g.N + kTBN - 1) / kTBN;
    const int m0 = (local_tile / tiles_n) * kTBM, n0 = (local_tile % tiles_n) * kTBN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_slab = (g.K + kTBK - 1) / kTBK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTConsumerWarps);
        }
        mbar_init_fence();
    }
    __syncthreads();

    auto issue_slab = [&](int slab) {   // thread 0 only
        const int s = slab % kTStages;
        unsigned char* a_tile = smem + (size_t)s * kTStageBytes;
        unsigned char* b_tile = a_tile + kTBM * kTBK * 8;
        mbar_expect_tx(&full[s], (unsigned)kTStageBytes);
        tma_load_2d(a_tile, &g.a, slab * kTBK, m0, &full[s]);
        tma_load_2d(b_tile, &g.b, slab * kTBK, n0, &full[s]);
    };
    if (tid == 0) {
        tma_prefetch_desc(&g.a);
        tma_prefetch_desc(&g.b);
        for (int i = 0; i < kTStages && i < n_slab; ++i) issue_slab(i);
    }

    const int wm = warp >> 1, wn = warp & 1;   // 4 x 2 warps, warp tile 32 x 32
    const int gq = lane >> 2, tq = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int i = 0; i < n_slab; ++i) {
        const int s = i % kTStages;
        mbar_wait(&full[s], (unsigned)(i / kTStages) & 1u);
        const unsigned char* a_tile = smem + (size_t)s * kTStageBytes;
        const unsigned char* b_tile = a_tile + kTBM * kTBK * 8;
#pragma unroll
        for (int kk = 0; kk < kTBK; kk += 4) {
            double fa[4], fb[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) fa[ii] = swz_ld(a_tile, wm * 32 + ii * 8 + gq, kk + tq);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) fb[jj] = swz_ld(b_tile, wn * 32 + jj * 8 + gq, kk + tq);
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma_m8n8k4(acc[ii][jj][0], acc[ii][jj][1], fa[ii], fb[jj]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        // refill the stage consumed one iteration ago with the slab kTStages further on
        if (tid == 0 && i >= 1 && i - 1 + kTStages < n_slab) {
            const int prev = i - 1;
            mbar_wait(&empty[prev % kTStages], (unsigned)(prev / kTStages) & 1u);
            issue_slab(prev + kTStages);
        }
    }

    // ---- epilogue -------------------------------------------------------------------------------
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

// Yk[n][m] = y[(ps*nd + m)*na + a],  n = ps*na + a  (row pitch ldk >= nd): the detector block of a band re-laid
// K-fast per detector column, the B operand of the adjoint product (per (p, s): an [nd][na] -> [na][nd] transpose).
template <typename T>
__global__ void __launch_bounds__(256)
detector_to_kfast_kernel(const T* __restrict__ y, int na, int nd, int Nn, int ldk, T* __restrict__ yk) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)Nn * nd) return;
    const int m = (int)(idx % nd), n = (int)(idx / nd);
    const int a = n % na, ps = n / na;
    yk[(size_t)n * ldk + m] = y[((size_t)ps * nd + m) * na + a];
}

}  // namespace surfh

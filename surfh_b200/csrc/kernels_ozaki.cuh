// k4 / k4^T on the 5th-generation tensor cores (both dtypes): the spectral response as an error-free
// integer-sliced ("Ozaki scheme") product on tcgen05.mma kind::i8 with int32 accumulators in tensor memory.
//
//   forward   y[m, n]  = sum_k W [m, k]  * G [n, k]     m = detector wavelength l', k = (l, b), n = (p, s, a)
//   adjoint   Gt[n, k] = sum_m Wt[k, m]  * Yk[n, m]
//
// Replaces jax_utils.wblur_subSampling + the alpha decimation and jax_utils.wblur_t + np.repeat
// (surfh/ToolsDir/jax_utils.py:72-91; surfh/Models/spectroModelChannel.py:229, 242-252), like kernels_gemm_tma.cuh,
// whose DMMA kernel is bounded by the FP64 pipe (36 TFLOP/s on B200: tcgen05 has no f64 kind).
// The arithmetic is restated in numpy, with its error bounds, by the CPU test tests/test_oracle_ozaki.py.
//
// Arithmetic.  Every row x of an operand is written as
//     x[k] = 2^(e-6) * sum_{p<S} d_p[k] * 2^(-7p),     d_p[k] integer, |d_p[k]| <= 64   (int8)
// with e = the row's binary exponent (|x[k]| < 2^e): d_0 = rint(x 2^(6-e)), and every further digit takes the next
// 7 bits of the remainder (all steps exact in fp64).  S digits keep 6 + 7(S-1) bits below the row maximum.
// The product of a row of A and a row of B is then
//     2^(eA-6) 2^(eB-6) * sum_t 2^(-7t) L_t,    L_t = sum_{p+q=t} sum_k dA_p[k] dB_q[k]
// and every L_t is an exact integer: |L_t| <= (t+1) K 64^2 < 2^31 for K < 65536 / (t+1).  Levels t >= S are
// dropped (they sit below the digits' own truncation).  S = 8 (36 int8 products) keeps 55 bits below every row's
// maximum: the fp64 product to 1e-15 ... 1e-14 depending on the rows' dynamic range (7 digits: ~1e-12); the
// accumulation itself has no rounding at all.  The fp32 operator uses S = 4 (27 bits, 10 products).
//
// Kernel: persistent, one CTA per SM (6 warps), clusters of CL CTAs; a CTA computes 128 x 64 tiles of C.
//   warp 0 (one lane)  TMA producer: per 64-deep k-block, S + S `cp.async.bulk.tensor.3d` boxes (A digit p:
//                      128 rows x 64 bytes, B digit q: 64 rows x 64 bytes, 64-byte swizzle) into a 2-stage ring,
//                      completion by transaction bytes on the stage's `full` mbarrier; the CTAs of a cluster own
//                      neighbouring column tiles and each multicasts 1 / CL of every A tile to the cluster; digit
//                      tiles of A that are entirely zero (per-tile bit mask, see ozaki_tile_mask_kernel) are skipped;
//                      the producer runs ahead into the next tile of the CTA's schedule;
//   warp 1 (one lane)  MMA issuer: `tcgen05.mma.cta_group::1.kind::i8` (M 128, K 32); digit pair (p, q) accumulates
//                      into the level-(p+q) accumulator = 64 columns of tensor memory (S x 64 <= 512 columns), and
//                      one instruction of N = 256 covers A digit p against FOUR stacked B digits (four levels);
//                      `tcgen05.commit` (multicast to the cluster) releases the stage (`empty`) and, after the last
//                      k-block, publishes the accumulators (`tmem_full`); `tmem_empty` (the epilogue has read them)
//                      gates the next tile's first product;
//   warps 2-5          epilogue: `tcgen05.ld` 32 lanes x 8 columns of all S levels at a time, Horner sum of the
//                      levels in fp64, the two power-of-two row scales, store through the two offset tables (the
//                      detector layout [P,S,L',na] is the reference's; the slit-space layout is the gather /
//                      scatter kernels').
// Loading all S digits of both operands once per k-block and running the S(S+1)/2 products out of shared memory
// keeps the L2 -> SM traffic at 12 KB x S per 1792 tensor cycles (8 KB x S with the A multicast); that feed, not the
// tensor pipe, is what bounds the kernel (profiles/r02_ozaki.md).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels_gemm.cuh"
#include "kernels_gemm_tma.cuh"

namespace surfh {

#ifndef SURFH_OZ_TEST_BK
#define SURFH_OZ_TEST_BK 64
#endif
constexpr int kOzBM = 128, kOzBN = 64, kOzBK = SURFH_OZ_TEST_BK;      // tile of C; k-block in int8 elements (= bytes)
constexpr int kOzStages = 2;
constexpr int kOzThreads = 192;                         // producer warp, MMA warp, 4 epilogue warps
constexpr int kOzMaxDigits = 8;
constexpr int kOzATile = kOzBM * kOzBK, kOzBTile = kOzBN * kOzBK;   // bytes of one digit's tile

__host__ __device__ constexpr size_t ozaki_stage_bytes(int S) { return (size_t)S * (kOzATile + kOzBTile); }
__host__ __device__ constexpr size_t ozaki_smem_bytes(int S) { return 1024 + kOzStages * ozaki_stage_bytes(S) + 64; }

struct OzakiProblem {
    CUtensorMap a;   // int8 digits of A: [S][M][Kp], dims {K, M, S}, box {64, 128 / CL, 1}, 64-byte swizzle
    CUtensorMap b;   // int8 digits of B: [S][N][Kp], box {64, 64, 1}
    int M, N, K;
    int n_major;             // tile order: 0 = row tile outer (B re-traversed per row tile), 1 = column-tile cluster outer
                             // (A re-traversed): pick the one that re-reads the SMALLER operand, it stays in L2
    const double* sa;        // [M] 2^(eA - 6)
    const double* sb;        // [N] 2^(eB - 6)
    const uint8_t* amask;    // [ceil(M / 128)][ceil(K / 64)] or NULL: bit p set = digit p of that A tile is not all zero
    void* C;                 // C(m, n) at C[cM[m] + cN[n]], elements of the kernel's output type
    const int32_t* cM;
    const int32_t* cN;
};

struct OzakiBatch {
    int count;
    int tile_start[kMaxGemmGroup + 1];
    OzakiProblem p[kMaxGemmGroup];
    int32_t* dump;   // debugging: when non-null, CTA 0 also writes its raw level accumulators [S][128][64]
};

// Host side: one wave of co-resident clusters for the persistent kernel (0 when the query fails).
template <typename Kernel> inline int ozaki_max_resident_ctas(Kernel kernel, int cluster, size_t smem_bytes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cluster * 64));
    cfg.blockDim = dim3(kOzThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n * cluster;
}

// ---- digit extraction ---------------------------------------------------------------------------------------
// One CTA per row, all rows of up to kMaxGemmGroup matrices in one launch: row maximum (block reduction) -> exponent
// e -> S int8 digit planes + the scale 2^(e-6).  A thread owns 4 consecutive elements per sweep: two 16-byte loads
// (8-byte for fp32), one 4-byte store per digit plane, both coalesced across the warp.  Rounding to the nearest
// integer is the add-and-subtract of 1.5 * 2^52, whose sum also carries the integer in its low word.
struct OzSliceJob {
    const void* x;        // rows at pitch ld (elements, even)
    int8_t* digits;       // [S][rows][Kp], Kp multiple of 16 and >= K; the pad is written as zero
    double* scale;        // [rows]
    size_t ld;
    int rows, K, Kp;
};
struct OzSliceBatch {
    int count;
    int row_start[kMaxGemmGroup + 1];
    OzSliceJob j[kMaxGemmGroup];
};

template <typename TIn> __device__ __forceinline__ void oz_load4(const TIn* x, int k0, int K, double* v);
template <> __device__ __forceinline__ void oz_load4<double>(const double* x, int k0, int K, double* v) {
    if (k0 + 3 < K) {
        const double2 a = *reinterpret_cast<const double2*>(x + k0), b = *reinterpret_cast<const double2*>(x + k0 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = k0 + j < K ? x[k0 + j] : 0.0;
    }
}
template <> __device__ __forceinline__ void oz_load4<float>(const float* x, int k0, int K, double* v) {
    if (k0 + 3 < K) {
        const float2 a = *reinterpret_cast<const float2*>(x + k0), b = *reinterpret_cast<const float2*>(x + k0 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = k0 + j < K ? (double)x[k0 + j] : 0.0;
    }
}

template <int S, typename TIn>
__global__ void __launch_bounds__(256)
ozaki_slice_rows_kernel(const __grid_constant__ OzSliceBatch batch) {
    __shared__ double red[8];
    int ji = 0;
    while (ji + 1 < batch.count && (int)blockIdx.x >= batch.row_start[ji + 1]) ++ji;
    const OzSliceJob& job = batch.j[ji];
    const int row = blockIdx.x - batch.row_start[ji], K = job.K, Kp = job.Kp;
    const TIn* x = static_cast<const TIn*>(job.x) + (size_t)row * job.ld;
    double amax = 0.0;
    int bad = 0;   // a NaN or an infinity in the row: fmax would drop the NaN and the digits of either are meaningless
    for (int k0 = 4 * threadIdx.x; k0 < K; k0 += 4 * 256) {
        double v[4];
        oz_load4<TIn>(x, k0, K, v);
        const double m = fmax(fmax(fabs(v[0]), fabs(v[1])), fmax(fabs(v[2]), fabs(v[3])));
        bad |= !(fabs(v[0]) + fabs(v[1]) + fabs(v[2]) + fabs(v[3]) <= 1.7976931348623157e308);
        amax = fmax(amax, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    bad = __syncthreads_or(bad);
#pragma unroll
    for (int w = 0; w < 8; ++w) amax = fmax(amax, red[w]);
    // |x| < 2^e (ilogb(amax) + 1); an all-zero row keeps e = 0; rows below 2^-1000 are cut relative to 2^-1000 (the
    // scales 2^(e-6) and 2^(6-e) stay normal numbers)
    const int e = amax > 0.0 ? max(ilogb(amax) + 1, -1000) : 0;
    // a row holding a NaN / infinity poisons its whole row (or column) of the product, like the fp64 product would
    if (threadIdx.x == 0) job.scale[row] = bad ? __longlong_as_double(0x7ff8000000000000ll) : scalbn(1.0, e - 6);
    const double up = scalbn(1.0, 6 - e);             // exact power of two
    constexpr double kMagic = 6755399441055744.0;     // 1.5 * 2^52
    const size_t plane = (size_t)job.rows * Kp;
    int8_t* drow = job.digits + (size_t)row * Kp;
    for (int k0 = 4 * threadIdx.x; k0 < Kp; k0 += 4 * 256) {
        double v[4];
        oz_load4<TIn>(x, k0, K, v);
        uint32_t word[S];
#pragma unroll
        for (int p = 0; p < S; ++p) word[p] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double r = v[j] * up;                          // |r| < 64
#pragma unroll
            for (int p = 0; p < S; ++p) {
                const double t = r + kMagic;               // round to nearest; the integer sits in the low word
                word[p] |= ((uint32_t)__double2loint(t) & 0xffu) << (8 * j);
                const double d = t - kMagic;
                r = (r - d) * 128.0;                       // |r - d| <= 0.5 -> |next| <= 64, exact
            }
        }
#pragma unroll
        for (int p = 0; p < S; ++p) *reinterpret_cast<uint32_t*>(drow + (size_t)p * plane + k0) = word[p];
    }
}

// Which digit tiles of a constant operand are entirely zero?  The line-spread function decays away from its peak, so its
// leading digits vanish outside a band around the diagonal: digit 0 is non-zero in 12-28 % of the 128 x 64 tiles of the
// MRS responses, digit 1 in 40-55 %.  The product skips the loads and the instructions of those tiles.  One CTA per
// tile (grid: k-blocks x row tiles, 128 threads = rows); k-block 0 is always marked dense (its products open the
// accumulators).
template <int S>
__global__ void __launch_bounds__(128)
ozaki_tile_mask_kernel(const int8_t* __restrict__ digits, int rows, int K, int Kp, uint8_t* __restrict__ mask) {
    const int kb = blockIdx.x, mt = blockIdx.y, row = mt * kOzBM + threadIdx.x;
    unsigned bits = 0;
    if (row < rows) {
        const int k0 = kb * kOzBK, k1 = min(K, k0 + kOzBK);
#pragma unroll
        for (int p = 0; p < S; ++p) {
            const int8_t* d = digits + ((size_t)p * rows + row) * Kp;
            bool any = false;
            for (int k = k0; k < k1; k += 4) any = any || *reinterpret_cast<const uint32_t*>(d + k) != 0u;   // Kp % 16 == 0, pad = 0
            bits |= any ? (1u << p) : 0u;
        }
    }
    unsigned all = 0;
#pragma unroll
    for (int p = 0; p < S; ++p) all |= __syncthreads_or((bits >> p) & 1u) ? (1u << p) : 0u;
    if (threadIdx.x == 0) mask[(size_t)mt * gridDim.x + kb] = (uint8_t)(kb == 0 ? (1u << S) - 1u : all);
}

// ---- tcgen05 / TMEM / TMA wrappers --------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, void* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// the same box delivered to the same shared-memory offset of every CTA of the cluster named in `mask`; each
// destination's mbarrier (same offset) receives the bytes
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, void* bar,
                                               unsigned short mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(void* smem_result, unsigned cols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned cols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ... on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(void* bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, int8 x int8 -> int32, M 128 x N 64 x K 32
__device__ __forceinline__ void tc_mma_i8(unsigned d_tmem, uint64_t a_desc, uint64_t b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base lane + i)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile written by TMA with the 64-byte swizzle: rows of 64 bytes,
// 8-row groups 512 bytes apart (SBO), descriptor version 1 (sm_100), layout type 4 = SWIZZLE_64B.
// (cute::UMMA::SmemDescriptor: start >> 4 in bits [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version in
//  [46,48), layout type in [61,64).)
__device__ __forceinline__ uint64_t umma_desc_k_sw64(const void* tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;             // LBO: unused for swizzled K-major operands (canonical value 1)
    d |= (uint64_t)(512 >> 4) << 32;    // SBO
    d |= (uint64_t)1 << 46;             // version
    d |= (uint64_t)4 << 61;             // SWIZZLE_64B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D s32 (2 << 4), A / B signed int8 (1 << 7, 1 << 10), both
// K-major (bits 15, 16 clear), N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr unsigned oz_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(kOzBM >> 4) << 24);
}

#ifdef SURFH_OZAKI_WATCHDOG
__device__ __forceinline__ void oz_wait(void* bar, unsigned parity, int what) {
    unsigned done;
    long long spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1ll << 22)) {
            printf("ozaki watchdog: block %d thread %d stuck waiting on barrier %d parity %u\n", (int)blockIdx.x,
                   (int)threadIdx.x, what, parity);
            __trap();
        }
    } while (!done);
}
#else
__device__ __forceinline__ void oz_wait(void* bar, unsigned parity, int) { mbar_wait(bar, parity); }
#endif

__device__ __forceinline__ void tmem_ld8(unsigned taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}

// One tile of the grouped launch: which problem, which 128 x 64 block of its C.
struct OzTile {
    int pi, m0, n0, n_kb;
};
template <int CL> __device__ __forceinline__ OzTile oz_decode_tile(const OzakiBatch& batch, int tile) {
    OzTile t;
    t.pi = 0;
    while (t.pi + 1 < batch.count && tile >= batch.tile_start[t.pi + 1]) ++t.pi;
    const OzakiProblem& g = batch.p[t.pi];
    const int local_tile = tile - batch.tile_start[t.pi];
    const int tiles_n = ((g.N + kOzBN - 1) / kOzBN + CL - 1) / CL * CL;   // column tiles, padded to whole clusters
    if (g.n_major) {
        const int tiles_m = (g.M + kOzBM - 1) / kOzBM, group = local_tile / CL;       // group = (column cluster, row tile)
        t.m0 = (group % tiles_m) * kOzBM;
        t.n0 = ((group / tiles_m) * CL + local_tile % CL) * kOzBN;
    } else {
        t.m0 = (local_tile / tiles_n) * kOzBM;
        t.n0 = (local_tile % tiles_n) * kOzBN;
    }
    t.n_kb = (g.K + kOzBK - 1) / kOzBK;
    return t;
}

// Static schedule of the persistent kernel: in round r cluster c takes tile group r * n_clusters + c (even rounds) or
// r * n_clusters + n_clusters - 1 - c (odd rounds).  The tiles are sorted by contraction length, longest first, so the
// boustrophedon evens out the load of the clusters; every role of every CTA of a cluster walks the same sequence.
template <int CL> struct OzSchedule {
    int n_groups, n_clusters, c, rank, round;
    __device__ OzSchedule(int n_tiles, int cta_rank)
        : n_groups(n_tiles / CL), n_clusters((int)gridDim.x / CL), c((int)blockIdx.x / CL), rank(cta_rank), round(0) {}
    // next tile of this CTA, or -1
    __device__ int next() {
        while (round * n_clusters < n_groups) {
            const int g = round * n_clusters + ((round & 1) ? n_clusters - 1 - c : c);
            ++round;
            if (g < n_groups) return g * CL + rank;
        }
        return -1;
    }
};

// CL: CTAs per cluster.  The CL CTAs of a cluster own CL neighbouring column tiles of ONE row tile: each loads 1/CL
// of every A digit tile and multicasts it to the whole cluster (the L2 -> SM traffic, which bounds the kernel,
// drops from 12 KB to (8 / CL + 4) KB per digit and k-block); tile_start counts CTAs (CL per cluster).
// PERSISTENT: the grid is one wave of co-resident clusters (cudaOccupancyMaxActiveClusters); every cluster walks its
// share of the tile groups (OzSchedule).  Tensor memory, barriers and the stage ring live for the whole
// kernel; the producer runs ahead into the next tile while the epilogue drains the accumulators (`tmem_empty` tells the
// MMA issuer when it may overwrite them), so only the tensor-memory read of the epilogue is not hidden.
template <int S, int CL, typename TOut>
__global__ void __launch_bounds__(kOzThreads, 1)
ozaki_gemm_kernel(const __grid_constant__ OzakiBatch batch) {
    static_assert(S >= 1 && S <= kOzMaxDigits, "digit count");
    static_assert(CL == 1 || CL == 2 || CL == 4, "cluster size");
    constexpr unsigned short kClMask = (unsigned short)((1u << CL) - 1u);
    constexpr int kARows = kOzBM / CL;                 // rows of every A digit tile this CTA fetches for the cluster
    constexpr unsigned kTmemCols = S * kOzBN <= 64 ? 64 : S * kOzBN <= 128 ? 128 : S * kOzBN <= 256 ? 256 : 512;
    constexpr size_t kStage = ozaki_stage_bytes(S);
    extern __shared__ unsigned char oz_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + kOzStages * kStage);
    unsigned long long* empty = full + kOzStages;
    unsigned long long* tmem_full = empty + kOzStages;
    unsigned long long* tmem_empty = tmem_full + 1;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(tmem_empty + 1);

    const int n_tiles = batch.tile_start[batch.count];
    const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;   // = tile % CL for every tile of this CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kOzStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);   // every CTA of the cluster writes into this stage
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);       // one arrival per epilogue warp
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // the peers' barriers are initialised before anything is multicast to them
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            unsigned kbg = 0;   // k-blocks issued so far by this CTA: ring position and phase
            OzSchedule<CL> sched(n_tiles, cta_rank);
            for (int tile = sched.next(); tile >= 0; tile = sched.next()) {
                const OzTile t = oz_decode_tile<CL>(batch, tile);
                const OzakiProblem& g = batch.p[t.pi];
                for (int kb = 0; kb < t.n_kb; ++kb, ++kbg) {
                    const unsigned s = kbg % kOzStages;
                    oz_wait(&empty[s], ((kbg / kOzStages) & 1u) ^ 1u, 0);   // first round: passes at once
                    unsigned char* a_tiles = smem + s * kStage;
                    unsigned char* b_tiles = a_tiles + S * kOzATile;
                    // digit tiles of A that are all zero are neither loaded nor multiplied (same mask in every CTA
                    // of the cluster: they share the row tile)
                    const unsigned present = g.amask ? __ldg(g.amask + (size_t)(t.m0 / kOzBM) * t.n_kb + kb) : 0xffu;
                    mbar_expect_tx(&full[s], (unsigned)(S * kOzBTile + __popc(present & ((1u << S) - 1u)) * kOzATile));
#pragma unroll
                    for (int p = 0; p < S; ++p) {
                        if (present & (1u << p)) {
                            if (CL == 1)
                                tma_load_3d(a_tiles + p * kOzATile, &g.a, kb * kOzBK, t.m0, p, &full[s]);
                            else
                                tma_load_3d_mc(a_tiles + p * kOzATile + cta_rank * kARows * kOzBK, &g.a, kb * kOzBK,
                                               t.m0 + cta_rank * kARows, p, &full[s], kClMask);
                        }
                        tma_load_3d(b_tiles + p * kOzBTile, &g.b, kb * kOzBK, t.n0, p, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            unsigned kbg = 0, done = 0;
            OzSchedule<CL> sched(n_tiles, cta_rank);
            for (int tile = sched.next(); tile >= 0; tile = sched.next(), ++done) {
                const OzTile t = oz_decode_tile<CL>(batch, tile);
                const uint8_t* amask = batch.p[t.pi].amask;
                if (amask) amask += (size_t)(t.m0 / kOzBM) * t.n_kb;
                oz_wait(tmem_empty, (done & 1u) ^ 1u, 3);   // the epilogue has drained the previous tile's accumulators
                tc_fence_after();
                for (int kb = 0; kb < t.n_kb; ++kb, ++kbg) {
                    const unsigned s = kbg % kOzStages;
                    oz_wait(&full[s], (kbg / kOzStages) & 1u, 1);
                    tc_fence_after();
                    const unsigned char* a_tiles = smem + s * kStage;
                    const unsigned char* b_tiles = a_tiles + S * kOzATile;
                    const uint64_t a0 = umma_desc_k_sw64(a_tiles), b0 = umma_desc_k_sw64(b_tiles);
                    const unsigned present = amask ? __ldg(amask + kb) : 0xffu;   // k-block 0 is always dense
#ifndef SURFH_OZ_TEST_NO_MMA
#pragma unroll
                    for (int ks = 0; ks < kOzBK / 32; ++ks) {
#pragma unroll
                        for (int p = 0; p < S; ++p) {
                            if (!(present & (1u << p))) continue;
                            // A digit p meets the B digits q = 0 .. S-1-p, whose products belong to the levels p .. S-1:
                            // CONSECUTIVE 64-column accumulators.  The digit tiles of B are consecutive in shared memory
                            // (64 rows x 64 bytes each, i.e. one tall K-major matrix), so up to four of them are ONE
                            // instruction of N = 256: the A tile is read from shared memory once per four products
                            // (an N = 64 instruction is bound by the 128 B/clk shared-memory port, not by the tensor pipe)
                            const uint64_t ad = a0 + (uint64_t)((p * kOzATile + ks * 32) >> 4);
#pragma unroll
                            for (int q0 = 0; q0 < S - p; q0 += 4) {
                                const int cnt = (S - p - q0) < 4 ? (S - p - q0) : 4;
                                const uint64_t bd = b0 + (uint64_t)((q0 * kOzBTile + ks * 32) >> 4);
                                const unsigned acc = (kb > 0 || ks > 0 || p > 0) ? 1u : 0u;   // the p = 0 products open every level
                                tc_mma_i8(tmem_base + (unsigned)((p + q0) * kOzBN), ad, bd, oz_idesc(cnt * kOzBN), acc);
                            }
                        }
                    }
#endif
                    // the stage is free once these MMAs have read it -- in every CTA of the cluster
                    if (CL == 1) tc_commit(&empty[s]); else tc_commit_mc(&empty[s], kClMask);
                    if (kb == t.n_kb - 1) tc_commit(tmem_full);   // ... and the accumulators are final
                }
            }
        }
    } else {
        // ---- epilogue: warp w owns tensor-memory lanes 32 (w % 4) .. + 31 = rows of the tile ---------------
        const int quad = warp & 3;
        const unsigned lane_base = tmem_base + ((unsigned)(quad * 32) << 16);
        unsigned done = 0;
        OzSchedule<CL> sched(n_tiles, cta_rank);
        for (int tile = sched.next(); tile >= 0; tile = sched.next(), ++done) {
            const OzTile t = oz_decode_tile<CL>(batch, tile);
            const OzakiProblem& g = batch.p[t.pi];
            const int m = t.m0 + quad * 32 + lane;
            const bool m_ok = m < g.M;
            const double sa = m_ok ? __ldg(g.sa + m) : 0.0;
            const int32_t cm = m_ok ? __ldg(g.cM + m) : 0;
            const bool dump = batch.dump != nullptr && tile == 0;
            oz_wait(tmem_full, done & 1u, 2);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < kOzBN / 8; ++c) {
                // all S levels of 8 columns in flight, one wait
                uint32_t v[S][8];
#pragma unroll
                for (int lv = 0; lv < S; ++lv) tmem_ld8(lane_base + (unsigned)(lv * kOzBN + c * 8), v[lv]);
                tmem_ld_wait();
                if (c == kOzBN / 8 - 1) {   // the accumulators are in registers: the next tile's products may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                }
                if (dump) {
#pragma unroll
                    for (int lv = 0; lv < S; ++lv)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            batch.dump[((size_t)lv * kOzBM + quad * 32 + lane) * kOzBN + c * 8 + j] = (int32_t)v[lv][j];
                }
                double acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = (double)(int32_t)v[S - 1][j];
#pragma unroll
                for (int lv = S - 2; lv >= 0; --lv)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fma(acc[j], 0.0078125, (double)(int32_t)v[lv][j]);   // Horner in 2^-7
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = t.n0 + c * 8 + j;
                    if (m_ok && n < g.N) static_cast<TOut*>(g.C)[(size_t)cm + __ldg(g.cN + n)] = (TOut)(acc[j] * sa * __ldg(g.sb + n));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace surfh

// Stand-alone bring-up / timing harness of the int8-sliced tcgen05 contraction (kernels_ozaki.cuh): no torch, no
// Python.  Build (here, no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DSURFH_OZAKI_WATCHDOG \
//        -I surfh_b200/csrc tools/ozaki_test.cu -o tools/_build/ozaki_test
// Run (GPU box):  tools/_build/ozaki_test M N K S [reps]
// Checks: (1) the digits reproduce the operands, (2) the raw int32 level accumulators of tile 0 against an integer
// reference, (3) sampled entries of C against a long-double reference, (4) time against the same flops of DGEMM.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "kernels_ozaki.cuh"

using namespace surfh;

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            std::exit(2);                                                                         \
        }                                                                                         \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap digits_map(const void* base, int K, int rows, int Kp, int S, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    CUtensorMap m;
    const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t gstride[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * rows};
    const cuuint32_t box[3] = {(cuuint32_t)kOzBK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, kOzBK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        std::fprintf(stderr, "cuTensorMapEncodeTiled failed (%d)\n", (int)r);
        std::exit(2);
    }
    return m;
}

template <int S, int CL>
static void run(int M, int N, int K, int reps) {
    const int Kp = (K + 15) / 16 * 16;
    std::mt19937_64 rng(1234);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<double> A((size_t)M * K), B((size_t)N * K);
    // rows with a wide dynamic range, like a line-spread function row (peak + small tails) and a spectrum
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            const double d = (double)(k - (m * (long long)K) / M) / 6.0;
            A[(size_t)m * K + k] = 1.0 / (1.0 + d * d) * (1.0 + 0.1 * nd(rng));
        }
    for (size_t i = 0; i < B.size(); ++i) B[i] = nd(rng) * std::exp(2.0 * nd(rng));
    double *dA, *dB, *dC, *dsa, *dsb;
    int8_t *dAd, *dBd;
    int32_t *dcM, *dcN, *ddump;
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, (size_t)M * N * 8));
    CK(cudaMalloc(&dsa, M * 8));
    CK(cudaMalloc(&dsb, N * 8));
    CK(cudaMalloc(&dAd, (size_t)S * M * Kp));
    CK(cudaMalloc(&dBd, (size_t)S * N * Kp));
    CK(cudaMalloc(&dcM, M * 4));
    CK(cudaMalloc(&dcN, N * 4));
    CK(cudaMalloc(&ddump, (size_t)S * kOzBM * kOzBN * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    std::vector<int32_t> cM(M), cN(N);
    for (int m = 0; m < M; ++m) cM[m] = m * N;
    for (int n = 0; n < N; ++n) cN[n] = n;
    CK(cudaMemcpy(dcM, cM.data(), M * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dcN, cN.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 8));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    OzSliceBatch sa_job, sb_job;
    std::memset(&sa_job, 0, sizeof(sa_job));
    std::memset(&sb_job, 0, sizeof(sb_job));
    sa_job.count = 1; sa_job.row_start[1] = M; sa_job.j[0] = OzSliceJob{dA, dAd, dsa, (size_t)K, M, K, Kp};
    sb_job.count = 1; sb_job.row_start[1] = N; sb_job.j[0] = OzSliceJob{dB, dBd, dsb, (size_t)K, N, K, Kp};
    ozaki_slice_rows_kernel<S, double><<<M, 256>>>(sa_job);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) ozaki_slice_rows_kernel<S, double><<<N, 256>>>(sb_job);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms_slice;
    CK(cudaEventElapsedTime(&ms_slice, e0, e1));
    ms_slice /= reps;

    // (1) digits reproduce the operands
    std::vector<int8_t> Ad((size_t)S * M * Kp), Bd((size_t)S * N * Kp);
    std::vector<double> sa(M), sb(N);
    CK(cudaMemcpy(Ad.data(), dAd, Ad.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Bd.data(), dBd, Bd.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sa.data(), dsa, M * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sb.data(), dsb, N * 8, cudaMemcpyDeviceToHost));
    double worst_digit = 0;
    for (int m = 0; m < M; m += 7)
        for (int k = 0; k < K; ++k) {
            long double v = 0;
            for (int p = S - 1; p >= 0; --p) v = v / 128.0L + Ad[((size_t)p * M + m) * Kp + k];
            const double err = std::fabs((double)(v * sa[m] - A[(size_t)m * K + k])) / (sa[m] * 64);
            worst_digit = std::fmax(worst_digit, err);
        }
    std::printf("digits: worst |x - sum digits| / 2^e = %.3e (expected <= %.3e)\n", worst_digit, std::ldexp(0.5, -6 - 7 * (S - 1)));

    OzakiBatch batch;
    std::memset(&batch, 0, sizeof(batch));
    batch.count = 1;
    batch.tile_start[0] = 0;
    batch.tile_start[1] = ((M + kOzBM - 1) / kOzBM) * (((N + kOzBN - 1) / kOzBN + CL - 1) / CL * CL);
    OzakiProblem& g = batch.p[0];
    g.a = digits_map(dAd, K, M, Kp, S, kOzBM / CL);
    g.b = digits_map(dBd, K, N, Kp, S, kOzBN);
    uint8_t* dmask = nullptr;
    const int m_tiles = (M + kOzBM - 1) / kOzBM, n_kb = (K + kOzBK - 1) / kOzBK;
    if (!std::getenv("OZ_NOMASK")) {
        CK(cudaMalloc(&dmask, (size_t)m_tiles * n_kb));
        ozaki_tile_mask_kernel<S><<<dim3(n_kb, m_tiles), 128>>>(dAd, M, K, Kp, dmask);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> hm((size_t)m_tiles * n_kb);
        CK(cudaMemcpy(hm.data(), dmask, hm.size(), cudaMemcpyDeviceToHost));
        long long present = 0;
        for (uint8_t b : hm) present += __builtin_popcount(b);
        std::printf("A digit tiles present: %.1f %% of %lld\n", 100.0 * present / ((double)hm.size() * S), (long long)hm.size() * S);
    }
    g.amask = dmask;
    g.M = M; g.N = N; g.K = K; g.sa = dsa; g.sb = dsb; g.C = dC; g.cM = dcM; g.cN = dcN;
    batch.dump = ddump;
    CK(cudaFuncSetAttribute(ozaki_gemm_kernel<S, CL, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ozaki_smem_bytes(S)));
    const int resident = ozaki_max_resident_ctas(ozaki_gemm_kernel<S, CL, double>, CL, ozaki_smem_bytes(S));
    std::printf("persistent grid: %d co-resident CTAs for %d tiles\n", resident, batch.tile_start[1]);
    auto launch = [&]() {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(resident > 0 && resident < batch.tile_start[1] ? resident : batch.tile_start[1]);
        cfg.blockDim = dim3(kOzThreads);
        cfg.dynamicSmemBytes = ozaki_smem_bytes(S);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<S, CL, double>, batch));
    };
    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    // (2) raw level accumulators of tile 0
    std::vector<int32_t> dump((size_t)S * kOzBM * kOzBN);
    CK(cudaMemcpy(dump.data(), ddump, dump.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0, checked = 0;
    for (int t = 0; t < S; ++t)
        for (int i = 0; i < kOzBM && i < M; i += 3)
            for (int j = 0; j < kOzBN && j < N; j += 5) {
                long long ref = 0;
                for (int p = 0; p <= t; ++p) {
                    const int q = t - p;
                    const int8_t* ap = &Ad[((size_t)p * M + i) * Kp];
                    const int8_t* bq = &Bd[((size_t)q * N + j) * Kp];
                    for (int k = 0; k < K; ++k) ref += (int)ap[k] * (int)bq[k];
                }
                const long long got = dump[((size_t)t * kOzBM + i) * kOzBN + j];
                ++checked;
                if (got != ref) {
                    if (bad < 8) std::printf("  level %d (%d, %d): got %lld expected %lld\n", t, i, j, got, ref);
                    ++bad;
                }
            }
    std::printf("tile 0 level accumulators: %lld of %lld sampled entries differ\n", bad, checked);

    // (3) sampled entries of C
    std::vector<double> C((size_t)M * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    std::uniform_int_distribution<int> um(0, M - 1), un(0, N - 1);
    long double num = 0, den = 0;
    double worst = 0;
    for (int it = 0; it < 20000; ++it) {
        const int m = it < 64 ? (it % 2 ? M - 1 : 0) : um(rng), n = it < 64 ? (it % 3 ? N - 1 : 0) : un(rng);
        long double ref = 0;
        for (int k = 0; k < K; ++k) ref += (long double)A[(size_t)m * K + k] * (long double)B[(size_t)n * K + k];
        const long double d = (long double)C[(size_t)m * N + n] - ref;
        num += d * d;
        den += ref * ref;
        worst = std::fmax(worst, (double)std::fabs((double)d));
    }
    std::printf("C (M %d N %d K %d, %d digits): relative L2 of 20000 sampled entries %.3e, worst abs %.3e, rms |C| %.3e\n", M, N, K, S,
                (double)std::sqrt((double)(num / den)), worst, (double)std::sqrt((double)(den / 20000)));

    // (4) timing
    batch.dump = nullptr;
    for (int r = 0; r < 2; ++r) launch();
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double flops = 2.0 * M * N * K;
    std::printf("time: gemm %.3f ms = %.1f TFLOP/s fp64-equivalent (%.2f POP/s int8 over %d products), slicing B %.3f ms\n", ms,
                flops / ms * 1e-9, flops * (S * (S + 1) / 2) / ms * 1e-12, S * (S + 1) / 2, ms_slice);
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? std::atoi(argv[1]) : 128, N = argc > 2 ? std::atoi(argv[2]) : 64, K = argc > 3 ? std::atoi(argv[3]) : 64;
    const int S = argc > 4 ? std::atoi(argv[4]) : 1, reps = argc > 5 ? std::atoi(argv[5]) : 5;
    const int CL = argc > 6 ? std::atoi(argv[6]) : 1;
    switch (S * 10 + CL) {
        case 21: run<2, 1>(M, N, K, reps); break;
        case 22: run<2, 2>(M, N, K, reps); break;
        case 24: run<2, 4>(M, N, K, reps); break;
        case 71: run<7, 1>(M, N, K, reps); break;
        case 72: run<7, 2>(M, N, K, reps); break;
        case 74: run<7, 4>(M, N, K, reps); break;
        case 81: run<8, 1>(M, N, K, reps); break;
        case 82: run<8, 2>(M, N, K, reps); break;
        case 84: run<8, 4>(M, N, K, reps); break;
        default: std::fprintf(stderr, "S must be 2, 7 or 8 and CL 1, 2 or 4\n"); return 1;
    }
    return 0;
}

// C ABI of surfh_b200 (see include/surfh_b200.h): handle, tables, cuFFT plans, the forward /
// adjoint pipelines and the CG vector primitives.  sm_100a only, no CPU fallback.
#include <cuda_runtime.h>
#include <cufft.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/surfh_b200.h"
#include "common.cuh"
#include "host_util.cuh"
#include "fft_plan.cuh"
#include "kernels_cg.cuh"
#include "kernels_gemm.cuh"
#include "kernels_gemm_tma.cuh"
#include "kernels_ozaki.cuh"
#include "kernels_lmm.cuh"
#include "kernels_precond.cuh"
#include "kernels_shepard.cuh"
#include "kernels_slit.cuh"

namespace surfh {

enum Stage {
    ST_RFFT_MAPS = 0, ST_LMM_OTF_FWD, ST_IRFFT_CUBE, ST_SLIT_GATHER, ST_GEMM_FWD,
    ST_GEMM_ADJ, ST_SLIT_SCATTER, ST_RFFT_CUBE, ST_LMM_OTF_ADJ, ST_IRFFT_MAPS, ST_MEMSET, ST_CG, ST_COUNT
};
// the four FFT stages are prefixed "chirpz_" (hand-written kernels) or "cufft_" at read time
static const char* kStageNames[ST_COUNT] = {
    "rfft_maps", "lmm_otf_fwd", "irfft_cube", "slit_gather", "spectral_gemm_fwd",
    "spectral_gemm_adj", "slit_scatter", "rfft_cube", "lmm_otf_adj", "irfft_maps", "memset", "cg_fused"};
static const char* kStageNamesOwnFft[ST_COUNT] = {
    "chirpz_rfft_maps", "lmm_otf_fwd", "chirpz_irfft_cube", "slit_gather", "spectral_gemm_fwd",
    "spectral_gemm_adj", "slit_scatter", "chirpz_rfft_cube", "lmm_otf_adj", "chirpz_irfft_maps", "memset", "cg_fused"};
static const char* kStageNamesCufft[ST_COUNT] = {
    "cufft_rfft_maps", "lmm_otf_fwd", "cufft_irfft_cube", "slit_gather", "spectral_gemm_fwd",
    "spectral_gemm_adj", "slit_scatter", "cufft_rfft_cube", "lmm_otf_adj", "cufft_irfft_maps", "memset", "cg_fused"};

}  // namespace surfh

using namespace surfh;

// ------------------------------------------------------------------------------------------------
struct surfh_model {
    int dtype = SURFH_F64;
    int K = 0, Na = 0, Nb = 0, Nl = 0, Nh = 0, chunk = 0;
    size_t plane = 0, nf = 0, nfp = 0;
    bool finalized = false;
    std::string last_error;
    int64_t launches = 0, own_launches = 0;
    int device = 0;
    bool own_fft_names = false;

    // profiling
    bool profiling = false;
    struct Rec { int stage; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    double stage_bytes[ST_COUNT] = {0}, stage_flops[ST_COUNT] = {0};
    int stage_launches[ST_COUNT] = {0};

    virtual ~surfh_model() {
        for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    }
    virtual void set_otf(int l_start, int l_count, const void* src) = 0;
    virtual void add_band(const surfh_band_desc* b) = 0;
    virtual void finalize() = 0;
    virtual int64_t input_size() const = 0;
    virtual int64_t output_size() const = 0;
    virtual int64_t workspace_bytes() const = 0;
    virtual void contraction_info(int32_t* mode, int32_t* digits, double* executed_fraction) const = 0;
    virtual void forward(const void* x, void* y, cudaStream_t st) = 0;
    virtual void adjoint(const void* y, void* x, int mode, cudaStream_t st) = 0;
    virtual void fwadj(const void* x, void* out, int mode, void* yscratch, cudaStream_t st) = 0;
    virtual void maps_to_cube(const void* maps, float* cube, cudaStream_t st) = 0;
    virtual void forward_host(const double* x, double* y) = 0;
    virtual void adjoint_host(const double* y, double* x, int mode) = 0;
    virtual void cg_regularise_dot(const void* d, void* q, double mu_s, double mu_r, double* s, cudaStream_t st) = 0;
    virtual void laplacian_axpby(const void* x, void* out, double a, double b, cudaStream_t st) = 0;
    virtual void cg_start(const void* b, const void* q, void* r, void* d, double* s, cudaStream_t st) = 0;
    virtual void cg_update(void* x, void* r, void* d, const void* q, double* s, cudaStream_t st) = 0;
    virtual void cg_refresh(int phase, void* x, void* r, void* d, const void* b, const void* qx, double* s,
                            cudaStream_t st) = 0;
    virtual void criterion_terms(const void* y, const void* hx, int64_t n, const void* x, double* out,
                                 cudaStream_t st) = 0;
    virtual void cg_dot_x_b_plus_r(const void* x, const void* b, const void* r, double* out, cudaStream_t st) = 0;
    virtual void axpy_device_scalar(void* y, const void* x, int64_t n, const double* s, int idx, cudaStream_t st) = 0;
    virtual void precond_build(const double* w, double mu_s, double mu_r, int joint) = 0;
    virtual void precond_apply(const void* r, void* z, cudaStream_t st) = 0;
    virtual void pcg_update(int phase, void* x, void* r, const void* d, const void* q, const void* b, double* s,
                            cudaStream_t st) = 0;
    virtual void pcg_direction(const void* r, const void* z, void* d, double* s, int first, cudaStream_t st) = 0;

    struct Scope {
        surfh_model* m; int stage; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
        Scope(surfh_model* m_, int stage_, cudaStream_t st_, double bytes, double flops, int n_launch, bool own)
            : m(m_), stage(stage_), st(st_) {
            m->launches += n_launch;
            if (own) m->own_launches += n_launch;
            if (m->profiling) {
                cudaEventCreate(&a);
                cudaEventCreate(&b);
                cudaEventRecord(a, st);
                m->stage_bytes[stage] += bytes;
                m->stage_flops[stage] += flops;
                m->stage_launches[stage] += n_launch;
            }
        }
        ~Scope() {
            if (m->profiling) {
                cudaEventRecord(b, st);
                m->recs.push_back({stage, a, b});
            }
        }
    };
};

namespace surfh {

template <typename T> struct BandT {
    int P, S, na, nb, srf, A, B, l0, nl, nd, ncol, Nn, KB, mode, det_start;
    int KBp = 0;  // KB rounded up to even: row pitch of the LSF and of G (TMA wants 16-byte multiples)
    int row_lo = 0, row_hi = 0;  // cube rows this band reads (gather) or writes (either adjoint table)
    int64_t footprint = 0;       // distinct cube pixels the bilinear taps of ALL pointings read (union)
    int64_t out_offset, out_size;
    DevBuf slit_a0, slit_b0, slit_w, lsf, grid_base, grid_frac;
    DevBuf csr_pix[2], csr_ptr[2], csr_col[2], csr_val[2];
    int csr_rows[2] = {0, 0};
    int64_t csr_nnz[2] = {0, 0};
    DevBuf t_ident, t_wrow, t_gK, t_gN, t_yM, t_yN;
    // TMA contraction (fp64): transposed LSF copy, K-fast detector block, epilogue tables in n' order, tensor maps
    DevBuf lsf_t, yk;
    int ndp = 0;                       // nd rounded up to even: row pitch of lsf_t / yk (16-byte multiple)
    CUtensorMap map_w, map_g, map_wt, map_yk;
    bool tma_ready = false;
    // int8-sliced tcgen05 contraction (kernels_ozaki.cuh): digit planes [S][rows][pitch] + per-row scales of the LSF
    // and its transpose (cut once), of the slit-space vector G and of the K-fast detector block (cut per call)
    DevBuf oz_w, oz_wt, oz_g, oz_yk, oz_sw, oz_swt, oz_sg, oz_syk;
    DevBuf oz_mask_w, oz_mask_wt;       // per 128 x 64 tile of the LSF / its transpose: which digits are not all zero
    int oz_kq = 0, oz_ndq = 0;          // digit row pitches: KB / nd rounded up to 16 bytes
    CUtensorMap ozmap_w, ozmap_g, ozmap_wt, ozmap_yk;
    bool oz_ready = false;
    DevBuf G;  // [nl][ncol] slit-space vector (forward G / adjoint Gt), columns in INTERNAL order
    // The ABI (and the detector) order slit-space columns as ((p*S + s)*na + a)*nb + b; internally they are
    // stored as ((p*na + a)*S + s)*nb + b, so that the 32 consecutive cube pixels a warp of the scatter owns
    // (and the 32 outputs a warp of the gather produces) touch one contiguous run of G instead of one run of
    // nb per slit.
    int internal_col(int c) const {
        const int bb = c % nb;
        int r = c / nb;
        const int aa = r % na;
        r /= na;
        const int ss = r % S, pp = r / S;
        return ((pp * na + aa) * S + ss) * nb + bb;
    }
    // Element (l, column, b) of G lives at column * g_col + l * g_l + b (see SlitTables): K-fast per detector
    // column in the detector's order for bands with a spectral response, [l][n'][b] for beta-sum bands.
    int g_col = 0, g_l = 0, g_psa = 0;
    // detector sample n = (p*S + s)*na + a  ->  its column of G
    int column_of_sample(int n) const {
        if (g_psa) return n;
        const int aa = n % na, ps = n / na, ss = ps % S, pp = ps / S;
        return (pp * na + aa) * S + ss;
    }
    // offset in plane 0 of G of the ABI slit-space column c = ((p*S + s)*na + a)*nb + b
    int32_t g_offset(int c) const { return (int32_t)((int64_t)column_of_sample(c / nb) * g_col + c % nb); }

    SlitTables<T> slit_tables() const {
        SlitTables<T> t;
        t.slit_a0 = slit_a0.as<int32_t>();
        t.slit_b0 = slit_b0.as<int32_t>();
        t.slit_w = slit_w.as<T>();
        t.grid_base = grid_base.as<int32_t>();
        t.grid_frac = grid_frac.as<T>();
        t.P = P; t.S = S; t.na = na; t.nb = nb; t.srf = srf; t.A = A; t.B = B; t.ncol = ncol;
        t.g_col = g_col; t.g_l = g_l; t.g_psa = g_psa;
        return t;
    }
    CsrTable<T> csr(int mode) const {
        CsrTable<T> c;
        c.row_pixel = csr_pix[mode].as<int32_t>();
        c.slice_ptr = csr_ptr[mode].as<int64_t>();
        c.col = csr_col[mode].as<int32_t>();
        c.val = csr_val[mode].as<T>();
        c.n_rows = csr_rows[mode];
        return c;
    }
};

template <typename T> struct FftTraits;
template <> struct FftTraits<float> {
    static constexpr cufftType R2C = CUFFT_R2C, C2R = CUFFT_C2R;
    static cufftResult fwd(cufftHandle p, float* in, float2* out) { return cufftExecR2C(p, in, out); }
    static cufftResult inv(cufftHandle p, float2* in, float* out) { return cufftExecC2R(p, in, out); }
};
template <> struct FftTraits<double> {
    static constexpr cufftType R2C = CUFFT_D2Z, C2R = CUFFT_Z2D;
    static cufftResult fwd(cufftHandle p, double* in, double2* out) { return cufftExecD2Z(p, in, out); }
    static cufftResult inv(cufftHandle p, double2* in, double* out) { return cufftExecZ2D(p, in, out); }
};

// GEMM tile configuration per dtype
template <typename T> struct GemmCfg;
template <> struct GemmCfg<double> { static constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4; };
template <> struct GemmCfg<float> { static constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8; };

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SURFH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) throw Error(SURFH_ECUDA, "cuTensorMapEncodeTiled is not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// fp64 matrix [rows][inner] with `inner` contiguous and `pitch` elements between rows -> tiles of box_rows x 16
// doubles (128 bytes) with the 128-byte swizzle; out-of-range elements read as zero
static CUtensorMap tensor_map_2d_f64(const void* base, int inner, int rows, size_t pitch, int box_rows) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)kTBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim, gstride, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(SURFH_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// int8 digit planes [S][rows][pitch] -> boxes of box_rows x 64 bytes of one digit, 64-byte swizzle, zero fill
static CUtensorMap tensor_map_digits(const void* base, int inner, int rows, int pitch, int digits, int box_rows) {
    CUtensorMap m;
    const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)digits};
    const cuuint64_t gstride[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)rows};
    const cuuint32_t box[3] = {(cuuint32_t)kOzBK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstride, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(SURFH_ECUDA, "cuTensorMapEncodeTiled (digits) failed (" + std::to_string((int)r) + ")");
    return m;
}

// How the spectral response is evaluated.  GEMM_OZAKI (default, both dtypes): int8-sliced product on tcgen05
// (kernels_ozaki.cuh) -- fp64: SURFH_OZAKI_DIGITS = 6, 7 or 8 (default) digits; fp32: 4 digits.
// SURFH_F64_GEMM = ozaki | tma (DMMA fed by TMA, kernels_gemm_tma.cuh) | mma (round 1's offset-table DMMA kernel);
// SURFH_F32_GEMM = ozaki | tf32 (3xTF32 mma.sync) | simt (FFMA).
enum { GEMM_LEGACY = 0, GEMM_TMA = 1, GEMM_OZAKI = 2, GEMM_SIMT = 3 };
static int gemm_mode_env(bool f64) {
    const char* e = std::getenv(f64 ? "SURFH_F64_GEMM" : "SURFH_F32_GEMM");
    if (!e || !*e || std::strcmp(e, "ozaki") == 0) return GEMM_OZAKI;
    if (f64) {
        if (std::strcmp(e, "tma") == 0 || std::strcmp(e, "dmma") == 0) return GEMM_TMA;
        if (std::strcmp(e, "mma") == 0) return GEMM_LEGACY;
        throw Error(SURFH_EINVAL, std::string("SURFH_F64_GEMM must be ozaki, tma or mma, not ") + e);
    }
    if (std::strcmp(e, "tf32") == 0 || std::strcmp(e, "mma") == 0 || std::strcmp(e, "tensor") == 0) return GEMM_LEGACY;
    if (std::strcmp(e, "simt") == 0) return GEMM_SIMT;
    throw Error(SURFH_EINVAL, std::string("SURFH_F32_GEMM must be ozaki, tf32 or simt, not ") + e);
}
static int ozaki_digits_env(bool f64) {
    if (!f64) return 4;   // 6 + 3 x 7 = 27 bits below the row maximum
    const char* e = std::getenv("SURFH_OZAKI_DIGITS");
    if (!e || !*e) return 8;
    const int d = std::atoi(e);
    if (d < 6 || d > 8) throw Error(SURFH_EINVAL, "SURFH_OZAKI_DIGITS must be 6, 7 or 8");
    return d;
}
constexpr int kOzCluster = 2;   // CTAs per cluster of the sliced contraction (A digit tiles multicast)

template <typename T> struct ModelImpl : surfh_model {
    using C = cplx_t<T>;
    DevBuf otf;       // [Nl][nfp] complex
    DevBuf tpl;       // [K][Nl] real, pre-scaled by 1/(Na*Nb)
    DevBuf tpl_raw;   // [K][Nl] real, unscaled (maps_to_cube)
    DevBuf xhat;      // [K][nfp] complex
    DevBuf spec;      // [chunk][nfp] complex
    DevBuf cubebuf;   // [chunk][cplane] real
    size_t cplane = 0;  // elements between planes of the working cube: `plane` rounded up to even with the hand-written
                        // FFT, so that every row pair starts 16-byte aligned (one TMA bulk copy stages it)
    DevBuf fft_work;
    // second lane of the chunk pipeline (two chunks in flight on two streams, see forward / adjoint)
    DevBuf spec2, cubebuf2, zbuf2;
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_acc[2] = {nullptr, nullptr};
    bool two_lanes = false;
    DevBuf zbuf;      // [max(chunk, K)][z_plane] complex: intermediate of the hand-written FFT passes
    OwnFft2d<T> ownfft;
    DevBuf plane_pairs;                   // [Nl] int2 (first row pair, count) each plane's FFT must cover
    std::vector<int> plane_pair_cnt;      // host copy of the counts
    int fft_backend = SURFH_FFT_AUTO;
    bool use_own_fft = false;
    bool prune_rows = true;  // SURFH_FFT_PRUNE=0 transforms every row (A/B measurements)
    DevBuf y_internal, x_stage, y_stage, dbl_stage;
    DevBuf cg_partial, cg_ticket;
    std::vector<std::unique_ptr<BandT<T>>> bands;
    std::vector<std::pair<int, int>> ranges;  // union of band wavelength windows
    std::map<std::pair<int, int>, cufftHandle> plans;  // (kind, batch) -> plan
    size_t fft_work_bytes = 0;
    std::vector<uint8_t> otf_set;

    ~ModelImpl() override {
        for (auto& kv : plans) cufftDestroy(kv.second);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        for (cudaEvent_t e : {ev_fork, ev_join, ev_acc[0], ev_acc[1]})
            if (e) cudaEventDestroy(e);
    }

    void init(const surfh_model_desc* d) {
        SURFH_REQUIRE(d->n_alpha > 1 && d->n_beta > 1 && d->n_lambda > 0, "bad cube shape");
        SURFH_REQUIRE(d->n_templates >= 0 && d->n_templates <= kMaxTemplates, "n_templates must be in [0, 8]");
        SURFH_CUDA(cudaGetDevice(&device));
        K = d->n_templates; Na = d->n_alpha; Nb = d->n_beta; Nl = d->n_lambda; Nh = Nb / 2 + 1;
        plane = (size_t)Na * Nb;
        nf = (size_t)Na * Nh;
        nfp = (nf + 15) / 16 * 16;
        chunk = d->chunk > 0 ? d->chunk : 0;
        fft_backend = d->fft_backend;
        if (const char* e = std::getenv("SURFH_FFT_BACKEND")) {  // A/B switch for benchmarks and tests
            if (!std::strcmp(e, "cufft")) fft_backend = SURFH_FFT_CUFFT;
            else if (!std::strcmp(e, "own")) fft_backend = SURFH_FFT_OWN;
            else if (!std::strcmp(e, "auto")) fft_backend = SURFH_FFT_AUTO;
            else throw Error(SURFH_EINVAL, "SURFH_FFT_BACKEND must be auto, own or cufft");
        }
        SURFH_REQUIRE(fft_backend == SURFH_FFT_AUTO || fft_backend == SURFH_FFT_CUFFT || fft_backend == SURFH_FFT_OWN,
                      "unknown fft_backend");
        SURFH_REQUIRE(fft_backend != SURFH_FFT_OWN || OwnFft2d<T>::supported(Na, Nb),
                      "fft_backend = own needs both map axes <= 512 pixels");
        if (const char* e = std::getenv("SURFH_FFT_PRUNE")) prune_rows = std::strcmp(e, "0") != 0;
        use_own_fft = fft_backend == SURFH_FFT_OWN || (fft_backend == SURFH_FFT_AUTO && OwnFft2d<T>::supported(Na, Nb));
        own_fft_names = use_own_fft;
        otf.alloc((size_t)Nl * nfp * sizeof(C));
        SURFH_CUDA(cudaMemset(otf.p, 0, otf.bytes));
        otf_set.assign(Nl, 0);
        if (K > 0) {
            SURFH_REQUIRE(d->templates != nullptr, "templates missing");
            std::vector<double> host((size_t)K * Nl);
            SURFH_CUDA(cudaMemcpy(host.data(), d->templates, host.size() * sizeof(double), cudaMemcpyDefault));
            upload_converted<T>(tpl_raw, host.data(), host.size());
            const double scale = 1.0 / ((double)Na * (double)Nb);
            for (auto& v : host) v *= scale;
            upload_converted<T>(tpl, host.data(), host.size());
        }
        cg_partial.alloc(sizeof(double) * 2 * kCgMaxBlocks);
        cg_ticket.alloc(sizeof(unsigned int));
        SURFH_CUDA(cudaMemset(cg_ticket.p, 0, sizeof(unsigned int)));
    }

    void set_otf(int l_start, int l_count, const void* src) override {
        SURFH_REQUIRE(l_start >= 0 && l_count > 0 && l_start + l_count <= Nl, "otf plane range out of bounds");
        SURFH_REQUIRE(src != nullptr, "otf pointer is NULL");
        cudaPointerAttributes attr;
        const double2* dsrc = nullptr;
        DevBuf tmp;
        cudaError_t e = cudaPointerGetAttributes(&attr, src);
        if (e == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
            dsrc = reinterpret_cast<const double2*>(src);
        } else {
            cudaGetLastError();
            tmp.alloc((size_t)l_count * nf * sizeof(double2));
            SURFH_CUDA(cudaMemcpy(tmp.p, src, tmp.bytes, cudaMemcpyHostToDevice));
            dsrc = tmp.as<double2>();
        }
        dim3 grid(ceil_div(nfp, 256), l_count);
        // the hand-written FFT keeps every half spectrum transposed ([Nh][Na]: contiguous columns); cuFFT's is [Na][Nh]
        otf_convert_kernel<T><<<grid, 256>>>(dsrc, otf.as<C>() + (size_t)l_start * nfp, nf, nfp, l_count, Na, Nh,
                                             use_own_fft ? 1 : 0);
        SURFH_CUDA(cudaGetLastError());
        SURFH_CUDA(cudaDeviceSynchronize());
        for (int l = l_start; l < l_start + l_count; ++l) otf_set[l] = 1;
    }

    static void check_csr(const surfh_csr& c, int64_t npix, int ncol, const char* what) {
        SURFH_REQUIRE(c.n_rows >= 0 && c.nnz >= 0, std::string(what) + ": negative size");
        if (c.n_rows == 0) return;
        SURFH_REQUIRE(c.row_pixel && c.row_ptr && c.col && c.val, std::string(what) + ": NULL table");
        SURFH_REQUIRE(c.row_ptr[0] == 0 && c.row_ptr[c.n_rows] == c.nnz, std::string(what) + ": row_ptr inconsistent");
        for (int r = 0; r < c.n_rows; ++r) {
            SURFH_REQUIRE(c.row_pixel[r] >= 0 && c.row_pixel[r] < npix, std::string(what) + ": pixel out of range");
            SURFH_REQUIRE(c.row_ptr[r + 1] >= c.row_ptr[r], std::string(what) + ": row_ptr not monotone");
            SURFH_REQUIRE(r == 0 || c.row_pixel[r] > c.row_pixel[r - 1], std::string(what) + ": rows not strictly increasing");
        }
        for (int64_t e = 0; e < c.nnz; ++e)
            SURFH_REQUIRE(c.col[e] >= 0 && c.col[e] < ncol, std::string(what) + ": column out of range");
    }

    void add_band(const surfh_band_desc* d) override {
        SURFH_REQUIRE(!finalized, "add_band after finalize");
        SURFH_REQUIRE(d->n_pointing > 0 && d->n_slit > 0 && d->na > 0 && d->nb > 0 && d->srf > 0, "bad band sizes");
        SURFH_REQUIRE(d->local_a > 1 && d->local_b > 1 && d->n_det > 0 && d->n_wave > 0, "bad band sizes");
        SURFH_REQUIRE(d->wave_start >= 0 && d->wave_start + d->n_wave <= Nl, "band wavelength window outside the cube");
        SURFH_REQUIRE(d->slit_a0 && d->slit_b0 && d->slit_w && d->grid_base && d->grid_frac, "NULL band table");
        SURFH_REQUIRE(d->spectral_mode == SURFH_SPECTRAL_LSF || d->spectral_mode == SURFH_SPECTRAL_BETA_SUM,
                      "unknown spectral_mode");
        SURFH_REQUIRE(d->spectral_mode == SURFH_SPECTRAL_BETA_SUM || d->lsf, "NULL lsf table");
        SURFH_REQUIRE(d->spectral_mode == SURFH_SPECTRAL_LSF ||
                          (d->det_start >= 0 && d->det_start + d->n_wave <= d->n_det),
                      "beta-sum band: local wavelengths outside the band's output rows");
        auto b = std::make_unique<BandT<T>>();
        b->P = d->n_pointing; b->S = d->n_slit; b->na = d->na; b->nb = d->nb; b->srf = d->srf;
        b->A = d->local_a; b->B = d->local_b; b->l0 = d->wave_start; b->nl = d->n_wave; b->nd = d->n_det;
        b->mode = d->spectral_mode; b->det_start = d->det_start;
        b->Nn = b->P * b->S * b->na;
        b->ncol = b->Nn * b->nb;
        b->KB = b->nl * b->nb;
        b->KBp = (b->KB + 1) / 2 * 2;
        if (b->mode == SURFH_SPECTRAL_LSF) { b->g_col = b->KBp; b->g_l = b->nb; b->g_psa = 1; }
        else { b->g_col = b->nb; b->g_l = b->ncol; b->g_psa = 0; }
        b->out_offset = d->out_offset;
        b->out_size = (int64_t)b->Nn * b->nd;
        SURFH_REQUIRE(d->out_offset >= 0, "negative out_offset");
        SURFH_REQUIRE(b->mode == SURFH_SPECTRAL_BETA_SUM ||
                          ((int64_t)b->nd * b->KBp < (1ll << 31) && (int64_t)b->Nn * b->KBp < (1ll << 31) &&
                           b->out_size < (1ll << 31)), "band too large for 32-bit operand offsets");
        const int AB = b->A * b->B;
        for (int s = 0; s < b->S; ++s) {
            SURFH_REQUIRE(d->slit_a0[s] >= 0 && d->slit_a0[s] < b->A, "slit_a0 out of the local grid");
            SURFH_REQUIRE(d->slit_b0[s] >= 0 && d->slit_b0[s] + b->nb <= b->B, "slit columns out of the local grid");
        }
        for (int64_t q = 0; q < (int64_t)b->P * AB; ++q) {
            const int32_t off = d->grid_base[q];
            // the reference raises ValueError when a pointing's FoV leaves the cube
            SURFH_REQUIRE(off >= 0 && off / Nb < Na - 1 && off % Nb < Nb - 1,
                          "One of the requested xi is out of bounds (a pointing's field of view leaves the cube)");
        }
        check_csr(d->adj_exact, (int64_t)plane, b->ncol, "adj_exact");
        check_csr(d->adj_reference, (int64_t)plane, b->ncol, "adj_reference");
        // hull of the cube rows the band touches: the FFT passes skip every row pair outside it
        b->row_lo = Na - 1;
        b->row_hi = 0;
        for (int64_t q = 0; q < (int64_t)b->P * AB; ++q) {
            const int i = d->grid_base[q] / Nb;
            b->row_lo = std::min(b->row_lo, i);
            b->row_hi = std::max(b->row_hi, i + 1);
        }
        {   // union footprint of the four taps over every pointing: what one gather launch reads per plane
            std::vector<uint8_t> seen(plane, 0);
            for (int64_t q = 0; q < (int64_t)b->P * AB; ++q) {
                const int32_t off = d->grid_base[q];
                seen[off] = seen[off + 1] = seen[off + Nb] = seen[off + Nb + 1] = 1;
            }
            for (uint8_t v : seen) b->footprint += v;
        }
        for (const surfh_csr* c : {&d->adj_exact, &d->adj_reference})
            if (c->n_rows > 0) {
                b->row_lo = std::min(b->row_lo, c->row_pixel[0] / Nb);
                b->row_hi = std::max(b->row_hi, c->row_pixel[c->n_rows - 1] / Nb);
            }

        upload_converted<int32_t>(b->slit_a0, d->slit_a0, b->S);
        upload_converted<int32_t>(b->slit_b0, d->slit_b0, b->S);
        upload_converted<T>(b->slit_w, d->slit_w, (size_t)b->S * b->nb);
        if (b->mode == SURFH_SPECTRAL_LSF) {   // rows at pitch KBp (the pad element, if any, is zero)
            std::vector<double> w((size_t)b->nd * b->KBp, 0.0);
            for (int m = 0; m < b->nd; ++m)
                std::copy(d->lsf + (size_t)m * b->KB, d->lsf + (size_t)(m + 1) * b->KB, w.begin() + (size_t)m * b->KBp);
            upload_converted<T>(b->lsf, w.data(), w.size());
        }
        upload_converted<int32_t>(b->grid_base, d->grid_base, (size_t)b->P * AB);
        upload_converted<T>(b->grid_frac, d->grid_frac, (size_t)b->P * AB * 2);
        const surfh_csr* cs[2] = {&d->adj_exact, &d->adj_reference};
        for (int m = 0; m < 2; ++m) {
            b->csr_rows[m] = cs[m]->n_rows;
            b->csr_nnz[m] = cs[m]->nnz;
            if (cs[m]->n_rows == 0) continue;
            upload_converted<int32_t>(b->csr_pix[m], cs[m]->row_pixel, cs[m]->n_rows);
            // CSR -> sliced ELL (see kernels_slit.cuh), columns in the internal slit-space order
            const int n_rows = cs[m]->n_rows, n_slices = (n_rows + 31) / 32;
            std::vector<int64_t> slice_ptr((size_t)n_slices + 1, 0);
            for (int sl = 0; sl < n_slices; ++sl) {
                int64_t width = 0;
                for (int r = sl * 32; r < std::min(n_rows, sl * 32 + 32); ++r)
                    width = std::max<int64_t>(width, cs[m]->row_ptr[r + 1] - cs[m]->row_ptr[r]);
                slice_ptr[(size_t)sl + 1] = slice_ptr[(size_t)sl] + 32 * width;
            }
            std::vector<int32_t> col((size_t)slice_ptr.back(), 0);
            std::vector<double> val((size_t)slice_ptr.back(), 0.0);
            for (int r = 0; r < n_rows; ++r) {
                const int64_t at = slice_ptr[(size_t)(r / 32)] + r % 32;
                for (int64_t e = cs[m]->row_ptr[r], k = 0; e < cs[m]->row_ptr[r + 1]; ++e, ++k) {
                    col[(size_t)(at + 32 * k)] = b->g_offset(cs[m]->col[e]);
                    val[(size_t)(at + 32 * k)] = cs[m]->val[e];
                }
            }
            upload_converted<int64_t>(b->csr_ptr[m], slice_ptr.data(), slice_ptr.size());
            upload_converted<int32_t>(b->csr_col[m], col.data(), col.size());
            upload_converted<T>(b->csr_val[m], val.data(), val.size());
        }
        // GEMM offset tables
        if (b->mode == SURFH_SPECTRAL_LSF) {
            const int nmax = std::max(b->KB, b->nd);
            std::vector<int32_t> v(nmax);
            for (int i = 0; i < nmax; ++i) v[i] = i;
            upload_converted<int32_t>(b->t_ident, v.data(), nmax);
            v.assign(b->nd, 0);
            for (int i = 0; i < b->nd; ++i) v[i] = i * b->KBp;
            upload_converted<int32_t>(b->t_wrow, v.data(), b->nd);
            v.assign(b->KB, 0);
            for (int k = 0; k < b->KB; ++k) v[k] = (k / b->nb) * b->g_l + k % b->nb;   // = k in the K-fast layout
            upload_converted<int32_t>(b->t_gK, v.data(), b->KB);
            v.assign(b->Nn, 0);
            for (int n = 0; n < b->Nn; ++n) v[n] = b->column_of_sample(n) * b->g_col;  // n = (p*S + s)*na + a
            upload_converted<int32_t>(b->t_gN, v.data(), b->Nn);
            v.assign(b->nd, 0);
            for (int m = 0; m < b->nd; ++m) v[m] = m * b->na;
            upload_converted<int32_t>(b->t_yM, v.data(), b->nd);
            v.assign(b->Nn, 0);
            for (int n = 0; n < b->Nn; ++n) v[n] = (n / b->na) * (b->nd * b->na) + n % b->na;
            upload_converted<int32_t>(b->t_yN, v.data(), b->Nn);
        }
        // K-fast layout: Nn rows at pitch g_col = KBp; beta-sum layout: nl planes of ncol
        const size_t g_elems = b->mode == SURFH_SPECTRAL_LSF ? (size_t)b->Nn * b->KBp : (size_t)b->nl * b->ncol;
        b->G.alloc(g_elems * sizeof(T));
        SURFH_CUDA(cudaMemset(b->G.p, 0, b->G.bytes));   // the pad column is never written: keep it finite (zero)
        if (b->mode == SURFH_SPECTRAL_LSF) {
            // K-fast operands of the adjoint product: the transposed LSF and the re-laid detector block
            b->ndp = (b->nd + 1) / 2 * 2;
            std::vector<double> wt((size_t)b->KB * b->ndp, 0.0);
            for (int m = 0; m < b->nd; ++m)
                for (int k = 0; k < b->KB; ++k) wt[(size_t)k * b->ndp + m] = d->lsf[(size_t)m * b->KB + k];
            upload_converted<T>(b->lsf_t, wt.data(), wt.size());
            b->yk.alloc((size_t)b->Nn * b->ndp * sizeof(T));
            if (std::is_same<T, double>::value) {
                // tensor maps of the TMA-fed DMMA contraction (kernels_gemm_tma.cuh)
                b->map_w = tensor_map_2d_f64(b->lsf.p, b->KB, b->nd, (size_t)b->KBp, kTBM);
                b->map_g = tensor_map_2d_f64(b->G.p, b->KB, b->Nn, (size_t)b->g_col, kTBN);
                b->map_wt = tensor_map_2d_f64(b->lsf_t.p, b->nd, b->KB, (size_t)b->ndp, kTBM);
                b->map_yk = tensor_map_2d_f64(b->yk.p, b->nd, b->Nn, (size_t)b->ndp, kTBN);
                b->tma_ready = true;
            }
            if (gemm_mode == GEMM_OZAKI) prepare_ozaki(*b);
        }
        bands.push_back(std::move(b));
    }

    cufftHandle plan(int kind, int batch) {
        auto key = std::make_pair(kind, batch);
        auto it = plans.find(key);
        if (it != plans.end()) return it->second;
        throw Error(SURFH_ESTATE, "internal: FFT plan missing");
    }

    // kind 0: real planes (dist plane) -> complex planes (dist nfp); kind 1: the inverse
    void make_plan(int kind, int batch) {
        auto key = std::make_pair(kind, batch);
        if (plans.count(key) || batch <= 0) return;
        cufftHandle p;
        SURFH_FFT(cufftCreate(&p));
        SURFH_FFT(cufftSetAutoAllocation(p, 0));
        int n[2] = {Na, Nb};
        int rembed[2] = {Na, Nb};
        int cembed[2] = {Na, Nh};
        size_t ws = 0;
        if (kind == 0)
            SURFH_FFT(cufftMakePlanMany(p, 2, n, rembed, 1, (int)plane, cembed, 1, (int)nfp, FftTraits<T>::R2C, batch, &ws));
        else
            SURFH_FFT(cufftMakePlanMany(p, 2, n, cembed, 1, (int)nfp, rembed, 1, (int)plane, FftTraits<T>::C2R, batch, &ws));
        fft_work_bytes = std::max(fft_work_bytes, ws);
        plans[key] = p;
    }

    void finalize() override {
        SURFH_REQUIRE(!bands.empty(), "no band added");
        // union of wavelength windows
        std::vector<std::pair<int, int>> w;
        for (auto& b : bands) w.emplace_back(b->l0, b->l0 + b->nl);
        std::sort(w.begin(), w.end());
        ranges.clear();
        for (auto& r : w) {
            if (!ranges.empty() && r.first <= ranges.back().second)
                ranges.back().second = std::max(ranges.back().second, r.second);
            else
                ranges.push_back(r);
        }
        for (auto& r : ranges)
            for (int l = r.first; l < r.second; ++l)
                SURFH_REQUIRE(otf_set[l], "OTF plane " + std::to_string(l) + " needed by a band was never uploaded");
        int longest = 0;
        for (auto& r : ranges) longest = std::max(longest, r.second - r.first);
        if (chunk <= 0) {
            // default: bound the two chunk buffers to ~2 GiB
            const size_t per_l = nfp * sizeof(C) + plane * sizeof(T);
            chunk = (int)std::max<size_t>(1, std::min<size_t>(512, ((size_t)2 << 30) / per_l));
        }
        if (use_own_fft) chunk = std::min(chunk, FftK<T, 256>::MAX_PLANES);  // planes per pruned FFT launch
        chunk = std::min(chunk, longest);
        spec.alloc((size_t)chunk * nfp * sizeof(C));
        SURFH_CUDA(cudaMemset(spec.p, 0, spec.bytes));
        cplane = use_own_fft ? (plane + 1) / 2 * 2 : plane;
        cubebuf.alloc(((size_t)chunk * cplane + Nb) * sizeof(T));   // + one row of slack: the bulk copy of a plane's
                                                                      // last (unpaired) row reads a full pair
        if (K > 0) {
            xhat.alloc((size_t)K * nfp * sizeof(C));
            SURFH_CUDA(cudaMemset(xhat.p, 0, xhat.bytes));
        }
        if (use_own_fft) {
            ownfft.init(Na, Nb);
            zbuf.alloc((size_t)std::max(chunk, K) * ownfft.z_plane() * sizeof(C));
            // Optional (SURFH_STREAMS=2, off by default): two chunks in flight, chunk i on lane i % 2 (own stream, own
            // spectrum / cube / intermediate buffers), so that the HBM-bound template x OTF stream of one chunk could
            // share the SMs with the LSU-bound slit gather / scatter of its neighbour.  Measured on C4: 33.7 ms per
            // application against 33.1 ms on one stream -- the persistent FFT kernels need whole SMs (225 KB of shared
            // memory, every register), so the other lane's small CTAs only delay their start.  Kept for A/B runs.
            int n_chunks = 0;
            for (auto& r : ranges) n_chunks += ceil_div(r.second - r.first, chunk);
            const char* e_streams = std::getenv("SURFH_STREAMS");
            two_lanes = n_chunks > 1 && e_streams && std::atoi(e_streams) == 2;
            if (two_lanes) {
                spec2.alloc(spec.bytes);
                SURFH_CUDA(cudaMemset(spec2.p, 0, spec2.bytes));
                cubebuf2.alloc(cubebuf.bytes);
                zbuf2.alloc((size_t)chunk * ownfft.z_plane() * sizeof(C));
                SURFH_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
                SURFH_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
                SURFH_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
                SURFH_CUDA(cudaEventCreateWithFlags(&ev_acc[0], cudaEventDisableTiming));
                SURFH_CUDA(cudaEventCreateWithFlags(&ev_acc[1], cudaEventDisableTiming));
            }
            std::vector<int2> pr(Nl, make_int2(0, 0));
            plane_pair_cnt.assign(Nl, 0);
            for (int l = 0; l < Nl; ++l) {
                int lo = Na, hi = -1;
                for (auto& b : bands)
                    if (l >= b->l0 && l < b->l0 + b->nl) { lo = std::min(lo, b->row_lo); hi = std::max(hi, b->row_hi); }
                if (hi < lo) continue;
                pr[l] = make_int2(lo / 2, hi / 2 - lo / 2 + 1);
                plane_pair_cnt[l] = pr[l].y;
            }
            plane_pairs.alloc((size_t)Nl * sizeof(int2));
            SURFH_CUDA(cudaMemcpy(plane_pairs.p, pr.data(), (size_t)Nl * sizeof(int2), cudaMemcpyHostToDevice));
        } else {
            if (K > 0) {
                make_plan(0, K);
                make_plan(1, K);
            }
            for (auto& r : ranges) {
                const int len = r.second - r.first;
                if (len >= chunk) { make_plan(0, chunk); make_plan(1, chunk); }
                if (len % chunk) { make_plan(0, len % chunk); make_plan(1, len % chunk); }
            }
            fft_work.alloc(fft_work_bytes);
            for (auto& kv : plans) SURFH_FFT(cufftSetWorkArea(kv.second, fft_work.p));
        }
        // opt in to > 48 KB dynamic shared memory for the GEMM kernels
        using G = GemmCfg<T>;
        const int smem = (int)otgemm_smem_bytes<T, G::BM, G::BN, G::BK>();
        SURFH_CUDA(cudaFuncSetAttribute(otgemm_kernel<T, G::BM, G::BN, G::BK, G::TM, G::TN, true, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        SURFH_CUDA(cudaFuncSetAttribute(otgemm_kernel<T, G::BM, G::BN, G::BK, G::TM, G::TN, false, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        set_ozaki_attributes();
        if (std::is_same<T, float>::value) {
            SURFH_CUDA(cudaFuncSetAttribute(sgemm_tf32x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sgemm_smem_bytes<true, true>()));
            SURFH_CUDA(cudaFuncSetAttribute(sgemm_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sgemm_smem_bytes<false, false>()));
        }
        if (std::is_same<T, double>::value) {
            SURFH_CUDA(cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmemBytes));
            SURFH_CUDA(cudaFuncSetAttribute(dgemm_mma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)dgemm_smem_bytes<true, true>()));
            SURFH_CUDA(cudaFuncSetAttribute(dgemm_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)dgemm_smem_bytes<false, false>()));
        }
        SURFH_CUDA(cudaDeviceSynchronize());
        finalized = true;
    }

    int64_t input_size() const override { return (int64_t)(K > 0 ? K : Nl) * (int64_t)plane; }
    int64_t output_size() const override {
        int64_t m = 0;
        for (auto& b : bands) m = std::max(m, b->out_offset + b->out_size);
        return m;
    }
    int64_t workspace_bytes() const override {
        int64_t t = otf.bytes + tpl.bytes + tpl_raw.bytes + xhat.bytes + spec.bytes + cubebuf.bytes + fft_work.bytes + zbuf.bytes + spec2.bytes + cubebuf2.bytes + zbuf2.bytes +
                    y_internal.bytes + x_stage.bytes + y_stage.bytes + dbl_stage.bytes;
        for (auto& b : bands)
            t += b->lsf.bytes + b->lsf_t.bytes + b->yk.bytes + b->G.bytes + b->grid_base.bytes + b->grid_frac.bytes +
                 b->oz_w.bytes + b->oz_wt.bytes + b->oz_g.bytes + b->oz_yk.bytes + b->oz_mask_w.bytes + b->oz_mask_wt.bytes +
                 b->csr_col[0].bytes + b->csr_val[0].bytes + b->csr_col[1].bytes + b->csr_val[1].bytes;
        return t + precond_inv.bytes;
    }

    void contraction_info(int32_t* mode, int32_t* digits, double* executed_fraction) const override {
        if (executed_fraction) *executed_fraction = 1.0;
        bool any_lsf = false;
        for (auto& b : bands) any_lsf = any_lsf || b->mode == SURFH_SPECTRAL_LSF;
        if (!any_lsf) {   // beta-sum bands only (MRSBlurred): there is no contraction
            if (mode) *mode = -1;
            if (digits) *digits = 0;
            return;
        }
        const bool oz = ozaki_usable();
        if (mode) *mode = oz ? GEMM_OZAKI : (gemm_mode == GEMM_OZAKI ? (std::is_same<T, double>::value ? GEMM_TMA : GEMM_LEGACY) : gemm_mode);
        if (digits) *digits = oz ? ozaki_digits : 0;
        if (executed_fraction && oz && oz_skip_zero_tiles && oz_products_dense > 0)
            *executed_fraction = oz_products_executed / oz_products_dense;
    }

    // ---- launch helpers ---------------------------------------------------------------------
    DevBuf& spec_of(int lane) { return lane ? spec2 : spec; }
    DevBuf& cube_of(int lane) { return lane ? cubebuf2 : cubebuf; }
    DevBuf& zbuf_of(int lane) { return lane ? zbuf2 : zbuf; }
    template <int KK> void launch_lmm_fwd(int c0, int nl, cudaStream_t st, int lane) {
        dim3 grid(ceil_div(nfp, 256), ceil_div(nl, kLmmLsub));
        lmm_otf_fwd_kernel<T, KK><<<grid, 256, 0, st>>>(xhat.as<C>(), otf.as<C>() + (size_t)c0 * nfp, tpl.as<T>(), Nl,
                                                        c0, nl, nfp, spec_of(lane).template as<C>());
    }
    template <int KK> void launch_lmm_adj(int c0, int nl, bool accumulate, cudaStream_t st, int lane) {
        dim3 grid(ceil_div(nfp, 32)), block(32, kAdjLanes);
        lmm_otf_adj_kernel<T, KK><<<grid, block, 0, st>>>(spec_of(lane).template as<C>(), otf.as<C>() + (size_t)c0 * nfp,
                                                          tpl.as<T>(), Nl, c0, nl, nfp, xhat.as<C>(), accumulate ? 1 : 0);
    }
    template <int KK> void launch_maps_to_cube(const T* maps, float* cube, cudaStream_t st) {
        dim3 grid(ceil_div(plane, 256), ceil_div(Nl, 32));
        maps_to_cube_kernel<T, KK><<<grid, 256, 0, st>>>(maps, tpl_raw.as<T>(), Nl, Nl, plane, cube);
    }
#define SURFH_DISPATCH_K(fn, ...)                                                          \
    switch (K) {                                                                           \
        case 1: fn<1>(__VA_ARGS__); break;                                                 \
        case 2: fn<2>(__VA_ARGS__); break;                                                 \
        case 3: fn<3>(__VA_ARGS__); break;                                                 \
        case 4: fn<4>(__VA_ARGS__); break;                                                 \
        case 5: fn<5>(__VA_ARGS__); break;                                                 \
        case 6: fn<6>(__VA_ARGS__); break;                                                 \
        case 7: fn<7>(__VA_ARGS__); break;                                                 \
        case 8: fn<8>(__VA_ARGS__); break;                                                 \
        default: throw Error(SURFH_EINVAL, "unsupported template count");                  \
    }

    int fft_launches() const { return use_own_fft ? 2 : 1; }
    // Algorithmic bytes of one batched 2-D transform: the half spectrum plus the real rows that matter
    // (all of them, or the pruned row pairs of planes [c0, c0 + batch) of the working cube).
    double fft_bytes(int batch, int prune_c0) const {
        double rows = (double)batch * Na;
        if (use_own_fft && prune_rows && prune_c0 >= 0) {
            rows = 0;
            for (int l = prune_c0; l < prune_c0 + batch; ++l) rows += 2.0 * plane_pair_cnt[l];
        }
        return (double)batch * nf * sizeof(C) + rows * Nb * sizeof(T);
    }
    // Flops of the chirp-z evaluation (hand-written backend): per 1-D transform two M-point FFTs at
    // 5 M log2 M, the filter product (6 M) and the two chirp products (6 N each).
    double fft_flops(int batch, int prune_c0) const {
        if (!use_own_fft) return 0.0;
        auto one = [](int m, int n) { return 10.0 * m * std::log2((double)m) + 6.0 * m + 12.0 * n; };
        double pairs = (double)batch * ((Na + 1) / 2);
        if (prune_rows && prune_c0 >= 0) {
            pairs = 0;
            for (int l = prune_c0; l < prune_c0 + batch; ++l) pairs += plane_pair_cnt[l];
        }
        return pairs * one(ownfft.axis_b->m, Nb) + (double)batch * Nh * one(ownfft.axis_a.m, Na);
    }
    // prune_c0 >= 0: the real side is the working cube of planes [prune_c0, prune_c0 + batch), of which only
    // the rows some band touches matter (C2R: produced; R2C: non-zero)
    // real_stride: elements between the real planes (0 = `plane`: caller-owned contiguous maps / cubes)
    void fft_exec(int kind, int batch, void* in, void* out, cudaStream_t st, int prune_c0 = -1, size_t real_stride = 0,
                  int lane = 0) {
        if (real_stride == 0) real_stride = plane;
        if (use_own_fft) {
            const int2* pr = nullptr;
            long long n_pairs = 0;
            if (prune_c0 >= 0 && prune_rows) {
                pr = plane_pairs.as<int2>() + prune_c0;
                for (int l = prune_c0; l < prune_c0 + batch; ++l) n_pairs += plane_pair_cnt[l];
            }
            if (kind == 0)
                ownfft.r2c(reinterpret_cast<const T*>(in), real_stride, reinterpret_cast<C*>(out), nfp,
                           zbuf_of(lane).template as<C>(), batch, st, true, pr, n_pairs,
                           /*row_slack=*/in == cubebuf.p || in == cubebuf2.p);
            else
                ownfft.c2r(reinterpret_cast<const C*>(in), nfp, reinterpret_cast<T*>(out), real_stride,
                           zbuf_of(lane).template as<C>(), batch, st, true, pr, n_pairs);
            return;
        }
        cufftHandle p = plan(kind, batch);
        SURFH_FFT(cufftSetStream(p, st));
        if (kind == 0)
            SURFH_FFT(FftTraits<T>::fwd(p, reinterpret_cast<T*>(in), reinterpret_cast<C*>(out)));
        else
            SURFH_FFT(FftTraits<T>::inv(p, reinterpret_cast<C*>(in), reinterpret_cast<T*>(out)));
    }

    GemmArgs<T> gemm_args(BandT<T>& b, T* y, bool adjoint) {
        GemmArgs<T> g;
        if (!adjoint) {  // y = W . G
            g.M = b.nd; g.N = b.Nn; g.K = b.KB;
            g.A = b.lsf.template as<T>(); g.aM = b.t_wrow.template as<int32_t>(); g.aK = b.t_ident.template as<int32_t>();
            g.B = b.G.template as<T>(); g.bK = b.t_gK.template as<int32_t>(); g.bN = b.t_gN.template as<int32_t>();
            g.C = y + b.out_offset; g.cM = b.t_yM.template as<int32_t>(); g.cN = b.t_yN.template as<int32_t>();
        } else {  // Gt = W^T . y
            g.M = b.KB; g.N = b.Nn; g.K = b.nd;
            g.A = b.lsf.template as<T>(); g.aM = b.t_ident.template as<int32_t>(); g.aK = b.t_wrow.template as<int32_t>();
            g.B = y + b.out_offset; g.bK = b.t_yM.template as<int32_t>(); g.bN = b.t_yN.template as<int32_t>();
            g.C = b.G.template as<T>(); g.cM = b.t_gK.template as<int32_t>(); g.cN = b.t_gN.template as<int32_t>();
        }
        return g;
    }

    // SIMT path (fp32): one launch per band
    void gemm_simt(BandT<T>& b, T* y, bool adjoint, cudaStream_t st) {
        using G = GemmCfg<T>;
        GemmArgs<T> g = gemm_args(b, y, adjoint);
        dim3 grid(ceil_div(g.N, G::BN), ceil_div(g.M, G::BM));
        const double bytes = sizeof(T) * ((double)b.nd * b.KB + (double)b.nl * b.ncol + (double)b.out_size);
        Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, 2.0 * g.M * g.N * g.K, 1, true);
        const int threads = (G::BM / G::TM) * (G::BN / G::TN);
        const size_t smem = otgemm_smem_bytes<T, G::BM, G::BN, G::BK>();
        if (!adjoint)
            otgemm_kernel<T, G::BM, G::BN, G::BK, G::TM, G::TN, true, true><<<grid, threads, smem, st>>>(g);
        else
            otgemm_kernel<T, G::BM, G::BN, G::BK, G::TM, G::TN, false, false><<<grid, threads, smem, st>>>(g);
        SURFH_CUDA(cudaGetLastError());
    }

    // FP64 tensor path: all bands in one grouped launch (per group of kMaxGemmGroup bands)
    void gemm_grouped_f64(double* y, bool adjoint, cudaStream_t st);
    // FP32 tensor path (3xTF32 split), same grouping; SURFH_F32_GEMM=simt selects the FFMA kernel
    void gemm_grouped_f32(float* y, bool adjoint, cudaStream_t st);
    int gemm_mode = gemm_mode_env(std::is_same<T, double>::value);   // read once per handle, before the bands are added
    int ozaki_digits = ozaki_digits_env(std::is_same<T, double>::value);
    void gemm_grouped_f64_tma(double* y, bool adjoint, cudaStream_t st);
    // ---- the sliced contraction on tcgen05 (kernels_ozaki.cuh) -------------------------------------------
    // cuts the rows of up to kMaxGemmGroup matrices into int8 digit planes + row scales: one launch, one CTA per row
    struct SliceJobs {
        OzSliceBatch batch;
        SliceJobs() { batch.count = 0; batch.row_start[0] = 0; }
        void add(const T* x, int rows, int k, size_t ld, DevBuf& digits, int pitch, DevBuf& scale) {
            OzSliceJob& j = batch.j[batch.count];
            j.x = x; j.digits = digits.as<int8_t>(); j.scale = scale.as<double>(); j.ld = ld; j.rows = rows; j.K = k; j.Kp = pitch;
            batch.row_start[batch.count + 1] = batch.row_start[batch.count] + rows;
            batch.count++;
        }
        bool full() const { return batch.count == kMaxGemmGroup; }
        template <int S> void launch(cudaStream_t st) {
            if (batch.count == 0) return;
            ozaki_slice_rows_kernel<S, T><<<batch.row_start[batch.count], 256, 0, st>>>(batch);
            batch.count = 0;
        }
    };
    template <typename F> void with_digits(F&& f) {   // f(std::integral_constant<int, S>) for this handle's digit count
        if (std::is_same<T, float>::value) return f(std::integral_constant<int, 4>());
        if (ozaki_digits == 6) return f(std::integral_constant<int, 6>());
        if (ozaki_digits == 7) return f(std::integral_constant<int, 7>());
        return f(std::integral_constant<int, 8>());
    }
    void prepare_ozaki(BandT<T>& b) {
        const int S = ozaki_digits;
        // every level sum_{p+q=t} sum_k dA dB must fit an int32: (t + 1) K 64^2 < 2^31
        if ((int64_t)S * std::max(b.KB, b.nd) * 4096 >= ((int64_t)1 << 31)) return;   // this band keeps the older kernels
        b.oz_kq = (b.KB + 15) / 16 * 16;
        b.oz_ndq = (b.nd + 15) / 16 * 16;
        b.oz_w.alloc((size_t)S * b.nd * b.oz_kq);
        b.oz_wt.alloc((size_t)S * b.KB * b.oz_ndq);
        b.oz_g.alloc((size_t)S * b.Nn * b.oz_kq);
        b.oz_yk.alloc((size_t)S * b.Nn * b.oz_ndq);
        b.oz_sw.alloc((size_t)b.nd * sizeof(double));
        b.oz_swt.alloc((size_t)b.KB * sizeof(double));
        b.oz_sg.alloc((size_t)b.Nn * sizeof(double));
        b.oz_syk.alloc((size_t)b.Nn * sizeof(double));
        // the LSF and its transpose are constant: cut them into digits once
        with_digits([&](auto s_) {
            SliceJobs jobs;
            jobs.add(b.lsf.template as<T>(), b.nd, b.KB, (size_t)b.KBp, b.oz_w, b.oz_kq, b.oz_sw);
            jobs.add(b.lsf_t.template as<T>(), b.KB, b.nd, (size_t)b.ndp, b.oz_wt, b.oz_ndq, b.oz_swt);
            jobs.template launch<decltype(s_)::value>(0);
        });
        // the response decays away from its peak: its leading digits vanish outside a band, those tiles are skipped
        b.oz_mask_w.alloc((size_t)ceil_div(b.nd, kOzBM) * ceil_div(b.KB, kOzBK));
        b.oz_mask_wt.alloc((size_t)ceil_div(b.KB, kOzBM) * ceil_div(b.nd, kOzBK));
        with_digits([&](auto s_) {
            constexpr int SS = decltype(s_)::value;
            ozaki_tile_mask_kernel<SS><<<dim3(ceil_div(b.KB, kOzBK), ceil_div(b.nd, kOzBM)), 128>>>(
                b.oz_w.template as<int8_t>(), b.nd, b.KB, b.oz_kq, b.oz_mask_w.template as<uint8_t>());
            ozaki_tile_mask_kernel<SS><<<dim3(ceil_div(b.nd, kOzBK), ceil_div(b.KB, kOzBM)), 128>>>(
                b.oz_wt.template as<int8_t>(), b.KB, b.nd, b.oz_ndq, b.oz_mask_wt.template as<uint8_t>());
        });
        SURFH_CUDA(cudaGetLastError());
        SURFH_CUDA(cudaDeviceSynchronize());
        {   // bookkeeping for surfh_contraction_info: digit products executed / of the dense scheme, both directions
            const double tiles_n = ceil_div(ceil_div(b.Nn, kOzBN), kOzCluster) * kOzCluster;
            for (DevBuf* mk : {&b.oz_mask_w, &b.oz_mask_wt}) {
                std::vector<uint8_t> hm(mk->bytes);
                SURFH_CUDA(cudaMemcpy(hm.data(), mk->p, mk->bytes, cudaMemcpyDeviceToHost));
                for (uint8_t bits : hm) {
                    for (int p = 0; p < S; ++p)
                        if (bits & (1u << p)) oz_products_executed += tiles_n * (S - p);
                    oz_products_dense += tiles_n * (S * (S + 1) / 2);
                }
            }
        }
        b.ozmap_w = tensor_map_digits(b.oz_w.p, b.KB, b.nd, b.oz_kq, S, kOzBM / kOzCluster);
        b.ozmap_g = tensor_map_digits(b.oz_g.p, b.KB, b.Nn, b.oz_kq, S, kOzBN);
        b.ozmap_wt = tensor_map_digits(b.oz_wt.p, b.nd, b.KB, b.oz_ndq, S, kOzBM / kOzCluster);
        b.ozmap_yk = tensor_map_digits(b.oz_yk.p, b.nd, b.Nn, b.oz_ndq, S, kOzBN);
        b.oz_ready = true;
    }
    double oz_products_executed = 0, oz_products_dense = 0;   // digit-tile products: with zero tiles skipped / dense
    int oz_n_major_fwd = std::getenv("SURFH_OZAKI_MMAJOR") ? 0 : 1;   // A/B switch of the forward product's tile order
    int oz_resident_ctas = 0;   // one wave of co-resident clusters: the persistent contraction's grid
    bool oz_skip_zero_tiles = !(std::getenv("SURFH_OZAKI_DENSE") && std::atoi(std::getenv("SURFH_OZAKI_DENSE")) == 1);   // A/B switch
    void set_ozaki_attributes() {
        with_digits([&](auto s_) {
            constexpr int SS = decltype(s_)::value;
            SURFH_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel<SS, kOzCluster, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)ozaki_smem_bytes(SS)));
            oz_resident_ctas = ozaki_max_resident_ctas(ozaki_gemm_kernel<SS, kOzCluster, T>, kOzCluster, ozaki_smem_bytes(SS));
        });
    }
    bool ozaki_usable() const {
        if (gemm_mode != GEMM_OZAKI) return false;
        for (auto& b : bands)
            if (b->mode == SURFH_SPECTRAL_LSF && !b->oz_ready) return false;
        return true;
    }
    template <int S> void ozaki_run(T* y, bool adjoint, cudaStream_t st) {
        std::vector<size_t> lsf_bands;
        for (size_t i = 0; i < bands.size(); ++i)
            if (bands[i]->mode == SURFH_SPECTRAL_LSF) lsf_bands.push_back(i);
        // longest contractions first: the block scheduler hands tiles out in blockIdx order
        std::stable_sort(lsf_bands.begin(), lsf_bands.end(), [&](size_t x, size_t y2) {
            const int kx = adjoint ? bands[x]->nd : bands[x]->KB, ky = adjoint ? bands[y2]->nd : bands[y2]->KB;
            return kx > ky;
        });
        {
            // the per-call operand -> int8 digit planes + row scales (forward: the slit-space vector G; adjoint: the
            // detector block, first re-laid K-fast per detector column)
            double bytes = 0;
            for (size_t j : lsf_bands) {
                const BandT<T>& b = *bands[j];
                bytes += adjoint ? (2.0 * sizeof(T) + S) * (double)b.out_size : (sizeof(T) + S) * (double)b.Nn * b.KB;
            }
            Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, 0,
                     (adjoint ? (int)lsf_bands.size() : 0) + ceil_div(lsf_bands.size(), kMaxGemmGroup), true);
            SliceJobs jobs;
            for (size_t j : lsf_bands) {
                BandT<T>& b = *bands[j];
                if (adjoint) {
                    const size_t n = (size_t)b.Nn * b.nd;
                    detector_to_kfast_kernel<T><<<ceil_div(n, 256), 256, 0, st>>>(y + b.out_offset, b.na, b.nd, b.Nn, b.ndp,
                                                                                   b.yk.template as<T>());
                    jobs.add(b.yk.template as<T>(), b.Nn, b.nd, (size_t)b.ndp, b.oz_yk, b.oz_ndq, b.oz_syk);
                } else {
                    jobs.add(b.G.template as<T>(), b.Nn, b.KB, (size_t)b.g_col, b.oz_g, b.oz_kq, b.oz_sg);
                }
                if (jobs.full()) jobs.template launch<S>(st);
            }
            jobs.template launch<S>(st);
            SURFH_CUDA(cudaGetLastError());
        }
        for (size_t first = 0; first < lsf_bands.size(); first += kMaxGemmGroup) {
            OzakiBatch batch;
            batch.count = 0;
            batch.tile_start[0] = 0;
            batch.dump = nullptr;
            double bytes = 0, flops = 0;
            for (size_t j = first; j < std::min(lsf_bands.size(), first + (size_t)kMaxGemmGroup); ++j) {
                BandT<T>& b = *bands[lsf_bands[j]];
                OzakiProblem& g = batch.p[batch.count];
                if (!adjoint) {   // y = W . G
                    g.a = b.ozmap_w; g.b = b.ozmap_g; g.M = b.nd; g.N = b.Nn; g.K = b.KB;
                    g.n_major = oz_n_major_fwd;   // the LSF digits are the smaller operand: they are the ones re-read
                    g.sa = b.oz_sw.template as<double>(); g.sb = b.oz_sg.template as<double>();
                    g.amask = oz_skip_zero_tiles ? b.oz_mask_w.template as<uint8_t>() : nullptr;
                    g.C = y + b.out_offset; g.cM = b.t_yM.template as<int32_t>(); g.cN = b.t_yN.template as<int32_t>();
                } else {          // Gt = Wt . Yk
                    g.a = b.ozmap_wt; g.b = b.ozmap_yk; g.M = b.KB; g.N = b.Nn; g.K = b.nd;
                    g.n_major = 0;                // the detector block's digits are small: re-read per row tile
                    g.sa = b.oz_swt.template as<double>(); g.sb = b.oz_syk.template as<double>();
                    g.amask = oz_skip_zero_tiles ? b.oz_mask_wt.template as<uint8_t>() : nullptr;
                    g.C = b.G.p; g.cM = b.t_ident.template as<int32_t>(); g.cN = b.t_gN.template as<int32_t>();
                }
                const int tiles_n = ceil_div(ceil_div(g.N, kOzBN), kOzCluster) * kOzCluster;
                batch.tile_start[batch.count + 1] = batch.tile_start[batch.count] + ceil_div(g.M, kOzBM) * tiles_n;
                batch.count++;
                bytes += (double)S * ((double)b.nd * b.KB + (double)b.Nn * (adjoint ? b.nd : b.KB)) + sizeof(T) * (double)g.M * g.N;
                flops += 2.0 * g.M * g.N * g.K;
            }
            Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, flops, 1, true);
            cudaLaunchConfig_t cfg = {};
            const int n_tiles = batch.tile_start[batch.count];
            cfg.gridDim = dim3((unsigned)(oz_resident_ctas > 0 ? std::min(n_tiles, oz_resident_ctas) : n_tiles));
            cfg.blockDim = dim3(kOzThreads);
            cfg.dynamicSmemBytes = ozaki_smem_bytes(S);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = kOzCluster;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            SURFH_CUDA(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<S, kOzCluster, T>, batch));
        }
    }

    void beta_sum(BandT<T>& b, T* y, bool adjoint, cudaStream_t st) {
        const size_t n = (size_t)b.nl * b.Nn * (adjoint ? b.nb : 1);
        const double bytes = sizeof(T) * ((double)b.nl * b.ncol + (double)b.nl * b.Nn);
        Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, (double)b.nl * b.ncol, 1, true);
        if (!adjoint)
            beta_sum_fwd_kernel<T><<<ceil_div(n, 256), 256, 0, st>>>(b.G.template as<T>(), b.nl, b.Nn, b.nb, b.S, b.na,
                                                                   b.det_start, y + b.out_offset);
        else
            beta_sum_adj_kernel<T><<<ceil_div(n, 256), 256, 0, st>>>(y + b.out_offset, b.nl, b.Nn, b.nb, b.S, b.na,
                                                                   b.det_start, b.G.template as<T>());
        SURFH_CUDA(cudaGetLastError());
    }

    void gemm_all(T* y, bool adjoint, cudaStream_t st) {
        bool any_lsf = false;
        for (auto& bp : bands) {
            if (bp->mode == SURFH_SPECTRAL_BETA_SUM) beta_sum(*bp, y, adjoint, st);
            else any_lsf = true;
        }
        if (!any_lsf) return;
        if (ozaki_usable()) {
            with_digits([&](auto s_) { ozaki_run<decltype(s_)::value>(y, adjoint, st); });
        } else if (std::is_same<T, double>::value) {
            gemm_grouped_f64(reinterpret_cast<double*>(y), adjoint, st);
        } else if (gemm_mode != GEMM_SIMT) {
            gemm_grouped_f32(reinterpret_cast<float*>(y), adjoint, st);
        } else {
            for (auto& bp : bands)
                if (bp->mode == SURFH_SPECTRAL_LSF) gemm_simt(*bp, y, adjoint, st);
        }
    }

#ifndef SURFH_GATHER_LB
#define SURFH_GATHER_LB 4
#endif
#ifndef SURFH_SCATTER_LB
#define SURFH_SCATTER_LB 4
#endif
    static constexpr int kLB = SURFH_GATHER_LB;    // wavelengths per thread in the gather ...
    static constexpr int kLBs = SURFH_SCATTER_LB;  // ... and in the scatter

    void gather_chunk(int c0, int c1, cudaStream_t st, int lane = 0) {
        for (auto& bp : bands) {
            BandT<T>& b = *bp;
            const int lo = std::max(c0, b.l0), hi = std::min(c1, b.l0 + b.nl);
            if (lo >= hi) continue;
            const int nl = hi - lo;
            // one launch serves every pointing of the band and reads the union of their footprints once
            const double bytes = sizeof(T) * ((double)nl * b.footprint + (double)nl * b.ncol);
            Scope sc(this, ST_SLIT_GATHER, st, bytes, 8.0 * nl * b.ncol * b.srf, 1, true);
            const int pp = std::min(b.P, 4), chunks = 4 / pp;  // see the kernel: 4 warps = pp pointings x chunks
            dim3 grid(ceil_div(b.S * b.na * b.nb, 32 * chunks), ceil_div(nl, kLB));
            slit_gather_kernel<T, kLB><<<grid, 128, 0, st>>>(cube_of(lane).template as<T>() + (size_t)(lo - c0) * cplane, cplane, Nb, nl,
                                                             b.slit_tables(),
                                                             b.G.template as<T>() + (size_t)(lo - b.l0) * b.g_l);
            SURFH_CUDA(cudaGetLastError());
        }
    }

    void scatter_chunk(int c0, int c1, int mode, cudaStream_t st, int lane = 0) {
        if (use_own_fft && prune_rows) {
            double rows = 0;
            for (int l = c0; l < c1; ++l) rows += 2.0 * plane_pair_cnt[l];
            Scope sc(this, ST_MEMSET, st, rows * Nb * sizeof(T), 0, 1, true);
            dim3 grid(32, c1 - c0);
            zero_row_hull_kernel<T><<<grid, 256, 0, st>>>(cube_of(lane).template as<T>(), cplane, Na, Nb, plane_pairs.as<int2>() + c0);
            SURFH_CUDA(cudaGetLastError());
        } else {
            Scope sc(this, ST_MEMSET, st, (double)(c1 - c0) * plane * sizeof(T), 0, 1, false);
            SURFH_CUDA(cudaMemsetAsync(cube_of(lane).p, 0, (size_t)(c1 - c0) * cplane * sizeof(T), st));
        }
        for (auto& bp : bands) {
            BandT<T>& b = *bp;
            const int lo = std::max(c0, b.l0), hi = std::min(c1, b.l0 + b.nl);
            if (lo >= hi || b.csr_rows[mode] == 0) continue;
            const int nl = hi - lo;
            dim3 grid(ceil_div(b.csr_rows[mode], 128), ceil_div(nl, kLBs));
            const double bytes = sizeof(T) * ((double)nl * b.ncol + 2.0 * nl * b.csr_rows[mode]);
            Scope sc(this, ST_SLIT_SCATTER, st, bytes, 2.0 * nl * (double)b.csr_nnz[mode], 1, true);
            slit_scatter_kernel<T, kLBs><<<grid, 128, 0, st>>>(b.G.template as<T>() + (size_t)(lo - b.l0) * b.g_l, b.g_l,
                                                              nl, b.csr(mode),
                                                              cube_of(lane).template as<T>() + (size_t)(lo - c0) * cplane, cplane);
            SURFH_CUDA(cudaGetLastError());
        }
    }

    void require_ready() const {
        if (!finalized) throw Error(SURFH_ESTATE, "surfh_finalize has not been called");
    }

    // Chunk i of the pipeline runs on lane i % 2: the caller's stream or the handle's auxiliary stream, each with its
    // own spectrum / cube / FFT-intermediate buffers.  fork: the auxiliary stream waits for everything queued so far
    // on the caller's stream; join: the caller's stream waits for the auxiliary one.
    void lanes_fork(cudaStream_t st) {
        if (!two_lanes) return;
        SURFH_CUDA(cudaEventRecord(ev_fork, st));
        SURFH_CUDA(cudaStreamWaitEvent(aux_stream, ev_fork, 0));
    }
    void lanes_join(cudaStream_t st) {
        if (!two_lanes) return;
        SURFH_CUDA(cudaEventRecord(ev_join, aux_stream));
        SURFH_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    }

    void forward(const void* xv, void* yv, cudaStream_t st) override {
        require_ready();
        SURFH_REQUIRE(xv && yv, "NULL buffer");
        const T* x = reinterpret_cast<const T*>(xv);
        T* y = reinterpret_cast<T*>(yv);
        if (K > 0) {
            Scope sc(this, ST_RFFT_MAPS, st, fft_bytes(K, -1), fft_flops(K, -1), fft_launches(), use_own_fft);
            fft_exec(0, K, const_cast<T*>(x), xhat.p, st);
        }
        lanes_fork(st);
        int i = 0;
        for (auto& r : ranges) {
            for (int c0 = r.first; c0 < r.second; c0 += chunk, ++i) {
                const int c1 = std::min(r.second, c0 + chunk), nl = c1 - c0;
                const int lane = two_lanes ? (i & 1) : 0;
                cudaStream_t ls = lane ? aux_stream : st;
                if (K > 0) {
                    Scope sc(this, ST_LMM_OTF_FWD, ls, (double)nl * nf * sizeof(C) * 2 + (double)K * nf * sizeof(C),
                             (4.0 * K + 6.0) * nl * nf, 1, true);
                    SURFH_DISPATCH_K(launch_lmm_fwd, c0, nl, ls, lane);
                    SURFH_CUDA(cudaGetLastError());
                } else {
                    {
                        Scope sc(this, ST_RFFT_CUBE, ls, fft_bytes(nl, -1), fft_flops(nl, -1), fft_launches(), use_own_fft);
                        fft_exec(0, nl, const_cast<T*>(x) + (size_t)c0 * plane, spec_of(lane).p, ls, -1, 0, lane);
                    }
                    Scope sc(this, ST_LMM_OTF_FWD, ls, (double)nl * nf * sizeof(C) * 3, 6.0 * nl * nf, 1, true);
                    const size_t n = (size_t)nl * nfp;
                    otf_mul_kernel<T, false><<<ceil_div(n, 256), 256, 0, ls>>>(
                        spec_of(lane).template as<C>(), otf.as<C>() + (size_t)c0 * nfp, n, (T)(1.0 / ((double)Na * Nb)));
                    SURFH_CUDA(cudaGetLastError());
                }
                {
                    Scope sc(this, ST_IRFFT_CUBE, ls, fft_bytes(nl, c0), fft_flops(nl, c0), fft_launches(), use_own_fft);
                    fft_exec(1, nl, spec_of(lane).p, cube_of(lane).p, ls, c0, cplane, lane);
                }
                gather_chunk(c0, c1, ls, lane);
            }
        }
        lanes_join(st);
        gemm_all(y, false, st);
    }

    void adjoint(const void* yv, void* xv, int mode, cudaStream_t st) override {
        require_ready();
        SURFH_REQUIRE(xv && yv, "NULL buffer");
        SURFH_REQUIRE(mode == SURFH_ADJ_EXACT || mode == SURFH_ADJ_REFERENCE, "unknown adjoint mode");
        const T* y = reinterpret_cast<const T*>(yv);
        T* x = reinterpret_cast<T*>(xv);
        gemm_all(const_cast<T*>(y), true, st);
        if (K == 0) {
            Scope sc(this, ST_MEMSET, st, (double)Nl * plane * sizeof(T), 0, 1, false);
            SURFH_CUDA(cudaMemsetAsync(x, 0, (size_t)Nl * plane * sizeof(T), st));
        }
        lanes_fork(st);
        bool first = true;
        int i = 0;
        for (auto& r : ranges) {
            for (int c0 = r.first; c0 < r.second; c0 += chunk, ++i) {
                const int c1 = std::min(r.second, c0 + chunk), nl = c1 - c0;
                const int lane = two_lanes ? (i & 1) : 0;
                cudaStream_t ls = lane ? aux_stream : st;
                scatter_chunk(c0, c1, mode, ls, lane);
                {
                    Scope sc(this, ST_RFFT_CUBE, ls, fft_bytes(nl, c0), fft_flops(nl, c0), fft_launches(), use_own_fft);
                    fft_exec(0, nl, cube_of(lane).p, spec_of(lane).p, ls, c0, cplane, lane);
                }
                if (K > 0) {
                    // the reduction over wavelengths accumulates chunk after chunk into the K map spectra: chunk i's
                    // stream waits for chunk i - 1's accumulation (fixed summation order: bitwise reproducible)
                    if (two_lanes && !first) SURFH_CUDA(cudaStreamWaitEvent(ls, ev_acc[lane ^ 1], 0));
                    {
                        Scope sc(this, ST_LMM_OTF_ADJ, ls, (double)nl * nf * sizeof(C) * 2 + (double)K * nf * sizeof(C),
                                 (4.0 * K + 6.0) * nl * nf, 1, true);
                        SURFH_DISPATCH_K(launch_lmm_adj, c0, nl, !first, ls, lane);
                        SURFH_CUDA(cudaGetLastError());
                    }
                    if (two_lanes) SURFH_CUDA(cudaEventRecord(ev_acc[lane], ls));
                } else {
                    {
                        Scope sc(this, ST_LMM_OTF_ADJ, ls, (double)nl * nf * sizeof(C) * 3, 6.0 * nl * nf, 1, true);
                        const size_t n = (size_t)nl * nfp;
                        otf_mul_kernel<T, true><<<ceil_div(n, 256), 256, 0, ls>>>(
                            spec_of(lane).template as<C>(), otf.as<C>() + (size_t)c0 * nfp, n, (T)(1.0 / ((double)Na * Nb)));
                        SURFH_CUDA(cudaGetLastError());
                    }
                    Scope sc(this, ST_IRFFT_CUBE, ls, fft_bytes(nl, -1), fft_flops(nl, -1), fft_launches(), use_own_fft);
                    fft_exec(1, nl, spec_of(lane).p, x + (size_t)c0 * plane, ls, -1, 0, lane);
                }
                first = false;
            }
        }
        lanes_join(st);
        if (K > 0) {
            Scope sc(this, ST_IRFFT_MAPS, st, fft_bytes(K, -1), fft_flops(K, -1), fft_launches(), use_own_fft);
            fft_exec(1, K, xhat.p, x, st);
        }
    }

    void fwadj(const void* x, void* out, int mode, void* yscratch, cudaStream_t st) override {
        require_ready();
        if (!yscratch) {
            y_internal.ensure((size_t)output_size() * sizeof(T));
            yscratch = y_internal.p;
        }
        forward(x, yscratch, st);
        adjoint(yscratch, out, mode, st);
    }

    void maps_to_cube(const void* maps, float* cube, cudaStream_t st) override {
        SURFH_REQUIRE(K > 0, "maps_to_cube needs templates");
        SURFH_REQUIRE(maps && cube, "NULL buffer");
        Scope sc(this, ST_LMM_OTF_FWD, st, 0, 0, 1, true);
        SURFH_DISPATCH_K(launch_maps_to_cube, reinterpret_cast<const T*>(maps), cube, st);
        SURFH_CUDA(cudaGetLastError());
    }

    // ---- host-buffer entry points -------------------------------------------------------------
    void to_device(const double* host, DevBuf& dev, size_t n, cudaStream_t st) {
        dev.ensure(n * sizeof(T));
        if (std::is_same<T, double>::value) {
            SURFH_CUDA(cudaMemcpyAsync(dev.p, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            dbl_stage.ensure(n * sizeof(double));
            SURFH_CUDA(cudaMemcpyAsync(dbl_stage.p, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
            convert_kernel<double, T><<<ceil_div(n, 256), 256, 0, st>>>(dbl_stage.as<double>(), dev.as<T>(), n);
            SURFH_CUDA(cudaGetLastError());
            own_launches++; launches++;
        }
    }
    void to_host(DevBuf& dev, double* host, size_t n, cudaStream_t st) {
        if (std::is_same<T, double>::value) {
            SURFH_CUDA(cudaMemcpyAsync(host, dev.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        } else {
            dbl_stage.ensure(n * sizeof(double));
            convert_kernel<T, double><<<ceil_div(n, 256), 256, 0, st>>>(dev.as<T>(), dbl_stage.as<double>(), n);
            SURFH_CUDA(cudaGetLastError());
            own_launches++; launches++;
            SURFH_CUDA(cudaMemcpyAsync(host, dbl_stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        SURFH_CUDA(cudaStreamSynchronize(st));
    }
    void forward_host(const double* x, double* y) override {
        require_ready();
        SURFH_REQUIRE(x && y, "NULL buffer");
        cudaStream_t st = 0;
        const size_t ni = (size_t)input_size(), no = (size_t)output_size();
        to_device(x, x_stage, ni, st);
        // the output staging is shared with adjoint_host's input and a partial handle (band / wavelength
        // shard) does not write every element: zero it at every call so non-owned slices are zeros
        y_stage.ensure(no * sizeof(T));
        SURFH_CUDA(cudaMemsetAsync(y_stage.p, 0, no * sizeof(T), st));
        forward(x_stage.p, y_stage.p, st);
        to_host(y_stage, y, no, st);
    }
    void adjoint_host(const double* y, double* x, int mode) override {
        require_ready();
        SURFH_REQUIRE(x && y, "NULL buffer");
        cudaStream_t st = 0;
        const size_t ni = (size_t)input_size(), no = (size_t)output_size();
        to_device(y, y_stage, no, st);
        x_stage.ensure(ni * sizeof(T));
        adjoint(y_stage.p, x_stage.p, mode, st);
        to_host(x_stage, x, ni, st);
    }

    // ---- CG -----------------------------------------------------------------------------------
    CgScratch scratch() { return CgScratch{cg_partial.as<double>(), cg_ticket.as<unsigned int>()}; }
    int cg_grid(size_t n) const { return (int)std::min<size_t>(kCgMaxBlocks, (n + kCgThreads - 1) / kCgThreads); }
    int n_maps() const { return K > 0 ? K : Nl; }

    void cg_regularise_dot(const void* d, void* q, double mu_s, double mu_r, double* s, cudaStream_t st) override {
        SURFH_REQUIRE(d && q && s, "NULL buffer");
        const size_t n = (size_t)input_size();
        Scope sc(this, ST_CG, st, 3.0 * n * sizeof(T), 10.0 * n, 1, true);
        cg_regularise_dot_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(d),
                                                                       reinterpret_cast<T*>(q), n_maps(), Na, Nb, mu_s,
                                                                       mu_r, s, scratch());
        SURFH_CUDA(cudaGetLastError());
    }
    void laplacian_axpby(const void* x, void* out, double a, double b, cudaStream_t st) override {
        SURFH_REQUIRE(x && out && x != out, "laplacian: NULL or aliased buffers");
        const size_t n = (size_t)input_size();
        Scope sc(this, ST_CG, st, 3.0 * n * sizeof(T), 7.0 * n, 1, true);
        laplacian_axpby_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(out),
                                                                     n_maps(), Na, Nb, a, b);
        SURFH_CUDA(cudaGetLastError());
    }
    void cg_start(const void* b, const void* q, void* r, void* d, double* s, cudaStream_t st) override {
        SURFH_REQUIRE(b && q && r && d && s, "NULL buffer");
        const size_t n = (size_t)input_size();
        Scope sc(this, ST_CG, st, 4.0 * n * sizeof(T), 3.0 * n, 1, true);
        cg_start_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(b), reinterpret_cast<const T*>(q),
                                                              reinterpret_cast<T*>(r), reinterpret_cast<T*>(d), n, s,
                                                              SURFH_CG_NSCALARS, scratch());
        SURFH_CUDA(cudaGetLastError());
    }
    void cg_update(void* x, void* r, void* d, const void* q, double* s, cudaStream_t st) override {
        SURFH_REQUIRE(x && r && d && q && s, "NULL buffer");
        const size_t n = (size_t)input_size();
        Scope sc(this, ST_CG, st, 9.0 * n * sizeof(T), 8.0 * n, 2, true);
        cg_step_kernel<T, false><<<cg_grid(n), kCgThreads, 0, st>>>(
            reinterpret_cast<T*>(x), reinterpret_cast<T*>(r), reinterpret_cast<const T*>(d),
            reinterpret_cast<const T*>(q), nullptr, n, s, SURFH_CG_NSCALARS, scratch());
        cg_direction_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(r), reinterpret_cast<T*>(d),
                                                                  n, s);
        SURFH_CUDA(cudaGetLastError());
    }
    void cg_refresh(int phase, void* x, void* r, void* d, const void* b, const void* qx, double* s,
                    cudaStream_t st) override {
        const size_t n = (size_t)input_size();
        if (phase == 0) {
            SURFH_REQUIRE(x && d && s, "NULL buffer");
            Scope sc(this, ST_CG, st, 3.0 * n * sizeof(T), 2.0 * n, 1, true);
            cg_axpy_alpha_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<T*>(x),
                                                                       reinterpret_cast<const T*>(d), n, s);
        } else {
            SURFH_REQUIRE(r && d && b && qx && s, "NULL buffer");
            Scope sc(this, ST_CG, st, 6.0 * n * sizeof(T), 6.0 * n, 2, true);
            cg_step_kernel<T, true><<<cg_grid(n), kCgThreads, 0, st>>>(
                nullptr, reinterpret_cast<T*>(r), nullptr, reinterpret_cast<const T*>(qx),
                reinterpret_cast<const T*>(b), n, s, SURFH_CG_NSCALARS, scratch());
            cg_direction_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(r),
                                                                      reinterpret_cast<T*>(d), n, s);
        }
        SURFH_CUDA(cudaGetLastError());
    }
    void cg_dot_x_b_plus_r(const void* x, const void* b, const void* r, double* out, cudaStream_t st) override;

    // ---- Fourier-domain block preconditioner (kernels_precond.cuh) ------------------------------
    DevBuf precond_inv;   // [K*K][nfp] real
    bool precond_ready = false;
    template <int KK> void launch_precond_gram(const double* w, double* gram) {
        dim3 grid(ceil_div(nfp, 32)), block(32, kGramLanes);
        precond_gram_kernel<T, KK><<<grid, block>>>(otf.as<C>(), tpl_raw.as<T>(), Nl, w, Nl, nfp, gram);
    }
    template <int KK> void launch_precond_factor(const double* gram, double mu_s, double mu_r, int power) {
        precond_factor_kernel<T, KK><<<ceil_div(nfp, 128), 128>>>(gram, nf, nfp, Na, Nb, Nh, use_own_fft ? 1 : 0, mu_s, mu_r,
                                                                  power, precond_inv.as<T>());
    }
    template <int KK> void launch_precond_apply(cudaStream_t st) {
        precond_apply_kernel<T, KK><<<ceil_div(nfp, 256), 256, 0, st>>>(precond_inv.as<T>(), xhat.as<C>(), nfp,
                                                                        (T)(1.0 / ((double)Na * (double)Nb)));
    }
    void precond_build(const double* w, double mu_s, double mu_r, int joint) override {
        require_ready();
        SURFH_REQUIRE(K > 0, "the Fourier-domain preconditioner needs templates (LMM model)");
        SURFH_REQUIRE(w != nullptr, "NULL wavelength weights");
        SURFH_REQUIRE(mu_s >= 0.0 && mu_r >= 0.0 && mu_s + mu_r > 0.0, "bad hyper-parameters");
        for (int l = 0; l < Nl; ++l)
            SURFH_REQUIRE(w[l] >= 0.0 && (w[l] == 0.0 || otf_set[l]), "weight on a wavelength whose OTF was never uploaded");
        DevBuf dw, gram;
        dw.alloc((size_t)Nl * sizeof(double));
        SURFH_CUDA(cudaMemcpy(dw.p, w, (size_t)Nl * sizeof(double), cudaMemcpyHostToDevice));
        gram.alloc((size_t)(K * (K + 1) / 2) * nfp * sizeof(double));
        precond_inv.alloc((size_t)K * K * nfp * sizeof(T));
        SURFH_DISPATCH_K(launch_precond_gram, dw.as<double>(), gram.as<double>());
        SURFH_CUDA(cudaGetLastError());
        SURFH_DISPATCH_K(launch_precond_factor, gram.as<double>(), mu_s, mu_r, joint ? 2 : 1);
        SURFH_CUDA(cudaGetLastError());
        SURFH_CUDA(cudaDeviceSynchronize());
        own_launches += 2; launches += 2;
        precond_ready = true;
    }
    // z = P r : K-map R2C, per-bin K x K product, K-map C2R (xhat is the scratch spectrum)
    void precond_apply(const void* r, void* z, cudaStream_t st) override {
        require_ready();
        SURFH_REQUIRE(precond_ready, "surfh_precond_build has not been called");
        SURFH_REQUIRE(r && z, "NULL buffer");
        {
            Scope sc(this, ST_RFFT_MAPS, st, fft_bytes(K, -1), fft_flops(K, -1), fft_launches(), use_own_fft);
            fft_exec(0, K, const_cast<void*>(r), xhat.p, st);
        }
        {
            Scope sc(this, ST_CG, st, (double)K * nf * sizeof(C) * 2 + (double)K * K * nf * sizeof(T), 4.0 * K * K * nf, 1, true);
            SURFH_DISPATCH_K(launch_precond_apply, st);
            SURFH_CUDA(cudaGetLastError());
        }
        Scope sc(this, ST_IRFFT_MAPS, st, fft_bytes(K, -1), fft_flops(K, -1), fft_launches(), use_own_fft);
        fft_exec(1, K, xhat.p, z, st);
    }
    // phase 0: x += alpha d, r -= alpha q (alpha = s[0] / s[1], s[0] = rho_z);  phase 1: r = b - q (q = Q x: refresh)
    void pcg_update(int phase, void* x, void* r, const void* d, const void* q, const void* b, double* s,
                    cudaStream_t st) override {
        const size_t n = (size_t)input_size();
        SURFH_REQUIRE(r && q && s, "NULL buffer");
        Scope sc(this, ST_CG, st, 6.0 * n * sizeof(T), 6.0 * n, 1, true);
        if (phase == 0) {
            SURFH_REQUIRE(x && d, "NULL buffer");
            cg_step_kernel<T, false><<<cg_grid(n), kCgThreads, 0, st>>>(
                reinterpret_cast<T*>(x), reinterpret_cast<T*>(r), reinterpret_cast<const T*>(d),
                reinterpret_cast<const T*>(q), nullptr, n, s, SURFH_CG_NSCALARS, scratch());
        } else {
            SURFH_REQUIRE(b, "NULL buffer");
            cg_step_kernel<T, true><<<cg_grid(n), kCgThreads, 0, st>>>(
                nullptr, reinterpret_cast<T*>(r), nullptr, reinterpret_cast<const T*>(q),
                reinterpret_cast<const T*>(b), n, s, SURFH_CG_NSCALARS, scratch());
        }
        SURFH_CUDA(cudaGetLastError());
    }
    // rho_z' = <r, z>, beta = rho_z' / rho_z, d = z + beta d  (first: d = z)
    void pcg_direction(const void* r, const void* z, void* d, double* s, int first, cudaStream_t st) override {
        const size_t n = (size_t)input_size();
        SURFH_REQUIRE(r && z && d && s, "NULL buffer");
        Scope sc(this, ST_CG, st, 5.0 * n * sizeof(T), 4.0 * n, 2, true);
        if (first)
            pcg_dot_kernel<T, true><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(r),
                                                                        reinterpret_cast<const T*>(z), n, s, scratch());
        else
            pcg_dot_kernel<T, false><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(r),
                                                                         reinterpret_cast<const T*>(z), n, s, scratch());
        cg_direction_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(z), reinterpret_cast<T*>(d), n, s);
        SURFH_CUDA(cudaGetLastError());
    }
    void axpy_device_scalar(void* y, const void* x, int64_t n, const double* s, int idx, cudaStream_t st) override {
        SURFH_REQUIRE(y && x && s && n >= 0 && idx >= 0, "axpy: bad argument");
        if (n == 0) return;
        Scope sc(this, ST_CG, st, 3.0 * n * sizeof(T), 2.0 * n, 1, true);
        axpy_device_scalar_kernel<T><<<cg_grid((size_t)n), kCgThreads, 0, st>>>(reinterpret_cast<T*>(y),
                                                                                reinterpret_cast<const T*>(x), (size_t)n, s, idx);
        SURFH_CUDA(cudaGetLastError());
    }
    void criterion_terms(const void* y, const void* hx, int64_t n, const void* x, double* out, cudaStream_t st) override {
        SURFH_REQUIRE(out, "NULL buffer");
        const size_t nmax = std::max<size_t>((size_t)std::max<int64_t>(n, 0), x ? (size_t)input_size() : 0);
        Scope sc(this, ST_CG, st, 2.0 * n * sizeof(T), 3.0 * n, 1, true);
        criterion_kernel<T><<<cg_grid(std::max<size_t>(nmax, 1)), kCgThreads, 0, st>>>(
            reinterpret_cast<const T*>(y), reinterpret_cast<const T*>(hx), (size_t)std::max<int64_t>(n, 0),
            reinterpret_cast<const T*>(x), n_maps(), Na, Nb, out, scratch());
        SURFH_CUDA(cudaGetLastError());
    }
};

template <typename T>
void ModelImpl<T>::cg_dot_x_b_plus_r(const void* x, const void* b, const void* r, double* out, cudaStream_t st) {
    SURFH_REQUIRE(x && b && r && out, "NULL buffer");
    const size_t n = (size_t)input_size();
    Scope sc(this, ST_CG, st, 3.0 * n * sizeof(T), 2.0 * n, 1, true);
    cg_dot_x_b_plus_r_kernel<T><<<cg_grid(n), kCgThreads, 0, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<const T*>(b),
                                                                   reinterpret_cast<const T*>(r), n, out, scratch());
    SURFH_CUDA(cudaGetLastError());
}

template <> void ModelImpl<float>::gemm_grouped_f64(double*, bool, cudaStream_t) {
    throw Error(SURFH_ESTATE, "internal: fp64 GEMM on an fp32 model");
}
template <> void ModelImpl<double>::gemm_grouped_f32(float*, bool, cudaStream_t) {
    throw Error(SURFH_ESTATE, "internal: fp32 GEMM on an fp64 model");
}

template <> void ModelImpl<float>::gemm_grouped_f32(float* y, bool adjoint, cudaStream_t st) {
    std::vector<size_t> lsf_bands;
    for (size_t i = 0; i < bands.size(); ++i)
        if (bands[i]->mode == SURFH_SPECTRAL_LSF) lsf_bands.push_back(i);
    // longest tiles first (a tile's cost is its contraction length): the block scheduler hands tiles out in
    // blockIdx order, so the tail of the grouped launch is made of the short ones
    std::stable_sort(lsf_bands.begin(), lsf_bands.end(), [&](size_t x, size_t y) {
        const int kx = adjoint ? bands[x]->nd : bands[x]->KB, ky = adjoint ? bands[y]->nd : bands[y]->KB;
        return kx > ky;
    });
    for (size_t first = 0; first < lsf_bands.size(); first += kMaxGemmGroup) {
        GemmBatchF batch;
        batch.count = 0;
        batch.tile_start[0] = 0;
        double bytes = 0, flops = 0;
        for (size_t j = first; j < std::min(lsf_bands.size(), first + (size_t)kMaxGemmGroup); ++j) {
            BandT<float>& b = *bands[lsf_bands[j]];
            GemmArgs<float> g = gemm_args(b, y, adjoint);
            batch.p[batch.count] = g;
            batch.tile_start[batch.count + 1] =
                batch.tile_start[batch.count] + ceil_div(g.M, kFBM) * ceil_div(g.N, kFBN);
            batch.count++;
            bytes += sizeof(float) * ((double)b.nd * b.KB + (double)b.nl * b.ncol + (double)b.out_size);
            flops += 2.0 * g.M * g.N * g.K;
        }
        Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, flops, 1, true);
        const int tiles = batch.tile_start[batch.count];
        if (!adjoint)
            sgemm_tf32x3_kernel<true, true><<<tiles, 256, sgemm_smem_bytes<true, true>(), st>>>(batch);
        else
            sgemm_tf32x3_kernel<false, false><<<tiles, 256, sgemm_smem_bytes<false, false>(), st>>>(batch);
        SURFH_CUDA(cudaGetLastError());
    }
}

template <> void ModelImpl<float>::gemm_grouped_f64_tma(double*, bool, cudaStream_t) {
    throw Error(SURFH_ESTATE, "internal: fp64 GEMM on an fp32 model");
}

template <> void ModelImpl<double>::gemm_grouped_f64_tma(double* y, bool adjoint, cudaStream_t st) {
    std::vector<size_t> lsf_bands;
    for (size_t i = 0; i < bands.size(); ++i)
        if (bands[i]->mode == SURFH_SPECTRAL_LSF) lsf_bands.push_back(i);
    std::stable_sort(lsf_bands.begin(), lsf_bands.end(), [&](size_t x, size_t y2) {
        const int kx = adjoint ? bands[x]->nd : bands[x]->KB, ky = adjoint ? bands[y2]->nd : bands[y2]->KB;
        return kx > ky;
    });
    if (adjoint) {
        // detector blocks -> K-fast per detector column (the B operand of Gt = Wt . Yk)
        double bytes = 0;
        for (size_t j : lsf_bands) bytes += 2.0 * sizeof(double) * (double)bands[j]->out_size;
        Scope sc(this, ST_GEMM_ADJ, st, bytes, 0, (int)lsf_bands.size(), true);
        for (size_t j : lsf_bands) {
            BandT<double>& b = *bands[j];
            const size_t n = (size_t)b.Nn * b.nd;
            detector_to_kfast_kernel<double><<<ceil_div(n, 256), 256, 0, st>>>(y + b.out_offset, b.na, b.nd, b.Nn, b.ndp,
                                                                                b.yk.as<double>());
        }
        SURFH_CUDA(cudaGetLastError());
    }
    for (size_t first = 0; first < lsf_bands.size(); first += kMaxGemmGroup) {
        GemmTmaBatch batch;
        batch.count = 0;
        batch.tile_start[0] = 0;
        double bytes = 0, flops = 0;
        for (size_t j = first; j < std::min(lsf_bands.size(), first + (size_t)kMaxGemmGroup); ++j) {
            BandT<double>& b = *bands[lsf_bands[j]];
            GemmTmaProblem& g = batch.p[batch.count];
            if (!adjoint) {   // y = W . G
                g.a = b.map_w; g.b = b.map_g; g.M = b.nd; g.N = b.Nn; g.K = b.KB;
                g.C = y + b.out_offset; g.cM = b.t_yM.as<int32_t>(); g.cN = b.t_yN.as<int32_t>();
            } else {          // Gt = Wt . Yk
                g.a = b.map_wt; g.b = b.map_yk; g.M = b.KB; g.N = b.Nn; g.K = b.nd;
                g.C = b.G.as<double>(); g.cM = b.t_ident.as<int32_t>(); g.cN = b.t_gN.as<int32_t>();
            }
            batch.tile_start[batch.count + 1] = batch.tile_start[batch.count] + ceil_div(g.M, kTBM) * ceil_div(g.N, kTBN);
            batch.count++;
            bytes += sizeof(double) * ((double)b.nd * b.KB + (double)b.nl * b.ncol + (double)b.out_size);
            flops += 2.0 * g.M * g.N * g.K;
        }
        Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, flops, 1, true);
        dgemm_tma_kernel<<<batch.tile_start[batch.count], kTThreads, kTSmemBytes, st>>>(batch);
        SURFH_CUDA(cudaGetLastError());
    }
}

template <> void ModelImpl<double>::gemm_grouped_f64(double* y, bool adjoint, cudaStream_t st) {
    if (gemm_mode != GEMM_LEGACY) return gemm_grouped_f64_tma(y, adjoint, st);
    std::vector<size_t> lsf_bands;
    for (size_t i = 0; i < bands.size(); ++i)
        if (bands[i]->mode == SURFH_SPECTRAL_LSF) lsf_bands.push_back(i);
    // longest tiles first (a tile's cost is its contraction length): the block scheduler hands tiles out in
    // blockIdx order, so the tail of the grouped launch is made of the short ones
    std::stable_sort(lsf_bands.begin(), lsf_bands.end(), [&](size_t x, size_t y) {
        const int kx = adjoint ? bands[x]->nd : bands[x]->KB, ky = adjoint ? bands[y]->nd : bands[y]->KB;
        return kx > ky;
    });
    for (size_t first = 0; first < lsf_bands.size(); first += kMaxGemmGroup) {
        GemmBatch batch;
        batch.count = 0;
        batch.tile_start[0] = 0;
        double bytes = 0, flops = 0;
        for (size_t j = first; j < std::min(lsf_bands.size(), first + (size_t)kMaxGemmGroup); ++j) {
            BandT<double>& b = *bands[lsf_bands[j]];
            GemmArgs<double> g = gemm_args(b, y, adjoint);
            batch.p[batch.count] = g;
            batch.tile_start[batch.count + 1] =
                batch.tile_start[batch.count] + ceil_div(g.M, kDBM) * ceil_div(g.N, kDBN);
            batch.count++;
            bytes += sizeof(double) * ((double)b.nd * b.KB + (double)b.nl * b.ncol + (double)b.out_size);
            flops += 2.0 * g.M * g.N * g.K;
        }
        Scope sc(this, adjoint ? ST_GEMM_ADJ : ST_GEMM_FWD, st, bytes, flops, 1, true);
        const int tiles = batch.tile_start[batch.count];
        if (!adjoint)
            dgemm_mma_kernel<true, true><<<tiles, 256, dgemm_smem_bytes<true, true>(), st>>>(batch);
        else
            dgemm_mma_kernel<false, false><<<tiles, 256, dgemm_smem_bytes<false, false>(), st>>>(batch);
        SURFH_CUDA(cudaGetLastError());
    }
}

}  // namespace surfh

// ------------------------------------------------------------------------------------------------
// A handle belongs to the device that was current at surfh_create: every entry point runs with that device
// current (and restores the caller's on exit), so a multi-GPU process cannot launch on the wrong GPU.
namespace {
struct DeviceScope {
    int prev = -1;
    bool switched = false;
    explicit DeviceScope(int want) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != want) {
            if (cudaSetDevice(want) != cudaSuccess) throw Error(SURFH_ECUDA, "cudaSetDevice(handle's device) failed");
            switched = true;
        }
    }
    ~DeviceScope() {
        if (switched) cudaSetDevice(prev);
    }
};
}  // namespace

#define SURFH_API_BEGIN(h)                                   \
    if (!(h)) return SURFH_EINVAL;                           \
    try {                                                    \
        DeviceScope surfh_device_scope_((h)->device);
#define SURFH_API_END(h)                                     \
    }                                                        \
    catch (const surfh::Error& e) {                          \
        (h)->last_error = e.what();                          \
        return e.code;                                       \
    }                                                        \
    catch (const std::bad_alloc&) {                          \
        (h)->last_error = "host allocation failed";          \
        return SURFH_ENOMEM;                                 \
    }                                                        \
    catch (const std::exception& e) {                        \
        (h)->last_error = e.what();                          \
        return SURFH_EINVAL;                                 \
    }                                                        \
    return SURFH_OK;

// Stand-alone batched 2-D real FFT pair through the hand-written kernels (plans cached per shape).
namespace {
template <typename T> struct FftCacheEntry {
    surfh::OwnFft2d<T> fft;
    surfh::DevBuf z;
};
template <typename T> int rfft2_impl(int na, int nb, int batch, int inverse, const void* in, void* out, cudaStream_t st) {
    using Cx = surfh::cplx_t<T>;
    // plans (device tables + scratch) are per (device, shape); the cache is process-wide, so calls are
    // serialised: concurrent callers would otherwise share one scratch buffer
    static std::map<std::tuple<int, int, int>, std::unique_ptr<FftCacheEntry<T>>> cache;
    static std::mutex cache_mutex;
    if (!surfh::OwnFft2d<T>::supported(na, nb)) throw Error(SURFH_EINVAL, "surfh_rfft2: axes must be in [2, 512]");
    if (batch <= 0 || !in || !out) throw Error(SURFH_EINVAL, "surfh_rfft2: bad batch or NULL buffer");
    int device = 0;
    SURFH_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(cache_mutex);
    auto& e = cache[std::make_tuple(device, na, nb)];
    if (!e) {
        e = std::make_unique<FftCacheEntry<T>>();
        e->fft.init(na, nb);
    }
    e->z.ensure((size_t)batch * e->fft.z_plane() * sizeof(Cx));
    const size_t rp = (size_t)na * nb, sp = (size_t)na * (nb / 2 + 1);
    if (!inverse) e->fft.r2c(reinterpret_cast<const T*>(in), rp, reinterpret_cast<Cx*>(out), sp, e->z.template as<Cx>(), batch, st, false);
    else e->fft.c2r(reinterpret_cast<const Cx*>(in), sp, reinterpret_cast<T*>(out), rp, e->z.template as<Cx>(), batch, st, false);
    return SURFH_OK;
}
}  // namespace

extern "C" {

int surfh_abi_version(void) { return SURFH_ABI_VERSION; }

int surfh_create(const surfh_model_desc* desc, surfh_handle* out) {
    if (!desc || !out) {
        g_create_error = "NULL argument";
        return SURFH_EINVAL;
    }
    *out = nullptr;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Error(SURFH_ECUDA, std::string("no CUDA device available (surfh_b200 has no CPU fallback): ") +
                                         cudaGetErrorString(e));
        if (desc->dtype == SURFH_F64) {
            auto m = std::make_unique<ModelImpl<double>>();
            m->dtype = SURFH_F64;
            m->init(desc);
            *out = m.release();
        } else if (desc->dtype == SURFH_F32) {
            auto m = std::make_unique<ModelImpl<float>>();
            m->dtype = SURFH_F32;
            m->init(desc);
            *out = m.release();
        } else {
            throw Error(SURFH_EINVAL, "dtype must be SURFH_F32 or SURFH_F64");
        }
    } catch (const surfh::Error& e) {
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return SURFH_EINVAL;
    }
    return SURFH_OK;
}

int surfh_set_otf(surfh_handle h, int32_t l_start, int32_t l_count, const void* otf) {
    SURFH_API_BEGIN(h) h->set_otf(l_start, l_count, otf);
    SURFH_API_END(h)
}
int surfh_add_band(surfh_handle h, const surfh_band_desc* band) {
    SURFH_API_BEGIN(h)
    if (!band) throw Error(SURFH_EINVAL, "NULL band descriptor");
    h->add_band(band);
    SURFH_API_END(h)
}
int surfh_finalize(surfh_handle h) {
    SURFH_API_BEGIN(h) h->finalize();
    SURFH_API_END(h)
}
void surfh_destroy(surfh_handle h) { delete h; }
const char* surfh_last_error(surfh_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int64_t surfh_input_size(surfh_handle h) { return h ? h->input_size() : -1; }
int64_t surfh_output_size(surfh_handle h) { return h ? h->output_size() : -1; }
int64_t surfh_workspace_bytes(surfh_handle h) { return h ? h->workspace_bytes() : -1; }

int surfh_forward(surfh_handle h, const void* x, void* y, void* stream) {
    SURFH_API_BEGIN(h) h->forward(x, y, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_adjoint(surfh_handle h, const void* y, void* x, int32_t mode, void* stream) {
    SURFH_API_BEGIN(h) h->adjoint(y, x, mode, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_fwadj(surfh_handle h, const void* x, void* out, int32_t mode, void* y_scratch, void* stream) {
    SURFH_API_BEGIN(h) h->fwadj(x, out, mode, y_scratch, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_maps_to_cube(surfh_handle h, const void* maps, float* cube, void* stream) {
    SURFH_API_BEGIN(h) h->maps_to_cube(maps, cube, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_forward_host(surfh_handle h, const double* x, double* y) {
    SURFH_API_BEGIN(h) h->forward_host(x, y);
    SURFH_API_END(h)
}
int surfh_adjoint_host(surfh_handle h, const double* y, double* x, int32_t mode) {
    SURFH_API_BEGIN(h) h->adjoint_host(y, x, mode);
    SURFH_API_END(h)
}

int surfh_cg_regularise_dot(surfh_handle h, const void* d, void* q, double mu_s, double mu_r, double* s, void* stream) {
    SURFH_API_BEGIN(h) h->cg_regularise_dot(d, q, mu_s, mu_r, s, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_laplacian_axpby(surfh_handle h, const void* x, void* out, double a, double b, void* stream) {
    SURFH_API_BEGIN(h) h->laplacian_axpby(x, out, a, b, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_cg_start(surfh_handle h, const void* b, const void* q, void* r, void* d, double* s, void* stream) {
    SURFH_API_BEGIN(h) h->cg_start(b, q, r, d, s, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_cg_update(surfh_handle h, void* x, void* r, void* d, const void* q, double* s, void* stream) {
    SURFH_API_BEGIN(h) h->cg_update(x, r, d, q, s, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_cg_refresh(surfh_handle h, int32_t phase, void* x, void* r, void* d, const void* b, const void* qx, double* s,
                     void* stream) {
    SURFH_API_BEGIN(h) h->cg_refresh(phase, x, r, d, b, qx, s, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_criterion_terms(surfh_handle h, const void* y, const void* hx, int64_t n, const void* x, double* s_out,
                          void* stream) {
    SURFH_API_BEGIN(h) h->criterion_terms(y, hx, n, x, s_out, (cudaStream_t)stream);
    SURFH_API_END(h)
}

int surfh_cg_dot_x_b_plus_r(surfh_handle h, const void* x, const void* b, const void* r, double* s_out, void* stream) {
    SURFH_API_BEGIN(h) h->cg_dot_x_b_plus_r(x, b, r, s_out, (cudaStream_t)stream);
    SURFH_API_END(h)
}

int surfh_axpy_device_scalar(surfh_handle h, void* y, const void* x, int64_t n, const double* s, int32_t idx, void* stream) {
    SURFH_API_BEGIN(h) h->axpy_device_scalar(y, x, n, s, idx, (cudaStream_t)stream);
    SURFH_API_END(h)
}

int surfh_precond_build(surfh_handle h, const double* w_lambda, double mu_s, double mu_r, int32_t joint) {
    SURFH_API_BEGIN(h) h->precond_build(w_lambda, mu_s, mu_r, joint);
    SURFH_API_END(h)
}
int surfh_precond_apply(surfh_handle h, const void* r, void* z, void* stream) {
    SURFH_API_BEGIN(h) h->precond_apply(r, z, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_pcg_update(surfh_handle h, int32_t phase, void* x, void* r, const void* d, const void* q, const void* b, double* s,
                     void* stream) {
    SURFH_API_BEGIN(h) h->pcg_update(phase, x, r, d, q, b, s, (cudaStream_t)stream);
    SURFH_API_END(h)
}
int surfh_pcg_direction(surfh_handle h, const void* r, const void* z, void* d, double* s, int32_t first, void* stream) {
    SURFH_API_BEGIN(h) h->pcg_direction(r, z, d, s, first, (cudaStream_t)stream);
    SURFH_API_END(h)
}

int surfh_rfft2(int32_t dtype, int32_t n_alpha, int32_t n_beta, int32_t batch, int32_t inverse, const void* in, void* out,
                void* stream) {
    try {
        if (dtype == SURFH_F64) return rfft2_impl<double>(n_alpha, n_beta, batch, inverse, in, out, (cudaStream_t)stream);
        if (dtype == SURFH_F32) return rfft2_impl<float>(n_alpha, n_beta, batch, inverse, in, out, (cudaStream_t)stream);
        throw Error(SURFH_EINVAL, "dtype must be SURFH_F32 or SURFH_F64");
    } catch (const surfh::Error& e) {
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return SURFH_EINVAL;
    }
}

int surfh_shepard(const float* alpha_coord, const float* lambda_coord, const float* values, int32_t n_in,
                  const float* alpha_mesh, const float* lambda_mesh, int32_t n_out, float p, float alpha,
                  float pixel_cutoff, float alpha_res, float lambda_res, float epsilon, float* out, void* stream) {
    try {
        if (n_in < 0 || n_out <= 0 || !alpha_mesh || !lambda_mesh || !out || (n_in > 0 && (!alpha_coord || !lambda_coord || !values)))
            throw Error(SURFH_EINVAL, "surfh_shepard: bad size or NULL buffer");
        if (!(alpha_res != 0.f) || !(lambda_res != 0.f)) throw Error(SURFH_EINVAL, "surfh_shepard: zero resolution");
        shepard_kernel<<<ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream>>>(
            alpha_coord, lambda_coord, values, n_in, alpha_mesh, lambda_mesh, n_out, p, alpha, pixel_cutoff,
            1.f / alpha_res, 1.f / lambda_res, epsilon, out);
        SURFH_CUDA(cudaGetLastError());
        return SURFH_OK;
    } catch (const surfh::Error& e) {
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return SURFH_EINVAL;
    }
}

int64_t surfh_launch_count(surfh_handle h) { return h ? h->launches : -1; }
int64_t surfh_own_launch_count(surfh_handle h) { return h ? h->own_launches : -1; }

int surfh_contraction_info(surfh_handle h, int32_t* mode, int32_t* digits, double* executed_fraction) {
    if (!h) return SURFH_EINVAL;
    h->contraction_info(mode, digits, executed_fraction);
    return SURFH_OK;
}

int surfh_profile_enable(surfh_handle h, int32_t on) {
    if (!h) return SURFH_EINVAL;
    for (auto& r : h->recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->recs.clear();
    for (int i = 0; i < ST_COUNT; ++i) {
        h->stage_bytes[i] = 0;
        h->stage_flops[i] = 0;
        h->stage_launches[i] = 0;
    }
    h->profiling = on != 0;
    return SURFH_OK;
}

int surfh_profile_read(surfh_handle h, int32_t cap, const char** names, float* ms, double* bytes, double* flops,
                       int32_t* launches) {
    if (!h || cap < 0) return SURFH_EINVAL;
    cudaDeviceSynchronize();
    float acc[ST_COUNT] = {0};
    for (auto& r : h->recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) acc[r.stage] += t;
    }
    int n = 0;
    for (int i = 0; i < ST_COUNT && n < cap; ++i) {
        if (h->stage_launches[i] == 0) continue;
        if (names) names[n] = h->own_fft_names ? kStageNamesOwnFft[i] : kStageNamesCufft[i];
        if (ms) ms[n] = acc[i];
        if (bytes) bytes[n] = h->stage_bytes[i];
        if (flops) flops[n] = h->stage_flops[i];
        if (launches) launches[n] = h->stage_launches[i];
        ++n;
    }
    return n;
}

}  // extern "C"

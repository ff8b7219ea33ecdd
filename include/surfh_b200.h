/* surfh_b200 -- C ABI of the B200-native LMM instrument operator and CG primitives.
 *
 * This is the drop-in boundary for ONE hot path of sidiso/surfh: the linear-mixing-model
 * MIRI-MRS operator `spectroSigRLSCT` (forward / adjoint / fwadj) and the vector work of the
 * conjugate-gradient loop that drives it.  The reference has no FFI for this path (it is pure
 * Python + Cython + JAX); each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SURFH_E* code otherwise; the message is
 *     available from surfh_last_error(); no exception crosses the boundary;
 *   - one handle per GPU (the device current at surfh_create; every entry point makes that device
 *     current for the duration of the call and restores the caller's); a handle is not thread-safe;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it, nothing synchronises unless stated;
 *   - "real" means the handle's dtype: SURFH_F64 -> double, SURFH_F32 -> float; complex means
 *     interleaved (re, im) pairs of that type;
 *   - pointers documented as [device] must be device memory of the handle's GPU and are
 *     caller-owned; pointers documented as [any] may be host or device (cudaMemcpyDefault);
 *     descriptor tables are copied at the call, the caller may free them afterwards;
 *   - there is no CPU fallback: without a CUDA device surfh_create fails.
 */
#ifndef SURFH_B200_H
#define SURFH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURFH_ABI_VERSION 5

enum { SURFH_F32 = 0, SURFH_F64 = 1 };

/* Adjoint flavour.  The reference's `gridding_t` (spectroModelChannel.py:180-199) is a second
 * bilinear interpolation, not the transpose of `gridding`; SURFH_ADJ_REFERENCE reproduces it,
 * SURFH_ADJ_EXACT applies the true transpose (dot-test to rounding). */
enum { SURFH_ADJ_EXACT = 0, SURFH_ADJ_REFERENCE = 1 };
enum { SURFH_SPECTRAL_LSF = 0, SURFH_SPECTRAL_BETA_SUM = 1 };
/* 2-D real FFT pair of the cube planes (jax_utils.py:30-41).  AUTO = the hand-written chirp-z
 * kernels when both map axes are <= 1024 pixels, cuFFT otherwise. */
enum { SURFH_FFT_AUTO = 0, SURFH_FFT_CUFFT = 1, SURFH_FFT_OWN = 2 };

enum {
    SURFH_OK = 0,
    SURFH_EINVAL = -1,   /* bad argument / inconsistent descriptor          */
    SURFH_ECUDA = -2,    /* CUDA runtime error                              */
    SURFH_ECUFFT = -3,   /* cuFFT error                                     */
    SURFH_ESTATE = -4,   /* call sequence violated (e.g. not finalised)     */
    SURFH_ENOMEM = -5
};

typedef struct surfh_model surfh_model;
typedef surfh_model* surfh_handle;

/* Whole-model description: what spectroSigRLSCT.__init__ receives
 * (surfh/Models/spectroModel.py:40-135) minus the per-band part. */
typedef struct {
    int32_t dtype;        /* SURFH_F32 / SURFH_F64: arithmetic type of every kernel           */
    int32_t n_templates;  /* K; 0 = no LMM (templates=None, input is the [n_lambda,N,N] cube) */
    int32_t n_alpha;      /* cube pixels along alpha (axis 0)                                 */
    int32_t n_beta;       /* cube pixels along beta  (axis 1, contiguous)                     */
    int32_t n_lambda;     /* cube wavelengths                                                 */
    int32_t chunk;        /* wavelengths per pipeline chunk (0 = library default)             */
    const double* templates; /* [any] [K, n_lambda] row-major, NULL when K == 0               */
    int32_t fft_backend;  /* SURFH_FFT_*; the environment variable SURFH_FFT_BACKEND
                             (auto | own | cufft) overrides it                                */
} surfh_model_desc;

/* Sparse table of one adjoint flavour for one band, pointings merged: cube pixel -> weighted
 * entries of the slit-space vector G[(p, s, a, b)].  Rows are the cube pixels that receive
 * anything (compact list), in increasing pixel order. */
typedef struct {
    int32_t n_rows;
    int64_t nnz;
    const int32_t* row_pixel; /* [n_rows]   flat cube pixel i*n_beta + j                      */
    const int64_t* row_ptr;   /* [n_rows+1]                                                   */
    const int32_t* col;       /* [nnz]      ((p*S + s)*na + a)*nb + b                         */
    const double* val;        /* [nnz]                                                        */
} surfh_csr;

/* One IFU band: everything `Channel.__init__` derives
 * (surfh/Models/spectroModelChannel.py:27-108) flattened into tables. */
typedef struct {
    int32_t n_pointing;  /* P                                                                  */
    int32_t n_slit;      /* S                                                                  */
    int32_t na;          /* detector pixels along a slit = ceil(npix_alpha / srf)              */
    int32_t nb;          /* cube pixels across a slit                                          */
    int32_t srf;         /* super-resolution factor along alpha                                */
    int32_t local_a;     /* A: local grid rows                                                 */
    int32_t local_b;     /* B: local grid columns                                              */
    int32_t wave_start;  /* first cube wavelength of the band (wslice.start)                   */
    int32_t n_wave;      /* Lambda  = wslice.stop - wslice.start                               */
    int32_t n_det;       /* Lambda' = detector wavelength samples                              */
    int32_t spectral_mode; /* SURFH_SPECTRAL_LSF: y = sum_{l,b} lsf * G (spectroSigRLSCT);
                            SURFH_SPECTRAL_BETA_SUM: y[l] = sum_b G[l] (no spectral response, one
                            output row per cube wavelength: MRSBlurred, spectro_blind.py:191-207);
                            then n_det = rows of the band's y block and lsf may be NULL          */
    int32_t det_start;   /* beta-sum mode: y row of the first local wavelength (0 unless sharded) */
    int64_t out_offset;  /* element offset of this band in the caller's y vector (_idx[band])  */
    const int32_t* slit_a0;   /* [S]      first local row of the slit                          */
    const int32_t* slit_b0;   /* [S]      first local column of the slit                       */
    const double* slit_w;     /* [S, nb]  beta edge weights (Slicer.get_slit_weights)          */
    const double* lsf;        /* [n_det, n_wave, nb] spectral response (SpectralBlur.psfs)     */
    const int32_t* grid_base; /* [P, A*B] flat cube index of the upper-left bilinear tap       */
    const double* grid_frac;  /* [P, A*B, 2] normalised distances (t_alpha, t_beta)            */
    surfh_csr adj_exact;      /* transpose of gather                                           */
    surfh_csr adj_reference;  /* the reference's gridding_t chain                              */
} surfh_band_desc;

/* ---- lifetime ---------------------------------------------------------------------------- */
int surfh_abi_version(void);
/* replaces spectroSigRLSCT.__init__ (spectroModel.py:40-135) */
int surfh_create(const surfh_model_desc* desc, surfh_handle* out);
/* upload OTF planes [l_start, l_start+l_count) : complex128 [l_count, n_alpha, n_beta/2+1]
 * ([any]; converted to the handle dtype).  Replaces `self.sotf = sotf` (spectroModel.py:51). */
int surfh_set_otf(surfh_handle h, int32_t l_start, int32_t l_count, const void* otf_c128);
/* replaces Channel.__init__ (spectroModelChannel.py:27-108); bands keep the order of calls */
int surfh_add_band(surfh_handle h, const surfh_band_desc* band);
/* build FFT plans and workspaces; must precede any compute call */
int surfh_finalize(surfh_handle h);
void surfh_destroy(surfh_handle h);
const char* surfh_last_error(surfh_handle h); /* h may be NULL: error of the last failed create */

/* ---- shapes ------------------------------------------------------------------------------ */
int64_t surfh_input_size(surfh_handle h);  /* K*N*N, or n_lambda*N*N without LMM               */
int64_t surfh_output_size(surfh_handle h); /* highest out_offset + band size over added bands  */
int64_t surfh_workspace_bytes(surfh_handle h);

/* ---- operator, device buffers ------------------------------------------------------------ */
/* y[out_offset_b ...] = H_b x for every added band.  x: [device] real [K,N,N]; y: [device] real.
 * Replaces spectroSigRLSCT.forward (spectroModel.py:158-170). */
int surfh_forward(surfh_handle h, const void* x, void* y, void* stream);
/* x = sum_b H_b^T y_b (mode: SURFH_ADJ_*).  Replaces spectroSigRLSCT.adjoint (:173-185). */
int surfh_adjoint(surfh_handle h, const void* y, void* x, int32_t mode, void* stream);
/* out = H^T H x without leaving the device; y_scratch: [device] real [output_size] or NULL to use
 * an internal buffer.  Replaces aljabr.LinOp.fwadj as inherited by spectroSigRLSCT. */
int surfh_fwadj(surfh_handle h, const void* x, void* out, int32_t mode, void* y_scratch, void* stream);
/* cube = T x in float32, result export.  Replaces spectroSigRLSCT.mapsToCube
 * (spectroModel.py:190-192 -> cythons_files.pyx:424-440).  maps: [device] real; cube: [device] float */
int surfh_maps_to_cube(surfh_handle h, const void* maps, float* cube, void* stream);

/* ---- the 2-D real FFT pair on its own --------------------------------------------------- */
/* Batched un-normalised 2-D real transforms by the hand-written chirp-z kernels that the operator
 * uses for its cube planes: inverse == 0: real [batch, n_alpha, n_beta] -> half-complex
 * [batch, n_alpha, n_beta/2+1] (numpy.fft.rfft2); inverse != 0: the reverse, scaled by
 * n_alpha*n_beta (numpy.fft.irfft2 * n_alpha*n_beta).  Replaces the rfftn / irfftn calls of
 * surfh/ToolsDir/jax_utils.py:30-41 (python_utils.py:41-71) up to the "ortho" factor.
 * in, out: [device], contiguous; axes in [2, 1024]; errors via surfh_last_error(NULL). */
int surfh_rfft2(int32_t dtype, int32_t n_alpha, int32_t n_beta, int32_t batch, int32_t inverse, const void* in, void* out,
                void* stream);

/* ---- operator, host buffers (the reference-facing call: numpy in, numpy out) --------------- */
/* double host arrays whatever the handle dtype; copies through pinned staging, synchronises. */
int surfh_forward_host(surfh_handle h, const double* x, double* y);
int surfh_adjoint_host(surfh_handle h, const double* y, double* x, int32_t mode);

/* ---- conjugate-gradient vector work (qmm.lcg as called from fusion_CT.py:194-232) ---------- */
/* Scalars live in a caller-owned [device] double array `s` of at least SURFH_CG_NSCALARS + max_iter
 * + 2 entries: s[0]=rho=<r,r>  s[1]=<d,q>  s[2]=alpha  s[3]=beta  s[4]=iteration counter,
 * s[SURFH_CG_NSCALARS + i] = grad_norm history.  No call below synchronises with the host. */
#define SURFH_CG_NSCALARS 8
/* q = mu_s*q + mu_r*(D_r^T D_r + D_c^T D_c) d ; s[1] = <d,q>.  n_maps*[n_alpha,n_beta] vectors.
 * Replaces the NpDiff_r / NpDiff_c hessp terms (fusion_CT.py:16-43) and the <d,Qd> dot. */
int surfh_cg_regularise_dot(surfh_handle h, const void* d, void* q, double mu_s, double mu_r, double* s,
                            void* stream);
/* out = a*out + b*L(x) on n_maps*[n_alpha,n_beta] vectors, L = circular 5-point Laplacian (x != out;
 * a == 0 does not read out).  Replaces Difference_Operator_Joint.D / D_t / DtD = L, L, L.L
 * (fusion_CT.py:45-63: the udft.laplacian(2) impulse response applied in Fourier space). */
int surfh_laplacian_axpby(surfh_handle h, const void* x, void* out, double a, double b, void* stream);
/* r = b - q ; d = r ; s[0] = <r,r> ; history[0] = s[0] ; counter = 0 (lcg initialisation) */
int surfh_cg_start(surfh_handle h, const void* b, const void* q, void* r, void* d, double* s, void* stream);
/* alpha = s[0]/s[1]; x += alpha d; r -= alpha q; rho' = <r,r>; beta = rho'/rho; d = r + beta d;
 * s[0] = rho'; history appended.  One lcg iteration after q = Q d. */
int surfh_cg_update(surfh_handle h, void* x, void* r, void* d, const void* q, double* s, void* stream);
/* same but the residual is recomputed exactly: r = b - q_x where q_x = Q x_new is supplied by the
 * caller in a second phase (lcg's periodic refresh): phase 0: x += alpha d.  phase 1: r = b - qx;
 * rho' ; beta ; d = r + beta d. */
int surfh_cg_refresh(surfh_handle h, int32_t phase, void* x, void* r, void* d, const void* b, const void* qx,
                     double* s, void* stream);
/* s_out[0] = sum (y - hx)^2 over n elements, s_out[1] = sum (D_r x)^2 + (D_c x)^2  (criterion pieces,
 * fusion_CT.py:242-265).  hx/y may be NULL to skip the data term, x may be NULL to skip the prior. */
int surfh_criterion_terms(surfh_handle h, const void* y, const void* hx, int64_t n, const void* x, double* s_out,
                          void* stream);

/* s_out[0] = <x, b + r> on n_maps*[n_alpha,n_beta] vectors.  With r = b - Q x (the CG residual) the
 * quadratic criterion is J(x) = mu_s |y|^2 / 2 - <x, b + r> / 2: `get_crit_val` on the current iterate
 * (fusion_CT.py:242-265 as called from the lcg callback, :164-192) without the extra forward pass. */
int surfh_cg_dot_x_b_plus_r(surfh_handle h, const void* x, const void* b, const void* r, double* s_out, void* stream);

/* y += s[idx] * x on n reals, the scalar read on the device (e.g. idx = 2: the step length alpha of the last
 * surfh_cg_update).  Keeps H x_k = H x_{k-1} + alpha H d alongside the iterate, so the criterion of
 * fusion_CT.py:242-265 is evaluated on the running iterate from (y, H x_k, x_k) without applying H again
 * (H d is the y_scratch of surfh_fwadj). */
int surfh_axpy_device_scalar(surfh_handle h, void* y, const void* x, int64_t n, const double* s, int32_t idx,
                             void* stream);

/* ---- Fourier-domain block preconditioner ------------------------------ */
/* Builds, per spatial frequency f, P(f) = ( mu_s * sum_l w_l |OTF_l(f)|^2 T_l T_l^T + mu_r * d(f)^p I )^-1
 * (K x K, p = 1: separated gradients, p = 2: joint), d = eigenvalue of the circular 5-point Laplacian.
 * Replaces the per-frequency Hessian of `Model_WCT` (surfh/Models/mixing.py:131-207, `hess_spec_freq`) and its
 * bin-wise inverse `Inv_Regul_Fusion_Model3` (surfh/ToolsDir/fusion_mixing.py:401-438,
 * algorithms.make_iHtH_spectro), used here as the `precond` of qmm.lcg.  w_lambda: [host] [n_lambda] mean gain
 * of the detector sampling at each cube wavelength (0 where no band observes).  Synchronises. */
int surfh_precond_build(surfh_handle h, const double* w_lambda, double mu_s, double mu_r, int32_t joint);
/* z = P r on n_maps*[n_alpha,n_beta] device vectors: K-map FFT, per-bin K x K product, K-map inverse FFT */
int surfh_precond_apply(surfh_handle h, const void* r, void* z, void* stream);
/* Preconditioned lcg iteration, scalars as for surfh_cg_*: s[0] holds rho_z = <r, z> between calls, s[5] too.
 * phase 0: alpha = s[0]/s[1]; x += alpha d; r -= alpha q; <r,r> appended to the history.
 * phase 1: r = b - q with q = Q x (exact residual refresh), <r,r> appended. */
int surfh_pcg_update(surfh_handle h, int32_t phase, void* x, void* r, const void* d, const void* q, const void* b, double* s,
                     void* stream);
/* rho_z' = <r, z>; beta = rho_z'/rho_z (0 when first != 0); d = z + beta d */
int surfh_pcg_direction(surfh_handle h, const void* r, const void* z, void* d, double* s, int32_t first, void* stream);

/* ---- distortion-correction pre-processing -------------------------------------------------- */
/* Exponential modified-Shepard interpolation of n_in irregular samples (alpha, lambda, value) onto n_out grid
 * points (float32, [device] pointers): out = sum w v / sum w, w = exp(-alpha * d^p) for d <= pixel_cutoff,
 * d = pixel distance + epsilon; 0 where no sample is within the cutoff.  Replaces
 * surfh/ToolsDir/shepard_interpolation.pyx:77-141 (`exponential_modified_shepard`) as called by
 * surfh/Preprocessing/distorsion_correction.py:55-98.  Errors via surfh_last_error(NULL). */
int surfh_shepard(const float* alpha_coord, const float* lambda_coord, const float* values, int32_t n_in,
                  const float* alpha_mesh, const float* lambda_mesh, int32_t n_out, float p, float alpha,
                  float pixel_cutoff, float alpha_res, float lambda_res, float epsilon, float* out, void* stream);

/* ---- instrumentation --------------------------------------------------------------------- */
/* number of kernels (own + cuFFT exec calls) this handle has launched since creation */
int64_t surfh_launch_count(surfh_handle h);
int64_t surfh_own_launch_count(surfh_handle h);
/* how this handle evaluates the spectral response (wblur_subSampling / wblur_t, surfh/ToolsDir/jax_utils.py:72-91):
 * *mode 2 = int8-sliced error-free product on tcgen05 tensor cores with *digits int8 digits per operand (default:
 * 8 digits in fp64, 4 in fp32), 1 = FP64 DMMA fed by TMA, 0 = mma.sync kernels of round 1 (DMMA / 3xTF32), 3 = FFMA,
 * -1 = no band has a spectral response (beta-sum bands only).
 * *executed_fraction (may be NULL): share of the S (S + 1) / 2 digit products per tile and k-block that is actually run --
 * the leading digits of the line-spread function vanish outside a band around its peak and those all-zero tiles are
 * skipped (SURFH_OZAKI_DENSE=1 runs them all); 1 for the other modes.
 * Selected per process by SURFH_F64_GEMM / SURFH_F32_GEMM / SURFH_OZAKI_DIGITS when the handle is created. */
int surfh_contraction_info(surfh_handle h, int32_t* mode, int32_t* digits, double* executed_fraction);
/* per-stage CUDA-event timing of the calls made while enabled: enable, run, then read.
 * Output arrays of capacity `cap` (any may be NULL): stage name ("chirpz_*" = hand-written FFT passes,
 * "cufft_*" = library FFT), summed milliseconds, algorithmic bytes, flops and kernel launches.
 * Returns the number of stages filled (synchronises the device). */
int surfh_profile_enable(surfh_handle h, int32_t on);
int surfh_profile_read(surfh_handle h, int32_t cap, const char** names, float* ms, double* bytes, double* flops,
                       int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* SURFH_B200_H */

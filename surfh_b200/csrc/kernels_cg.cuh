// k5: fused conjugate-gradient vector work.  Every scalar (rho, <d,Qd>, alpha, beta, the
// gradient-norm history, the iteration counter) lives in a small device array so that a whole
// solve is enqueued without a single host synchronisation.
//
// Replaces the numpy vector algebra of qmm.lcg as driven by
// surfh/Simulation/fusion_CT.py:194-232 and the circular first-difference regulariser
// NpDiff_r / NpDiff_c (fusion_CT.py:16-43), whose normal operator is the 5-point stencil
//     (D_r^T D_r + D_c^T D_c) x = 4 x - x[i-1] - x[i+1] - x[j-1] - x[j+1]   (circular).
//
// Reductions are deterministic: per-thread partial -> warp shuffle -> fixed-order block sum ->
// per-block partial in global memory -> the LAST block to finish (ticket counter) sums the
// partials in index order.  Vectors are K*N^2 elements (~1.5 M at most): L2-resident, so each
// kernel is a single sweep whose cost is launch latency, hence the fusion.
#pragma once
#include "common.cuh"

namespace surfh {

constexpr int kCgThreads = 256;
constexpr int kCgMaxBlocks = 1024;

struct CgScratch {
    double* partial;         // [2 * kCgMaxBlocks]
    unsigned int* ticket;    // [1], zero between kernels
};

// Last-block finalisation helper: returns true in thread 0 of the last block, with `total`
// holding the ordered sum of partial[0 .. gridDim.x).
__device__ __forceinline__ bool finish_reduction(double block_value, double* partial, unsigned int* ticket,
                                                 double* smem, double& total) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_value;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // ordered: thread j sums partial[j], partial[j + T], ... then the fixed-order block sum
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += partial[i];
    total = block_sum(v, smem);
    if (threadIdx.x == 0) *ticket = 0u;
    return threadIdx.x == 0;
}

// Step length rho / <d,Qd>.  A zero curvature (x0 already the solution: r = d = 0) or a zero residual would
// give 0/0: the step is then 0, x stays put and the gradient-norm history records 0, which the host's
// stopping test sees.
__device__ __forceinline__ double cg_alpha(const double* s) {
    const double rho = s[0], curv = s[1];
    return (curv > 0.0 && rho > 0.0) ? rho / curv : 0.0;
}

// q = mu_s * q + mu_r * stencil(d);  s[1] = <d, q>
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
cg_regularise_dot_kernel(const T* __restrict__ d, T* __restrict__ q, int n_maps, int na, int nb, double mu_s,
                         double mu_r, double* __restrict__ s, CgScratch sc) {
    __shared__ double smem[32];
    const size_t npix = (size_t)na * nb, n = npix * n_maps;
    double part = 0.0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t k = idx / npix, p = idx - k * npix;
        const int i = (int)(p / nb), j = (int)(p - (size_t)i * nb);
        const T* m = d + k * npix;
        const int im = i == 0 ? na - 1 : i - 1, ip = i == na - 1 ? 0 : i + 1;
        const int jm = j == 0 ? nb - 1 : j - 1, jp = j == nb - 1 ? 0 : j + 1;
        const double c = (double)m[p];
        const double lap = 4.0 * c - (double)m[(size_t)im * nb + j] - (double)m[(size_t)ip * nb + j] -
                           (double)m[(size_t)i * nb + jm] - (double)m[(size_t)i * nb + jp];
        const double v = mu_s * (double)q[idx] + mu_r * lap;
        q[idx] = (T)v;
        part += c * (double)(T)v;
    }
    const double b = block_sum(part, smem);
    double total;
    if (finish_reduction(b, sc.partial, sc.ticket, smem, total)) s[1] = total;
}

// out = a * out + b * L(x),  L = circular 5-point Laplacian of every map (the centred impulse response
// udft.laplacian(2) that Difference_Operator_Joint applies in Fourier space, fusion_CT.py:45-63;
// also D_r^T D_r + D_c^T D_c).  a == 0 never reads `out`.
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
laplacian_axpby_kernel(const T* __restrict__ x, T* __restrict__ out, int n_maps, int na, int nb, double a, double b) {
    const size_t npix = (size_t)na * nb, n = npix * n_maps;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t k = idx / npix, p = idx - k * npix;
        const int i = (int)(p / nb), j = (int)(p - (size_t)i * nb);
        const T* m = x + k * npix;
        const int im = i == 0 ? na - 1 : i - 1, ip = i == na - 1 ? 0 : i + 1;
        const int jm = j == 0 ? nb - 1 : j - 1, jp = j == nb - 1 ? 0 : j + 1;
        const double lap = 4.0 * (double)m[p] - (double)m[(size_t)im * nb + j] - (double)m[(size_t)ip * nb + j] -
                           (double)m[(size_t)i * nb + jm] - (double)m[(size_t)i * nb + jp];
        out[idx] = (T)((a == 0.0 ? 0.0 : a * (double)out[idx]) + b * lap);
    }
}

// r = b - q ; d = r ; s[0] = <r,r> ; history[0] = s[0] ; s[4] = 0
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
cg_start_kernel(const T* __restrict__ b, const T* __restrict__ q, T* __restrict__ r, T* __restrict__ d, size_t n,
                double* __restrict__ s, int nscal, CgScratch sc) {
    __shared__ double smem[32];
    double part = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const T v = (T)((double)b[i] - (double)q[i]);
        r[i] = v;
        d[i] = v;
        part += (double)v * (double)v;
    }
    const double bs = block_sum(part, smem);
    double total;
    if (finish_reduction(bs, sc.partial, sc.ticket, smem, total)) {
        s[0] = total;
        s[4] = 0.0;
        s[nscal] = total;
    }
}

// alpha = s[0]/s[1]; x += alpha d; r -= alpha q (or r = b - qx when REFRESH); rho' = <r,r>;
// last block: s[2]=alpha, s[3]=beta=rho'/rho, s[0]=rho', history, counter.
template <typename T, bool REFRESH>
__global__ void __launch_bounds__(kCgThreads)
cg_step_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ d, const T* __restrict__ q,
               const T* __restrict__ b, size_t n, double* __restrict__ s, int nscal, CgScratch sc) {
    __shared__ double smem[32];
    const double rho = s[0];
    const double alpha = cg_alpha(s);
    double part = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T rv;
        if (REFRESH) {
            rv = (T)((double)b[i] - (double)q[i]);  // q holds Q x_new here
        } else {
            x[i] = (T)((double)x[i] + alpha * (double)d[i]);
            rv = (T)((double)r[i] - alpha * (double)q[i]);
        }
        r[i] = rv;
        part += (double)rv * (double)rv;
    }
    const double bs = block_sum(part, smem);
    double total;
    if (finish_reduction(bs, sc.partial, sc.ticket, smem, total)) {
        const int it = (int)s[4] + 1;
        if (!REFRESH) s[2] = alpha;
        s[3] = rho > 0.0 ? total / rho : 0.0;
        s[0] = total;
        s[4] = (double)it;
        s[nscal + it] = total;
    }
}

// x += alpha d only (first phase of a refresh iteration)
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
cg_axpy_alpha_kernel(T* __restrict__ x, const T* __restrict__ d, size_t n, double* __restrict__ s) {
    const double alpha = cg_alpha(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) s[2] = alpha;  // nobody reads s[2] in this kernel
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = (T)((double)x[i] + alpha * (double)d[i]);
}

// y += s[idx] * x on n elements (the detector-space companion of x += alpha d: H x_k kept by linearity)
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
axpy_device_scalar_kernel(T* __restrict__ y, const T* __restrict__ x, size_t n, const double* __restrict__ s, int idx) {
    const double a = s[idx];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (T)((double)y[i] + a * (double)x[i]);
}

// Preconditioned CG: rho_z' = <r, z> (z = P r);  last block: beta = rho_z' / rho_z (s[5], 0 on the first call:
// FIRST), s[3] = beta, s[5] = rho_z', and s[0] = rho_z' so that the next step length of the plain-CG update
// kernels, s[0] / s[1], is rho_z / <d, Q d>.
template <typename T, bool FIRST>
__global__ void __launch_bounds__(kCgThreads)
pcg_dot_kernel(const T* __restrict__ r, const T* __restrict__ z, size_t n, double* __restrict__ s, CgScratch sc) {
    __shared__ double smem[32];
    double part = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        part += (double)r[i] * (double)z[i];
    const double bs = block_sum(part, smem);
    double total;
    if (finish_reduction(bs, sc.partial, sc.ticket, smem, total)) {
        const double old = FIRST ? 0.0 : s[5];
        s[3] = old > 0.0 ? total / old : 0.0;
        s[5] = total;
        s[0] = total;
    }
}

// d = r + beta d
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
cg_direction_kernel(const T* __restrict__ r, T* __restrict__ d, size_t n, const double* __restrict__ s) {
    const double beta = s[3];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = (T)((double)r[i] + beta * (double)d[i]);
}

// out[0] = <x, b + r>: with r = b - Q x the quadratic criterion is J(x) = c - <x, b + r> / 2,
// c = mu_s |y|^2 / 2, so the CG state gives J(x_k) without another forward pass.
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
cg_dot_x_b_plus_r_kernel(const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ r, size_t n,
                         double* __restrict__ out, CgScratch sc) {
    __shared__ double smem[32];
    double part = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        part += (double)x[i] * ((double)b[i] + (double)r[i]);
    const double bs = block_sum(part, smem);
    double total;
    if (finish_reduction(bs, sc.partial, sc.ticket, smem, total)) out[0] = total;
}

// out[0] = sum (y - hx)^2 ; out[1] = sum (D_r x)^2 + (D_c x)^2
template <typename T>
__global__ void __launch_bounds__(kCgThreads)
criterion_kernel(const T* __restrict__ y, const T* __restrict__ hx, size_t n, const T* __restrict__ x, int n_maps,
                 int na, int nb, double* __restrict__ out, CgScratch sc) {
    __shared__ double smem[32];
    double p0 = 0.0, p1 = 0.0;
    if (y != nullptr && hx != nullptr) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const double e = (double)y[i] - (double)hx[i];
            p0 += e * e;
        }
    }
    if (x != nullptr) {
        const size_t npix = (size_t)na * nb, nx = npix * n_maps;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nx;
             idx += (size_t)gridDim.x * blockDim.x) {
            const size_t k = idx / npix, p = idx - k * npix;
            const int i = (int)(p / nb), j = (int)(p - (size_t)i * nb);
            const T* m = x + k * npix;
            const int im = i == 0 ? na - 1 : i - 1, jm = j == 0 ? nb - 1 : j - 1;
            const double c = (double)m[p];
            const double dr = (double)m[(size_t)im * nb + j] - c, dc = (double)m[(size_t)i * nb + jm] - c;
            p1 += dr * dr + dc * dc;
        }
    }
    const double b0 = block_sum(p0, smem);
    const double b1 = block_sum(p1, smem);
    // two reductions share the ticket: finish the first with the second riding in partial[+max]
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        sc.partial[blockIdx.x] = b0;
        sc.partial[kCgMaxBlocks + blockIdx.x] = b1;
        __threadfence();
        is_last = (atomicAdd(sc.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double v0 = 0.0, v1 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        v0 += sc.partial[i];
        v1 += sc.partial[kCgMaxBlocks + i];
    }
    v0 = block_sum(v0, smem);
    v1 = block_sum(v1, smem);
    if (threadIdx.x == 0) {
        out[0] = v0;
        out[1] = v1;
        *sc.ticket = 0u;
    }
}

}  // namespace surfh

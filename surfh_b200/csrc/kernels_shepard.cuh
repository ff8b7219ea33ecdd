// kS: exponential modified-Shepard interpolation of one slit's irregular detector samples onto the regular
// (alpha, lambda) grid of the model -- the core of the distortion-correction pre-processing.
//
// Replaces surfh/ToolsDir/shepard_interpolation.pyx:77-141 (`exponential_modified_shepard`, single-threaded
// Cython, float32) as called by surfh/Preprocessing/distorsion_correction.py:55-98, 106-181:
//     out[g] = sum_k w_k v_k / sum_k w_k ,  w_k = exp(-alpha * d_k^p)  for d_k <= cutoff,
//     d_k = sqrt(((a_k - A_g) / alpha_res)^2 + ((l_k - L_g) / lambda_res)^2) + epsilon        (pixel units)
// and 0 where no sample lies within the cutoff.  Same float32 arithmetic, same summation order (k ascending)
// per output point; the samples are streamed through shared memory in tiles so that every sample is read from
// HBM once per CTA.  One thread per output point: ~20 k points x ~25 k samples per slit = 5e8 pair tests, FP32.
#pragma once
#include "common.cuh"

namespace surfh {

constexpr int kShepardTile = 1024;

__global__ void __launch_bounds__(256)
shepard_kernel(const float* __restrict__ a_in, const float* __restrict__ l_in, const float* __restrict__ v_in, int n_in,
               const float* __restrict__ a_mesh, const float* __restrict__ l_mesh, int n_out, float p, float alpha,
               float cutoff, float inv_ares, float inv_lres, float eps, float* __restrict__ out) {
    __shared__ float sa[kShepardTile], sl[kShepardTile], sv[kShepardTile];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = g < n_out;
    const float qa = live ? a_mesh[g] : 0.f, ql = live ? l_mesh[g] : 0.f;
    float num = 0.f, den = 0.f;
    for (int base = 0; base < n_in; base += kShepardTile) {
        const int n = min(kShepardTile, n_in - base);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            sa[i] = a_in[base + i];
            sl[i] = l_in[base + i];
            sv[i] = v_in[base + i];
        }
        __syncthreads();
        if (live) {
            for (int k = 0; k < n; ++k) {
                const float d1 = (sa[k] - qa) * inv_ares, d2 = (sl[k] - ql) * inv_lres;
                const float dist = sqrtf(d1 * d1 + d2 * d2) + eps;
                if (dist <= cutoff) {
                    const float w = expf(-alpha * (p == 2.f ? dist * dist : powf(dist, p)));
                    num += w * sv[k];
                    den += w;
                }
            }
        }
    }
    if (live) out[g] = den != 0.f ? num / den : 0.f;
}

}  // namespace surfh

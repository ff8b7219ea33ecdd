// Shared helpers for the surfh_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace surfh {

constexpr int kMaxTemplates = 8;

template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };
template <typename T> using cplx_t = typename Cplx<T>::type;

template <typename T> __device__ __forceinline__ cplx_t<T> make_c(T re, T im);
template <> __device__ __forceinline__ float2 make_c<float>(float re, float im) { return make_float2(re, im); }
template <> __device__ __forceinline__ double2 make_c<double>(double re, double im) { return make_double2(re, im); }

// a * b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// conj(a) * b
template <typename C> __device__ __forceinline__ C cmul_conj(C a, C b) {
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.x * b.y - a.y * b.x;
    return r;
}

// Streaming (read-once) 128-bit / 64-bit loads: bypass L1 allocation, keep L2 default policy.
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum of one double per thread (blockDim.x multiple of 32, <= 1024).
// Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 doubles */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Ampere-style asynchronous global -> shared copies (LDGSTS): no register staging, many in flight.
template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async moves 4, 8 or 16 bytes");
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- Hopper/Blackwell bulk asynchronous copy (TMA, non-tensor form) + mbarrier -----------------------------
// One thread moves a whole contiguous run global -> shared with a single instruction; completion is signalled
// on an mbarrier in shared memory by transaction bytes.  Source, destination and size must be multiples of 16.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace surfh

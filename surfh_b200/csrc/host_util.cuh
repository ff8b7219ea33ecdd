// Host-side helpers shared by the C ABI translation unit: error type, checking macros, device buffers.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/surfh_b200.h"

namespace surfh {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SURFH_CUDA(expr)                                                                                 \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            throw Error(SURFH_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                \
    } while (0)
#define SURFH_FFT(expr)                                                                                  \
    do {                                                                                                 \
        cufftResult r_ = (expr);                                                                         \
        if (r_ != CUFFT_SUCCESS) throw Error(SURFH_ECUFFT, std::string(#expr) + ": cufft error " + std::to_string((int)r_)); \
    } while (0)
#define SURFH_REQUIRE(cond, msg)                                                                         \
    do {                                                                                                 \
        if (!(cond)) throw Error(SURFH_EINVAL, std::string(msg));                                        \
    } while (0)

static thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    void alloc(size_t n) {
        release();
        if (n == 0) return;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) throw Error(SURFH_ENOMEM, "cudaMalloc(" + std::to_string(n) + " bytes): " + cudaGetErrorString(e));
        bytes = n;
    }
    void ensure(size_t n) {
        if (bytes < n) alloc(n);
    }
    void ensure_zeroed(size_t n) {  // zero-filled on (re)allocation only
        if (bytes < n) {
            alloc(n);
            if (cudaMemset(p, 0, n) != cudaSuccess) throw Error(SURFH_ECUDA, "cudaMemset failed");
        }
    }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

template <typename U, typename V> static void upload_converted(DevBuf& dst, const V* src, size_t n) {
    std::vector<U> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = (U)src[i];
    dst.alloc(n * sizeof(U));
    SURFH_CUDA(cudaMemcpy(dst.p, tmp.data(), n * sizeof(U), cudaMemcpyHostToDevice));
}

}  // namespace surfh

// Host side of the hand-written 2-D real FFT pair (kernels_fft.cuh): table construction and launches.
//
// Replaces the cuFFT plans for the reference's `rfftn` / `irfftn(norm="ortho")` calls
// (surfh/ToolsDir/jax_utils.py:30-41, python_utils.py:41-71) whenever both map axes are <= 512
// pixels (every size the reference uses: 251, 301, 501); the transforms are un-normalised like cuFFT's.
#pragma once
#include <cmath>
#include <vector>

#include "host_util.cuh"
#include "kernels_fft.cuh"

namespace surfh {

// Dispatch on the chirp-z length (a power of two in [256, 1024]).
#define SURFH_DISPATCH_M(m, ...)                                       \
    switch (m) {                                                       \
        case 256: { constexpr int MM = 256; __VA_ARGS__; } break;      \
        case 512: { constexpr int MM = 512; __VA_ARGS__; } break;      \
        case 1024: { constexpr int MM = 1024; __VA_ARGS__; } break;    \
        default: throw Error(SURFH_EINVAL, "unsupported chirp-z length"); \
    }

template <typename S, typename D>
__global__ void fft_convert_kernel(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (D)src[i];
}

// Tables of one axis length.
template <typename T> struct FftAxis {
    using C = cplx_t<T>;
    int n = 0, m = 0;
    DevBuf chirp, filt, tw;

    // smallest supported chirp-z length for n points: n <= m/2 (so 2n-1 <= m); 0 when n > 512
    static int pick_m(int n) {
        for (int m = 256; m <= 1024; m *= 2)
            if (n <= m / 2) return m;
        return 0;
    }

    template <int M, typename Pass> static void set_pass_attr() {
        SURFH_CUDA(cudaFuncSetAttribute(fft_pass_kernel<T, M, Pass>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)FftK<T, M>::SMEM_BYTES));
    }
    template <int M> static void set_smem_attr() {
        set_pass_attr<M, RowsR2C<T, M, false>>();
        set_pass_attr<M, RowsR2C<T, M, true>>();
        set_pass_attr<M, ColsPass<T, M, false, false>>();
        set_pass_attr<M, ColsPass<T, M, true, false>>();
        set_pass_attr<M, ColsPass<T, M, false, true>>();
        set_pass_attr<M, ColsPass<T, M, true, true>>();
        set_pass_attr<M, RowsC2R<T, M>>();
    }

    void init(int n_) {
        n = n_;
        m = pick_m(n);
        if (m == 0) throw Error(SURFH_EINVAL, "axis too long for the hand-written FFT (max 512)");
        const double pi = 3.14159265358979323846;
        // everything is evaluated in double, the filter spectrum with the double instantiation of the
        // very FFT code that consumes it, and only then rounded to T
        const int pts = m / 32, r3 = 32 / pts, ntw = m + pts * r3;
        std::vector<double2> h_tw(ntw), h_chirp(n), h_b(m);
        auto root = [&](long long e) {  // exp(-2 pi i e / m)
            const double a = -2.0 * pi * (double)(e % m) / (double)m;
            return make_double2(std::cos(a), std::sin(a));
        };
        for (int k1 = 0; k1 < pts; ++k1)
            for (int t = 0; t < 32; ++t) h_tw[k1 * 32 + t] = root((long long)k1 * t);
        for (int q2 = 0; q2 < pts; ++q2)
            for (int n2 = 0; n2 < r3; ++n2) h_tw[m + q2 * r3 + n2] = root((long long)pts * n2 * q2);  // exp(-2 pi i n2 q2 / 32)
        for (int j = 0; j < m; ++j) h_b[j] = make_double2(0.0, 0.0);
        for (int j = 0; j < n; ++j) {
            const long long q = ((long long)j * j) % (2ll * n);  // exp(-i pi j^2 / n) has period 2n in j^2
            const double a = -pi * (double)q / (double)n;
            h_chirp[j] = make_double2(std::cos(a), std::sin(a));
            const double2 bj = make_double2(std::cos(a) / m, -std::sin(a) / m);  // conj(chirp) / M
            h_b[j] = bj;
            if (j) h_b[m - j] = bj;
        }
        DevBuf d_tw, d_b, d_filt;
        d_tw.alloc(ntw * sizeof(double2));
        d_b.alloc(m * sizeof(double2));
        d_filt.alloc(m * sizeof(double2));
        SURFH_CUDA(cudaMemcpy(d_tw.p, h_tw.data(), ntw * sizeof(double2), cudaMemcpyHostToDevice));
        SURFH_CUDA(cudaMemcpy(d_b.p, h_b.data(), m * sizeof(double2), cudaMemcpyHostToDevice));
        SURFH_DISPATCH_M(m, {
            using KD = FftK<double, MM>;
            const int bytes = (int)((KD::OFF_BUF + KD::BUF) * sizeof(double2));
            SURFH_CUDA(cudaFuncSetAttribute(fft_filter_kernel<double, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            FftPlan1d<double> pd;
            pd.chirp = nullptr; pd.filt = nullptr; pd.tw = d_tw.as<double2>(); pd.n = n;
            fft_filter_kernel<double, MM><<<1, 32, bytes>>>(d_b.as<double2>(), pd, d_filt.as<double2>());
            set_smem_attr<MM>();
        });
        SURFH_CUDA(cudaGetLastError());
        DevBuf d_chirp;
        d_chirp.alloc(n * sizeof(double2));
        SURFH_CUDA(cudaMemcpy(d_chirp.p, h_chirp.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
        chirp.alloc(n * sizeof(C));
        filt.alloc(m * sizeof(C));
        tw.alloc(ntw * sizeof(C));
        fft_convert_kernel<double, T><<<ceil_div(2 * n, 256), 256>>>(d_chirp.as<double>(), chirp.as<T>(), (size_t)2 * n);
        fft_convert_kernel<double, T><<<ceil_div(2 * m, 256), 256>>>(d_filt.as<double>(), filt.as<T>(), (size_t)2 * m);
        fft_convert_kernel<double, T><<<ceil_div(2 * ntw, 256), 256>>>(d_tw.as<double>(), tw.as<T>(), (size_t)2 * ntw);
        SURFH_CUDA(cudaGetLastError());
        SURFH_CUDA(cudaDeviceSynchronize());
    }

    FftPlan1d<T> plan() const {
        FftPlan1d<T> p;
        p.chirp = chirp.as<C>();
        p.filt = filt.as<C>();
        p.tw = tw.as<C>();
        p.n = n;
        return p;
    }
};

// Batched 2-D real <-> half-complex transforms of [na][nb] planes.
template <typename T> struct OwnFft2d {
    using C = cplx_t<T>;
    int na = 0, nb = 0, nh = 0;
    FftAxis<T> axis_a, axis_b_store;
    const FftAxis<T>* axis_b = nullptr;
    bool ready = false;
    int n_sm = 148;

    static bool supported(int na, int nb) { return FftAxis<T>::pick_m(na) && FftAxis<T>::pick_m(nb) && na > 1 && nb > 1; }

    void init(int na_, int nb_) {
        na = na_; nb = nb_; nh = nb / 2 + 1;
        int dev = 0;
        SURFH_CUDA(cudaGetDevice(&dev));
        SURFH_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        axis_a.init(na);
        if (nb == na) axis_b = &axis_a;
        else { axis_b_store.init(nb); axis_b = &axis_b_store; }
        ready = true;
    }

    // complex elements per plane of the intermediate buffer
    int ypitch() const { return (na + 1) / 2 * 2; }
    size_t z_plane() const { return std::max((size_t)((na + 1) / 2) * nb, (size_t)nh * ypitch()); }

    FftShape shape(size_t real_plane, size_t spec_plane, int batch, const int2* pair_range) const {
        FftShape s;
        s.na = na; s.nb = nb; s.nh = nh; s.npair = (na + 1) / 2;
        s.real_plane = real_plane; s.spec_plane = spec_plane; s.z_plane = z_plane(); s.ypitch = ypitch(); s.batch = batch;
        s.pair_range = pair_range;
        if (pair_range && batch > FftK<T, 256>::MAX_PLANES) throw Error(SURFH_EINVAL, "pruned FFT launch: too many planes");
        return s;
    }

    // persistent launch: one CTA per SM, each of its G warps walking its own items with a grid stride
    template <int MM, typename Pass> void launch(const Pass& pass, long long n_items, const FftAxis<T>& ax, cudaStream_t st) const {
        using K = FftK<T, MM>;
        if (n_items >= (1ll << 31)) throw Error(SURFH_EINVAL, "FFT batch too large");
        const long long ctas = (n_items + K::G - 1) / K::G;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ctas, n_sm));
        fft_pass_kernel<T, MM, Pass><<<grid, K::NT, K::SMEM_BYTES, st>>>(pass, ax.plan());
    }

    // in: real [batch] planes (stride real_plane) -> spec (stride spec_plane): [na][nh], or [nh][na] when
    // `transposed`; z: scratch.
    // pair_range: [device] per-plane (first row pair, count) outside of which the input rows are zero, or NULL;
    // n_pairs: sum of the counts (host copy), ignored without pair_range
    // row_slack: the caller's real buffer has at least one row of readable slack after its last plane (the
    //             operator's working cube), so an unpaired last row may be staged as a full pair
    void r2c(const T* in, size_t real_plane, C* spec, size_t spec_plane, C* z, int batch, cudaStream_t st,
             bool transposed, const int2* pair_range = nullptr, long long n_pairs = 0, bool row_slack = false) const {
        const FftShape s = shape(real_plane, spec_plane, batch, pair_range);
        const long long row_items = pair_range ? n_pairs : (long long)batch * s.npair;
        // one TMA bulk copy per row pair when every pair is 16-byte aligned and a multiple of 16 bytes long
        const bool aligned = sizeof(T) == 8 && real_plane % 2 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 &&
                             (na % 2 == 0 || row_slack);
        SURFH_DISPATCH_M(axis_b->m, {
            if (aligned) {
                RowsR2C<T, MM, true> pass{in, z, s};
                launch<MM>(pass, row_items, *axis_b, st);
            } else {
                RowsR2C<T, MM, false> pass{in, z, s};
                launch<MM>(pass, row_items, *axis_b, st);
            }
        });
        SURFH_DISPATCH_M(axis_a.m, {
            if (transposed) {
                ColsPass<T, MM, false, true> pass{z, spec, s};
                launch<MM>(pass, (long long)batch * nh, axis_a, st);
            } else {
                ColsPass<T, MM, false, false> pass{z, spec, s};
                launch<MM>(pass, (long long)batch * nh, axis_a, st);
            }
        });
        SURFH_CUDA(cudaGetLastError());
    }

    // spec -> out: real planes; z: scratch
    // pair_range: per-plane row pairs of the output that are wanted (the others are left untouched), or NULL
    void c2r(const C* spec, size_t spec_plane, T* out, size_t real_plane, C* z, int batch, cudaStream_t st,
             bool transposed, const int2* pair_range = nullptr, long long n_pairs = 0) const {
        const FftShape s = shape(real_plane, spec_plane, batch, pair_range);
        const long long row_items = pair_range ? n_pairs : (long long)batch * s.npair;
        SURFH_DISPATCH_M(axis_a.m, {
            if (transposed) {
                ColsPass<T, MM, true, true> pass{spec, z, s};
                launch<MM>(pass, (long long)batch * nh, axis_a, st);
            } else {
                ColsPass<T, MM, true, false> pass{spec, z, s};
                launch<MM>(pass, (long long)batch * nh, axis_a, st);
            }
        });
        SURFH_DISPATCH_M(axis_b->m, {
            RowsC2R<T, MM> pass{z, out, s};
            launch<MM>(pass, row_items, *axis_b, st);
        });
        SURFH_CUDA(cudaGetLastError());
    }
};

}  // namespace surfh

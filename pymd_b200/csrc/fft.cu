// See fft.cuh.  Replaces numpy.fft in myfft.Fourier1D / iFourier1D (reference myfft.py:15-52).
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "fft.cuh"

namespace pymd {

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

template <int R>
__device__ __forceinline__ void butterfly(double2* v);

template <>
__device__ __forceinline__ void butterfly<2>(double2* v) {
    const double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(double2* v) {
    const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    const double2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<8>(double2* v) {
    const double h = 0.70710678118654752440;
    double2 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        a[r] = cadd(v[r], v[r + 4]);
        b[r] = csub(v[r], v[r + 4]);
    }
    b[1] = make_double2(h * (b[1].x + b[1].y), h * (b[1].y - b[1].x));     // * (1-i)/sqrt2
    b[2] = mul_mi(b[2]);                                                   // * (-i)
    b[3] = make_double2(h * (b[3].y - b[3].x), -h * (b[3].x + b[3].y));    // * (-1-i)/sqrt2
    butterfly<4>(a);
    butterfly<4>(b);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[2 * q] = a[q];
        v[2 * q + 1] = b[q];
    }
}

template <int R>
__global__ void __launch_bounds__(128) fft_pass_kernel(const void* __restrict__ src, void* __restrict__ dst,
                                                       int n, int Ns, long long cols, int conj_in, int conj_out,
                                                       const double2* __restrict__ tw, int twN, int loadmode,
                                                       int storemode, long long src_ld, long long dst_ld,
                                                       double scale) {
    const int j = blockIdx.x;
    const int k = j % Ns;
    const int base = (j / Ns) * Ns * R + k;
    const int nr = n / R;
    double2 w[R];
    if (Ns > 1) {
        const int step = twN / (Ns * R);
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = tw[(size_t)r * k * step];
    }
    for (long long col = blockIdx.y * (long long)blockDim.x + threadIdx.x; col < cols;
         col += (long long)gridDim.y * blockDim.x) {
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long idx = j + (long long)r * nr;
            if (loadmode == LOAD_C) {
                v[r] = static_cast<const double2*>(src)[idx * src_ld + col];
            } else {
                const double* xr = static_cast<const double*>(src);
                v[r] = make_double2(xr[(2 * idx) * src_ld + col], xr[(2 * idx + 1) * src_ld + col]);
            }
            if (conj_in) v[r].y = -v[r].y;
        }
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        }
        butterfly<R>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long o = base + (long long)r * Ns;
            double2 out = make_double2(scale * v[r].x, (conj_out ? -scale : scale) * v[r].y);
            if (storemode == STORE_C) {
                static_cast<double2*>(dst)[o * dst_ld + col] = out;
            } else {
                double* dr = static_cast<double*>(dst);
                dr[(2 * o) * dst_ld + col] = out.x;
                dr[(2 * o + 1) * dst_ld + col] = out.y;
            }
        }
    }
}

struct TwKey {
    int dev, N;
    bool operator<(const TwKey& o) const { return dev < o.dev || (dev == o.dev && N < o.N); }
};
std::map<TwKey, double2*> g_tw;
std::mutex g_tw_mu;

}  // namespace

const double2* twiddle_table(int N) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto it = g_tw.find({dev, N});
    if (it != g_tw.end()) return it->second;
    std::vector<double2> h(N);
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int t = 0; t < N; ++t) {
        const long double a = -two_pi * (long double)t / (long double)N;
        h[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    double2* d = nullptr;
    if (cudaMalloc(&d, sizeof(double2) * N) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), sizeof(double2) * N, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    g_tw[{dev, N}] = d;
    return d;
}

int fft_columns(const void* src, void* dst, double2* bufA, double2* bufB, const FftPlan& p,
                cudaStream_t stream, int* npasses) {
    PYMD_REQUIRE(p.n >= 2 && (p.n & (p.n - 1)) == 0, "fft: length %d is not a power of two >= 2", p.n);
    PYMD_REQUIRE(p.cols > 0, "fft: no columns");
    const double2* tw = twiddle_table(p.n);
    PYMD_REQUIRE(tw, "fft: twiddle table allocation failed");
    int radices[32], np = 0, rem = p.n;
    while (rem >= 8) { radices[np++] = 8; rem /= 8; }
    if (rem > 1) radices[np++] = rem;
    PYMD_REQUIRE(np == 1 || (bufA && (np == 2 || bufB)), "fft: work buffers required");
    int Ns = 1;
    const void* cur = src;
    double2* pp[2] = {bufA, bufB};
    for (int i = 0; i < np; ++i) {
        const int R = radices[i];
        const bool first = i == 0, last = i == np - 1;
        void* out = last ? dst : static_cast<void*>(pp[i & 1]);
        const int lm = first ? p.load : LOAD_C, sm = last ? p.store : STORE_C;
        const long long sld = first ? p.src_ld : p.cols, dld = last ? p.dst_ld : p.cols;
        const double sc = last ? p.scale : 1.0;
        long long yb = (p.cols + 127) / 128;
        if (yb > 65535) yb = 65535;
        dim3 grid(p.n / R, (unsigned)yb);
        // ifft(x) = conj(fft(conj(x))): conjugate on the first load and on the last store
        const int ci = (p.sign > 0 && first) ? 1 : 0, co = (p.sign > 0 && last) ? 1 : 0;
        switch (R) {
            case 8:
                fft_pass_kernel<8><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc);
                break;
            case 4:
                fft_pass_kernel<4><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc);
                break;
            default:
                fft_pass_kernel<2><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc);
        }
        PYMD_CUDA_CHECK(cudaGetLastError());
        cur = out;
        Ns *= R;
    }
    if (npasses) *npasses = np;
    return PYMD_OK;
}

}  // namespace pymd

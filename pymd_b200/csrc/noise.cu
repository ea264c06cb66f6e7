// Coloured-noise generation on device (reference noise.py:38-88 phnoise, 135-190 enoise,
// 252-284 vargau; myfft.py:35-52 iFourier1D) and the counter-based white-noise source.
//
// Per bath and trajectory:   X_f = L_f xi_f   (L_f = U_f diag(sqrt(max(lambda_f,0))), host-made
// exactly as the reference builds its eigen-decomposition),  Hermitian mirror with the
// reference's Nyquist/DC convention (only their real parts survive N.real),  then
// noise(t) = Re sum_j Y_j exp(-2 pi i j t/N) / (N dt).
// Because the output is real, the length-N transform is done as ONE half-length complex FFT:
//   Z_k = E_k + i O_k,  E_k = Y_k + conj(Y_{N/2-k}),  O_k = (Y_k - conj(Y_{N/2-k})) W_N^k,
//   z = FFT_{N/2}(Z),  noise(2m) = Re z_m, noise(2m+1) = Im z_m.
#include "../../include/pymd_b200.h"
#include "fft.cuh"

namespace pymd {

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t lo, uint32_t hi) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

template <int NORMAL>
__global__ void philox_kernel(double* out, long long rows, int ldb, uint32_t k0, uint32_t k1,
                              uint32_t stream_id, long long traj0) {
    const size_t total = (size_t)rows * ldb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long r = idx / ldb;
        const unsigned long long tr = (unsigned long long)(traj0 + (long long)(idx % ldb));
        uint32_t x[4];
        philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)tr, stream_id ^ (uint32_t)(tr >> 32) * 0x85EBCA6Bu,
                      k0, k1, x);
        const double u1 = u53(x[0], x[1]);
        if (NORMAL) {
            const double u2 = u53(x[2], x[3]);
            out[idx] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        } else {
            out[idx] = u1;
        }
    }
}

// ------------------------------------------------------------------ shaping + pre-processing
constexpr int SX = 32, SY = 8, JC = 64;

__global__ void __launch_bounds__(SX* SY) shape_kernel(const double* __restrict__ lre, const double* __restrict__ lim,
                                                       const double* __restrict__ xi, int N, int nc, int ldb,
                                                       int i0, int i1, const double2* __restrict__ twN,
                                                       double2* __restrict__ Z) {
    __shared__ double x1[JC][SX], x2[JC][SX];
    const int kp = blockIdx.x;              // pair (k1, k2) = (kp, N/2 - kp), kp in [0, N/4]
    const int half = N / 2;
    const int k1 = kp, k2 = half - kp;
    const int n = blockIdx.y * SX + threadIdx.x;
    const int rows = i1 - i0;
    const int tid = threadIdx.y * SX + threadIdx.x;

    // each thread owns up to MAXI output components: i = i0 + threadIdx.y + SY * q
    constexpr int MAXI = 8;
    for (int ib = i0; ib < i1; ib += SY * MAXI) {
        double a1r[MAXI], a1i[MAXI], a2r[MAXI], a2i[MAXI];
#pragma unroll
        for (int q = 0; q < MAXI; ++q) a1r[q] = a1i[q] = a2r[q] = a2i[q] = 0.0;
        for (int j0 = 0; j0 < nc; j0 += JC) {
            const int jn = min(JC, nc - j0);
            __syncthreads();
            for (int e = tid; e < jn * SX; e += SX * SY) {
                const int jj = e / SX, xx = e % SX;
                const size_t col = (size_t)blockIdx.y * SX + xx;
                x1[jj][xx] = xi[((size_t)k1 * nc + j0 + jj) * ldb + col];
                x2[jj][xx] = xi[((size_t)k2 * nc + j0 + jj) * ldb + col];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < MAXI; ++q) {
                const int i = ib + threadIdx.y + SY * q;
                if (i >= i1) break;
                const double* l1r = lre + ((size_t)k1 * nc + i) * nc + j0;
                const double* l2r = lre + ((size_t)k2 * nc + i) * nc + j0;
                double s1 = 0.0, s2 = 0.0;
                for (int jj = 0; jj < jn; ++jj) {
                    s1 += l1r[jj] * x1[jj][threadIdx.x];
                    s2 += l2r[jj] * x2[jj][threadIdx.x];
                }
                a1r[q] += s1;
                a2r[q] += s2;
                if (lim) {
                    const double* l1i = lim + ((size_t)k1 * nc + i) * nc + j0;
                    const double* l2i = lim + ((size_t)k2 * nc + i) * nc + j0;
                    double t1 = 0.0, t2 = 0.0;
                    for (int jj = 0; jj < jn; ++jj) {
                        t1 += l1i[jj] * x1[jj][threadIdx.x];
                        t2 += l2i[jj] * x2[jj][threadIdx.x];
                    }
                    a1i[q] += t1;
                    a2i[q] += t2;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const int i = ib + threadIdx.y + SY * q;
            if (i >= i1) break;
            double y1r = a1r[q], y1i = a1i[q], y2r = a2r[q], y2i = a2i[q];
            if (kp == 0) { y1i = 0.0; y2i = 0.0; }          // DC and Nyquist rows: real part only
            // Z_{k1}
            {
                const double er = y1r + y2r, ei = y1i - y2i;               // Y1 + conj(Y2)
                const double dr = y1r - y2r, di = y1i + y2i;               // Y1 - conj(Y2)
                const double2 w = twN[k1];
                const double orr = dr * w.x - di * w.y, oi = dr * w.y + di * w.x;
                Z[((size_t)k1 * rows + (i - i0)) * ldb + n] = make_double2(er - oi, ei + orr);
            }
            if (kp != 0 && k2 != k1) {
                const double er = y2r + y1r, ei = y2i - y1i;               // Y2 + conj(Y1)
                const double dr = y2r - y1r, di = y2i + y1i;               // Y2 - conj(Y1)
                const double2 w = twN[k2];
                const double orr = dr * w.x - di * w.y, oi = dr * w.y + di * w.x;
                Z[((size_t)k2 * rows + (i - i0)) * ldb + n] = make_double2(er - oi, ei + orr);
            }
        }
    }
}

int fft_num_passes(int n) {
    int np = 0;
    while (n >= 8) { ++np; n /= 8; }
    if (n > 1) ++np;
    return np;
}

}  // namespace pymd

using namespace pymd;

extern "C" {

size_t pymd_noise_work_bytes(int nmd, int nrows, int ldb) {
    return 2 * (size_t)(nmd / 2) * nrows * ldb * sizeof(double2);
}

int pymd_noise_generate(const double* lre, const double* lim, const double* xi, int nmd, int nc, int ldb,
                        int i0, int i1, double dt, double* out, void* work, void* stream) {
    PYMD_REQUIRE(lre && xi && out && work, "noise_generate: null argument");
    PYMD_REQUIRE(nmd >= 4 && (nmd & (nmd - 1)) == 0, "noise_generate: nmd=%d must be a power of two >= 4", nmd);
    PYMD_REQUIRE(nc > 0 && ldb > 0 && ldb % 32 == 0, "noise_generate: nc=%d ldb=%d", nc, ldb);
    PYMD_REQUIRE(0 <= i0 && i0 < i1 && i1 <= nc, "noise_generate: row range [%d, %d) of %d", i0, i1, nc);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int half = nmd / 2, rows = i1 - i0;
    const double2* twN = twiddle_table(nmd);
    PYMD_REQUIRE(twN, "noise_generate: twiddle table allocation failed");
    // Z goes to W0; the passes ping-pong W1, W0, W1, ...; the last one writes the real rows of `out`
    double2* W0 = static_cast<double2*>(work);
    double2* W1 = W0 + (size_t)half * rows * ldb;
    dim3 grid(half / 2 + 1, ldb / SX), blk(SX, SY);
    shape_kernel<<<grid, blk, 0, st>>>(lre, lim, xi, nmd, nc, ldb, i0, i1, twN, W0);
    PYMD_CUDA_CHECK(cudaGetLastError());
    FftPlan p{};
    p.n = half; p.cols = (long long)rows * ldb; p.sign = -1;
    p.load = LOAD_C; p.store = STORE_REAL_ROWS;
    p.src_ld = p.cols; p.dst_ld = (long long)nc * ldb;
    p.scale = 1.0 / ((double)nmd * dt);
    return fft_columns(W0, out + (size_t)i0 * ldb, W1, W0, p, st, nullptr);
}

int pymd_philox_normal(double* out, long long rows, int ldb, uint64_t seed, uint32_t stream_id,
                       long long traj0, void* stream) {
    PYMD_REQUIRE(out && rows > 0 && ldb > 0, "philox_normal: bad argument");
    philox_kernel<1><<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        out, rows, ldb, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, traj0);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

int pymd_philox_uniform(double* out, long long rows, int ldb, uint64_t seed, uint32_t stream_id,
                        long long traj0, void* stream) {
    PYMD_REQUIRE(out && rows > 0 && ldb > 0, "philox_uniform: bad argument");
    philox_kernel<0><<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        out, rows, ldb, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, traj0);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // extern "C"

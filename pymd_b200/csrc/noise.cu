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

// NORMAL = 0: one uniform per (row, trajectory), counter (row, trajectory).
// NORMAL = 1: Box-Muller with BOTH outputs: the counter (row pair r/2, trajectory) yields the normals of
// rows 2(r/2) (cosine branch) and 2(r/2)+1 (sine branch), halving the Philox / log / sqrt work per number.
template <int NORMAL>
__global__ void philox_kernel(double* out, long long rows, int ldb, uint32_t k0, uint32_t k1,
                              uint32_t stream_id, long long traj0) {
    const long long prow = NORMAL ? (rows + 1) / 2 : rows;
    const size_t total = (size_t)prow * ldb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long r = idx / ldb;
        const int n = (int)(idx % ldb);
        const unsigned long long tr = (unsigned long long)(traj0 + n);
        uint32_t x[4];
        philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)tr, stream_id ^ (uint32_t)(tr >> 32) * 0x85EBCA6Bu,
                      k0, k1, x);
        const double u1 = u53(x[0], x[1]);
        if (NORMAL) {
            const double u2 = u53(x[2], x[3]);
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            out[(size_t)(2 * r) * ldb + n] = rad * cs;
            if ((long long)(2 * r + 1) < rows) out[(size_t)(2 * r + 1) * ldb + n] = rad * sn;
        } else {
            out[idx] = u1;
        }
    }
}

// ------------------------------------------------------------------ shaping + pre-processing
constexpr int SX = 32, SY = 8, JC = 64;

__global__ void __launch_bounds__(SX* SY) shape_kernel(const double* __restrict__ lre, const double* __restrict__ lim,
                                                       const double* __restrict__ xi, int N, int nc, int ldb,
                                                       int n0, int cw, const double2* __restrict__ twN,
                                                       double2* __restrict__ Z) {
    // scalar fallback for nc > 64; all component rows, trajectories [n0, n0 + cw); Z is [N/2][nc][cw]
    __shared__ double x1[JC][SX], x2[JC][SX];
    const int i0 = 0, i1 = nc;
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
                const size_t col = (size_t)n0 + (size_t)blockIdx.y * SX + xx;
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
                Z[((size_t)k1 * rows + (i - i0)) * cw + n] = make_double2(er - oi, ei + orr);
            }
            if (kp != 0 && k2 != k1) {
                const double er = y2r + y1r, ei = y2i - y1i;               // Y2 + conj(Y1)
                const double dr = y2r - y1r, di = y2i + y1i;               // Y2 - conj(Y1)
                const double2 w = twN[k2];
                const double orr = dr * w.x - di * w.y, oi = dr * w.y + di * w.x;
                Z[((size_t)k2 * rows + (i - i0)) * cw + n] = make_double2(er - oi, ei + orr);
            }
        }
    }
}


// ------------------------------------------------------------------ shaping on the FP64 tensor pipe
// One CTA per frequency pair (k1, k2 = N/2 - k1): the four nc x nc operand planes Re/Im L_{k1},
// Re/Im L_{k2} are staged in shared memory once (row pitch = 4 mod 16 doubles: conflict-free A
// fragments) and applied to the white noise of every trajectory of the column chunk: warp w
// takes n8 tiles w, w+8, ...; the B fragments (xi rows of the two frequencies) come straight from
// global memory into registers, Y = L xi is 4 DMMA chains per m8 tile (Re/Im of the two
// frequencies), and the Hermitian pre-processing of the half-length real transform is the epilogue.
// 16 warps for the large complex operands (electron bath: 128 registers, no spills; 123 vs 128 ms for 4096
// trajectories of 60 dofs), 8 warps otherwise (the real 64-dof variant would spill at 128 registers)
template <int KKMAX, bool CPLX>
__host__ __device__ constexpr int shape_threads() { return KKMAX == 16 && CPLX ? 512 : 256; }
template <int KKMAX, bool CPLX>
__global__ void __launch_bounds__((shape_threads<KKMAX, CPLX>()), 1) shape_dmma_kernel(const double* __restrict__ lre, const double* __restrict__ lim,
                                                            const double* __restrict__ xi, int N, int nc, int ldb,
                                                            int n0, int cw, const double2* __restrict__ twN,
                                                            double2* __restrict__ Z) {
    extern __shared__ __align__(16) double sh[];
    const int half = N / 2, kp = blockIdx.x, k1 = kp, k2 = half - kp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const int MT = (nc + 7) / 8, KK = (nc + 3) / 4;
    const int LP = 4 * KK + 4 + ((4 * KK) % 16 == 0 ? 0 : 16 - (4 * KK) % 16);   // = 4 mod 16
    const int plane = 8 * MT * LP;
    double* A1r = sh;
    double* A1i = sh + plane;
    double* A2r = sh + 2 * plane;
    double* A2i = sh + 3 * plane;
    constexpr int SHAPE_THREADS = shape_threads<KKMAX, CPLX>();
    for (int e = tid; e < 8 * MT * 4 * KK; e += SHAPE_THREADS) {
        const int i = e / (4 * KK), j = e % (4 * KK);
        const bool in = i < nc && j < nc;
        const size_t o1 = ((size_t)k1 * nc + i) * nc + j, o2 = ((size_t)k2 * nc + i) * nc + j;
        A1r[i * LP + j] = in ? lre[o1] : 0.0;
        A2r[i * LP + j] = in ? lre[o2] : 0.0;
        if (CPLX) {
            A1i[i * LP + j] = in ? lim[o1] : 0.0;
            A2i[i * LP + j] = in ? lim[o2] : 0.0;
        }
    }
    __syncthreads();
    const double2 w1 = twN[k1], w2 = twN[k2];
    for (int nt = warp; nt < cw / 8; nt += SHAPE_THREADS / 32) {
        const int col = n0 + nt * 8;
        double b1[KKMAX], b2[KKMAX];
#pragma unroll
        for (int kk = 0; kk < KKMAX; ++kk) {
            const int j = 4 * kk + lc;
            const bool in = kk < KK && j < nc;
            b1[kk] = in ? xi[((size_t)k1 * nc + j) * ldb + col + lr] : 0.0;
            b2[kk] = in ? xi[((size_t)k2 * nc + j) * ldb + col + lr] : 0.0;
        }
        for (int mt = 0; mt < MT; ++mt) {
            double y1r[2] = {0.0, 0.0}, y1i[2] = {0.0, 0.0}, y2r[2] = {0.0, 0.0}, y2i[2] = {0.0, 0.0};
            const int ao = (8 * mt + lr) * LP + lc;
#pragma unroll
            for (int kk = 0; kk < KKMAX; ++kk) {
                if (kk < KK) {
                    dmma884(y1r[0], y1r[1], A1r[ao + 4 * kk], b1[kk]);
                    dmma884(y2r[0], y2r[1], A2r[ao + 4 * kk], b2[kk]);
                    if (CPLX) {
                        dmma884(y1i[0], y1i[1], A1i[ao + 4 * kk], b1[kk]);
                        dmma884(y2i[0], y2i[1], A2i[ao + 4 * kk], b2[kk]);
                    }
                }
            }
            const int i = 8 * mt + lr;
            if (i < nc) {
                double2 z1[2], z2[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    double a1r = y1r[c], a1i = y1i[c], a2r = y2r[c], a2i = y2i[c];
                    if (kp == 0) { a1i = 0.0; a2i = 0.0; }          // DC and Nyquist rows: real part only
                    {
                        const double er = a1r + a2r, ei = a1i - a2i;               // Y1 + conj(Y2)
                        const double dr = a1r - a2r, di = a1i + a2i;               // Y1 - conj(Y2)
                        const double orr = dr * w1.x - di * w1.y, oi = dr * w1.y + di * w1.x;
                        z1[c] = make_double2(er - oi, ei + orr);
                    }
                    {
                        const double er = a2r + a1r, ei = a2i - a1i;               // Y2 + conj(Y1)
                        const double dr = a2r - a1r, di = a2i + a1i;               // Y2 - conj(Y1)
                        const double orr = dr * w2.x - di * w2.y, oi = dr * w2.y + di * w2.x;
                        z2[c] = make_double2(er - oi, ei + orr);
                    }
                }
                // Z is [N/2][nc][cw] for this chunk of trajectories
                double2* d1 = Z + ((size_t)k1 * nc + i) * cw + nt * 8 + 2 * lc;
                d1[0] = z1[0]; d1[1] = z1[1];
                if (kp != 0 && k2 != k1) {
                    double2* d2 = Z + ((size_t)k2 * nc + i) * cw + nt * 8 + 2 * lc;
                    d2[0] = z2[0]; d2[1] = z2[1];
                }
            }
        }
    }
}

template <int KKMAX, bool CPLX>
int launch_shape(const double* lre, const double* lim, const double* xi, int nmd, int nc, int ldb, int n0, int cw,
                 const double2* twN, double2* Z, cudaStream_t st) {
    const int MT = (nc + 7) / 8, KK = (nc + 3) / 4;
    const int LP = 4 * KK + 4 + ((4 * KK) % 16 == 0 ? 0 : 16 - (4 * KK) % 16);
    const size_t smem = (size_t)4 * 8 * MT * LP * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(shape_dmma_kernel<KKMAX, CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    shape_dmma_kernel<KKMAX, CPLX><<<nmd / 4 + 1, shape_threads<KKMAX, CPLX>(), smem, st>>>(lre, lim, xi, nmd, nc, ldb, n0, cw, twN, Z);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace pymd

using namespace pymd;

extern "C" {

size_t pymd_noise_work_bytes(int nmd, int nc, int ncols) {
    return 2 * (size_t)(nmd / 2) * nc * ncols * sizeof(double2);
}

// xi: [nf][nc][xi_ld] white noise whose column 0 is trajectory n0 of the output; out: [nmd][nc][ldb]
static int noise_generate_impl(const double* lre, const double* lim, const double* xi, int xi_ld, int xi_n0, int nmd, int nc,
                               int ldb, int n0, int n1, double dt, double* out, void* work, void* stream, long long out_ld = 0) {
    PYMD_REQUIRE(lre && xi && out && work, "noise_generate: null argument");
    PYMD_REQUIRE(nmd >= 4 && (nmd & (nmd - 1)) == 0, "noise_generate: nmd=%d must be a power of two >= 4", nmd);
    PYMD_REQUIRE(nc > 0 && ldb > 0 && ldb % 32 == 0 && xi_ld % 8 == 0, "noise_generate: nc=%d ldb=%d xi_ld=%d", nc, ldb, xi_ld);
    PYMD_REQUIRE(0 <= n0 && n0 < n1 && n1 <= ldb && n0 % 8 == 0 && (n1 - n0) % 8 == 0,
                 "noise_generate: trajectory range [%d, %d) of %d (multiples of 8)", n0, n1, ldb);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int half = nmd / 2, cw = n1 - n0;
    PYMD_REQUIRE(xi_n0 >= 0 && xi_n0 + cw <= xi_ld, "noise_generate: white-noise columns [%d, %d) of %d", xi_n0, xi_n0 + cw, xi_ld);
    const double2* twN = twiddle_table(nmd);
    PYMD_REQUIRE(twN, "noise_generate: twiddle table allocation failed");
    // Z goes to W0; the passes ping-pong W1, W0, W1, ...; the last one writes the real rows of `out`
    double2* W0 = static_cast<double2*>(work);
    double2* W1 = W0 + (size_t)half * nc * cw;
    int rc;
    if (nc <= 64) {
        // tensor-pipe shaping: all operand planes of a frequency pair fit in shared memory
        if (nc <= 16) rc = lim ? launch_shape<4, true>(lre, lim, xi, nmd, nc, xi_ld, xi_n0, cw, twN, W0, st)
                               : launch_shape<4, false>(lre, lim, xi, nmd, nc, xi_ld, xi_n0, cw, twN, W0, st);
        else rc = lim ? launch_shape<16, true>(lre, lim, xi, nmd, nc, xi_ld, xi_n0, cw, twN, W0, st)
                      : launch_shape<16, false>(lre, lim, xi, nmd, nc, xi_ld, xi_n0, cw, twN, W0, st);
        if (rc) return rc;
    } else {
        PYMD_REQUIRE(cw % SX == 0, "noise_generate: trajectory chunks must be multiples of %d for nc > 64", SX);
        dim3 grid(half / 2 + 1, cw / SX), blk(SX, SY);
        shape_kernel<<<grid, blk, 0, st>>>(lre, lim, xi, nmd, nc, xi_ld, xi_n0, cw, twN, W0);
        PYMD_CUDA_CHECK(cudaGetLastError());
    }
    FftPlan p{};
    p.n = half; p.cols = (long long)nc * cw; p.sign = -1;
    p.load = LOAD_C; p.store = STORE_REAL_ROWS;
    p.src_ld = p.cols; p.dst_ld = out_ld ? out_ld : (long long)nc * ldb;
    p.dst_cw = cw; p.dst_cs = ldb;
    p.scale = 1.0 / ((double)nmd * dt);
    return fft_columns(W0, out + n0, W1, W0, p, st, nullptr);
}

int pymd_noise_generate(const double* lre, const double* lim, const double* xi, int nmd, int nc, int ldb,
                        int n0, int n1, double dt, double* out, void* work, void* stream) {
    return noise_generate_impl(lre, lim, xi, ldb, n0, nmd, nc, ldb, n0, n1, dt, out, work, stream);
}

int pymd_noise_generate_chunk(const double* lre, const double* lim, const double* xi_chunk, int cw, int nmd, int nc, int ldb,
                              int n0, double dt, double* out, void* work, void* stream) {
    return noise_generate_impl(lre, lim, xi_chunk, cw, 0, nmd, nc, ldb, n0, n0 + cw, dt, out, work, stream);
}

int pymd_noise_generate_chunk_ld(const double* lre, const double* lim, const double* xi_chunk, int cw, int nmd, int nc, int ldb,
                                 int n0, double dt, double* out, long long out_row_stride, void* work, void* stream) {
    PYMD_REQUIRE(out_row_stride >= (long long)nc * ldb && out_row_stride % 2 == 0, "noise_generate: output row stride %lld", out_row_stride);
    return noise_generate_impl(lre, lim, xi_chunk, cw, 0, nmd, nc, ldb, n0, n0 + cw, dt, out, work, stream, out_row_stride);
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

// Far field of long memory kernels for the per-step pipeline (any nd, any nc): a uniformly partitioned
// frequency-domain delay line (phbath.py:192-202 evaluated as a fast convolution).
//
// The general path advances in super-blocks of L steps (L a power of two near sqrt(8 ml)).  For a super-block
// that starts at time u0 (history rows p_c(0 .. u0-1) appended) the part of the friction tails that only needs
// rows OLDER than the super-block,
//       F[r] = dt * sum_{n < u0} K[u0 + r + 1 - n] . p_c(n),       r = 0 .. L-1,
// is split over the history blocks B_p = rows [u0 - (p+1) L, u0 - p L), p = 0 .. P-1: block p meets the taps
// k = (p+1) L + 1 + d, d = r - m in (-L, L), i.e. one length-2L circular convolution per block, dof and trajectory.
//   * the spectrum of a block (2L-point FFT of its L rows, zero padded) is computed ONCE, when the block is
//     complete, and kept in a ring of P spectra (the delay line);
//   * per super-block and frequency f = 0 .. L:  Yhat_f = sum_p Ghat_p(f) Xhat_{p}(f) is ONE real tensor-core
//     product with the partitions stacked along K (block form [[Gr, -Gi], [Gi, Gr]], K = P * 2 ncp), batched over
//     the L+1 frequencies through the TMA-fed DMMA GEMM of gemm.cu;
//   * an inverse FFT gives F.  Rows inside the super-block are handled by the direct block-Toeplitz products of
//     engine.cu (at most L taps instead of ml).
// Per trajectory-step this executes about nc^2 (8 ml / L + L) flops per bath instead of 2 nc^2 ml: 14x fewer for
// the stretch shape of config 4 (nc = 81, ml = 2000, L = 128).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

#include "engine.cuh"
#include "fft.cuh"

namespace pymd {

namespace {

// Z[m][j*ldb + n] = p_c(first + m)[j][n] for m < L (0 where the row is older than the run or than the ring), 0 for m >= L
__global__ void fdl_pack_kernel(const double* ring, int R, int ncp, int ldb, long long first, long long hcount, int L,
                                double2* Z) {
    const size_t cols = (size_t)ncp * ldb, total = (size_t)2 * L * cols;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / cols);
        const size_t c = idx % cols;
        double v = 0.0;
        const long long row = first + m;
        if (m < L && row >= 0 && row < hcount && row >= hcount - R) v = ring[(size_t)(row % R) * cols + c];
        Z[idx] = make_double2(v, 0.0);
    }
}

// S[f][slot][part][j][n] = part(Zhat[f][j*ldb + n]),  f <= L
__global__ void fdl_store_kernel(const double2* Zhat, double* S, int L, int P, int slot, int ncp, int ldb) {
    const size_t cols = (size_t)ncp * ldb, total = (size_t)(L + 1) * cols;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(idx / cols);
        const size_t c = idx % cols;
        const double2 z = Zhat[idx];
        double* dst = S + (((size_t)f * P + slot) * 2) * cols + c;
        dst[0] = z.x;
        dst[cols] = z.y;
    }
}

// full Hermitian spectrum from Y[f][part][i][n], f <= L:  Z[f] = Y[f], Z[2L - f] = conj Y[f]
__global__ void fdl_spec_kernel(const double* Y, double2* Z, int L, int nc, int ldb) {
    const size_t cols = (size_t)nc * ldb, total = (size_t)2 * L * cols;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / cols);
        const size_t c = idx % cols;
        const int f = m <= L ? m : 2 * L - m;
        const double* src = Y + ((size_t)f * 2) * cols + c;
        const double re = src[0], im = (m == 0 || m == L) ? 0.0 : src[cols];
        Z[idx] = make_double2(re, m <= L ? im : -im);
    }
}

// F[r][i][n] = Re z[r][i*ldb + n], r < L
__global__ void fdl_out_kernel(const double2* z, double* F, int L, int nc, int ldb) {
    const size_t total = (size_t)L * nc * ldb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        F[idx] = z[idx].x;
}

int fdl_pick_L(int ml) {
    const double t = std::sqrt(8.0 * ml);
    int L = 32;
    while (L < 256 && std::fabs(std::log2((double)(2 * L)) - std::log2(t)) < std::fabs(std::log2((double)L) - std::log2(t))) L *= 2;
    return L;
}

}  // namespace

void fdl_free(BathHost& H) {
    cudaFree(H.d_fdlA); cudaFree(H.d_fdlS); cudaFree(H.d_fdlY); cudaFree(H.d_fdlF); cudaFree(H.d_fdlZ);
    H.d_fdlA = H.d_fdlS = H.d_fdlY = H.d_fdlF = H.d_fdlZ = nullptr;
    H.fdl_L = H.fdl_P = 0;
    H.fdl_valid = false;
}

// The delay line pays off for memories of a few partitions; its operand table (L+1) x 2nc x 2 P ncp doubles and the
// spectra ring must stay small beside the ensemble (testing knob PYMD_FDL=0 keeps the direct product).
bool fdl_wanted(const pymd_ctx* c, const BathHost& H) {
    const char* e = getenv("PYMD_FDL");
    if (e && atoi(e) == 0) return false;
    if (H.ml < 64) return false;
    const int L = fdl_pick_L(H.ml), P = (H.ml - 3) / L + 1;
    if (H.R < L) return false;
    const double table = (double)(L + 1) * 2 * H.nc * ((double)P * 2 * H.ncp + 32) * 8;
    const double ring = (double)(L + 1) * P * 2 * H.ncp * c->ldb * 8;
    return table < 6e9 && ring < 24e9;
}

int fdl_build(pymd_ctx* c, int b) {
    BathHost& H = c->bath[b];
    fdl_free(H);
    const int nc = H.nc, ncp = H.ncp, L = fdl_pick_L(H.ml), P = (H.ml - 3) / L + 1, N2 = 2 * L;
    H.fdl_L = L; H.fdl_P = P;
    const long long Kc = (long long)P * 2 * ncp;                 // columns of A_f
    const long long lda = (Kc + 31) / 32 * 32;
    const int M = 2 * nc;
    // Ghat_p(f) = (1/2L) sum_d dt K[(p+1) L + 1 + d] exp(-2 pi i f d / 2L),  d in (-L, L), 1 <= k <= ml-1:
    // a 2L-point FFT of the circularly arranged taps per (p, i, j) (iterative radix 2 on the host)
    std::vector<std::complex<double>> w(N2);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < N2; ++k) w[k] = std::complex<double>((double)cosl(2.0L * pi * k / N2), (double)-sinl(2.0L * pi * k / N2));
    int lg = 0;
    while ((1 << lg) < N2) ++lg;
    std::vector<int> rev(N2);
    for (int i = 0; i < N2; ++i) {
        int r = 0;
        for (int bb = 0; bb < lg; ++bb) r |= ((i >> bb) & 1) << (lg - 1 - bb);
        rev[i] = r;
    }
    std::vector<double> A((size_t)(L + 1) * M * lda + 64, 0.0);
    std::vector<std::complex<double>> gh(N2);
    for (int p = 0; p < P; ++p) {
        const int kap = P - 1 - p;                               // partitions stored in reversed order (see fdl_advance)
        for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j) {
                bool any = false;
                for (int t = 0; t < N2; ++t) gh[t] = 0.0;
                for (int d = -(L - 1); d <= L - 1; ++d) {
                    const long long k = (long long)(p + 1) * L + 1 + d;
                    if (k < 1 || k > H.ml - 1) continue;
                    const double g = c->dt * H.kernel[((size_t)k * nc + i) * nc + j] / N2;
                    if (g != 0.0) any = true;
                    gh[rev[(d + N2) % N2]] = g;                  // bit-reversed input order
                }
                if (!any) continue;
                for (int len = 2; len <= N2; len <<= 1) {
                    const int half = len >> 1, step = N2 / len;
                    for (int s0 = 0; s0 < N2; s0 += len)
                        for (int t = 0; t < half; ++t) {
                            const std::complex<double> u = gh[s0 + t], v = gh[s0 + t + half] * w[t * step];
                            gh[s0 + t] = u + v;
                            gh[s0 + t + half] = u - v;
                        }
                }
                for (int f = 0; f <= L; ++f) {
                    double* Af = A.data() + (size_t)f * M * lda;
                    const size_t cr = (size_t)kap * 2 * ncp + j, ci = cr + ncp;
                    Af[(size_t)i * lda + cr] = gh[f].real();          // Yr += Gr Xr - Gi Xi
                    Af[(size_t)i * lda + ci] = -gh[f].imag();
                    Af[(size_t)(nc + i) * lda + cr] = gh[f].imag();   // Yi += Gi Xr + Gr Xi
                    Af[(size_t)(nc + i) * lda + ci] = gh[f].real();
                }
            }
    }
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_fdlA, A.size() * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMemcpy(H.d_fdlA, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
    const size_t cols = (size_t)ncp * c->ldb;
    const size_t sbytes = (size_t)(L + 1) * P * 2 * cols * sizeof(double);
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_fdlS, sbytes));
    PYMD_CUDA_CHECK(cudaMemset(H.d_fdlS, 0, sbytes));
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_fdlY, (size_t)(L + 1) * 2 * nc * c->ldb * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_fdlF, (size_t)L * nc * c->ldb * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_fdlZ, (size_t)3 * N2 * cols * sizeof(double2)));
    H.fdl_valid = false;
    return PYMD_OK;
}

// spectrum of the history block that starts at row `first` into ring slot `slot`
static int fdl_block(pymd_ctx* c, int b, long long first, int slot, cudaStream_t st) {
    BathHost& H = c->bath[b];
    const int L = H.fdl_L, N2 = 2 * L, ncp = H.ncp;
    const size_t cols = (size_t)ncp * c->ldb;
    double2* Z = reinterpret_cast<double2*>(H.d_fdlZ);
    double2* A = Z + (size_t)N2 * cols;
    double2* B = A + (size_t)N2 * cols;
    fdl_pack_kernel<<<148 * 8, 256, 0, st>>>(c->sd.b[b].ring, H.R, ncp, c->ldb, first, H.hcount, L, Z);
    FftPlan p{};
    p.n = N2; p.cols = (long long)cols; p.sign = -1;
    p.load = LOAD_C; p.store = STORE_C; p.src_ld = p.cols; p.dst_ld = p.cols; p.scale = 1.0;
    const int np = fft_num_passes(N2);
    double2* dst = (np % 2) ? A : B;
    int rc = fft_columns(Z, dst, A, B, p, st, nullptr);
    if (rc) return rc;
    fdl_store_kernel<<<148 * 8, 256, 0, st>>>(dst, H.d_fdlS, L, H.fdl_P, slot, ncp, c->ldb);
    c->launches += 2 + np;
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

// A new super-block starts now (H.hcount rows appended): bring the delay line up to date and compute its far field.
int fdl_advance(pymd_ctx* c, int b, cudaStream_t st) {
    BathHost& H = c->bath[b];
    const int L = H.fdl_L, P = H.fdl_P, N2 = 2 * L, nc = H.nc, ncp = H.ncp;
    int rc;
    if (H.fdl_valid && H.hcount == H.fdl_base + L) {
        // steady state: the block that has just been completed enters the ring
        H.fdl_head = (H.fdl_head + 1) % P;
        if ((rc = fdl_block(c, b, H.hcount - L, H.fdl_head, st))) return rc;
    } else {
        // (re)start: rebuild every block from the velocity ring, aligned to the current time
        H.fdl_head = 0;
        for (int p = 0; p < P; ++p)
            if ((rc = fdl_block(c, b, H.hcount - (long long)(p + 1) * L, ((-p) % P + P) % P, st))) return rc;
    }
    H.fdl_base = H.hcount;
    H.fdl_valid = true;
    // Yhat_f = A_f . S_f for all f: partitions in ring order, the reversed table slides with the head (engine.cu's
    // launch_tail does the same with the reversed kernel table): slot s <= head holds block p = head - s, column
    // block P-1-p = P-1-head+s; slots above the head wrap: p = head + P - s, column block s - head - 1.
    const long long Kc = (long long)P * 2 * ncp, lda = (Kc + 31) / 32 * 32;
    const int h = H.fdl_head, W2 = 2 * ncp;
    GemmArgs g{};
    g.a_base = H.d_fdlA; g.a_ld = lda; g.a_ls = 0; g.a_mod = 2 * nc;
    g.x_base = H.d_fdlS; g.x_idx = nullptr; g.x_ld = c->ldb;
    g.y_base = H.d_fdlY; g.y_idx = nullptr; g.y_ld = c->ldb;
    g.M = 2 * nc; g.N = c->ldb;
    g.nseg = 1;
    g.seg_lo[0] = 0; g.seg_hi[0] = (h + 1) * W2; g.seg_aoff[0] = (long long)(P - 1 - h) * W2;
    if (h + 1 < P) {
        g.nseg = 2;
        g.seg_lo[1] = (h + 1) * W2; g.seg_hi[1] = P * W2; g.seg_aoff[1] = -(long long)(h + 1) * W2;
    }
    g.alpha = 1.0; g.accumulate = 0;
    g.nbatch = L + 1;
    g.a_bs = (long long)2 * nc * lda; g.x_bs = (long long)P * 2 * ncp * c->ldb; g.y_bs = (long long)2 * nc * c->ldb;
    prof_begin(c, 4, st);
    rc = launch_gemm(g, st);
    if (rc) return rc;
    // inverse transform of the Hermitian spectrum; the first L outputs are the far field
    const size_t cols = (size_t)nc * c->ldb;
    double2* Z = reinterpret_cast<double2*>(H.d_fdlZ);
    double2* A = Z + (size_t)N2 * ncp * c->ldb;
    double2* B = A + (size_t)N2 * ncp * c->ldb;
    fdl_spec_kernel<<<148 * 8, 256, 0, st>>>(H.d_fdlY, Z, L, nc, c->ldb);
    FftPlan p{};
    p.n = N2; p.cols = (long long)cols; p.sign = +1;
    p.load = LOAD_C; p.store = STORE_C; p.src_ld = p.cols; p.dst_ld = p.cols; p.scale = 1.0;
    const int np = fft_num_passes(N2);
    double2* dst = (np % 2) ? A : B;
    rc = fft_columns(Z, dst, A, B, p, st, nullptr);
    if (rc) return rc;
    fdl_out_kernel<<<148 * 8, 256, 0, st>>>(dst, H.d_fdlF, L, nc, c->ldb);
    prof_end(c, st);
    c->launches += 3 + np;
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace pymd

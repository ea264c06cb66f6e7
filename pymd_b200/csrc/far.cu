// Far-field part of the memory-kernel friction by fast convolution (phbath.py:192-202).
//
// The fused path advances in super-blocks of kFarS = 32 steps.  For a super-block that starts
// at time u0 the contribution of the history OLDER than p(u0-1) to the friction tails of all 32
// steps,
//       F[r] = dt * sum_{a>=1} K[r+2+a] . p_c(u0-1-a),          r = 0..31,   r+2+a <= ml-1,
// is one linear convolution in time per (bath dof, trajectory).  It is evaluated here with a
// length-128 FFT: x[m] = p_c(u0-98+m) (m = 0..96, zero padded), y = IFFT( Ghat . FFT(x) ) where
// Ghat[f] (nc x nc complex, built on the host once per bath) is the transform of the shifted
// kernel g[d mod 128] = dt K[d+99], d = -96..31, so that F[r] = y[r] exactly (no wrap-around:
// the linear convolution is supported on [3, 195] and only y[99..130] is read).  This replaces a
// (32 nc) x (97 nc) block-Toeplitz GEMM per super-block (2*32*~82*nc^2 flops per trajectory) by
// two real FFTs per series and 65 small complex products (~0.19 Mflop for nc = 15, 6.5x fewer).
//
// One CTA owns 8 trajectories of one bath:
//   phase 1  thread (dof j, trajectory n): gathers its 97 ring entries (coalesced over n), runs a
//            64-point complex FFT entirely in registers (real sequence packed as even + i odd),
//            untangles the 65 bins and writes them to shared memory [f][re|im][j][n];
//   phase 2  warp w takes frequencies w, w+8, ...: Yhat_f = Ghat_f Xhat_f as 4 real DMMA.8x8x4
//            products per (m8, k4) fragment pair, A fragments (Ghat, pre-swizzled into fragment
//            order on the host) streamed from L2, B fragments from shared memory, result written
//            back in place;
//   phase 3  thread (dof i, trajectory n): re-tangles, inverse 64-point FFT in registers, stores
//            the 32 wanted outputs to F[r][i][n].
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "engine.cuh"

namespace pymd {

namespace {

constexpr int NF = 65;            // bins 0..64 of the length-128 real transform
constexpr int NTRAJ = 8;          // trajectories per CTA (one n8 tile)
constexpr int FAR_THREADS = 256;

__constant__ double2 c_w64[64];   // exp(-2 pi i k / 64)
__constant__ double2 c_w128[65];  // exp(-2 pi i k / 128)

// Position of bin f after the 32-point transform below (mixed radix 4, 4, 2, decimation in
// frequency: every stage scatters its outputs by residue).
__host__ __device__ constexpr int pos32(int f) { return (f & 3) * 8 + ((f >> 2) & 3) * 2 + (f >> 4); }

// 32-point complex FFT, radix 4 x 4 x 2 decimation in frequency, all indices compile-time: the
// arrays live in registers.  Input in natural order, output bin k at position pos32(k).
// SIGN = -1 forward, +1 inverse (unnormalised).
template <int SIGN>
__device__ __forceinline__ void fft32(double (&re)[32], double (&im)[32]) {
#pragma unroll
    for (int stage = 0; stage < 2; ++stage) {
        const int s = stage == 0 ? 32 : 8, q = s >> 2, tw = 64 / s;
#pragma unroll
        for (int b = 0; b < 32; b += s) {
#pragma unroll
            for (int j = 0; j < q; ++j) {
                const int i0 = b + j, i1 = i0 + q, i2 = i1 + q, i3 = i2 + q;
                const double t0r = re[i0] + re[i2], t0i = im[i0] + im[i2];
                const double t1r = re[i0] - re[i2], t1i = im[i0] - im[i2];
                const double t2r = re[i1] + re[i3], t2i = im[i1] + im[i3];
                const double dr = re[i1] - re[i3], di = im[i1] - im[i3];
                // t3 = -i d (forward), +i d (inverse)
                const double t3r = SIGN < 0 ? di : -di, t3i = SIGN < 0 ? -dr : dr;
                re[i0] = t0r + t2r; im[i0] = t0i + t2i;
                double y1r = t1r + t3r, y1i = t1i + t3i;
                double y2r = t0r - t2r, y2i = t0i - t2i;
                double y3r = t1r - t3r, y3i = t1i - t3i;
                const int e1 = j * tw, e2 = 2 * j * tw, e3 = 3 * j * tw;     // < 64 for both stages
                if (e1 != 0) {
                    const double c1 = c_w64[e1].x, s1 = SIGN < 0 ? c_w64[e1].y : -c_w64[e1].y;
                    const double c2 = c_w64[e2].x, s2 = SIGN < 0 ? c_w64[e2].y : -c_w64[e2].y;
                    const double c3 = c_w64[e3].x, s3 = SIGN < 0 ? c_w64[e3].y : -c_w64[e3].y;
                    double t;
                    t = y1r * c1 - y1i * s1; y1i = y1r * s1 + y1i * c1; y1r = t;
                    t = y2r * c2 - y2i * s2; y2i = y2r * s2 + y2i * c2; y2r = t;
                    t = y3r * c3 - y3i * s3; y3i = y3r * s3 + y3i * c3; y3r = t;
                }
                re[i1] = y1r; im[i1] = y1i;
                re[i2] = y2r; im[i2] = y2i;
                re[i3] = y3r; im[i3] = y3i;
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 32; b += 2) {
        const double ar = re[b], ai = im[b], br = re[b + 1], bi = im[b + 1];
        re[b] = ar + br; im[b] = ai + bi;
        re[b + 1] = ar - br; im[b + 1] = ai - bi;
    }
}

// untangle one conjugate pair of the packed real transform: X[f] from Z[f] = (zr, zi), Z[64-f] = (cr, ci)
__device__ __forceinline__ void untangle(double zr, double zi, double cr, double ci, int f, double* out, int plane) {
    const double ar = 0.5 * (zr + cr), ai = 0.5 * (zi - ci);
    const double br = 0.5 * (zr - cr), bi = 0.5 * (zi + ci);
    const double wr = c_w128[f].x, wi = c_w128[f].y;
    // X[f] = a - i w b = a + (wi br + wr bi) + i (wi bi - wr br)
    out[(2 * f) * plane] = ar + (wi * br + wr * bi);
    out[(2 * f + 1) * plane] = ai + (wi * bi - wr * br);
}
// the inverse step: Z'[f] = (Y[f] + conj Y[64-f]) + i conj(w128^f) (Y[f] - conj Y[64-f])
__device__ __forceinline__ void retangle(double yr, double yi, double cr, double ci, int f, double& zr, double& zi) {
    const double ar = yr + cr, ai = yi - ci, br = yr - cr, bi = yi + ci;
    const double wr = c_w128[f].x, wi = -c_w128[f].y;
    zr = ar - (wr * bi + wi * br);
    zi = ai + (wr * br - wi * bi);
}

struct FarBath {
    const double* ring;     // [R][ncp][ldb]
    const double* ghat;     // [NF][2][ncp/8][ncp/4][32] A fragments of Re/Im Ghat (1/128 folded in)
    double* F;              // [kFarS][nc][ldb]
    int R, nc, ncp, ml;
    int seg;                // window segment m: ages 1 + 97 m .. 97 (m + 1) before p(u0-1)
    int accumulate;         // add to F (segments m > 0)
    long long hc0;          // rows appended to the ring before p(u0)
};
struct FarParams {
    int ldb;
    FarBath b[PYMD_MAX_BATHS];
};

template <int NCP>
__global__ void __launch_bounds__(FAR_THREADS, 1) far_kernel(const FarParams P) {
    extern __shared__ __align__(16) double spec[];          // [NF][2][ncp][NTRAJ]
    const FarBath& B = P.b[blockIdx.y];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int ncp = NCP;
    const int nc = B.nc;
    const size_t ldb = P.ldb;
    const int col0 = blockIdx.x * NTRAJ;
    const int plane = ncp * NTRAJ;                          // doubles per (f, part)
    // a series (dof j, trajectory n) is shared by two threads: half h = 0 takes the even bins, h = 1
    // the odd bins of its 64-point complex transform (one radix-2 step, then a 32-point FFT each)
    const int e = tid & 127, h = tid >> 7;
    const int j = e / NTRAJ, n = e % NTRAJ;
    const bool series = j < nc;

    // padding rows (j >= nc) are B operands of phase 2 against zero A columns: keep them finite
    for (int i = tid; i < NF * 2 * (ncp - nc) * NTRAJ; i += FAR_THREADS) {
        const int fp = i / ((ncp - nc) * NTRAJ), r = i % ((ncp - nc) * NTRAJ);
        spec[fp * plane + nc * NTRAJ + r] = 0.0;
    }

    // ------------------------------------------------------------------ phase 1: forward
    if (series) {
        double re[32], im[32];
        // x[m] = p_c(u0-98+m) lives in ring slot (hc0-98+m) mod R; entries older than the memory
        // (age a = 97-m > ml-3) or older than the run (slot index < 0) are zero.  z[m] = x[2m] + i x[2m+1];
        // a[m] = (z[m] +- z[m+32]) w64^(h m) feeds the even (h = 0) / odd (h = 1) bins.
        const long long first = B.hc0 - 98 - 97LL * B.seg;
        int mlo = 100 + 97 * B.seg - B.ml;
        if (first + mlo < 0) mlo = first + 97 < 0 ? 97 : (int)(-first);
        int slot = (int)(((first % B.R) + B.R) % B.R);
        const size_t rowpitch = (size_t)ncp * ldb;
        const double* base = B.ring + (size_t)j * ldb + col0 + n;
        const double* src = base + (size_t)slot * rowpitch;
        const double sgn = h ? -1.0 : 1.0;
#pragma unroll
        for (int m = 0; m < 64; ++m) {                      // x[m]: first half of z
            double v = 0.0;
            if (m >= mlo) v = *src;
            src += rowpitch;
            if (++slot == B.R) { slot = 0; src = base; }
            if (m & 1) im[m >> 1] = v; else re[m >> 1] = v;
        }
#pragma unroll
        for (int m = 64; m < 97; ++m) {                     // x[m]: z[32..48], folded in with the sign
            double v = 0.0;
            if (m >= mlo) v = *src;
            src += rowpitch;
            if (++slot == B.R) { slot = 0; src = base; }
            if (m & 1) im[(m >> 1) - 32] += sgn * v; else re[(m >> 1) - 32] += sgn * v;
        }
        // (entries 17..31 have no partner: z[m+32] = 0 there.)  The odd half carries the twiddle w64^m.
        if (h) {
#pragma unroll
            for (int m = 1; m < 32; ++m) {
                const double c = c_w64[m].x, sn = c_w64[m].y;
                const double t = re[m] * c - im[m] * sn;
                im[m] = re[m] * sn + im[m] * c;
                re[m] = t;
            }
        }
        fft32<-1>(re, im);
        // untangle: X[f] = (Z[f] + conj Z[64-f])/2 - (i/2) w128^f (Z[f] - conj Z[64-f]); Z[f] of this
        // thread's parity sits at pos32((f - h) / 2)
        double* out = spec + j * NTRAJ + n;
        if (h == 0) {
#pragma unroll
            for (int f = 0; f <= 32; f += 2) {
                const int g = 64 - f;
                const double zr = re[pos32(f / 2)], zi = im[pos32(f / 2)];
                const double cr = re[pos32((g & 63) / 2)], ci = im[pos32((g & 63) / 2)];
                untangle(zr, zi, cr, ci, f, out, plane);
                if (g != f) untangle(cr, ci, zr, zi, g, out, plane);
            }
        } else {
#pragma unroll
            for (int f = 1; f < 32; f += 2) {
                const int g = 64 - f;
                const double zr = re[pos32((f - 1) / 2)], zi = im[pos32((f - 1) / 2)];
                const double cr = re[pos32((g - 1) / 2)], ci = im[pos32((g - 1) / 2)];
                untangle(zr, zi, cr, ci, f, out, plane);
                untangle(cr, ci, zr, zi, g, out, plane);
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: Yhat = Ghat Xhat
    {
        constexpr int MT = NCP / 8, KK = NCP / 4, NA = MT * KK;
        const int lr = lane >> 2, lc = lane & 3;
        // A fragments of the next frequency are fetched from L2 while the current one is multiplied
        double gr[NA], gi[NA], nr[NA], ni[NA];
        {
            const double* G = B.ghat + (size_t)(2 * warp) * NA * 32 + lane;
#pragma unroll
            for (int a = 0; a < NA; ++a) { nr[a] = G[a * 32]; ni[a] = G[(NA + a) * 32]; }
        }
        for (int f = warp; f < NF; f += FAR_THREADS / 32) {
#pragma unroll
            for (int a = 0; a < NA; ++a) { gr[a] = nr[a]; gi[a] = ni[a]; }
            if (f + FAR_THREADS / 32 < NF) {
                const double* G = B.ghat + (size_t)(2 * (f + FAR_THREADS / 32)) * NA * 32 + lane;
#pragma unroll
                for (int a = 0; a < NA; ++a) { nr[a] = G[a * 32]; ni[a] = G[(NA + a) * 32]; }
            }
            double* Xr = spec + (2 * f) * plane;
            double* Xi = Xr + plane;
            double accr[MT][2], acci[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) accr[mt][0] = accr[mt][1] = acci[mt][0] = acci[mt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
                const double br = Xr[(4 * kk + lc) * NTRAJ + lr], bi = Xi[(4 * kk + lc) * NTRAJ + lr];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    dmma884(accr[mt][0], accr[mt][1], gr[mt * KK + kk], br);
                    dmma884(acci[mt][0], acci[mt][1], gi[mt * KK + kk], br);
                    dmma884(accr[mt][0], accr[mt][1], -gi[mt * KK + kk], bi);
                    dmma884(acci[mt][0], acci[mt][1], gr[mt * KK + kk], bi);
                }
            }
            __syncwarp();       // every lane has read its B fragments of this frequency
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                *reinterpret_cast<double2*>(&Xr[(mt * 8 + lr) * NTRAJ + 2 * lc]) = make_double2(accr[mt][0], accr[mt][1]);
                *reinterpret_cast<double2*>(&Xi[(mt * 8 + lr) * NTRAJ + 2 * lc]) = make_double2(acci[mt][0], acci[mt][1]);
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: inverse
    {
        double re[32], im[32];
        const double* in = spec + j * NTRAJ + n;
        if (series) {
            // re-tangle this thread's parity: Z'[k] <- bins f = 2k + h and 64 - f
            if (h == 0) {
#pragma unroll
                for (int k = 0; k <= 16; ++k) {
                    const int f = 2 * k, g = 64 - f;
                    const double yr = in[(2 * f) * plane], yi = in[(2 * f + 1) * plane];
                    const double cr = in[(2 * g) * plane], ci = in[(2 * g + 1) * plane];
                    retangle(yr, yi, cr, ci, f, re[k], im[k]);
                    if (k != 0 && k != 16) retangle(cr, ci, yr, yi, g, re[32 - k], im[32 - k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int f = 2 * k + 1, g = 64 - f;
                    const double yr = in[(2 * f) * plane], yi = in[(2 * f + 1) * plane];
                    const double cr = in[(2 * g) * plane], ci = in[(2 * g + 1) * plane];
                    retangle(yr, yi, cr, ci, f, re[k], im[k]);
                    retangle(cr, ci, yr, yi, g, re[31 - k], im[31 - k]);
                }
            }
            fft32<+1>(re, im);          // E[m] = sum_k Z'[2k+h] w32^(-k m) at pos32(m)
        }
        __syncthreads();                // every thread has read its bins: the spectra can be overwritten
        // z'[m] = E_0[m] + conj(w64^m) E_1[m], m < 16: the odd half hands its term over through shared memory
        double2* xch = reinterpret_cast<double2*>(spec);     // [16][128]
        if (series && h) {
#pragma unroll
            for (int m = 0; m < kFarS / 2; ++m) {
                const double c = c_w64[m].x, sn = -c_w64[m].y;
                const double er = re[pos32(m)], ei = im[pos32(m)];
                xch[m * 128 + e] = make_double2(er * c - ei * sn, er * sn + ei * c);
            }
        }
        __syncthreads();
        if (series && !h) {
            double* dst = B.F + (size_t)j * ldb + col0 + n;
            const size_t rowpitch = (size_t)nc * ldb;
#pragma unroll
            for (int m = 0; m < kFarS / 2; ++m) {
                const double2 o = xch[m * 128 + e];
                double y0 = re[pos32(m)] + o.x, y1 = im[pos32(m)] + o.y;
                if (B.accumulate) {
                    y0 += dst[(size_t)(2 * m) * rowpitch];
                    y1 += dst[(size_t)(2 * m + 1) * rowpitch];
                }
                dst[(size_t)(2 * m) * rowpitch] = y0;
                dst[(size_t)(2 * m + 1) * rowpitch] = y1;
            }
        }
    }
}

unsigned long long g_tw_ready = 0;      // bit d: constant tables uploaded to device d

int init_twiddles() {
    int dev = 0;
    PYMD_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && (g_tw_ready >> dev) & 1ull) return PYMD_OK;
    double2 w64[64], w128[65];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < 64; ++k) {
        w64[k].x = (double)cosl(2.0L * pi * k / 64.0L);
        w64[k].y = (double)-sinl(2.0L * pi * k / 64.0L);
    }
    for (int k = 0; k <= 64; ++k) {
        w128[k].x = (double)cosl(2.0L * pi * k / 128.0L);
        w128[k].y = (double)-sinl(2.0L * pi * k / 128.0L);
    }
    PYMD_CUDA_CHECK(cudaMemcpyToSymbol(c_w64, w64, sizeof(w64)));
    PYMD_CUDA_CHECK(cudaMemcpyToSymbol(c_w128, w128, sizeof(w128)));
    if (dev < 64) g_tw_ready |= 1ull << dev;
    return PYMD_OK;
}

}  // namespace

// windows of 97 ring rows needed to cover the taps k = 3 .. ml-1
static int far_segments(int ml) { return std::max(1, (ml - 3 + 96) / 97); }

bool far_supported(const BathHost& H) {
    // a 97-entry window + 32 outputs fill the length-128 transform exactly: memories longer than 100
    // steps take one window (one launch) per 97 further taps; the in-kernel mid field of fused.cu holds
    // one tap (ncp x ncp) as at most 2 x 4 register fragments
    return H.ml >= kFarMinMl && H.ncp <= 16 && H.R >= H.ml - 2;
}

// Host side, once per bath: Ghat[f] = (1/128) sum_d dt K[d+99] exp(-2 pi i f d / 128), d = -96..31,
// stored as DMMA A fragments: [f][re|im][mt][kk][lane] = part(Ghat[f])[8 mt + lane/4][4 kk + lane%4].
int far_build(pymd_ctx* c, int b) {
    BathHost& H = c->bath[b];
    int rc = init_twiddles();
    if (rc) return rc;
    const int nc = H.nc, ncp = H.ncp, MT = ncp / 8, KK = ncp / 4;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<std::complex<double>> w(128);
    for (int k = 0; k < 128; ++k) w[k] = std::complex<double>((double)cosl(2.0L * pi * k / 128.0L), (double)-sinl(2.0L * pi * k / 128.0L));
    const int nseg = far_segments(H.ml);
    const size_t segsize = (size_t)NF * 2 * MT * KK * 32;
    std::vector<double> frag(segsize * nseg, 0.0);
    std::vector<std::complex<double>> gh(NF);
    for (int sg = 0; sg < nseg; ++sg)
        for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j) {
                for (int f = 0; f < NF; ++f) gh[f] = 0.0;
                for (int d = -96; d <= 31; ++d) {
                    const int k = d + 99 + 97 * sg;
                    if (k < 3 || k > H.ml - 1) continue;
                    const double g = c->dt * H.kernel[((size_t)k * nc + i) * nc + j] / 128.0;
                    for (int f = 0; f < NF; ++f) gh[f] += g * w[(((long long)f * d) % 128 + 128) % 128];
                }
                const int mt = i / 8, lr = i % 8, kk = j / 4, lc = j % 4, lane = lr * 4 + lc;
                for (int f = 0; f < NF; ++f) {
                    frag[sg * segsize + (((size_t)(2 * f) * MT + mt) * KK + kk) * 32 + lane] = gh[f].real();
                    frag[sg * segsize + (((size_t)(2 * f + 1) * MT + mt) * KK + kk) * 32 + lane] = gh[f].imag();
                }
            }
    if (H.d_ghat) cudaFree(H.d_ghat);
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_ghat, frag.size() * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMemcpy(H.d_ghat, frag.data(), frag.size() * sizeof(double), cudaMemcpyHostToDevice));
    // dt K[k], k = 2 .. 2+kMidK-1, in the same fragment order, for the rows of the current super-block
    std::vector<double> km((size_t)kMidK * MT * KK * 32, 0.0);
    for (int kidx = 0; kidx < kMidK; ++kidx) {
        const int k = kidx + 2;
        if (k > H.ml - 1) break;
        for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j)
                km[(((size_t)kidx * MT + i / 8) * KK + j / 4) * 32 + (i % 8) * 4 + j % 4] =
                    c->dt * H.kernel[((size_t)k * nc + i) * nc + j];
    }
    if (H.d_kmid) cudaFree(H.d_kmid);
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_kmid, km.size() * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMemcpy(H.d_kmid, km.data(), km.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (H.d_far) cudaFree(H.d_far);
    PYMD_CUDA_CHECK(cudaMalloc(&H.d_far, (size_t)kFarS * nc * c->ldb * sizeof(double)));
    H.far_valid = false;
    return PYMD_OK;
}

template <int NCP>
int far_launch_t(pymd_ctx* c, const FarParams& P, int nb, cudaStream_t st) {
    const size_t smem = (size_t)NF * 2 * NCP * NTRAJ * sizeof(double);
    static bool configured = false;
    if (!configured) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(far_kernel<NCP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    prof_begin(c, 4, st);
    far_kernel<NCP><<<dim3(c->ldb / NTRAJ, nb), FAR_THREADS, smem, st>>>(P);
    prof_end(c, st);
    c->launches++;
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

// Far fields of the super-block that starts now (ring state: hcount rows appended) for every bath
// in `mask`; one launch per distinct padded bath size.
int far_launch(pymd_ctx* c, unsigned mask, cudaStream_t st) {
    int maxseg = 1;
    for (int b = 0; b < c->nbath; ++b)
        if ((mask & (1u << b)) && c->bath[b].ml > 1) maxseg = std::max(maxseg, far_segments(c->bath[b].ml));
    for (int sg = 0; sg < maxseg; ++sg) {
        for (int ncp : {8, 16}) {
            FarParams P{};
            P.ldb = c->ldb;
            int nb = 0;
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (!(mask & (1u << b)) || H.ncp != ncp || sg >= far_segments(H.ml)) continue;
                if (sg > 0 && H.hcount < 2 + 97LL * sg) continue;       // window entirely before the run began
                FarBath& B = P.b[nb++];
                const size_t segsize = (size_t)NF * 2 * (ncp / 8) * (ncp / 4) * 32;
                B.ring = c->sd.b[b].ring; B.ghat = H.d_ghat + sg * segsize; B.F = H.d_far;
                B.R = H.R; B.nc = H.nc; B.ncp = H.ncp; B.ml = H.ml; B.hc0 = H.hcount;
                B.seg = sg; B.accumulate = sg > 0;
                H.far_base = H.hcount;
                H.far_valid = true;
            }
            if (!nb) continue;
            int rc = ncp == 8 ? far_launch_t<8>(c, P, nb, st) : far_launch_t<16>(c, P, nb, st);
            if (rc) return rc;
        }
    }
    return PYMD_OK;
}

}  // namespace pymd

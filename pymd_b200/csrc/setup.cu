// Set-up tables on the device: work that the reference does once per
// bath / per temperature on the host and that grows to seconds for long memory kernels.
//
//   pymd_gamt      phbath.gamt (phbath.py:221-254): the friction kernel in time as a cosine transform,
//                  K[k][e] = scale * sum_i coef(t_k, w_i) g[i][e], a [ml x nw] x [nw x nc^2] product whose
//                  left operand is generated on the fly (one cos / sincos per DMMA A-fragment element).
//   pymd_harmonic  harmonic.py:37-84: for every frequency D = [(w + i eta)^2 - K - Sigma(w)]^-1 by an
//                  in-place Gauss-Jordan inversion with partial pivoting in shared memory (one CTA per
//                  frequency), then DDos = -2 w Tr Im D and UU = Tr[D Pi D^dagger].
#include <math.h>

#include "../../include/pymd_b200.h"
#include "common.cuh"

namespace pymd {
namespace {

// ---------------------------------------------------------------------------------------------------
// gamt: warp = 8 time rows x 64 table columns; the k4 steps run over the frequency grid.
// eta == 0: coef = 2 cos(t w); eta != 0 (artificial damping, phbath.py:236-252):
//   coef = exp(-eta t) [w/(w - i eta) e^{-i w t} + w/(w + i eta) e^{+i w t}]
//        = 2 exp(-eta t) (w^2 cos(w t) + eta w sin(w t)) / (w^2 + eta^2)            (real; 0 at w = 0)
constexpr int GW = 4, GNJ = 8;

__global__ void __launch_bounds__(GW * 32) gamt_kernel(const double* __restrict__ tl, int ml, const double* __restrict__ w,
                                                       int nw, const double* __restrict__ g, int ne, double eta,
                                                       double scale, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, lr = lane >> 2, lc = lane & 3;
    const int k = (blockIdx.y * GW + wp) * 8 + lr;
    const int col0 = blockIdx.x * (8 * GNJ);
    const double tk = k < ml ? tl[k] : 0.0;
    const double dec = eta != 0.0 ? exp(-eta * tk) : 1.0;
    double acc[GNJ][2];
#pragma unroll
    for (int j = 0; j < GNJ; ++j) acc[j][0] = acc[j][1] = 0.0;
    for (int i0 = 0; i0 < nw; i0 += 4) {
        const int i = i0 + lc;
        double a = 0.0;
        if (k < ml && i < nw) {
            const double wi = w[i], ph = tk * wi;
            if (eta == 0.0) {
                a = 2.0 * cos(ph);
            } else {
                double s, c;
                sincos(ph, &s, &c);
                a = 2.0 * dec * (wi * wi * c + eta * wi * s) / (wi * wi + eta * eta);
            }
        }
#pragma unroll
        for (int j = 0; j < GNJ; ++j) {
            const int col = col0 + 8 * j + lr;
            const double b = (i < nw && col < ne) ? __ldg(g + (size_t)i * ne + col) : 0.0;
            dmma884(acc[j][0], acc[j][1], a, b);
        }
    }
    if (k < ml) {
#pragma unroll
        for (int j = 0; j < GNJ; ++j) {
            const int col = col0 + 8 * j + 2 * lc;
            if (col < ne) out[(size_t)k * ne + col] = scale * acc[j][0];
            if (col + 1 < ne) out[(size_t)k * ne + col + 1] = scale * acc[j][1];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// harmonic spectra: complex arithmetic on double2
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfnma(double2 a, double2 b, double2 c) {      // c - a b
    return make_double2(c.x - (a.x * b.x - a.y * b.y), c.y - (a.x * b.y + a.y * b.x));
}
__device__ __forceinline__ double2 cinv(double2 a) {                             // Smith's division 1/a
    if (fabs(a.x) >= fabs(a.y)) {
        const double r = a.y / a.x, d = a.x + a.y * r;
        return make_double2(1.0 / d, -r / d);
    }
    const double r = a.x / a.y, d = a.x * r + a.y;
    return make_double2(r / d, -1.0 / d);
}

constexpr int HT = 256;

// shared memory: A[n][n+1] double2 | rowk[n] double2 | colk[n] double2 | red[HT] double | piv[n] int
__global__ void __launch_bounds__(HT) harmonic_kernel(const double* __restrict__ wl, int nw, const double* __restrict__ dyn,
                                                      const double2* __restrict__ se, const double2* __restrict__ pi,
                                                      int n, double eta, double* __restrict__ ddos,
                                                      double* __restrict__ uu) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int ld = n + 1;
    double2* A = reinterpret_cast<double2*>(smraw);
    double2* rowk = A + (size_t)n * ld;
    double2* colk = rowk + n;
    double* red = reinterpret_cast<double*>(colk + n);
    int* piv = reinterpret_cast<int*>(red + HT);
    const int tid = threadIdx.x, lane = tid & 31;

    for (int f = blockIdx.x; f < nw; f += gridDim.x) {
        const double wv = wl[f];
        const double2 z = make_double2(wv * wv - eta * eta, 2.0 * wv * eta);            // (w + i eta)^2
        const double2* S = se + (size_t)f * n * n;
        for (int e = tid; e < n * n; e += HT) {
            const int i = e / n, j = e % n;
            const double2 s = S[e];
            double2 v = make_double2(-dyn[e] - s.x, -s.y);
            if (i == j) { v.x += z.x; v.y += z.y; }
            A[i * ld + j] = v;
        }
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            if (tid < 32) {                       // pivot: largest |A[r][k]|, r >= k (lowest r on ties)
                double best = -1.0;
                int br = k;
                for (int r = k + lane; r < n; r += 32) {
                    const double2 v = A[r * ld + k];
                    const double m = fabs(v.x) + fabs(v.y);
                    if (m > best) { best = m; br = r; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int orr = __shfl_xor_sync(0xffffffffu, br, o);
                    if (ob > best || (ob == best && orr < br)) { best = ob; br = orr; }
                }
                if (lane == 0) piv[k] = br;
            }
            __syncthreads();
            const int r = piv[k];
            if (r != k)
                for (int j = tid; j < n; j += HT) {
                    const double2 t = A[k * ld + j];
                    A[k * ld + j] = A[r * ld + j];
                    A[r * ld + j] = t;
                }
            __syncthreads();
            {
                const double2 pinv = cinv(A[k * ld + k]);
                for (int j = tid; j < n; j += HT) {
                    rowk[j] = j == k ? pinv : cmul(A[k * ld + j], pinv);
                    colk[j] = A[j * ld + k];
                }
            }
            __syncthreads();
            for (int e = tid; e < n * n; e += HT) {
                const int i = e / n, j = e % n;
                if (i == k) {
                    A[i * ld + j] = rowk[j];
                } else {
                    const double2 c = j == k ? make_double2(0.0, 0.0) : A[i * ld + j];
                    A[i * ld + j] = cfnma(colk[i], rowk[j], c);
                }
            }
            __syncthreads();
        }
        for (int k = n - 1; k >= 0; --k) {        // undo the row interchanges as column interchanges
            const int r = piv[k];
            if (r != k)
                for (int i = tid; i < n; i += HT) {
                    const double2 t = A[i * ld + k];
                    A[i * ld + k] = A[i * ld + r];
                    A[i * ld + r] = t;
                }
            __syncthreads();
        }
        // DDos = -2 w Tr Im D ; UU = Re sum_ik (D Pi)_ik conj(D_ik)
        double tr = 0.0, acc = 0.0;
        for (int i = tid; i < n; i += HT) tr += A[i * ld + i].y;
        if (pi) {
            const double2* Pf = pi + (size_t)f * n * n;
            for (int e = tid; e < n * n; e += HT) {
                const int i = e / n, kk = e % n;
                double2 t = make_double2(0.0, 0.0);
                for (int j = 0; j < n; ++j) {
                    const double2 d = A[i * ld + j], p = Pf[(size_t)j * n + kk];
                    t.x += d.x * p.x - d.y * p.y;
                    t.y += d.x * p.y + d.y * p.x;
                }
                const double2 d = A[i * ld + kk];
                acc += t.x * d.x + t.y * d.y;
            }
        }
        for (int pass = 0; pass < 2; ++pass) {
            red[tid] = pass == 0 ? tr : acc;
            __syncthreads();
            for (int o = HT / 2; o > 0; o >>= 1) {
                if (tid < o) red[tid] += red[tid + o];
                __syncthreads();
            }
            if (tid == 0) {
                if (pass == 0) ddos[f] = -2.0 * wv * red[0];
                else if (uu) uu[f] = red[0];
            }
            __syncthreads();
        }
    }
}

size_t harmonic_smem(int n) {
    return ((size_t)n * (n + 1) + 2 * (size_t)n) * sizeof(double2) + HT * sizeof(double) + (size_t)n * sizeof(int);
}

}  // namespace
}  // namespace pymd

extern "C" {

int pymd_gamt(const double* tl_dev, int ml, const double* w_dev, int nw, const double* g_dev, int ne,
              double eta_ad, double scale, double* out_dev, void* stream) {
    using namespace pymd;
    PYMD_REQUIRE(tl_dev && w_dev && g_dev && out_dev, "gamt: null pointer");
    PYMD_REQUIRE(ml > 0 && nw > 0 && ne > 0, "gamt: bad shape ml=%d nw=%d ne=%d", ml, nw, ne);
    dim3 grid((ne + 8 * GNJ - 1) / (8 * GNJ), (ml + 8 * GW - 1) / (8 * GW));
    gamt_kernel<<<grid, GW * 32, 0, static_cast<cudaStream_t>(stream)>>>(tl_dev, ml, w_dev, nw, g_dev, ne, eta_ad, scale,
                                                                         out_dev);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

int pymd_harmonic_max_n(void) {
    int n = 1;
    while (pymd::harmonic_smem(n + 1) <= 227 * 1024) ++n;
    return n;
}

int pymd_harmonic(const double* wl_dev, int nw, const double* dyn_dev, const double* se_dev, const double* pi_dev,
                  int n, double eta, double* ddos_dev, double* uu_dev, void* stream) {
    using namespace pymd;
    PYMD_REQUIRE(wl_dev && dyn_dev && se_dev && ddos_dev, "harmonic: null pointer");
    PYMD_REQUIRE(nw > 0 && n > 0, "harmonic: bad shape nw=%d n=%d", nw, n);
    PYMD_REQUIRE(n <= pymd_harmonic_max_n(), "harmonic: n=%d exceeds the shared-memory inversion limit %d", n,
                 pymd_harmonic_max_n());
    PYMD_REQUIRE((pi_dev == nullptr) == (uu_dev == nullptr), "harmonic: pi and uu go together");
    const size_t smem = harmonic_smem(n);
    PYMD_CUDA_CHECK(cudaFuncSetAttribute(harmonic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    PYMD_CUDA_CHECK(cudaGetDevice(&dev));
    PYMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
    const int grid = nw < sms * per_sm ? nw : sms * per_sm;
    harmonic_kernel<<<grid, HT, smem, static_cast<cudaStream_t>(stream)>>>(
        wl_dev, nw, dyn_dev, reinterpret_cast<const double2*>(se_dev), reinterpret_cast<const double2*>(pi_dev), n, eta,
        ddos_dev, uu_dev);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// pymd_eigh_factors: batched Hermitian / real-symmetric eigen-decomposition of the noise covariances
// (numpy.linalg.eigh at noise.py:69,172) by cyclic Jacobi rotations, one CTA per frequency with the matrix
// and the accumulated eigenvectors in shared memory.  A round applies the n/2 disjoint rotations of a
// round-robin tournament at once: J = [[c, s e^{i phi}], [-s e^{-i phi}, c]] on columns (p, q) of A and V,
// then J^H on rows (p, q) of A; n - 1 rounds make a sweep and convergence is quadratic in sweeps.  The
// epilogue orders the eigenvalues ascending (as eigh does) and writes L = U diag(sqrt(max(lambda, 0))).
namespace pymd {
namespace {

__device__ __forceinline__ double e_re(double a) { return a; }
__device__ __forceinline__ double e_re(double2 a) { return a.x; }
__device__ __forceinline__ double e_im(double) { return 0.0; }
__device__ __forceinline__ double e_im(double2 a) { return a.y; }
__device__ __forceinline__ double e_abs2(double a) { return a * a; }
__device__ __forceinline__ double e_abs2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ void e_make(double& o, double re, double) { o = re; }
__device__ __forceinline__ void e_make(double2& o, double re, double im) { o = make_double2(re, im); }
// c x - conj(s) y   and   s x + c y   (column pair) ;  c x - s y  and  conj(s) x + c y  (row pair)
__device__ __forceinline__ void rot_cols(double c, double s, double& x, double& y) {
    const double nx = c * x - s * y, ny = s * x + c * y;
    x = nx; y = ny;
}
__device__ __forceinline__ void rot_rows(double c, double s, double& x, double& y) { rot_cols(c, s, x, y); }
__device__ __forceinline__ void rot_cols(double c, double2 s, double2& x, double2& y) {
    const double2 nx = make_double2(c * x.x - (s.x * y.x + s.y * y.y), c * x.y - (s.x * y.y - s.y * y.x));
    const double2 ny = make_double2(s.x * x.x - s.y * x.y + c * y.x, s.x * x.y + s.y * x.x + c * y.y);
    x = nx; y = ny;
}
__device__ __forceinline__ void rot_rows(double c, double2 s, double2& x, double2& y) {
    const double2 nx = make_double2(c * x.x - (s.x * y.x - s.y * y.y), c * x.y - (s.x * y.y + s.y * y.x));
    const double2 ny = make_double2(s.x * x.x + s.y * x.y + c * y.x, s.x * x.y - s.y * x.x + c * y.y);
    x = nx; y = ny;
}

constexpr int ET = 256, EMAXSWEEP = 40;

// vglob: the eigenvector accumulator lives in a global scratch (one n x n tile per CTA, L1/L2 resident)
// instead of shared memory -- for matrices too large to keep both there.
template <typename E>
size_t eigh_smem(int n, bool vglob) {
    const int ne = n + (n & 1);
    return (vglob ? 1 : 2) * (size_t)n * (n + 1) * sizeof(E) +
           (size_t)(ne / 2) * (sizeof(E) + sizeof(double) + 2 * sizeof(int)) +
           (size_t)n * (sizeof(double) + sizeof(int)) + (ET + 2) * sizeof(double) + 64;
}
constexpr size_t kSmemMax = 227 * 1024;

// Input: either the matrices themselves (ncomb == 0: cre/cim are [nf][n][n]) or, to avoid building and
// uploading them, a linear combination of a small table of matrices (cre/cim are [ntab][n][n]):
//   C_f = herm( sum_{m < ncomb} coef[f][m] * table[idx[f][m]] ),  herm(M) = (M + M^H) / 2
// -- the electron covariance is three fixed matrices with frequency-dependent weights (noise.py:156-170),
// the phonon one two neighbouring entries of the gamma table with the flinterp weights (noise.py:61-66).
template <typename E>
__global__ void __launch_bounds__(ET) eigh_kernel(const double* __restrict__ cre, const double* __restrict__ cim,
                                                  const int* __restrict__ idx, const double* __restrict__ coef, int ncomb,
                                                  int nf, int n, double* __restrict__ lam, double* __restrict__ lre,
                                                  double* __restrict__ lim, double* __restrict__ ure,
                                                  double* __restrict__ uim, int* __restrict__ sweeps, E* vwork) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int ld = n + 1, ne = n + (n & 1), np = ne / 2, tid = threadIdx.x;
    E* A = reinterpret_cast<E*>(smraw);
    // V is kept TRANSPOSED (V[j * ldv + i] = component i of vector j): a rotation of two vectors runs over
    // contiguous elements, conflict-free in shared memory and coalesced in the global scratch
    const int ldv = vwork ? n : ld;
    E* V = vwork ? vwork + (size_t)blockIdx.x * n * n : A + (size_t)n * ld;
    E* rs = (vwork ? A : V) + (size_t)n * ld;                     // s e^{i phi} per pair
    double* rc = reinterpret_cast<double*>(rs + np);              // c per pair
    double* ev = rc + np;                                         // eigenvalues
    double* red = ev + n;                                         // [ET + 2]
    int* rp = reinterpret_cast<int*>(red + ET + 2);               // p per pair (-1: idle)
    int* rq = rp + np;
    int* rank = rq + np;
    constexpr bool CPLX = sizeof(E) == sizeof(double2);

    for (int f = blockIdx.x; f < nf; f += gridDim.x) {
        const size_t base = (size_t)f * n * n;
        for (int e = tid; e < n * n; e += ET) {
            const int i = e / n, j = e % n;
            if (ncomb == 0) {
                e_make(A[i * ld + j], cre[base + e], CPLX ? cim[base + e] : 0.0);
            } else {
                double re = 0.0, im = 0.0;
                for (int m = 0; m < ncomb; ++m) {
                    const double cf = 0.5 * coef[(size_t)f * ncomb + m];
                    const size_t tb = (size_t)idx[(size_t)f * ncomb + m] * n * n;
                    re += cf * (cre[tb + e] + cre[tb + (size_t)j * n + i]);
                    if (CPLX) im += cf * (cim[tb + e] - cim[tb + (size_t)j * n + i]);
                }
                e_make(A[i * ld + j], re, im);
            }
            e_make(V[i * ldv + j], i == j ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        int sweep = 0;
        for (; sweep < EMAXSWEEP; ++sweep) {
            // off-diagonal and total squared norms
            double off = 0.0, dia = 0.0;
            for (int e = tid; e < n * n; e += ET) {
                const int i = e / n, j = e % n;
                const double a2 = e_abs2(A[i * ld + j]);
                if (i == j) dia += a2; else off += a2;
            }
            for (int pass = 0; pass < 2; ++pass) {
                red[tid] = pass == 0 ? off : dia;
                __syncthreads();
                for (int o = ET / 2; o > 0; o >>= 1) {
                    if (tid < o) red[tid] += red[tid + o];
                    __syncthreads();
                }
                if (tid == 0) red[ET + pass] = red[0];
                __syncthreads();
            }
            if (red[ET] <= 1e-30 * (red[ET] + red[ET + 1])) break;        // uniform
            for (int r = 0; r < ne - 1; ++r) {
                if (tid < np) {
                    int p = tid == 0 ? ne - 1 : (r + tid) % (ne - 1);
                    int q = tid == 0 ? r : (r - tid + (ne - 1)) % (ne - 1);
                    if (p > q) { const int t = p; p = q; q = t; }
                    double c = 1.0;
                    E s;
                    e_make(s, 0.0, 0.0);
                    bool act = q < n;
                    if (act) {
                        const E b = A[p * ld + q];
                        const double ab2 = e_abs2(b);
                        act = ab2 > 0.0;
                        if (act) {
                            const double ab = sqrt(ab2);
                            const double tau = (e_re(A[q * ld + q]) - e_re(A[p * ld + p])) / (2.0 * ab);
                            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                            c = 1.0 / sqrt(1.0 + t * t);
                            const double sn = t * c / ab;
                            e_make(s, sn * e_re(b), sn * e_im(b));
                        }
                    }
                    rp[tid] = act ? p : -1;
                    rq[tid] = q;
                    rc[tid] = c;
                    rs[tid] = s;
                }
                __syncthreads();
                for (int e = tid; e < np * n; e += ET) {              // columns of A and V
                    const int k = e / n, i = e % n, p = rp[k], q = rq[k];
                    if (p < 0) continue;
                    rot_cols(rc[k], rs[k], A[i * ld + p], A[i * ld + q]);
                    rot_cols(rc[k], rs[k], V[p * ldv + i], V[q * ldv + i]);
                }
                __syncthreads();
                for (int e = tid; e < np * n; e += ET) {              // rows of A
                    const int k = e / n, j = e % n, p = rp[k], q = rq[k];
                    if (p < 0) continue;
                    E x = A[p * ld + j], y = A[q * ld + j];
                    rot_rows(rc[k], rs[k], x, y);
                    if (j == q) e_make(x, 0.0, 0.0);                  // the annihilated pair, exactly
                    if (j == p) e_make(y, 0.0, 0.0);
                    if (j == p) e_make(x, e_re(x), 0.0);              // the diagonal stays real
                    if (j == q) e_make(y, e_re(y), 0.0);
                    A[p * ld + j] = x;
                    A[q * ld + j] = y;
                }
                __syncthreads();
            }
        }
        if (tid == 0 && sweeps) sweeps[f] = sweep;
        for (int i = tid; i < n; i += ET) ev[i] = e_re(A[i * ld + i]);
        __syncthreads();
        for (int i = tid; i < n; i += ET) {
            int rk = 0;
            const double v = ev[i];
            for (int j = 0; j < n; ++j) rk += (ev[j] < v || (ev[j] == v && j < i)) ? 1 : 0;
            rank[i] = rk;
            lam[(size_t)f * n + rk] = v;
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += ET) {
            const int i = e / n, j = e % n, col = rank[j];
            const E v = V[j * ldv + i];
            const double sq = sqrt(fmax(ev[j], 0.0));
            lre[base + (size_t)i * n + col] = sq * e_re(v);
            if (CPLX) lim[base + (size_t)i * n + col] = sq * e_im(v);
            if (ure) ure[base + (size_t)i * n + col] = e_re(v);
            if (CPLX && uim) uim[base + (size_t)i * n + col] = e_im(v);
        }
        __syncthreads();
    }
}

// CTAs of a launch (each owns one global V tile when the matrices do not fit shared memory twice)
template <typename E>
int eigh_grid(int nf, int n, bool vglob) {
    const size_t smem = eigh_smem<E>(n, vglob);
    int per_sm = (int)(kSmemMax / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    return nf < 148 * per_sm ? nf : 148 * per_sm;
}

template <typename E>
int eigh_launch(const double* cre, const double* cim, const int* idx, const double* coef, int ncomb, int nf, int n,
                double* lam, double* lre, double* lim, double* ure, double* uim, int* sweeps, void* work,
                cudaStream_t st) {
    const bool vglob = eigh_smem<E>(n, false) > kSmemMax;
    PYMD_REQUIRE(!vglob || work, "eigh_factors: n=%d needs the scratch of pymd_eigh_work_bytes()", n);
    const size_t smem = eigh_smem<E>(n, vglob);
    PYMD_CUDA_CHECK(cudaFuncSetAttribute(eigh_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = eigh_grid<E>(nf, n, vglob);
    eigh_kernel<E><<<grid, ET, smem, st>>>(cre, cim, idx, coef, ncomb, nf, n, lam, lre, lim, ure, uim, sweeps,
                                           vglob ? static_cast<E*>(work) : nullptr);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace
}  // namespace pymd

extern "C" {

int pymd_eigh_max_n(int cplx) {
    int n = 1;
    while ((cplx ? pymd::eigh_smem<double2>(n + 1, true) : pymd::eigh_smem<double>(n + 1, true)) <= pymd::kSmemMax) ++n;
    return n;
}

size_t pymd_eigh_work_bytes(int nf, int n, int cplx) {
    using namespace pymd;
    if (nf <= 0 || n <= 0 || n > pymd_eigh_max_n(cplx)) return 0;
    if (cplx) return eigh_smem<double2>(n, false) > kSmemMax ? (size_t)eigh_grid<double2>(nf, n, true) * n * n * sizeof(double2) : 0;
    return eigh_smem<double>(n, false) > kSmemMax ? (size_t)eigh_grid<double>(nf, n, true) * n * n * sizeof(double) : 0;
}

int pymd_eigh_factors(const double* cre_dev, const double* cim_dev, const int* idx_dev, const double* coef_dev, int ncomb,
                      int nf, int n, double* lam_dev, double* lre_dev, double* lim_dev, double* ure_dev, double* uim_dev,
                      int* sweeps_dev, void* work_dev, void* stream) {
    using namespace pymd;
    PYMD_REQUIRE(cre_dev && lam_dev && lre_dev, "eigh_factors: null pointer");
    PYMD_REQUIRE(nf > 0 && n > 0, "eigh_factors: bad shape nf=%d n=%d", nf, n);
    PYMD_REQUIRE(ncomb >= 0 && (ncomb == 0 || (idx_dev && coef_dev)), "eigh_factors: a combination needs idx and coef");
    const int cplx = cim_dev != nullptr;
    PYMD_REQUIRE(!cplx || lim_dev, "eigh_factors: complex input needs lim");
    PYMD_REQUIRE(n <= pymd_eigh_max_n(cplx), "eigh_factors: n=%d exceeds the shared-memory limit %d", n,
                 pymd_eigh_max_n(cplx));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cplx)
        return eigh_launch<double2>(cre_dev, cim_dev, idx_dev, coef_dev, ncomb, nf, n, lam_dev, lre_dev, lim_dev, ure_dev,
                                    uim_dev, sweeps_dev, work_dev, st);
    return eigh_launch<double>(cre_dev, nullptr, idx_dev, coef_dev, ncomb, nf, n, lam_dev, lre_dev, nullptr, ure_dev, nullptr,
                               sweeps_dev, work_dev, st);
}

}  // extern "C"

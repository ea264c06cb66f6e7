// Stepping engine: context management and the per-step GEMM pipeline (mode 1).
//
// One md.vv step (reference md.py:315-358) for the whole ensemble is restructured as
//   stage_a : record ps/qs/etot, first force evaluation (heads on p(t)), currents,
//             history append, half kick / drift                     (md.py:320-343)
//   tail    : block-Toeplitz product  dt * sum_{k>=1} K[k] p(t+1-k)  (phbath.py:197-201, k>=1)
//   stage_b : second force evaluation -> predictor p1                (md.py:347-348)
//   stage_c : third force evaluation -> p(t+1), constraints          (md.py:349-358)
// using the identity of SURVEY.md section 8a: the three evaluations share one tail and
// differ only in the k=0 "head"; heads of all baths are summed into one nd x nd operand.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "engine.cuh"

namespace {
thread_local char g_err[512] = "";
}

void pymd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace pymd {

int launch_tail(pymd_ctx* c, int b, int T, int koff, double* out, cudaStream_t st, long long wmax = -1,
                int accumulate = 0);
int fused_supported(const pymd_ctx* c);
int fused_step(pymd_ctx* c, long long t0, int nsteps, cudaStream_t stream);
int local_supported(const pymd_ctx* c);
int local_step(pymd_ctx* c, long long t0, int nsteps, cudaStream_t stream);

struct Slots {
    int s[PYMD_MAX_BATHS];
};

// ------------------------------------------------------------------ stage kernels
// block = (32 trajectories, 8 row lanes); a thread walks rows y, y+8, ... of one trajectory.
constexpr int TX = 8, TY = 32;     // 8 trajectories x 32 row lanes per CTA: ldb / 8 CTAs (a 1024-trajectory ensemble fills 128 SMs)

__device__ __forceinline__ double block_rows_sum(double v, double (*red)[TX]) {
    red[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < TY; ++y) s += red[y][threadIdx.x];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(TX* TY) stage_a_kernel(const StepDev d, int tn, Slots slots) {
    __shared__ double red[TY][TX];
    const int n = blockIdx.x * TX + threadIdx.x;
    const size_t ldb = d.ldb;
    const double* fsel = d.samef[n] ? d.fpotn : d.fpotc;
    const double dt = d.dt, dt2 = d.dt * d.dt;
    double e = 0.0;
    double cur[PYMD_MAX_BATHS];
#pragma unroll
    for (int b = 0; b < PYMD_MAX_BATHS; ++b) cur[b] = 0.0;

    for (int i = threadIdx.y; i < d.nd; i += TY) {
        const size_t o = (size_t)i * ldb + n;
        const double pi = d.p[o], qi = d.q[o];
        double f = fsel[o];
        if (d.ps) d.ps[((size_t)tn * d.nd + i) * ldb + n] = pi;
        if (d.qs) d.qs[((size_t)tn * d.nd + i) * ldb + n] = qi;
        e += pi * pi;
#pragma unroll
        for (int b = 0; b < PYMD_MAX_BATHS; ++b) {
            if (b >= d.nbath) break;
            const BathDev& B = d.b[b];
            const int ip = B.inv[i];
            if (ip < 0) continue;
            const size_t ob = (size_t)ip * ldb + n;
            double fb = B.noise[((size_t)tn * B.nc + ip) * ldb + n] - B.g[ob];
            if (B.nonlocal) {
                fb -= B.tail_cur[ob];
                B.ring[((size_t)slots.s[b] * B.ncp + ip) * ldb + n] = pi;
            }
            if (B.fhis) B.fhis[((size_t)tn * B.nc + ip) * ldb + n] = fb;
            cur[b] += fb * pi;
            f += fb;
        }
        d.ph[o] = pi + f * dt / 2.0;
        d.qn[o] = qi + pi * dt + f * dt2 / 2.0;
    }
    e = block_rows_sum(e, red);
    if (threadIdx.y == 0) d.etot[(size_t)tn * ldb + n] = 0.5 * e;
#pragma unroll
    for (int b = 0; b < PYMD_MAX_BATHS; ++b) {
        if (b >= d.nbath) break;
        const double c = block_rows_sum(cur[b], red);
        if (threadIdx.y == 0) d.b[b].cur[(size_t)tn * ldb + n] = c;
    }
}

__global__ void __launch_bounds__(TX* TY) stage_b_kernel(const StepDev d, int tn1) {
    const int n = blockIdx.x * TX + threadIdx.x;
    const size_t ldb = d.ldb;
    for (int i = threadIdx.y; i < d.nd; i += TY) {
        const size_t o = (size_t)i * ldb + n;
        double f = d.fpotn[o];
#pragma unroll
        for (int b = 0; b < PYMD_MAX_BATHS; ++b) {
            if (b >= d.nbath) break;
            const BathDev& B = d.b[b];
            const int ip = B.inv[i];
            if (ip < 0) continue;
            const size_t ob = (size_t)ip * ldb + n;
            double fb = B.noise[((size_t)tn1 * B.nc + ip) * ldb + n];
            if (B.nonlocal) fb -= B.tail_next[ob];
            if (B.has_aq) fb += B.uq[ob];
            f += fb;
        }
        d.fbase[o] = f;
        d.p1[o] = d.ph[o] + d.dt * (f - d.h[o]) / 2.0;
    }
}

__global__ void __launch_bounds__(TX* TY) stage_c_kernel(const StepDev d) {
    __shared__ double red[TY][TX];
    const int n = blockIdx.x * TX + threadIdx.x;
    const size_t ldb = d.ldb;
    if (d.ncon == 0) {
        for (int i = threadIdx.y; i < d.nd; i += TY) {
            const size_t o = (size_t)i * ldb + n;
            d.p[o] = d.ph[o] + d.dt * (d.fbase[o] - d.h[o]) / 2.0;
            d.q[o] = d.qn[o];
        }
        if (threadIdx.y == 0) d.samef[n] = 1;
        return;
    }
    // unprojected update, kept in p[] / qn[]
    for (int i = threadIdx.y; i < d.nd; i += TY) {
        const size_t o = (size_t)i * ldb + n;
        d.p[o] = d.ph[o] + d.dt * (d.fbase[o] - d.h[o]) / 2.0;
    }
    __syncthreads();
    double dp[PYMD_MAX_CONSTRAINTS], dq[PYMD_MAX_CONSTRAINTS];
    for (int c = 0; c < d.ncon; ++c) {
        double sp = 0.0, sq = 0.0;
        for (int i = threadIdx.y; i < d.nd; i += TY) {
            const size_t o = (size_t)i * ldb + n;
            const double cv = d.con[(size_t)c * d.nd + i];
            sp += d.p[o] * cv;
            sq += d.qn[o] * cv;
        }
        dp[c] = block_rows_sum(sp, red);
        dq[c] = block_rows_sum(sq, red);
    }
    double md = 0.0;
    for (int i = threadIdx.y; i < d.nd; i += TY) {
        const size_t o = (size_t)i * ldb + n;
        const double p0 = d.p[o], q0 = d.qn[o];
        double pn = p0, qq = q0;
        for (int c = 0; c < d.ncon; ++c) {
            const double cv = d.con[(size_t)c * d.nd + i];
            pn -= dp[c] * cv;
            qq -= dq[c] * cv;
        }
        d.p[o] = pn;
        d.q[o] = qq;
        md = fmax(md, fabs(qq - q0));
    }
    red[threadIdx.y][threadIdx.x] = md;
    __syncthreads();
    if (threadIdx.y == 0) {
        double m = 0.0;
        for (int y = 0; y < TY; ++y) m = fmax(m, red[y][threadIdx.x]);
        d.samef[n] = (m < 10e-10) ? 1 : 0;   // md.sameq, md.py:733
    }
}

__global__ void fill_int_kernel(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// newest-first copy between a ring and a dense [nrows][nc][ldb] array
__global__ void history_copy_kernel(double* ring, double* dense, int R, int ncp, int nc, int ldb,
                                    long long hcount, int nrows, int to_ring) {
    const size_t total = (size_t)nrows * nc * ldb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int n = idx % ldb;
        const int i = (idx / ldb) % nc;
        const int k = idx / ((size_t)ldb * nc);
        const long long src = hcount - 1 - k;
        const int slot = (int)(((src % R) + R) % R);
        double* r = ring + ((size_t)slot * ncp + i) * ldb + n;
        if (to_ring) *r = dense[idx];
        else dense[idx] = (src >= 0) ? *r : 0.0;
    }
}

// ------------------------------------------------------------------ helpers
static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int upload(double** dst, const std::vector<double>& h) {
    if (*dst) cudaFree(*dst);
    PYMD_CUDA_CHECK(cudaMalloc(dst, h.size() * sizeof(double)));
    PYMD_CUDA_CHECK(cudaMemcpy(*dst, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return PYMD_OK;
}
static int upload_i(int** dst, const std::vector<int>& h) {
    if (*dst) cudaFree(*dst);
    PYMD_CUDA_CHECK(cudaMalloc(dst, h.size() * sizeof(int)));
    PYMD_CUDA_CHECK(cudaMemcpy(*dst, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    return PYMD_OK;
}

// Build every padded device operand and the workspace.  Called lazily by pymd_step.
int finalize(pymd_ctx* c) {
    if (c->finalized) return PYMD_OK;
    PYMD_REQUIRE(!c->dyn.empty(), "pymd_step: no dynamical matrix (md.potforce: 'no driver, no md')");
    PYMD_REQUIRE(c->sd.p && c->sd.q, "pymd_step: state not bound");
    PYMD_REQUIRE(c->sd.etot, "pymd_step: etot not bound");
    const int nd = c->nd, ldb = c->ldb;
    c->ldn = round_up(nd, 32);
    std::vector<double> dynp((size_t)nd * c->ldn, 0.0), mp((size_t)nd * c->ldn, 0.0);
    for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) dynp[(size_t)i * c->ldn + j] = c->dyn[(size_t)i * nd + j];

    size_t ws_doubles = (size_t)7 * nd * ldb;
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        PYMD_REQUIRE(H.set, "pymd_step: bath %d not set", b);
        PYMD_REQUIRE(c->sd.b[b].noise, "pymd_step: bath %d has no noise bound (call gnoi)", b);
        PYMD_REQUIRE(c->sd.b[b].cur, "pymd_step: bath %d has no current buffer bound", b);
        PYMD_REQUIRE(H.ml == 1 || c->sd.b[b].ring, "pymd_step: bath %d (ml=%d) has no history bound", b, H.ml);
        fdl_free(H);                              // rebuilt on the first per-step-pipeline call that wants it
        H.fdl_tried = false;
        const int nc = H.nc;
        H.lda = round_up(nc, 32);
        const double alpha = H.dt_scaled ? c->dt : 1.0;
        std::vector<double> head((size_t)nc * H.lda, 0.0), aq((size_t)nc * H.lda, 0.0);
        std::vector<int> inv(nd, -1);
        for (int i = 0; i < nc; ++i) inv[H.cids[i]] = i;
        for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j) {
                const double v = alpha * H.kernel[(size_t)i * nc + j];
                head[(size_t)i * H.lda + j] = v;
                mp[(size_t)H.cids[i] * c->ldn + H.cids[j]] += v;
                if (H.has_aq) aq[(size_t)i * H.lda + j] = H.aq[(size_t)i * nc + j];
            }
        int rc;
        if ((rc = upload(&H.d_head, head))) return rc;
        if (H.has_aq && (rc = upload(&H.d_aq, aq))) return rc;
        {
            // fragment-ordered copies for the warp-specialised step kernel (register-resident taps)
            const int MT = H.ncp / 8, KQ = H.ncp / 4;
            std::vector<double> hdf((size_t)MT * KQ * 32, 0.0), kinf((size_t)kBlockT * MT * KQ * 32, 0.0);
            for (int i = 0; i < nc; ++i)
                for (int j = 0; j < nc; ++j) {
                    const size_t fo = ((size_t)(i / 8) * KQ + j / 4) * 32 + (i % 8) * 4 + j % 4;
                    hdf[fo] = alpha * H.kernel[(size_t)i * nc + j];
                    for (int k = 1; k <= kBlockT && k < H.ml; ++k)
                        kinf[(size_t)(k - 1) * MT * KQ * 32 + fo] = c->dt * H.kernel[((size_t)k * nc + i) * nc + j];
                }
            if ((rc = upload(&H.d_hdf, hdf))) return rc;
            if ((rc = upload(&H.d_kinf, kinf))) return rc;
        }
        if ((rc = upload_i(&H.d_cids, H.cids))) return rc;
        if ((rc = upload_i(&H.d_inv, inv))) return rc;
        if (H.ml > 1) {
            // reversed, zero-padded kernel table: Krev[i][(kappa+padf)*ncp + j] = K[ml-1-kappa][i][j]
            H.padf = kBlockT + 2;
            const int backpad = 32 / H.ncp + 3;
            H.Lk = (long long)(H.padf + H.ml + backpad) * H.ncp;
            std::vector<double> krev((size_t)nc * H.Lk, 0.0);
            for (int k = 1; k < H.ml; ++k) {
                const int kappa = H.ml - 1 - k;
                for (int i = 0; i < nc; ++i)
                    for (int j = 0; j < nc; ++j)
                        krev[(size_t)i * H.Lk + (size_t)(kappa + H.padf) * H.ncp + j] =
                            H.kernel[((size_t)k * nc + i) * nc + j];
            }
            if ((rc = upload(&H.d_krev, krev))) return rc;
            std::vector<double> kin((size_t)kBlockT * nc * nc, 0.0);
            for (int k = 1; k <= kBlockT && k < H.ml; ++k)
                for (int e = 0; e < nc * nc; ++e)
                    kin[(size_t)(k - 1) * nc * nc + e] = c->dt * H.kernel[(size_t)k * nc * nc + e];
            if ((rc = upload(&H.d_kintra, kin))) return rc;
            if (far_supported(H)) {
                if ((rc = far_build(c, b))) return rc;
            } else if (H.d_ghat) {
                cudaFree(H.d_ghat); H.d_ghat = nullptr;
            }
            ws_doubles += (size_t)2 * kBlockT * nc * ldb;
        }
        ws_doubles += (size_t)2 * H.ncp * ldb;
    }
    int rc;
    if ((rc = upload(&c->d_dyn, dynp))) return rc;
    if ((rc = upload(&c->d_mp, mp))) return rc;
    c->mp_host = mp;
    if (c->ncon > 0 && (rc = upload(&c->d_con, c->con))) return rc;

    c->ws_bytes = ws_doubles * sizeof(double) + (size_t)ldb * sizeof(int);
    if (c->ws) cudaFree(c->ws);
    PYMD_CUDA_CHECK(cudaMalloc(&c->ws, c->ws_bytes));
    PYMD_CUDA_CHECK(cudaMemset(c->ws, 0, c->ws_bytes));
    double* w = static_cast<double*>(c->ws);
    StepDev& d = c->sd;
    const size_t nb = (size_t)nd * ldb;
    d.ph = w; w += nb;
    d.qn = w; w += nb;
    d.p1 = w; w += nb;
    d.fpotn = w; w += nb;
    d.fpotc = w; w += nb;
    d.fbase = w; w += nb;
    d.h = w; w += nb;
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        BathDev& B = d.b[b];
        B.nc = H.nc; B.ncp = H.ncp; B.ml = H.ml; B.R = H.R;
        B.has_aq = H.has_aq; B.nonlocal = H.ml > 1;
        B.inv = H.d_inv; B.cids = H.d_cids;
        B.g = w; w += (size_t)H.ncp * ldb;
        B.uq = w; w += (size_t)H.ncp * ldb;
        if (H.ml > 1) {
            B.tail_cur = w; w += (size_t)kBlockT * H.nc * ldb;
            B.tail_next = w; w += (size_t)kBlockT * H.nc * ldb;
            H.d_tailA = B.tail_cur;
            H.d_tailB = B.tail_next;
        } else {
            B.tail_cur = B.tail_next = nullptr;
        }
    }
    d.samef = reinterpret_cast<int*>(w);
    d.con = c->d_con;
    d.nd = nd; d.ldb = ldb; d.nbath = c->nbath; d.nmd = c->nmd; d.ncon = c->ncon; d.dt = c->dt;
    if (c->d_aqT) { cudaFree(c->d_aqT); c->d_aqT = nullptr; }
    if (c->d_con3) { cudaFree(c->d_con3); c->d_con3 = nullptr; }
    c->finalized = true;
    c->fpot_valid = false;
    c->fpotc_valid = false;
    c->tail_valid = false;
    return PYMD_OK;
}

// Y[M][ldb] = alpha * A[M][lda] X(rows gathered by idx)        (plain operands)
static GemmArgs plain_args(pymd_ctx* c, const double* A, int lda, int M, int K, const double* X, const int* xidx, double* Y,
                           double alpha, int accumulate) {
    GemmArgs g{};
    g.a_base = A; g.a_ld = lda; g.a_ls = 0; g.a_mod = M;
    g.x_base = X; g.x_idx = xidx; g.x_ld = c->ldb;
    g.y_base = Y; g.y_idx = nullptr; g.y_ld = c->ldb;
    g.M = M; g.N = c->ldb;
    g.nseg = 1; g.seg_lo[0] = 0; g.seg_hi[0] = K; g.seg_aoff[0] = 0;
    g.alpha = alpha; g.accumulate = accumulate;
    return g;
}
static int gemm_plain(pymd_ctx* c, const double* A, int lda, int M, int K, const double* X,
                      const int* xidx, double* Y, double alpha, int accumulate, cudaStream_t st) {
    const GemmArgs g = plain_args(c, A, lda, M, K, X, xidx, Y, alpha, accumulate);
    c->launches++;
    prof_begin(c, 2, st);
    const int rc = launch_gemm(g, st);
    prof_end(c, st);
    return rc;
}
// the products collected in m as ONE launch (profile class cls)
static int flush_multi(pymd_ctx* c, GemmMulti& m, int cls, cudaStream_t st) {
    if (m.n == 0) return PYMD_OK;
    c->launches++;
    prof_begin(c, cls, st);
    const int rc = launch_gemm_multi(m, st);
    prof_end(c, st);
    m.n = 0;
    return rc;
}

// out[s][i][n] = dt * sum_{lag>=0} K[s + koff + lag][i][:] . ring(newest - lag)[:][n],  s < T
// wmax >= 0: only the newest wmax ring rows; accumulate: out += product.
// operands of that product; false when no history row contributes (nothing to launch)
static bool tail_args(pymd_ctx* c, int b, int T, int koff, double* out, long long wmax, int accumulate, GemmArgs& g) {
    BathHost& H = c->bath[b];
    const int R = H.R, ncp = H.ncp;
    long long W = H.ml - koff;                 // lags 0 .. ml-1-koff contribute
    if (wmax >= 0 && W > wmax) W = wmax;
    if (W > H.hcount) W = H.hcount;
    if (W > R) W = R;
    if (W <= 0) return false;
    const int h = (int)((H.hcount - 1) % R);
    const int lo = h - (int)(W - 1);
    g = GemmArgs{};
    g.a_base = H.d_krev; g.a_ld = H.Lk; g.a_ls = -ncp; g.a_mod = H.nc;
    g.x_base = c->sd.b[b].ring; g.x_idx = nullptr; g.x_ld = c->ldb;
    g.y_base = out; g.y_idx = nullptr; g.y_ld = c->ldb;
    g.M = T * H.nc; g.N = c->ldb;
    const long long base = (long long)(H.ml - 1 - koff - h + H.padf) * ncp;
    if (lo >= 0) {
        g.nseg = 1;
        g.seg_lo[0] = lo * ncp; g.seg_hi[0] = (h + 1) * ncp; g.seg_aoff[0] = base;
    } else {
        g.nseg = 2;
        g.seg_lo[0] = 0; g.seg_hi[0] = (h + 1) * ncp; g.seg_aoff[0] = base;
        g.seg_lo[1] = (lo + R) * ncp; g.seg_hi[1] = R * ncp; g.seg_aoff[1] = base - (long long)R * ncp;
    }
    g.alpha = c->dt; g.accumulate = accumulate;
    return true;
}

int launch_tail(pymd_ctx* c, int b, int T, int koff, double* out, cudaStream_t st, long long wmax, int accumulate) {
    GemmArgs g{};
    if (!tail_args(c, b, T, koff, out, wmax, accumulate, g)) {
        if (!accumulate)
            PYMD_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)T * c->bath[b].nc * c->ldb * sizeof(double), st));
        return PYMD_OK;
    }
    c->launches++;
    prof_begin(c, 0, st);
    const int rc = launch_gemm(g, st);
    prof_end(c, st);
    return rc;
}

static int compute_fpotc(pymd_ctx* c, cudaStream_t st) {
    return gemm_plain(c, c->d_dyn, c->ldn, c->nd, c->nd, c->sd.q, nullptr, c->sd.fpotc, -1.0, 0, st);
}

static int generic_step(pymd_ctx* c, long long t0, int nsteps, cudaStream_t st) {
    StepDev& d = c->sd;
    const dim3 blk(TX, TY), grd(c->ldb / TX);
    int rc;
    if (!c->fpot_valid) {
        if ((rc = compute_fpotc(c, st))) return rc;
        fill_int_kernel<<<(c->ldb + 255) / 256, 256, 0, st>>>(d.samef, 0, c->ldb);
        c->launches++;
        c->fpot_valid = true;
        c->fpotc_valid = true;
    }
    if (c->ncon > 0 && !c->fpotc_valid) {      // coming from the fused path: refresh -K q(t)
        if ((rc = compute_fpotc(c, st))) return rc;
        c->fpotc_valid = true;
    }
    if (!c->tail_valid) {
        for (int b = 0; b < c->nbath; ++b)
            if (c->bath[b].ml > 1 && (rc = launch_tail(c, b, 1, 1, d.b[b].tail_cur, st))) return rc;
        c->tail_valid = true;
    }
    // Memory-kernel tails in blocks of GT steps: the part of tail(t0+s+1) that only needs history older
    // than the block, P[s] = dt sum_{k>=s+2} K[k] p_c(t0+s+1-k), is ONE block-Toeplitz GEMM with GT*nc
    // rows per bath (the velocity history is read once per GT steps instead of every step); the rows that
    // enter during the block are added per step by a small product over the s+1 newest ring rows.
    constexpr int GT = 8;
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        if (H.ml > 1 && !H.fdl_tried) {
            H.fdl_tried = true;
            if (fdl_wanted(c, H) && (rc = fdl_build(c, b))) return rc;
        }
    }
    for (int s0 = 0, T = 0; s0 < nsteps; s0 += T) {
        T = std::min(GT, nsteps - s0);
        // long memories: rows older than the current super-block of L steps come from the frequency-domain delay
        // line (fdl.cu); a block of GT steps never straddles two super-blocks
        long long r0[PYMD_MAX_BATHS] = {};
        for (int b = 0; b < c->nbath; ++b) {
            BathHost& H = c->bath[b];
            if (H.ml <= 1 || !H.fdl_L) continue;
            r0[b] = H.hcount - H.fdl_base;
            if (!H.fdl_valid || r0[b] < 0 || r0[b] >= H.fdl_L) {
                if ((rc = fdl_advance(c, b, st))) return rc;
                r0[b] = 0;
            }
            T = std::min<long long>(T, H.fdl_L - r0[b]);
        }
        double* Pblk[PYMD_MAX_BATHS] = {};
        for (int b = 0; b < c->nbath; ++b) {
            BathHost& H = c->bath[b];
            if (H.ml <= 1) continue;
            // the buffer that does not hold tail(t) of the step about to start
            const bool inA = d.b[b].tail_cur >= H.d_tailA && d.b[b].tail_cur < H.d_tailA + (size_t)kBlockT * H.nc * c->ldb;
            Pblk[b] = inA ? H.d_tailB : H.d_tailA;
            if (H.fdl_L) {
                PYMD_CUDA_CHECK(cudaMemcpyAsync(Pblk[b], H.d_fdlF + (size_t)r0[b] * H.nc * c->ldb,
                                                (size_t)T * H.nc * c->ldb * sizeof(double), cudaMemcpyDeviceToDevice, st));
                if ((rc = launch_tail(c, b, T, 2, Pblk[b], st, r0[b], 1))) return rc;
            } else if ((rc = launch_tail(c, b, T, 2, Pblk[b], st))) return rc;
        }
        for (int s = s0; s < s0 + T; ++s) {
            const long long t = t0 + s;
            const int tn = (int)(t % c->nmd), tn1 = (int)((t + 1) % c->nmd);
            Slots slots{};
            // heads of the first force evaluation: g_b = alpha K_b[0] p_c - Aq q_c.  The small products of a step
            // that do not depend on each other share a launch (gemm_multi_kernel).
            GemmMulti mm{};
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (mm.n == GemmMulti::kMax && (rc = flush_multi(c, mm, 2, st))) return rc;
                mm.a[mm.n++] = plain_args(c, H.d_head, H.lda, H.nc, H.nc, d.p, H.d_cids, d.b[b].g, 1.0, 0);
                slots.s[b] = H.ml > 1 ? (int)(H.hcount % H.R) : 0;
            }
            if ((rc = flush_multi(c, mm, 2, st))) return rc;
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (H.has_aq &&
                    (rc = gemm_plain(c, H.d_aq, H.lda, H.nc, H.nc, d.q, H.d_cids, d.b[b].g, -1.0, 1, st)))
                    return rc;
            }
            prof_begin(c, 3, st);
            stage_a_kernel<<<grd, blk, 0, st>>>(d, tn, slots);
            prof_end(c, st);
            c->launches++;
            for (int b = 0; b < c->nbath; ++b)
                if (c->bath[b].ml > 1) c->bath[b].hcount++;
            // shared tail of the 2nd/3rd evaluation and of the next step's first one: pre-block part
            // plus the taps k = 1 .. s-s0+1 on the rows appended since the block began
            for (int b = 0; b < c->nbath; ++b) {
                if (c->bath[b].ml <= 1) continue;
                double* row = Pblk[b] + (size_t)(s - s0) * c->bath[b].nc * c->ldb;
                if (mm.n == GemmMulti::kMax && (rc = flush_multi(c, mm, 0, st))) return rc;
                if (tail_args(c, b, 1, 1, row, s - s0 + 1, 1, mm.a[mm.n])) ++mm.n;
                d.b[b].tail_next = row;
            }
            if ((rc = flush_multi(c, mm, 0, st))) return rc;
            // potential force at q' (the only dyn product of the step when unconstrained), the q-coupled operators
            // at q' and the summed heads of the second evaluation
            mm.a[mm.n++] = plain_args(c, c->d_dyn, c->ldn, c->nd, c->nd, d.qn, nullptr, d.fpotn, -1.0, 0);
            mm.a[mm.n++] = plain_args(c, c->d_mp, c->ldn, c->nd, c->nd, d.ph, nullptr, d.h, 1.0, 0);
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (!H.has_aq) continue;
                if (mm.n == GemmMulti::kMax && (rc = flush_multi(c, mm, 2, st))) return rc;
                mm.a[mm.n++] = plain_args(c, H.d_aq, H.lda, H.nc, H.nc, d.qn, H.d_cids, d.b[b].uq, 1.0, 0);
            }
            if ((rc = flush_multi(c, mm, 2, st))) return rc;
            prof_begin(c, 3, st);
            stage_b_kernel<<<grd, blk, 0, st>>>(d, tn1);
            prof_end(c, st);
            if ((rc = gemm_plain(c, c->d_mp, c->ldn, c->nd, c->nd, d.p1, nullptr, d.h, 1.0, 0, st))) return rc;
            prof_begin(c, 3, st);
            stage_c_kernel<<<grd, blk, 0, st>>>(d);
            prof_end(c, st);
            c->launches += 2;
            if (c->ncon > 0 && (rc = compute_fpotc(c, st))) return rc;
            for (int b = 0; b < c->nbath; ++b)
                if (c->bath[b].ml > 1) d.b[b].tail_cur = d.b[b].tail_next;
        }
    }
    // leave tail(t) at the base of buffer A and buffer B free (the fused path uses them that way)
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        if (H.ml <= 1) continue;
        if (d.b[b].tail_cur != H.d_tailA)
            PYMD_CUDA_CHECK(cudaMemcpyAsync(H.d_tailA, d.b[b].tail_cur, (size_t)H.nc * c->ldb * sizeof(double),
                                            cudaMemcpyDeviceToDevice, st));
        d.b[b].tail_cur = H.d_tailA;
        d.b[b].tail_next = H.d_tailB;
    }
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace pymd

// =============================================================================== C ABI
using namespace pymd;

extern "C" {

const char* pymd_last_error(void) { return g_err; }
int pymd_version(void) { return 100; }

int pymd_bath_ncp(int nc) { return (nc + 7) / 8 * 8; }
int pymd_ring_len(int ml, int block) { return ml <= 1 ? 0 : ml - 1 + (block > kBlockT ? block : kBlockT); }

int pymd_ctx_create(int nd, int ldb, int nbath, double dt, int nmd, int device, pymd_ctx** out) {
    PYMD_REQUIRE(out, "ctx_create: null out");
    PYMD_REQUIRE(nd > 0 && ldb > 0 && ldb % 32 == 0, "ctx_create: nd=%d ldb=%d (ldb must be a multiple of 32)", nd, ldb);
    PYMD_REQUIRE(nbath >= 0 && nbath <= PYMD_MAX_BATHS, "ctx_create: nbath=%d (max %d)", nbath, PYMD_MAX_BATHS);
    PYMD_REQUIRE(nmd > 0 && dt > 0, "ctx_create: nmd=%d dt=%g", nmd, dt);
    int ndev = 0;
    PYMD_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    PYMD_REQUIRE(device >= 0 && device < ndev, "ctx_create: device %d of %d", device, ndev);
    PYMD_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    PYMD_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    PYMD_REQUIRE(prop.major == 10, "libpymd_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    pymd_ctx* c = new pymd_ctx();
    c->nd = nd; c->ldb = ldb; c->nbath = nbath; c->dt = dt; c->nmd = nmd; c->device = device;
    *out = c;
    return PYMD_OK;
}

int pymd_ctx_destroy(pymd_ctx* c) {
    if (!c) return PYMD_OK;
    for (auto& H : c->bath) {
        cudaFree(H.d_cids); cudaFree(H.d_inv); cudaFree(H.d_head); cudaFree(H.d_aq);
        fdl_free(H);
        cudaFree(H.d_krev); cudaFree(H.d_kintra); cudaFree(H.d_ghat); cudaFree(H.d_far); cudaFree(H.d_kmid); cudaFree(H.d_kinf); cudaFree(H.d_hdf);
    }
    cudaFree(c->d_dyn); cudaFree(c->d_mp); cudaFree(c->d_con); cudaFree(c->d_aqT); cudaFree(c->d_con3); cudaFree(c->ws);
    delete c;
    return PYMD_OK;
}

int pymd_set_dyn(pymd_ctx* c, const double* dyn) {
    PYMD_REQUIRE(c && dyn, "set_dyn: null argument");
    c->dyn.assign(dyn, dyn + (size_t)c->nd * c->nd);
    c->finalized = false;
    return PYMD_OK;
}

int pymd_set_constraints(pymd_ctx* c, const double* con, int n) {
    PYMD_REQUIRE(c, "set_constraints: null ctx");
    PYMD_REQUIRE(n >= 0 && n <= PYMD_MAX_CONSTRAINTS, "set_constraints: n=%d (max %d)", n, PYMD_MAX_CONSTRAINTS);
    PYMD_REQUIRE(n == 0 || con, "set_constraints: null vectors");
    c->ncon = n;
    c->con.assign((size_t)n * c->nd, 0.0);
    for (int i = 0; i < n; ++i) {
        double nn = 0.0;
        for (int j = 0; j < c->nd; ++j) nn += con[(size_t)i * c->nd + j] * con[(size_t)i * c->nd + j];
        nn = sqrt(nn);
        for (int j = 0; j < c->nd; ++j)
            c->con[(size_t)i * c->nd + j] = nn != 0.0 ? con[(size_t)i * c->nd + j] / nn : con[(size_t)i * c->nd + j];
    }
    c->finalized = false;
    return PYMD_OK;
}

int pymd_bath_set(pymd_ctx* c, int b, int nc, const int* cids, int ml, const double* kernel,
                  const double* aq, int dt_scaled) {
    PYMD_REQUIRE(c && cids && kernel, "bath_set: null argument");
    PYMD_REQUIRE(b >= 0 && b < c->nbath, "bath_set: bath index %d of %d", b, c->nbath);
    PYMD_REQUIRE(nc > 0 && nc <= c->nd && ml >= 1, "bath_set: nc=%d ml=%d", nc, ml);
    PYMD_REQUIRE(!(aq && ml != 1), "bath_set: the q-coupled operator needs ml == 1 (ebath.py:195-197)");
    BathHost& H = c->bath[b];
    std::vector<char> seen(c->nd, 0);
    for (int i = 0; i < nc; ++i) {
        PYMD_REQUIRE(cids[i] >= 0 && cids[i] < c->nd && !seen[cids[i]], "bath_set: bad/duplicate cid %d", cids[i]);
        seen[cids[i]] = 1;
    }
    H.set = true; H.nc = nc; H.ncp = pymd_bath_ncp(nc); H.ml = ml; H.dt_scaled = dt_scaled;
    H.has_aq = aq != nullptr;
    H.cids.assign(cids, cids + nc);
    H.kernel.assign(kernel, kernel + (size_t)ml * nc * nc);
    if (aq) H.aq.assign(aq, aq + (size_t)nc * nc); else H.aq.clear();
    c->finalized = false;
    return PYMD_OK;
}

int pymd_bind_state(pymd_ctx* c, double* p, double* q) {
    PYMD_REQUIRE(c && p && q, "bind_state: null argument");
    c->sd.p = p; c->sd.q = q;
    c->fpot_valid = false;
    return PYMD_OK;
}

int pymd_bind_history(pymd_ctx* c, int b, double* ring, int ring_len) {
    PYMD_REQUIRE(c && ring, "bind_history: null argument");
    PYMD_REQUIRE(b >= 0 && b < c->nbath && c->bath[b].set, "bind_history: bath %d not set", b);
    PYMD_REQUIRE(ring_len >= pymd_ring_len(c->bath[b].ml, 0), "bind_history: ring_len %d < %d", ring_len,
                 pymd_ring_len(c->bath[b].ml, 0));
    c->sd.b[b].ring = ring;
    c->bath[b].R = ring_len;
    c->bath[b].hcount = 0;
    c->bath[b].fdl_valid = false;
    c->finalized = false;
    return PYMD_OK;
}

int pymd_bind_noise(pymd_ctx* c, int b, const double* noise) {
    PYMD_REQUIRE(c && noise && b >= 0 && b < c->nbath, "bind_noise: bad argument");
    c->sd.b[b].noise = noise;
    c->sd.b[b].noise_ld = 0;
    return PYMD_OK;
}

int pymd_bind_noise_strided(pymd_ctx* c, int b, const double* noise, long long row_stride) {
    PYMD_REQUIRE(c && noise && b >= 0 && b < c->nbath, "bind_noise_strided: bad argument");
    PYMD_REQUIRE(row_stride >= (long long)c->bath[b].nc * c->ldb && row_stride % 2 == 0, "bind_noise_strided: row stride %lld", row_stride);
    c->sd.b[b].noise = noise;
    c->sd.b[b].noise_ld = row_stride == (long long)c->bath[b].nc * c->ldb ? 0 : row_stride;
    return PYMD_OK;
}

int pymd_set_noise_wrap_row(pymd_ctx* c, int row) {
    PYMD_REQUIRE(c && (row == 0 || row == c->nmd), "set_noise_wrap_row: row %d (0 or nmd)", row);
    c->sd.noise_wrap = row;
    return PYMD_OK;
}

int pymd_bind_outputs(pymd_ctx* c, double* ps, double* qs, double* etot) {
    PYMD_REQUIRE(c && etot, "bind_outputs: etot is required");
    c->sd.ps = ps; c->sd.qs = qs; c->sd.etot = etot;
    c->sd.rec_ld = 0;
    return PYMD_OK;
}

int pymd_bind_outputs_strided(pymd_ctx* c, double* ps, double* qs, double* etot, long long row_stride) {
    PYMD_REQUIRE(c && etot && ps && qs, "bind_outputs_strided: null argument");
    PYMD_REQUIRE(row_stride >= (long long)c->nd * c->ldb && row_stride % 2 == 0, "bind_outputs_strided: row stride %lld", row_stride);
    c->sd.ps = ps; c->sd.qs = qs; c->sd.etot = etot;
    c->sd.rec_ld = row_stride == (long long)c->nd * c->ldb ? 0 : row_stride;
    return PYMD_OK;
}

int pymd_bind_bath_outputs(pymd_ctx* c, int b, double* cur, double* fhis) {
    PYMD_REQUIRE(c && cur && b >= 0 && b < c->nbath, "bind_bath_outputs: bad argument");
    c->sd.b[b].cur = cur; c->sd.b[b].fhis = fhis;
    return PYMD_OK;
}

int pymd_reset_history(pymd_ctx* c, void* stream) {
    PYMD_REQUIRE(c, "reset_history: null ctx");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        if (H.ml > 1 && c->sd.b[b].ring)
            PYMD_CUDA_CHECK(cudaMemsetAsync(c->sd.b[b].ring, 0, (size_t)H.R * H.ncp * c->ldb * sizeof(double), st));
        H.hcount = 0;
        H.far_valid = false;
        H.fdl_valid = false;
    }
    c->tail_valid = false;
    return PYMD_OK;
}

int pymd_invalidate_cache(pymd_ctx* c) {
    PYMD_REQUIRE(c, "invalidate_cache: null ctx");
    c->fpot_valid = false;
    return PYMD_OK;
}

static int history_copy(pymd_ctx* c, int b, double* dense, int nrows, int to_ring, void* stream) {
    PYMD_REQUIRE(c && dense && b >= 0 && b < c->nbath, "history: bad argument");
    BathHost& H = c->bath[b];
    PYMD_REQUIRE(H.ml > 1 && c->sd.b[b].ring, "history: bath %d keeps no history", b);
    PYMD_REQUIRE(nrows > 0 && nrows <= H.R, "history: nrows=%d (ring %d)", nrows, H.R);
    if (to_ring) {
        PYMD_CUDA_CHECK(cudaMemsetAsync(c->sd.b[b].ring, 0, (size_t)H.R * H.ncp * c->ldb * sizeof(double),
                                        static_cast<cudaStream_t>(stream)));
        H.hcount = nrows;
        H.far_valid = false;
        H.fdl_valid = false;
        c->tail_valid = false;
    }
    history_copy_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        c->sd.b[b].ring, dense, H.R, H.ncp, H.nc, c->ldb, H.hcount, nrows, to_ring);
    c->launches++;
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}
int pymd_history_get(pymd_ctx* c, int b, double* out, int nrows, void* stream) {
    return history_copy(c, b, out, nrows, 0, stream);
}
int pymd_history_set(pymd_ctx* c, int b, const double* in, int nrows, void* stream) {
    return history_copy(c, b, const_cast<double*>(in), nrows, 1, stream);
}

int pymd_step(pymd_ctx* c, long long t0, int nsteps, int mode, void* stream) {
    PYMD_REQUIRE(c && nsteps >= 0 && t0 >= 0, "step: bad argument");
    PYMD_REQUIRE(mode >= 0 && mode <= 2, "step: mode %d", mode);
    PYMD_CUDA_CHECK(cudaSetDevice(c->device));
    int rc = finalize(c);
    if (rc) return rc;
    if (nsteps == 0) return PYMD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool strided = c->sd.rec_ld != 0 || c->sd.noise_wrap != 0;
    for (int b = 0; b < c->nbath; ++b) strided = strided || c->sd.b[b].noise_ld != 0;
    if (strided) {
        // overlaid records (ps / qs rows written over consumed noise rows): only the two-unit on-chip kernel reads
        // and writes them, and a call must not run past the end of the noise table
        PYMD_REQUIRE(mode != 1 && fused_supported(c), "step: strided record / noise layouts need the on-chip step kernel");
        PYMD_REQUIRE(c->sd.noise_wrap == 0 || (t0 % c->nmd) + nsteps <= c->nmd,
                     "step: with overlaid records a call must end at the end of the noise table (t0=%lld nsteps=%d nmd=%d)", t0, nsteps, c->nmd);
        return fused_step(c, t0, nsteps, st);
    }
    if (mode == 0) {
        mode = (fused_supported(c) || local_supported(c)) ? 2 : 1;
        // very long memories: the far field of the on-chip path re-transforms one 97-row window per 97 taps and launch,
        // the delay line of the per-step pipeline keeps the block spectra; measured at nd = 63, 4096 trajectories:
        // ml = 1000: 58 M (on-chip) vs 37 M traj-steps/s, ml = 4000: 17 M vs 30 M -- the lines cross near ml = 1800
        if (mode == 2 && fused_supported(c))
            for (int b = 0; b < c->nbath; ++b)
                if (c->bath[b].ml >= 2000 && fdl_wanted(c, c->bath[b])) mode = 1;
    }
    if (mode == 2) {
        // on-chip step kernels: fused.cu (nd <= 64), local.cu (64 < nd <= 96, one bath without memory)
        if (fused_supported(c)) return fused_step(c, t0, nsteps, st);
        PYMD_REQUIRE(local_supported(c), "step: the on-chip step kernels do not support this configuration");
        return local_step(c, t0, nsteps, st);
    }
    return generic_step(c, t0, nsteps, st);
}

long long pymd_launch_count(const pymd_ctx* c) { return c ? c->launches : 0; }

int pymd_profile_enable(pymd_ctx* c, int on) {
    PYMD_REQUIRE(c, "profile_enable: null ctx");
    c->prof = on != 0;
    return PYMD_OK;
}

int pymd_profile_read(pymd_ctx* c, double* ms4, long long* count4) {
    PYMD_REQUIRE(c && ms4 && count4, "profile_read: null argument");
    PYMD_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < PYMD_PROFILE_CLASSES; ++i) { ms4[i] = 0.0; count4[i] = 0; }
    for (auto& r : c->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.cls >= 0 && r.cls < PYMD_PROFILE_CLASSES) {
            ms4[r.cls] += ms;
            count4[r.cls]++;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    c->prof_recs.clear();
    return PYMD_OK;
}

int pymd_gemm(const double* a, int lda, const double* x, double* y, int M, int K, int ldb, double alpha,
              int accumulate, void* stream) {
    PYMD_REQUIRE(a && x && y && lda % 32 == 0 && lda >= K, "gemm: lda=%d must be a multiple of 32 and >= K=%d", lda, K);
    GemmArgs g{};
    g.a_base = a; g.a_ld = lda; g.a_ls = 0; g.a_mod = M;
    g.x_base = x; g.x_idx = nullptr; g.x_ld = ldb;
    g.y_base = y; g.y_idx = nullptr; g.y_ld = ldb;
    g.M = M; g.N = ldb;
    g.nseg = 1; g.seg_lo[0] = 0; g.seg_hi[0] = K; g.seg_aoff[0] = 0;
    g.alpha = alpha; g.accumulate = accumulate;
    return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

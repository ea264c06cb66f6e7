// On-chip step kernel for Markovian junctions with 64 < nd <= 96 degrees of freedom (mode 3):
// ONE bath without memory (ml == 1: the electron bath of examples/Alkane.py, ebath.py:180-199, or a
// Debye-damped phonon bath, phbath.py:86) coupled to any subset of the dofs.  Without memory kernel
// there is no history, so a launch covers a whole pymd_step call of any length.
//
// Same round structure as fused.cu (three sequential force evaluations per md.vv step, md.py:315-358),
// with 12 warps owning 8 operator rows each:
//   * Mp = scatter(alpha K[0]) (used three times per step) lives in registers as DMMA A fragments,
//     24 k-steps; dyn and the q-coupled operator AqT (used once per step) live in shared memory with a
//     row pitch of 100 doubles (= 4 mod 16: conflict-free A-fragment loads);
//   * head velocity / q' vectors in shared memory [96][N+4], N = 8 NT trajectories per CTA;
//   * kinetic energy and the bath current: one reduce-scatter over the warp's 8 rows per step, per-warp
//     slots summed in a fixed order after the round-A barrier (deterministic, no atomics);
//   * one constraint vector in closed form (md.py:739-762), -D q and AqT q corrected by linearity, the
//     sameq rule of md.py:733 kept per trajectory.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "engine.cuh"

namespace pymd {

int prepare_scattered_operands(pymd_ctx* c, int aq_bath);

namespace {

constexpr int LW = 12;        // warps per CTA
constexpr int LNDP = 96;      // padded rows (8 per warp)
constexpr int LKS = 24;       // k-steps of the nd x nd operators
constexpr int LPA = 100;      // row pitch of the shared-memory operators (doubles)

struct LocalParams {
    int nd, ldb, ldn, nmd, ncon, nsteps, nc, has_bath, carry_valid;
    long long t0;
    double dt, cmax;
    double *p, *q, *ps, *qs, *etot, *fpotn, *cur;
    int* samef;
    const double *dyn, *mp, *aqT, *con3, *noise;
    const int* inv;
};

__device__ __forceinline__ double2 lld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void lst2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// Reduce-scatter over the 8 lanes sharing lane%4 (lr = lane>>2): lane lr returns sum_lanes v[lr].
__device__ __forceinline__ double lrs8(const double (&v)[8], int lr) {
    const bool h4 = lr & 4, h2 = lr & 2, h1 = lr & 1;
    double k[4], m[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double keep = h4 ? v[i + 4] : v[i], send = h4 ? v[i] : v[i + 4];
        k[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double keep = h2 ? k[i + 2] : k[i], send = h2 ? k[i] : k[i + 2];
        m[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double keep = h1 ? m[1] : m[0], send = h1 ? m[0] : m[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

template <int NT, bool HASQ, bool HASCON>
__global__ void __launch_bounds__(LW * 32, 1) local_kernel(const LocalParams P) {
    constexpr int N = 8 * NT, LD = N + 4;
    extern __shared__ __align__(16) double sm[];
    double* opD = sm;                                   // [96][100]
    double* opQ = opD + LNDP * LPA;                     // [96][100] (HASQ)
    double* xb = opQ + (HASQ ? LNDP * LPA : 0);         // xp0, xp1, xq: [96][LD] each
    double* curp = xb + 3 * LNDP * LD;                  // [2][LW][N]   kinetic energy, bath current
    double* conp = curp + 2 * LW * N;                   // [8][LW][4]   constraint dots
    double* conv = conp + 8 * LW * 4;                   // [3][96]      c, D c, AqT c
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const int row = 8 * w + lr, col0 = blockIdx.x * N, nd = P.nd;
    const size_t ldb = P.ldb;
    const double dt = P.dt, dt2 = P.dt * P.dt;
    const bool live = row < nd;
    const int own = row * LD + 2 * lc;

    // ------------------------------------------------------------------ prologue
    for (int e = tid; e < LNDP * LPA; e += LW * 32) {
        const int i = e / LPA, j = e % LPA;
        const bool in = i < nd && j < nd;
        opD[e] = in ? P.dyn[(size_t)i * P.ldn + j] : 0.0;
        if (HASQ) opQ[e] = in ? P.aqT[(size_t)i * P.ldn + j] : 0.0;
    }
    for (int e = tid; e < 3 * LNDP * LD + 2 * LW * N + 8 * LW * 4 + 3 * LNDP; e += LW * 32) xb[e] = 0.0;
    __syncthreads();
    if (HASCON)
        for (int e = tid; e < 3 * nd; e += LW * 32) conv[(e / nd) * LNDP + e % nd] = P.con3[e];
    double aM[LKS];
#pragma unroll
    for (int k = 0; k < LKS; ++k) aM[k] = live ? P.mp[(size_t)row * P.ldn + 4 * k + lc] : 0.0;   // ldn >= 96: zero padded
    const int ip = (live && P.has_bath) ? P.inv[row] : -1;       // index of this row in the bath, or -1
    const double member = ip >= 0 ? 1.0 : 0.0;
    int xpc = 0, xpn = LNDP * LD;
    const int xq = 2 * LNDP * LD;
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double2 pv = lld2(&P.p[(size_t)row * ldb + col0 + 8 * j + 2 * lc]);
            const double2 qv = lld2(&P.q[(size_t)row * ldb + col0 + 8 * j + 2 * lc]);
            lst2(&xb[xpc + own + 8 * j], pv.x, pv.y);
            lst2(&xb[xq + own + 8 * j], qv.x, qv.y);
        }
    }
    __syncthreads();

    auto noise_row = [&](int trow, double (&z)[NT][2]) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            z[j][0] = z[j][1] = 0.0;
            if (ip >= 0) {
                const double2 v = lld2(&P.noise[((size_t)trow * P.nc + ip) * ldb + col0 + 8 * j + 2 * lc]);
                z[j][0] = v.x; z[j][1] = v.y;
            }
        }
    };
    // acc[j] += (rows of this warp of an operator in shared memory) . X[:, 8j..8j+7]
    auto matvec_smem2 = [&](const double* X, double (&a0)[NT][2], double (&a1)[NT][2]) {
        const double* Ad = opD + row * LPA + lc;
        const double* Aq = opQ + row * LPA + lc;
#pragma unroll
        for (int k = 0; k < LKS; ++k) {
            const double ad = Ad[4 * k];
            const double aq = HASQ ? Aq[4 * k] : 0.0;
            const double* xr = X + (4 * k + lc) * LD + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = xr[8 * j];
                dmma884(a0[j][0], a0[j][1], ad, b);
                if (HASQ) dmma884(a1[j][0], a1[j][1], aq, b);
            }
        }
    };
    auto matvec_reg = [&](const double* X, double (&acc)[NT][2]) {
#pragma unroll
        for (int k = 0; k < LKS; ++k) {
            const double* xr = X + (4 * k + lc) * LD + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[j][0], acc[j][1], aM[k], xr[8 * j]);
        }
    };

    double fpot[NT][2], uq[NT][2];              // -D q and AqT q at the current q, own rows
#pragma unroll
    for (int j = 0; j < NT; ++j) fpot[j][0] = fpot[j][1] = uq[j][0] = uq[j][1] = 0.0;
    matvec_smem2(xb + xq, fpot, uq);
#pragma unroll
    for (int j = 0; j < NT; ++j) { fpot[j][0] = -fpot[j][0]; fpot[j][1] = -fpot[j][1]; }
    if (HASCON && P.carry_valid && live) {      // md.potforce's cache (md.py:456-457) across launches
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = col0 + 8 * j + 2 * lc;
            if (P.samef[n]) fpot[j][0] = P.fpotn[(size_t)row * ldb + n];
            if (P.samef[n + 1]) fpot[j][1] = P.fpotn[(size_t)row * ldb + n + 1];
        }
    }
    __syncthreads();                            // every warp has read q before round A overwrites it

    // ------------------------------------------------------------------ time loop
    int tn = (int)(P.t0 % P.nmd);
    for (int s = 0; s < P.nsteps; ++s) {
        const int tn1 = tn + 1 == P.nmd ? 0 : tn + 1;
        // ============ round A: first force evaluation at time t, head p(t) ============
        double hA[NT][2], nz[NT][2];
        noise_row(tn, nz);
#pragma unroll
        for (int j = 0; j < NT; ++j) hA[j][0] = hA[j][1] = 0.0;
        matvec_reg(xb + xpc, hA);
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double2 pv = live ? lld2(&xb[xpc + own + 8 * j]) : make_double2(0.0, 0.0);
            const double2 qv = lld2(&xb[xq + own + 8 * j]);
            // bath force on this row: noise - alpha K[0] p (+ Aq q): all of Mp p and AqT q belong to the one bath
            const double g0 = nz[j][0] + uq[j][0] - hA[j][0], g1 = nz[j][1] + uq[j][1] - hA[j][1];
            const double f0 = fpot[j][0] + g0, f1 = fpot[j][1] + g1;
            if (live) {
                lst2(&xb[xpn + own + 8 * j], pv.x + f0 * dt / 2.0, pv.y + f1 * dt / 2.0);
                lst2(&xb[xq + own + 8 * j], qv.x + pv.x * dt + f0 * dt2 / 2.0, qv.y + pv.y * dt + f1 * dt2 / 2.0);
                if (P.ps) lst2(&P.ps[((size_t)tn * nd + row) * ldb + col0 + 8 * j + 2 * lc], pv.x, pv.y);
                if (P.qs) lst2(&P.qs[((size_t)tn * nd + row) * ldb + col0 + 8 * j + 2 * lc], qv.x, qv.y);
            }
            // reduce 4 values per n-tile (kinetic energy and current, two columns): two n-tiles per call
            const int h = j & 1;
            v[4 * h + 0] = pv.x * pv.x; v[4 * h + 1] = pv.y * pv.y;
            v[4 * h + 2] = member * pv.x * g0; v[4 * h + 3] = member * pv.y * g1;
            if (h == 1 || j == NT - 1) {
                const double r = lrs8(v, lr);            // lane lr: n-tile (j & ~1) + (lr >> 2), kind lr & 3
                const int jj = (j & ~1) + (lr >> 2), kind = lr & 3;
                if (jj < NT) curp[((kind >> 1) * LW + w) * N + 8 * jj + 2 * lc + (kind & 1)] = r;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.0;
            }
        }
        __syncthreads();                                                        // ---- S2
        for (int e = tid; e < 2 * N; e += LW * 32) {
            const int q = e / N, n = e % N;
            double sum = 0.0;
#pragma unroll
            for (int ww = 0; ww < LW; ++ww) sum += curp[(q * LW + ww) * N + n];
            if (q == 0) P.etot[(size_t)tn * ldb + col0 + n] = 0.5 * sum;
            else if (P.has_bath) P.cur[(size_t)tn * ldb + col0 + n] = sum;
        }
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1/2)

        // ============ round B: second evaluation at time t+1, head p(t+1/2) ============
        double fb[NT][2];
        {
            double aDq[NT][2], hB[NT][2];
            noise_row(tn1, nz);
#pragma unroll
            for (int j = 0; j < NT; ++j) aDq[j][0] = aDq[j][1] = hB[j][0] = hB[j][1] = uq[j][0] = uq[j][1] = 0.0;
            matvec_smem2(xb + xq, aDq, uq);
            matvec_reg(xb + xpc, hB);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                fpot[j][0] = -aDq[j][0];
                fpot[j][1] = -aDq[j][1];
                fb[j][0] = fpot[j][0] + uq[j][0] + nz[j][0];
                fb[j][1] = fpot[j][1] + uq[j][1] + nz[j][1];
                const double2 ph = lld2(&xb[xpc + own + 8 * j]);
                if (live) lst2(&xb[xpn + own + 8 * j], ph.x + dt * (fb[j][0] - hB[j][0]) / 2.0, ph.y + dt * (fb[j][1] - hB[j][1]) / 2.0);
            }
        }
        __syncthreads();                                                        // ---- S3
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p1, xpn: p(t+1/2)

        // ============ round C: third evaluation, head p1 -> p(t+1), constraint ============
        {
            double hC[NT][2], pn[NT][2], qn[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) hC[j][0] = hC[j][1] = 0.0;
            matvec_reg(xb + xpc, hC);
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 ph = lld2(&xb[xpn + own + 8 * j]);
                pn[j][0] = live ? ph.x + dt * (fb[j][0] - hC[j][0]) / 2.0 : 0.0;
                pn[j][1] = live ? ph.y + dt * (fb[j][1] - hC[j][1]) / 2.0 : 0.0;
                const double2 q0 = lld2(&xb[xq + own + 8 * j]);
                qn[j][0] = q0.x; qn[j][1] = q0.y;
            }
            if (HASCON) {
                const double cv = conv[row], dv = conv[LNDP + row], av = HASQ ? conv[2 * LNDP + row] : 0.0;
                constexpr int NG = (4 * NT + 7) / 8;
                double tot[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    double u[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int vi = 8 * g + i, j = vi >> 2, kind = vi & 3;       // d0, d1, g0, g1 of n-tile j
                        u[i] = j < NT ? (kind < 2 ? pn[j < NT ? j : 0][kind] : qn[j < NT ? j : 0][kind - 2]) * cv : 0.0;
                    }
                    conp[((8 * g + lr) * LW + w) * 4 + lc] = lrs8(u, lr);
                }
                __syncthreads();                                                // ---- S4
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    double sum = 0.0;
#pragma unroll
                    for (int ww = 0; ww < LW; ++ww) sum += conp[((8 * g + lr) * LW + ww) * 4 + lc];
                    tot[g] = sum;
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    double d[4];
#pragma unroll
                    for (int kind = 0; kind < 4; ++kind) {
                        const int vi = 4 * j + kind;
                        d[kind] = __shfl_sync(0xffffffffu, tot[vi >> 3], ((vi & 7) << 2) | lc);
                    }
                    pn[j][0] -= d[0] * cv;
                    pn[j][1] -= d[1] * cv;
                    qn[j][0] -= d[2] * cv;
                    qn[j][1] -= d[3] * cv;
                    // md.sameq (md.py:724-736): the cached force is kept while max_i |q(t+1)_i - q'_i| < 1e-9
                    if (!(fabs(d[2]) * P.cmax < 10e-10)) fpot[j][0] += d[2] * dv;
                    if (!(fabs(d[3]) * P.cmax < 10e-10)) fpot[j][1] += d[3] * dv;
                    if (HASQ) { uq[j][0] -= d[2] * av; uq[j][1] -= d[3] * av; }
                }
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    lst2(&xb[xpn + own + 8 * j], pn[j][0], pn[j][1]);
                    if (HASCON) lst2(&xb[xq + own + 8 * j], qn[j][0], qn[j][1]);
                }
            }
        }
        __syncthreads();                                                        // ---- S1
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1)
        tn = tn1;
    }

    // ------------------------------------------------------------------ epilogue
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            const double2 pv = lld2(&xb[xpc + own + 8 * j]);
            const double2 qv = lld2(&xb[xq + own + 8 * j]);
            lst2(&P.p[(size_t)row * ldb + col0 + n], pv.x, pv.y);
            lst2(&P.q[(size_t)row * ldb + col0 + n], qv.x, qv.y);
            lst2(&P.fpotn[(size_t)row * ldb + col0 + n], fpot[j][0], fpot[j][1]);
        }
    }
    for (int n = tid; n < N; n += LW * 32) P.samef[col0 + n] = 1;     // the stored force is the force at q(t)
}

template <int NT, bool HASQ, bool HASCON>
int launch_local(const LocalParams& lp, int grid, cudaStream_t st) {
    constexpr int N = 8 * NT, LD = N + 4;
    const size_t smem = ((size_t)(HASQ ? 2 : 1) * LNDP * LPA + 3 * LNDP * LD + 2 * LW * N + 8 * LW * 4 + 3 * LNDP) * sizeof(double);
    static bool cfg = false;
    if (!cfg) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(local_kernel<NT, HASQ, HASCON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = true;
    }
    local_kernel<NT, HASQ, HASCON><<<grid, LW * 32, smem, st>>>(lp);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

template <int NT>
int launch_local_nt(const LocalParams& lp, int grid, bool hasq, bool hascon, cudaStream_t st) {
    if (hasq && hascon) return launch_local<NT, true, true>(lp, grid, st);
    if (hasq) return launch_local<NT, true, false>(lp, grid, st);
    if (hascon) return launch_local<NT, false, true>(lp, grid, st);
    return launch_local<NT, false, false>(lp, grid, st);
}

}  // namespace

int local_supported(const pymd_ctx* c) {
    if (c->nd <= 64 || c->nd > LNDP || c->ncon > 1 || c->nbath > 1) return 0;
    if (c->nbath == 1 && (c->bath[0].ml != 1 || c->sd.b[0].fhis)) return 0;
    return 1;
}

int local_step(pymd_ctx* c, long long t0, int nsteps, cudaStream_t st) {
    StepDev& d = c->sd;
    const bool has_bath = c->nbath == 1;
    const bool hasq = has_bath && c->bath[0].has_aq;
    int rc = prepare_scattered_operands(c, hasq ? 0 : -1);
    if (rc) return rc;
    LocalParams lp{};
    lp.nd = c->nd; lp.ldb = c->ldb; lp.ldn = c->ldn; lp.nmd = c->nmd; lp.ncon = c->ncon; lp.nsteps = nsteps;
    lp.has_bath = has_bath ? 1 : 0;
    lp.nc = has_bath ? c->bath[0].nc : 0;
    lp.carry_valid = c->fpot_valid ? 1 : 0;
    lp.t0 = t0; lp.dt = c->dt;
    lp.cmax = 0.0;
    for (int i = 0; i < (c->ncon > 0 ? c->nd : 0); ++i) lp.cmax = std::max(lp.cmax, std::fabs(c->con[i]));
    lp.p = d.p; lp.q = d.q; lp.ps = d.ps; lp.qs = d.qs; lp.etot = d.etot; lp.fpotn = d.fpotn; lp.samef = d.samef;
    lp.cur = has_bath ? d.b[0].cur : nullptr;
    lp.dyn = c->d_dyn; lp.mp = c->d_mp; lp.aqT = c->d_aqT; lp.con3 = c->d_con3;
    lp.noise = has_bath ? d.b[0].noise : nullptr;
    lp.inv = has_bath ? c->bath[0].d_inv : nullptr;
    // trajectory tile: 16 per CTA unless 8 fills the SMs better
    const char* force = getenv("PYMD_FUSED_NT");
    int NT = (c->ldb % 16 == 0 && c->ldb / 16 >= 148) ? 2 : 1;
    if (force && (atoi(force) == 1 || (atoi(force) == 2 && c->ldb % 16 == 0))) NT = atoi(force);
    prof_begin(c, 1, st);
    rc = NT == 2 ? launch_local_nt<2>(lp, c->ldb / 16, hasq, c->ncon > 0, st)
                 : launch_local_nt<1>(lp, c->ldb / 8, hasq, c->ncon > 0, st);
    prof_end(c, st);
    if (rc) return rc;
    c->launches++;
    c->fpot_valid = true;
    c->fpotc_valid = false;
    return PYMD_OK;
}

}  // namespace pymd

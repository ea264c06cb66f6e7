// Fused on-chip step kernel (mode 2): up to 32 consecutive md.vv steps (reference md.py:315-358)
// per launch, in blocks of FT = 8, for a tile of 8*NT trajectories per CTA, without leaving the SM.
//
//   * The three nd x nd operators -- dyn (md.py:484), the summed k=0 bath heads
//     Mp = sum_b scatter(alpha_b K_b[0]) (phbath.py:197-201 / ebath.py:191,194) and the q-coupled
//     electron operator AqT (ebath.py:192-193) -- are held in REGISTERS as DMMA A-fragments: warp w
//     owns rows 8w..8w+7, so one mat-vec round is ks DMMA.8x8x4 per 8 trajectories per warp (all 16
//     k-steps, branch-free, when nd > 48: FULLK).
//   * State vectors (head velocity, q', half-kick, base force) live in shared memory, [64][N+4]
//     doubles; the B-fragment read (ld % 16 == 4) and the C-fragment write are conflict free.
//   * Memory-kernel friction on three time scales:
//       - taps that meet the history of the current block: contracted here from an in-smem block
//         history by the "tap" work items;
//       - taps that meet earlier blocks of the current super-block of 32 steps: mid_phase() at
//         every block start (mid mode), ring rows staged from HBM, taps prefetched from L2;
//       - everything older: the far field F of far.cu (FFT convolution, one launch per super-block),
//         added in mid_phase().
//     Without a far field (short memories, nc > 16, PYMD_FAR=0) one launch covers one block and the
//     block-Toeplitz GEMM (launch_tail, koff = 2) has produced the pre-block part P[s] beforehand.
//   * The three force evaluations of a step are sequential (each head depends on the previous
//     result): three rounds A/B/C separated by one __syncthreads each; the head buffer is
//     double buffered so no extra barrier protects it.
//   * Per-bath heat currents (md.py:342) need the per-bath force only through scalars
//     p.g_b; they are produced as warp-shuffle column reductions into per-warp slots and summed
//     in a fixed order (deterministic, no atomics).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "fused.cuh"

namespace pymd {

int launch_tail(pymd_ctx* c, int b, int T, int koff, double* out, cudaStream_t st, long long wmax = -1,
                int accumulate = 0);
// fused3.cu: two-unit variant of the warp-specialised kernel
size_t plan3(const pymd_ctx* c, FusedParams& fp, bool hasq, int NJ);
int launch_f3(const FusedParams& fp, size_t smem, int NJ, bool hasq, bool hascon, cudaStream_t st);

namespace {

// acc[j] += A(rows of this warp) . X[:, 8j..8j+7]   (A fragments in registers, X in smem)
// FULLK: all MAXKS k-steps without the per-step branch (operator columns >= nd are zero, the padding
// rows of X finite), so that the compiler can overlap the B loads of one k-step with the DMMAs of
// the previous one.
template <int NT, bool FULLK>
__device__ __forceinline__ void matvec1(const double (&a)[MAXKS], int ks, const double* X, int LD,
                                        double (&acc)[NT][2], int lr, int lc) {
#pragma unroll
    for (int k = 0; k < MAXKS; ++k) {
        if (FULLK || k < ks) {
            const double* xr = X + (4 * k + lc) * LD + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[j][0], acc[j][1], a[k], xr[8 * j]);
        }
    }
}
// two operators sharing one B operand
template <int NT, bool FULLK>
__device__ __forceinline__ void matvec2(const double (&a0)[MAXKS], const double (&a1)[MAXKS], int ks,
                                        const double* X, int LD, double (&acc0)[NT][2],
                                        double (&acc1)[NT][2], int lr, int lc) {
#pragma unroll
    for (int k = 0; k < MAXKS; ++k) {
        if (FULLK || k < ks) {
            const double* xr = X + (4 * k + lc) * LD + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double b = xr[8 * j];
                dmma884(acc0[j][0], acc0[j][1], a0[k], b);
                dmma884(acc1[j][0], acc1[j][1], a1[k], b);
            }
        }
    }
}

// Pre-block tail of one block of up to FT steps, computed in the kernel (mid mode):
//   P[s] = F[r0b+s] + dt sum_{a=0}^{W-1} K[s+2+a] p_c(t0-1-a),   W = rows since p(u0-1) (far.cu covers the rest).
// Warp w owns step s = w: A fragments (dt K[w+2+a], pre-swizzled on the host) come from L2 one ring
// row ahead of their use, the history rows are staged from the HBM ring through the free head
// buffer, 64/ncp rows at a time, with the next stage prefetched into registers while the current
// one is multiplied.  From the second block of a launch on, the seven newest rows p_c(t0-1 .. t0-7)
// are still in the shared-memory block history (slots 7 .. 1) and are multiplied in place, without
// staging or barriers.  MT x KQ = m8 x k4 fragments of one ncp x ncp tap.
template <int NT, int MT, int KQ, int NTH, bool COMPUTE>
__device__ __forceinline__ void mid_bath(const FusedParams& P, const FusedBath& B, double* stage, const double* hist,
                                         int nsblk, int g, int col0) {
    constexpr int N = 8 * NT, LD = N + 4, ncp = 8 * MT, nA = MT * KQ;
    constexpr int spst = NDP / ncp;                     // ring rows per stage
    constexpr int per_slot = ncp * (N / 2);             // double2 per ring-row tile
    constexpr int NPF = (spst * per_slot + NTH - 1) / (NTH);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const size_t ldb = P.ldb;
    const int r0b = P.r0 + g, slotb = (B.slot0 + g) % B.R;
    long long Wl = r0b + 1;
    if (Wl > P.hc0 + g) Wl = P.hc0 + g;
    if (Wl > B.ml - 2) Wl = B.ml - 2;
    const int W = (int)Wl;
    const bool mine = COMPUTE && w < nsblk;     // COMPUTE = false (bath warps of fused2): this warp only keeps the barriers
    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[mt][j][0] = acc[mt][j][1] = 0.0;
    // this thread's share of a stage: element e = tid + i * 256 -> (ring row sl, bath row j, column pair c2)
    double2 pf[NPF];
    auto fetch = [&](int a0) {
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
            const int e = tid + i * NTH;
            const int sl = e / per_slot, j = (e % per_slot) / (N / 2), c2 = e % (N / 2);
            pf[i] = make_double2(0.0, 0.0);
            if (COMPUTE && e < spst * per_slot && a0 + sl < W) {
                int slot = slotb - 1 - a0 - sl;
                if (slot < 0) slot += B.R;
                pf[i] = __ldcg(reinterpret_cast<const double2*>(&B.ring[((size_t)slot * ncp + j) * ldb + col0 + 2 * c2]));
            }
        }
    };
    const double* Anext = B.kmid + ((size_t)w * nA) * 32 + lane;
    double ac[nA], an[nA];
    const int nh = g > 0 ? min(W, FT - 1) : 0;          // rows taken from the block history
    if (W > nh) fetch(nh);
    if (W > 0 && mine) {
#pragma unroll
        for (int i = 0; i < nA; ++i) an[i] = __ldg(Anext + i * 32);
        Anext += nA * 32;
    }
    // products of one ring row whose B fragments start at H
    auto row_products = [&](const double* H, bool more) {
#pragma unroll
        for (int i = 0; i < nA; ++i) ac[i] = an[i];
        if (more) {
#pragma unroll
            for (int i = 0; i < nA; ++i) an[i] = __ldg(Anext + i * 32);
            Anext += nA * 32;
        }
#pragma unroll
        for (int kq = 0; kq < KQ; ++kq) {
            double bf[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = H[(4 * kq) * LD + 8 * j];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[mt][j][0], acc[mt][j][1], ac[mt * KQ + kq], bf[j]);
        }
    };
    if (mine)
        for (int a = 0; a < nh; ++a) row_products(hist + ((FT - 1 - a) * ncp + lc) * LD + lr, a + 1 < W);
    for (int a0 = nh; a0 < W; a0 += spst) {
        __syncthreads();                                // previous stage fully consumed
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
            const int e = tid + i * NTH;
            const int sl = e / per_slot, j = (e % per_slot) / (N / 2), c2 = e % (N / 2);
            if (COMPUTE && e < spst * per_slot) *reinterpret_cast<double2*>(&stage[(sl * ncp + j) * LD + 2 * c2]) = pf[i];
        }
        __syncthreads();
        if (a0 + spst < W) fetch(a0 + spst);
        if (mine) {
            const int na = min(spst, W - a0);
            for (int sl = 0; sl < na; ++sl) row_products(stage + (sl * ncp + lc) * LD + lr, a0 + sl + 1 < W);
        }
    }
    if (mine) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int rr = mt * 8 + lr;
            if (rr < B.nc) {
                const double* Fr = B.F + ((size_t)(r0b + w) * B.nc + rr) * ldb + col0 + 2 * lc;
                double* Pr = const_cast<double*>(B.P) + ((size_t)w * B.nc + rr) * ldb + col0 + 2 * lc;
                double2 f[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) f[j] = __ldcg(reinterpret_cast<const double2*>(Fr + 8 * j));
#pragma unroll
                for (int j = 0; j < NT; ++j) st2(Pr + 8 * j, f[j].x + acc[mt][j][0], f[j].y + acc[mt][j][1]);
            }
        }
    }
}

template <int NT, int NTH = FW * 32, bool COMPUTE = true>
__device__ __forceinline__ void mid_phase(const FusedParams& P, double* sm, int stage_off, int nsblk, int g, int col0) {
    for (int b = 0; b < P.nbath; ++b) {
        const FusedBath& B = P.b[b];
        if (!B.nonlocal) continue;
        if (B.ncp == 16) mid_bath<NT, 2, 4, NTH, COMPUTE>(P, B, sm + stage_off, sm + B.off_hist, nsblk, g, col0);
        else mid_bath<NT, 1, 2, NTH, COMPUTE>(P, B, sm + stage_off, sm + B.off_hist, nsblk, g, col0);      // far_supported: ncp is 8 or 16
    }
    __syncthreads();                                    // P visible to every warp; staging buffer free again
}

template <int NT, bool HASQ, bool HASCON, bool FULLK>
__global__ void __launch_bounds__(FW * 32, 1) fused_kernel(const FusedParams P) {
    constexpr int N = 8 * NT, LD = N + 4;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const int row = 8 * w + lr;                 // operator row owned by this thread
    const int col0 = blockIdx.x * N;            // first trajectory of this CTA
    const int nd = P.nd, ks = P.ks, nb = P.nbath;
    const size_t ldb = P.ldb;
    const double dt = P.dt, dt2 = P.dt * P.dt;
    const bool live = row < nd;
    const int own = row * LD + 2 * lc;          // this thread's entries: own + 8 j (+0,+1)

    // ------------------------------------------------------------------ prologue
    for (int i = tid; i < P.smem_doubles; i += FW * 32) sm[i] = 0.0;
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
        const FusedBath& B = P.b[b];
        if (B.direct) continue;
        int* CIDS = reinterpret_cast<int*>(sm + B.off_cids);
        int* INREM = reinterpret_cast<int*>(sm + B.off_inrem);
        for (int i = tid; i < B.ncp; i += FW * 32) {
            CIDS[i] = B.cids_g[i < B.nc ? i : B.nc - 1];
            INREM[i] = (P.rem >= 0 && i < B.nc) ? (P.b[P.rem].inv_g[B.cids_g[i]] >= 0) : 0;
        }
        if (b != P.rem) {
            double* HD = sm + B.off_hd;
            for (int e = tid; e < B.nc * B.nc; e += FW * 32)
                HD[(e / B.nc) * B.ldk + e % B.nc] = B.head_g[(size_t)(e / B.nc) * B.lda + e % B.nc];
        }
        if (B.nonlocal) {
            double* KIN = sm + B.off_kin;
            for (int e = tid; e < B.ntap * B.nc * B.nc; e += FW * 32) {
                const int k = e / (B.nc * B.nc), r = (e / B.nc) % B.nc, cc = e % B.nc;
                KIN[(k * B.ncp + r) * B.ldk + cc] = B.kin_g[e];
            }
        }
    }
    if (HASCON)
        for (int e = tid; e < 3 * P.ncon * nd; e += FW * 32) sm[P.off_conv + (e / nd) * NDP + e % nd] = P.con3[e];

    // per-thread bath membership of the owned row: index into the bath's dofs or -1; smem offset of
    // the row's noise-minus-tail entry (a zero row when the dof is not coupled to the bath)
    int ip[FB_MAX], nbo[FB_MAX], nbs[FB_MAX];      // nbs: double-buffer stride (0 for the zero row)
    double mrem = 0.0, maq = 0.0;               // 1.0 when the row belongs to the remainder / q-coupled bath
#pragma unroll
    for (int b = 0; b < FB_MAX; ++b) {
        ip[b] = (b < nb && live) ? P.b[b].inv_g[row] : -1;
        const bool member = ip[b] >= 0 && !P.b[b].direct;
        nbo[b] = member ? P.b[b].off_nb + ip[b] * LD + 2 * lc : P.off_zero + 2 * lc;
        nbs[b] = member ? P.b[b].ncp * LD : 0;
        if (ip[b] >= 0 && b == P.rem) mrem = 1.0;
        if (ip[b] >= 0 && b == P.aq) maq = 1.0;
    }
    // static work-item list of this warp: code = b | mt << 4 | kind << 8.  An item covers one m8 tile of
    // one bath for ALL n-tiles (A fragment loaded once per k-step, NT independent DMMA chains).  The
    // heavier tap items (kind 1) are dealt first, then the heads of the small baths (kind 0).
    // up to MAXIT items of 16 bits each in one register (a local array indexed by the loop counter would
    // live in local memory and cost a long-scoreboard stall per item)
    unsigned long long items = 0;
    int nit = 0;
    bool split_items = false;
    {
        int total = 0;
        for (int b = 0; b < nb; ++b) {
            if (!P.b[b].direct) total += P.b[b].ncp / 8;
            if (b != P.rem) total += P.b[b].ncp / 8;
        }
        split_items = NT >= 2 && !P.nosplit && 2 * total <= FW * MAXIT;
        const int nh = split_items ? 2 : 1;
        int item = 0;
        for (int b = 0; b < nb; ++b) {
            if (P.b[b].direct) continue;
            for (int mt = 0; mt < P.b[b].ncp / 8; ++mt)
                for (int h = 0; h < nh; ++h, ++item)
                    if (item % FW == w && nit < MAXIT)
                        items |= (unsigned long long)(b | (mt << 4) | (1 << 8) | ((h * (NT / 2)) << 12)) << (16 * nit++);
        }
        for (int b = 0; b < nb; ++b) {
            if (b == P.rem) continue;
            for (int mt = 0; mt < P.b[b].ncp / 8; ++mt)
                for (int h = 0; h < nh; ++h, ++item)
                    if (item % FW == w && nit < MAXIT)
                        items |= (unsigned long long)(b | (mt << 4) | ((h * (NT / 2)) << 12)) << (16 * nit++);
        }
    }

    // L2 prefetch of the noise rows two steps ahead (they stream from HBM, 8 nc_b N bytes per bath and
    // step): thread tid owns one 128-byte line of one (bath, row), or nothing
    const double* pf_base = nullptr;
    size_t pf_stride = 0;
    {
        constexpr int LPR = (N * 8 + 127) / 128;           // lines per row tile
        int e = tid;
        for (int b = 0; b < nb; ++b) {
            const int cnt = P.b[b].nc * LPR;
            if (e >= 0 && e < cnt) {
                pf_base = P.b[b].noise + (size_t)(e / LPR) * ldb + col0 + (e % LPR) * 16;
                pf_stride = (size_t)P.b[b].nc * ldb;
            }
            e -= cnt;
        }
    }

    int xpc = P.off_xp0, xpn = P.off_xp1;      // current / next head buffer (offsets into sm)
    const int xq = P.off_xq;
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            const double2 pv = ld2(&P.p[(size_t)row * ldb + col0 + n]);
            const double2 qv = ld2(&P.q[(size_t)row * ldb + col0 + n]);
            st2(&sm[xpc + own + 8 * j], pv.x, pv.y);
            st2(&sm[xq + own + 8 * j], qv.x, qv.y);
            // p_c(t0) opens the block history and is appended to the HBM ring (functions.rpadleft)
#pragma unroll
            for (int b = 0; b < FB_MAX; ++b) {
                if (ip[b] < 0 || P.b[b].direct) continue;
                const FusedBath& B = P.b[b];
                st2(&sm[B.off_hist + ip[b] * LD + n], pv.x, pv.y);
                if (B.nonlocal) st2(&B.ring[((size_t)B.slot0 * B.ncp + ip[b]) * ldb + col0 + n], pv.x, pv.y);
            }
        }
    }
    // noise minus tail at time t0: in smem for every bath but a local remainder bath (registers)
    double nzr[NT][2];
    zero_acc<NT>(nzr);
    {
        const int tn0 = (int)(P.t0 % P.nmd);
        for (int b = 0; b < nb; ++b) {
            const FusedBath& B = P.b[b];
            if (B.direct) {
                if (ip[b] >= 0) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double2 v = ld2(&B.noise[((size_t)tn0 * B.nc + ip[b]) * ldb + col0 + 8 * j + 2 * lc]);
                        nzr[j][0] = v.x; nzr[j][1] = v.y;
                    }
                }
                continue;
            }
            double* NB = sm + B.off_nb;                 // buffer 0
            for (int e = tid; e < B.nc * N; e += FW * 32) {
                const int r = e / N, n = e % N;
                double v = B.noise[((size_t)tn0 * B.nc + r) * ldb + col0 + n];
                if (B.nonlocal) v -= B.tail_cur[(size_t)r * ldb + col0 + n];
                NB[r * LD + n] = v;
            }
        }
    }
    __syncthreads();

    double fpot[NT][2], uq[NT][2];              // -D q and AqT q at the current q, own rows
    zero_acc<NT>(fpot);
    zero_acc<NT>(uq);
    int nbuf = 0;                               // NB buffer holding noise-minus-tail at time t
#ifdef PYMD_FUSED_TIMING
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define TICK(i) { const long long _n = clock64(); tacc[i] += _n - tlast; tlast = _n; }
#else
#define TICK(i)
#endif

    // ------------------------------------------------------------------ time loop
    int tn = (int)(P.t0 % P.nmd);               // noise row of time t (noise is periodic in nmd, phbath.py:196)
    for (int g0 = 0; g0 < P.nsteps; g0 += FT) {
    // ---- block of up to FT steps.  Pre-block tails first (mid mode), while the register file is
    // still free of the operator fragments, which are (re)loaded for every block.
    if (P.mid) {
        TICK(9)
        mid_phase<NT>(P, sm, xpn, min(FT, P.nsteps - g0), g0, col0);
        TICK(8)
    }
    double aD[MAXKS], aM[MAXKS], aQ[MAXKS];
#pragma unroll
    for (int k = 0; k < MAXKS; ++k) {
        const bool ok = k < ks && live;
        aD[k] = ok ? P.dyn[(size_t)row * P.ldn + 4 * k + lc] : 0.0;
        aM[k] = ok ? P.mp[(size_t)row * P.ldn + 4 * k + lc] : 0.0;
        aQ[k] = (HASQ && ok) ? P.aqT[(size_t)row * P.ldn + 4 * k + lc] : 0.0;
    }
    if (g0 == 0) {
        // forces at the initial q
        if (HASQ) matvec2<NT, FULLK>(aD, aQ, ks, sm + xq, LD, fpot, uq, lr, lc);
        else matvec1<NT, FULLK>(aD, ks, sm + xq, LD, fpot, lr, lc);
#pragma unroll
        for (int j = 0; j < NT; ++j) { fpot[j][0] = -fpot[j][0]; fpot[j][1] = -fpot[j][1]; }
        if (HASCON) {
            // constrained runs carry md.potforce's cache (md.py:456-457) across launches
            if (P.carry_valid && live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int n = col0 + 8 * j + 2 * lc;
                    if (P.samef[n]) fpot[j][0] = P.fpotn[(size_t)row * ldb + n];
                    if (P.samef[n + 1]) fpot[j][1] = P.fpotn[(size_t)row * ldb + n + 1];
                }
            }
        }
        __syncthreads();                        // every warp has read q before round A overwrites it
    }
    const int gend = min(P.nsteps, g0 + FT);
    for (int g = g0; g < gend; ++g) {
        const int s = g - g0;                   // step within the block
        const int tn_cur = tn;
        const int tn1 = tn + 1 == P.nmd ? 0 : tn + 1;
        if (pf_base) {
            const int tn2 = tn1 + 1 == P.nmd ? 0 : tn1 + 1;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + (size_t)tn2 * pf_stride));
        }

        // ============ round A: first force evaluation at time t, head p(t) ============
        double hA[NT][2];
        zero_acc<NT>(hA);
        matvec1<NT, FULLK>(aM, ks, sm + xpc, LD, hA, lr, lc);
        TICK(0)
        // stage 1: every shared-memory load and all arithmetic for all n-tiles (no stores in between, so
        // the loads of different n-tiles overlap); stage 2: the stores
        double phv[NT][2], qnv[NT][2], pvv[NT][2], qvv[NT][2], rsum[NT], c3a[NT], c3b[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            // padding rows (row >= nd) of the head buffers may hold staged history rows (mid_phase)
            const double2 pv = live ? ld2(&sm[xpc + own + 8 * j]) : make_double2(0.0, 0.0);
            const double2 qv = ld2(&sm[xq + own + 8 * j]);
            double f0 = fpot[j][0] + uq[j][0] - hA[j][0];
            double f1 = fpot[j][1] + uq[j][1] - hA[j][1];
            double v[8];                         // kinetic energy and up to three bath currents, two columns each
            v[0] = pv.x * pv.x;
            v[1] = pv.y * pv.y;
            c3a[j] = c3b[j] = 0.0;
#pragma unroll
            for (int b = 0; b < FB_MAX; ++b) {
                double2 nv = ld2(&sm[nbo[b] + nbuf * nbs[b] + 8 * j]);
                if (b < nb && P.b[b].direct && ip[b] >= 0) nv = make_double2(nzr[j][0], nzr[j][1]);
                f0 += nv.x;
                f1 += nv.y;
                double c0 = nv.x, c1 = nv.y;
                if (HASQ && b == P.aq) { c0 += maq * uq[j][0]; c1 += maq * uq[j][1]; }
                if (b == P.rem) { c0 -= mrem * hA[j][0]; c1 -= mrem * hA[j][1]; }
                if (b < 3) { v[2 + 2 * b] = pv.x * c0; v[3 + 2 * b] = pv.y * c1; }
                else { c3a[j] = pv.x * c0; c3b[j] = pv.y * c1; }
            }
            pvv[j][0] = pv.x; pvv[j][1] = pv.y;
            qvv[j][0] = qv.x; qvv[j][1] = qv.y;
            phv[j][0] = pv.x + f0 * dt / 2.0;
            phv[j][1] = pv.y + f1 * dt / 2.0;
            qnv[j][0] = qv.x + pv.x * dt + f0 * dt2 / 2.0;
            qnv[j][1] = qv.y + pv.y * dt + f1 * dt2 / 2.0;
            // column sums over this warp's 8 rows: reduce-scatter, lane lr ends with the sum of v[lr]
            rsum[j] = rs8(v, lr);
            if (nb > 3) { c3a[j] = colsum8(c3a[j]); c3b[j] = colsum8(c3b[j]); }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            if (live) {
                st2(&sm[xpn + own + 8 * j], phv[j][0], phv[j][1]);
                st2(&sm[xq + own + 8 * j], qnv[j][0], qnv[j][1]);
                if (P.ps) st2(&P.ps[((size_t)tn_cur * nd + row) * ldb + col0 + n], pvv[j][0], pvv[j][1]);
                if (P.qs) st2(&P.qs[((size_t)tn_cur * nd + row) * ldb + col0 + n], qvv[j][0], qvv[j][1]);
            }
            const int qn = lr >> 1;              // 0: kinetic energy, 1..3: bath qn-1
            if (qn <= nb) sm[P.off_curp + (((qn == 0) ? nb : qn - 1) * FW + w) * N + n + (lr & 1)] = rsum[j];
            if (nb > 3 && lr == 0) st2(&sm[P.off_curp + (3 * FW + w) * N + n], c3a[j], c3b[j]);
        }
        // noise of the direct bath at time t+1 (used from round B on)
#pragma unroll
        for (int b = 0; b < FB_MAX; ++b) {
            if (b < nb && P.b[b].direct && ip[b] >= 0) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double2 vv = ld2(&P.b[b].noise[((size_t)tn1 * P.b[b].nc + ip[b]) * ldb + col0 + 8 * j + 2 * lc]);
                    nzr[j][0] = vv.x; nzr[j][1] = vv.y;
                }
            }
        }
        __syncwarp();
        TICK(1)
        // ---- work items ----
        // An item is one m8 tile of one bath for NJ n-tiles starting at j0: all NT tiles, or half of
        // them when the item list is short enough to be split (every warp then gets a tap half and a
        // head half instead of four warps a whole tap item and four a whole head item).
        auto run_item = [&](auto njc, const int code) {
            constexpr int NJ = decltype(njc)::value;
            const int b = code & 15, mt = (code >> 4) & 15, j0 = (code >> 12) & 15;
            const FusedBath& B = P.b[b];
            const int rr = mt * 8 + lr;
            const bool valid = rr < B.nc;
            const int cb = 8 * j0;                       // first column of the item
            if ((code >> 8) & 1) {
                // noise minus memory-kernel tail at time t+1: xi(t+1) - P[s] - sum_{k=1}^{s+1} dt K[k] p_c(t+1-k)
                double2 xi[NJ], pt[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    xi[j] = make_double2(0.0, 0.0);
                    pt[j] = make_double2(0.0, 0.0);
                    if (valid) {
                        xi[j] = ld2(&B.noise[((size_t)tn1 * B.nc + rr) * ldb + col0 + cb + 8 * j + 2 * lc]);
                        if (B.nonlocal) pt[j] = ld2(&B.P[((size_t)s * B.nc + rr) * ldb + col0 + cb + 8 * j + 2 * lc]);
                    }
                }
                double acc[NJ][2];
                zero_acc<NJ>(acc);
                if (B.nonlocal) {
                    const int ntap = min(s + 1, B.ntap);
                    if (B.ncp == 16) {
                        // the usual lead size (nc <= 16): four k-steps per tap, unrolled, and two
                        // accumulator sets (odd / even taps) to halve the dependent DMMA chains
                        double acc2[NJ][2];
                        zero_acc<NJ>(acc2);
                        const double* A0 = sm + B.off_kin + rr * B.ldk + lc;
                        const double* H0 = sm + B.off_hist + lc * LD + lr + cb;
                        int k = 1;
                        for (; k + 1 <= ntap; k += 2) {
                            const double* A = A0 + (k - 1) * 16 * B.ldk;         // taps k and k + 1
                            const double* H = H0 + (s + 1 - k) * 16 * LD;        // p_c(t+1-k); one slot below: p_c(t-k)
#pragma unroll
                            for (int kk = 0; kk < 16; kk += 4) {
                                const double a0 = A[kk], a1 = A[16 * B.ldk + kk];
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
                                    dmma884(acc[j][0], acc[j][1], a0, H[kk * LD + 8 * j]);
                                    dmma884(acc2[j][0], acc2[j][1], a1, H[(kk - 16) * LD + 8 * j]);
                                }
                            }
                        }
                        if (k <= ntap) {
                            const double* A = A0 + (k - 1) * 16 * B.ldk;
                            const double* H = H0 + (s + 1 - k) * 16 * LD;
#pragma unroll
                            for (int kk = 0; kk < 16; kk += 4) {
                                const double a0 = A[kk];
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma884(acc[j][0], acc[j][1], a0, H[kk * LD + 8 * j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < NJ; ++j) { acc[j][0] += acc2[j][0]; acc[j][1] += acc2[j][1]; }
                    } else {
                        for (int k = 1; k <= ntap; ++k) {
                            const double* A = sm + B.off_kin + ((k - 1) * B.ncp + rr) * B.ldk + lc;
                            const double* H = sm + B.off_hist + ((s + 1 - k) * B.ncp + lc) * LD + lr + cb;
                            for (int kk = 0; kk < B.ncp; kk += 4) {
                                const double a = A[kk];
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma884(acc[j][0], acc[j][1], a, H[kk * LD + 8 * j]);
                            }
                        }
                    }
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const double t0v = pt[j].x + acc[j][0], t1v = pt[j].y + acc[j][1];
                        st2(&sm[B.off_nb + ((nbuf ^ 1) * B.ncp + rr) * LD + cb + 8 * j + 2 * lc], xi[j].x - t0v, xi[j].y - t1v);
                        if (B.nonlocal && g == P.nsteps - 1)
                            st2(&B.tail_cur[(size_t)rr * ldb + col0 + cb + 8 * j + 2 * lc], t0v, t1v);
                    }
                }
            } else {
                // head of a small bath: only the scalars p_c.g_b (currents) are needed
                const int hs = B.nonlocal ? s : 0;
                const int inrem = valid ? reinterpret_cast<const int*>(sm + B.off_inrem)[rr] : 0;
                double gh[NJ][2];
                zero_acc<NJ>(gh);
                const double* A = sm + B.off_hd + rr * B.ldk + lc;
                const double* H = sm + B.off_hist + (hs * B.ncp + lc) * LD + lr + cb;
                if (B.ncp == 16) {
#pragma unroll
                    for (int kk = 0; kk < 16; kk += 4) {
                        const double a = A[kk];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) dmma884(gh[j][0], gh[j][1], a, H[kk * LD + 8 * j]);
                    }
                } else {
                    for (int kk = 0; kk < B.ncp; kk += 4) {
                        const double a = A[kk];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) dmma884(gh[j][0], gh[j][1], a, H[kk * LD + 8 * j]);
                    }
                }
#pragma unroll
                for (int j2 = 0; j2 < NJ; j2 += 2) {
                    double v[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = j2 + h;
                        double s0 = 0.0, s1 = 0.0;
                        if (j < NJ && valid) {
                            const double2 pv = ld2(&sm[B.off_hist + (hs * B.ncp + rr) * LD + cb + 8 * j + 2 * lc]);
                            s0 = pv.x * gh[j < NJ ? j : 0][0];
                            s1 = pv.y * gh[j < NJ ? j : 0][1];
                        }
                        v[4 * h + 0] = s0; v[4 * h + 1] = s1;
                        v[4 * h + 2] = inrem ? s0 : 0.0; v[4 * h + 3] = inrem ? s1 : 0.0;
                    }
                    const double r = rs8(v, lr);     // lane lr: n-tile j2 + (lr>>2), kind lr&3 (s0, s1, r0, r1)
                    const int j = j2 + (lr >> 2), kind = lr & 3;
                    if (j < NJ) {
                        if (kind < 2) sm[P.off_curp + (b * FW + w) * N + cb + 8 * j + 2 * lc + kind] -= r;
                        else if (P.rem >= 0) sm[P.off_curp + (P.rem * FW + w) * N + cb + 8 * j + 2 * lc + kind - 2] += r;
                    }
                    __syncwarp();
                }
            }
        };
        for (int q = 0; q < nit; ++q) {
            const int code = (int)((items >> (16 * q)) & 0xffffu);
            if (NT >= 2 && split_items) run_item(std::integral_constant<int, (NT >= 2 ? NT / 2 : 1)>{}, code);
            else run_item(std::integral_constant<int, NT>{}, code);
        }
        TICK(2)
        __syncthreads();                                                        // ---- S2
        TICK(3)
        // currents and kinetic energy of time t: fixed-order sum of the per-warp slots
        for (int e = tid; e < (nb + 1) * N; e += FW * 32) {
            const int b = e / N, n = e % N;
            double sum = 0.0;
#pragma unroll
            for (int ww = 0; ww < FW; ++ww) sum += sm[P.off_curp + (b * FW + ww) * N + n];
            if (b < nb) P.b[b].cur[(size_t)tn_cur * ldb + col0 + n] = sum;
            else P.etot[(size_t)tn_cur * ldb + col0 + n] = 0.5 * sum;
        }
        nbuf ^= 1;
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1/2), xpn: free (held p(t))

        // ============ round B: second evaluation at time t+1, head p(t+1/2) ============
        double fb[NT][2];
        {
            double aDq[NT][2], hB[NT][2];
            zero_acc<NT>(aDq);
            zero_acc<NT>(hB);
            zero_acc<NT>(uq);
            if (HASQ) matvec2<NT, FULLK>(aD, aQ, ks, sm + xq, LD, aDq, uq, lr, lc);
            else matvec1<NT, FULLK>(aD, ks, sm + xq, LD, aDq, lr, lc);
            matvec1<NT, FULLK>(aM, ks, sm + xpc, LD, hB, lr, lc);
            TICK(4)
            double p1v[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                fpot[j][0] = -aDq[j][0];
                fpot[j][1] = -aDq[j][1];
                double f0 = fpot[j][0] + uq[j][0], f1 = fpot[j][1] + uq[j][1];
#pragma unroll
                for (int b = 0; b < FB_MAX; ++b) {
                    double2 nv = ld2(&sm[nbo[b] + nbuf * nbs[b] + 8 * j]);
                    if (b < nb && P.b[b].direct && ip[b] >= 0) nv = make_double2(nzr[j][0], nzr[j][1]);
                    f0 += nv.x;
                    f1 += nv.y;
                }
                fb[j][0] = f0;
                fb[j][1] = f1;
                const double2 ph = ld2(&sm[xpc + own + 8 * j]);
                p1v[j][0] = ph.x + dt * (f0 - hB[j][0]) / 2.0;
                p1v[j][1] = ph.y + dt * (f1 - hB[j][1]) / 2.0;
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) st2(&sm[xpn + own + 8 * j], p1v[j][0], p1v[j][1]);
            }
        }
        TICK(5)
        __syncthreads();                                                        // ---- S3
        TICK(3)
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p1, xpn: still holds p(t+1/2)

        // ============ round C: third evaluation, head p1 -> p(t+1), constraint ============
        {
            double hC[NT][2];
            zero_acc<NT>(hC);
            matvec1<NT, FULLK>(aM, ks, sm + xpc, LD, hC, lr, lc);
            TICK(6)
            double pn[NT][2], qn[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double2 ph = ld2(&sm[xpn + own + 8 * j]);
                pn[j][0] = live ? ph.x + dt * (fb[j][0] - hC[j][0]) / 2.0 : 0.0;
                pn[j][1] = live ? ph.y + dt * (fb[j][1] - hC[j][1]) / 2.0 : 0.0;
                if (HASCON) {
                    const double2 q0 = ld2(&sm[xq + own + 8 * j]);
                    qn[j][0] = q0.x; qn[j][1] = q0.y;
                }
            }
            if (HASCON) {
                // md.ApplyConstraint (md.py:739-762) with one constraint vector: the dots c.p', c.q' are
                // warp-reduced with a reduce-scatter, summed over the 8 warps after one barrier
                const double cv = sm[P.off_conv + row];
                const double dv = sm[P.off_conv + NDP + row];
                const double av = HASQ ? sm[P.off_conv + 2 * NDP + row] : 0.0;
                constexpr int NG = (4 * NT + 7) / 8;
                double tot[NG], part[NG];
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int vi = 8 * g + i, j = vi >> 2, kind = vi & 3;       // d0, d1, g0, g1 of n-tile j
                        double x = 0.0;
                        if (j < NT) x = (kind < 2 ? pn[j < NT ? j : 0][kind] : qn[j < NT ? j : 0][kind - 2]) * cv;
                        v[i] = x;
                    }
                    part[g] = rs8(v, lr);
                }
#pragma unroll
                for (int g = 0; g < NG; ++g) sm[P.off_conp + ((8 * g + lr) * FW + w) * 4 + lc] = part[g];
                TICK(7)
                __syncthreads();                                                // ---- S4
                TICK(3)
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    double sum = 0.0;
#pragma unroll
                    for (int ww = 0; ww < FW; ++ww) sum += sm[P.off_conp + ((8 * g + lr) * FW + ww) * 4 + lc];
                    tot[g] = sum;
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    // all-gather the four totals of n-tile j from the lanes that own them
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
                    // md.sameq (md.py:724-736): max_i |q(t+1)_i - q'_i| = |dq| max_i |c_i|; when it reaches 1e-9
                    // the cached force is dropped: -D q(t+1) = -D q' + dq (D c) by linearity
                    if (!(fabs(d[2]) * P.cmax < 10e-10)) fpot[j][0] += d[2] * dv;
                    if (!(fabs(d[3]) * P.cmax < 10e-10)) fpot[j][1] += d[3] * dv;
                    if (HASQ) { uq[j][0] -= d[2] * av; uq[j][1] -= d[3] * av; }      // AqT q(t+1)
                }
            }
            if (live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    st2(&sm[xpn + own + 8 * j], pn[j][0], pn[j][1]);
                    if (HASCON) st2(&sm[xq + own + 8 * j], qn[j][0], qn[j][1]);
                }
                // p_c(t+1) enters the block history / the HBM ring
#pragma unroll
                for (int b = 0; b < FB_MAX; ++b) {
                    if (ip[b] < 0 || P.b[b].direct) continue;
                    const FusedBath& B = P.b[b];
                    if (!B.nonlocal) {
#pragma unroll
                        for (int j = 0; j < NT; ++j) st2(&sm[B.off_hist + ip[b] * LD + 8 * j + 2 * lc], pn[j][0], pn[j][1]);
                    } else if (g + 1 < P.nsteps) {
                        const int hb = B.off_hist + (((s + 1) & (FT - 1)) * B.ncp + ip[b]) * LD + 2 * lc;
                        int slot = (B.slot0 + g + 1) % B.R;
                        double* rp = &B.ring[((size_t)slot * B.ncp + ip[b]) * ldb + col0 + 2 * lc];
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            st2(&sm[hb + 8 * j], pn[j][0], pn[j][1]);
                            st2(rp + 8 * j, pn[j][0], pn[j][1]);
                        }
                    }
                }
            }
        }
        TICK(7)
        __syncthreads();                                                        // ---- S1
        TICK(3)
        { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1)
        tn = tn1;
    }
    }

    // ------------------------------------------------------------------ epilogue
#ifdef PYMD_FUSED_TIMING
    if (P.timing && lane == 0)
        for (int i = 0; i < 10; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&P.timing[i]), (unsigned long long)tacc[i]);
#endif
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            const double2 pv = ld2(&sm[xpc + own + 8 * j]);
            const double2 qv = ld2(&sm[xq + own + 8 * j]);
            st2(&P.p[(size_t)row * ldb + col0 + n], pv.x, pv.y);
            st2(&P.q[(size_t)row * ldb + col0 + n], qv.x, qv.y);
            st2(&P.fpotn[(size_t)row * ldb + col0 + n], fpot[j][0], fpot[j][1]);
        }
    }
    for (int n = tid; n < N; n += FW * 32) P.samef[col0 + n] = 1;     // the stored force is the force at q(t)
}

// =====================================================================================================
// fused2: warp-specialised variant of the fused step kernel for N = 32 trajectories per CTA.
//   * 8 ROW warps run the three rounds of a step.  Only Mp lives in their registers; dyn and AqT are
//     fragment-ordered shared-memory operands [m8 tile][k4 step][lane] (one conflict-free LDS.64 each).
//   * 4 BATH warps each own one (bath, m8 tile) pair: the in-block taps dt K[1..8] and the head alpha K[0]
//     of that pair are REGISTER fragments (no shared-memory tables), and while the row warps run round A
//     they produce noise-minus-tail of the next time for their 8 bath rows and the head scalars of the
//     currents; after the round-A barrier they sum the per-warp current / energy slots and store them.
//   * barriers: S2 and S1 join all 12 warps; S3 and the constraint barrier are named barriers of the
//     256 row threads only.
// Used when every non-direct bath has ncp <= 16 and there are at most 4 (bath, m8 tile) pairs.
constexpr int RW2 = 8, TW2 = 12, NTH2 = TW2 * 32;
// register re-balancing between the two row-warp groups and the bath group (setmaxnreg works on groups of four
// warps): 2 x F2_ROW_REGS + F2_BATH_REGS <= 3 x 168
#ifndef F2_SETMAXNREG
#define F2_SETMAXNREG 0
#endif
#ifndef F2_ROW_REGS
#define F2_ROW_REGS 192
#endif
#ifndef F2_BATH_REGS
#define F2_BATH_REGS 120
#endif

__device__ __forceinline__ void row_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, 384;" ::: "memory"); }

// NBX: baths the per-row loops run over (3 when the junction has at most three baths, else FB_MAX)
template <int NT, bool HASQ, bool HASCON, int NBX>
__global__ void __launch_bounds__(NTH2, 1) fused2_kernel(const FusedParams P) {
    constexpr int N = 8 * NT, LD = N + 4;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const bool roww = w < RW2;
    const int row = 8 * (w & 7) + lr;
    const int col0 = blockIdx.x * N;
    const int nd = P.nd, nb = P.nbath;
    const size_t ldb = P.ldb;
    const double dt = P.dt, dt2 = P.dt * P.dt;
    const bool live = roww && row < nd;
    const int own = row * LD + 2 * lc;

    // ------------------------------------------------------------------ prologue (all warps)
    for (int i = tid; i < P.smem_doubles; i += NTH2) sm[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < 8 * MAXKS * 32; e += NTH2) {
        const int mt = e / (MAXKS * 32), k = (e / 32) % MAXKS, ln = e % 32;
        const int r = 8 * mt + (ln >> 2), cidx = 4 * k + (ln & 3);
        const bool in = r < nd && cidx < nd;
        sm[P.off_opD + e] = in ? P.dyn[(size_t)r * P.ldn + cidx] : 0.0;
        if (HASQ) sm[P.off_opQ + e] = in ? P.aqT[(size_t)r * P.ldn + cidx] : 0.0;
    }
    for (int b = 0; b < nb; ++b) {
        const FusedBath& B = P.b[b];
        if (B.direct) continue;
        int* INREM = reinterpret_cast<int*>(sm + B.off_inrem);
        for (int i = tid; i < B.ncp; i += NTH2)
            INREM[i] = (P.rem >= 0 && i < B.nc) ? (P.b[P.rem].inv_g[B.cids_g[i]] >= 0) : 0;
    }
    if (HASCON)
        for (int e = tid; e < 3 * P.ncon * nd; e += NTH2) sm[P.off_conv + (e / nd) * NDP + e % nd] = P.con3[e];

    int xpc = P.off_xp0, xpn = P.off_xp1;
    const int xq = P.off_xq;
    // noise minus tail at time t0 of the small baths (all threads)
    {
        const int tn0 = (int)(P.t0 % P.nmd);
        for (int b = 0; b < nb; ++b) {
            const FusedBath& B = P.b[b];
            if (B.direct) continue;
            double* NB = sm + B.off_nb;
            for (int e = tid; e < B.nc * N; e += NTH2) {
                const int r = e / N, n = e % N;
                double v = B.noise[((size_t)tn0 * B.nc + r) * ldb + col0 + n];
                if (B.nonlocal) v -= B.tail_cur[(size_t)r * ldb + col0 + n];
                NB[r * LD + n] = v;
            }
        }
    }
    int nbuf = 0;
    int tn = (int)(P.t0 % P.nmd);
#ifdef PYMD_FUSED_TIMING
    long long t2acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t2last = clock64();
#define TICK2(i) { const long long _n = clock64(); t2acc[i] += _n - t2last; t2last = _n; }
#else
#define TICK2(i)
#endif

    if (roww) {
    // ================================================================== ROW WARPS
#if F2_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F2_ROW_REGS));      // the bath warp group hands registers over
#endif
    // ---- row-warp state
    double aM[MAXKS];
#pragma unroll
    for (int k = 0; k < MAXKS; ++k) aM[k] = (live && 4 * k + lc < nd) ? P.mp[(size_t)row * P.ldn + 4 * k + lc] : 0.0;   // ldn may be < 64
    int ip[NBX], nbo[NBX], nbs[NBX];
    double mrem = 0.0, maq = 0.0;
#pragma unroll
    for (int b = 0; b < NBX; ++b) {
        ip[b] = (b < nb && live) ? P.b[b].inv_g[row] : -1;
        const bool memb = ip[b] >= 0 && !P.b[b].direct;
        nbo[b] = memb ? P.b[b].off_nb + ip[b] * LD + 2 * lc : P.off_zero + 2 * lc;
        nbs[b] = memb ? P.b[b].ncp * LD : 0;
        if (ip[b] >= 0 && b == P.rem) mrem = 1.0;
        if (ip[b] >= 0 && b == P.aq) maq = 1.0;
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            const double2 pv = ld2(&P.p[(size_t)row * ldb + col0 + n]);
            const double2 qv = ld2(&P.q[(size_t)row * ldb + col0 + n]);
            st2(&sm[xpc + own + 8 * j], pv.x, pv.y);
            st2(&sm[xq + own + 8 * j], qv.x, qv.y);
#pragma unroll
            for (int b = 0; b < NBX; ++b) {
                if (ip[b] < 0 || P.b[b].direct) continue;
                const FusedBath& B = P.b[b];
                st2(&sm[B.off_hist + ip[b] * LD + n], pv.x, pv.y);
                if (B.nonlocal) st2(&B.ring[((size_t)B.slot0 * B.ncp + ip[b]) * ldb + col0 + n], pv.x, pv.y);
            }
        }
    }
    // the direct (memory-less remainder) bath's noise row is read at the top of each round that needs it
    // (prefetched into L2 two steps ahead) instead of being held in registers across rounds
    const double* dnz = nullptr;
    size_t dnz_stride = 0;
    for (int b = 0; b < nb; ++b)
        if (P.b[b].direct && ip[b] >= 0) {
            dnz = P.b[b].noise + (size_t)ip[b] * ldb + col0 + 2 * lc;
            dnz_stride = (size_t)P.b[b].nc * ldb;
        }
    auto load_direct = [&](int trow, double (&z)[NT][2]) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            z[j][0] = z[j][1] = 0.0;
            if (dnz) {
                const double2 v = ld2(dnz + (size_t)trow * dnz_stride + 8 * j);
                z[j][0] = v.x; z[j][1] = v.y;
            }
        }
    };
    cta_barrier();
    // acc += (rows of this warp of a fragment-ordered shared-memory operator) . X
    auto smem_matvec2 = [&](const double* X, double (&a0)[NT][2], double (&a1)[NT][2]) {
        const double* Ad = sm + P.off_opD + (w * MAXKS) * 32 + lane;
        const double* Aq = sm + P.off_opQ + (w * MAXKS) * 32 + lane;
#pragma unroll
        for (int k = 0; k < MAXKS; ++k) {
            const double ad = Ad[k * 32];
            const double aq = HASQ ? Aq[k * 32] : 0.0;
            const double* xr = X + (4 * k + lc) * LD + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const double bv = xr[8 * j];
                dmma884(a0[j][0], a0[j][1], ad, bv);
                if (HASQ) dmma884(a1[j][0], a1[j][1], aq, bv);
            }
        }
    };

    double fpot[NT][2], uq[NT][2], fb[NT][2];
    zero_acc<NT>(fpot);
    zero_acc<NT>(uq);
    zero_acc<NT>(fb);
    for (int g0 = 0; g0 < P.nsteps; g0 += FT) {
        TICK2(7)
        if (P.mid) mid_phase<NT, RW2 * 32>(P, sm, P.off_xp1, min(FT, P.nsteps - g0), g0, col0);   // block starts find the head buffers in their initial roles
        TICK2(6)
        if (g0 == 0) {
            smem_matvec2(sm + xq, fpot, uq);
#pragma unroll
            for (int j = 0; j < NT; ++j) { fpot[j][0] = -fpot[j][0]; fpot[j][1] = -fpot[j][1]; }
            if (HASCON && P.carry_valid && live) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int n = col0 + 8 * j + 2 * lc;
                    if (P.samef[n]) fpot[j][0] = P.fpotn[(size_t)row * ldb + n];
                    if (P.samef[n + 1]) fpot[j][1] = P.fpotn[(size_t)row * ldb + n + 1];
                }
            }
            cta_barrier();
        }
        const int gend = min(P.nsteps, g0 + FT);
        for (int g = g0; g < gend; ++g) {
            const int s = g - g0;
            const int tn_cur = tn;
            const int tn1 = tn + 1 == P.nmd ? 0 : tn + 1;
            {
                // ============ round A: first force evaluation at time t, head p(t) ============
                double hA[NT][2], nzr[NT][2];
                load_direct(tn_cur, nzr);
                zero_acc<NT>(hA);
                matvec1<NT, true>(aM, MAXKS, sm + xpc, LD, hA, lr, lc);
                TICK2(0)
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int n = 8 * j + 2 * lc;
                    const double2 pv = live ? ld2(&sm[xpc + own + 8 * j]) : make_double2(0.0, 0.0);
                    const double2 qv = ld2(&sm[xq + own + 8 * j]);
                    double f0 = fpot[j][0] + uq[j][0] - hA[j][0];
                    double f1 = fpot[j][1] + uq[j][1] - hA[j][1];
                    double v[8];
                    double c3a = 0.0, c3b = 0.0;
                    v[0] = pv.x * pv.x;
                    v[1] = pv.y * pv.y;
#pragma unroll
                    for (int b = 0; b < NBX; ++b) {
                        double2 nv = ld2(&sm[nbo[b] + nbuf * nbs[b] + 8 * j]);
                        if (b < nb && P.b[b].direct && ip[b] >= 0) nv = make_double2(nzr[j][0], nzr[j][1]);
                        f0 += nv.x;
                        f1 += nv.y;
                        double c0 = nv.x, c1 = nv.y;
                        if (HASQ && b == P.aq) { c0 += maq * uq[j][0]; c1 += maq * uq[j][1]; }
                        if (b == P.rem) { c0 -= mrem * hA[j][0]; c1 -= mrem * hA[j][1]; }
                        if (b < 3) { v[2 + 2 * b] = pv.x * c0; v[3 + 2 * b] = pv.y * c1; }
                        else { c3a = pv.x * c0; c3b = pv.y * c1; }
                    }
                    if (live) {
                        st2(&sm[xpn + own + 8 * j], pv.x + f0 * dt / 2.0, pv.y + f1 * dt / 2.0);
                        st2(&sm[xq + own + 8 * j], qv.x + pv.x * dt + f0 * dt2 / 2.0, qv.y + pv.y * dt + f1 * dt2 / 2.0);
                        if (P.ps) st2(&P.ps[((size_t)tn_cur * nd + row) * ldb + col0 + n], pv.x, pv.y);
                        if (P.qs) st2(&P.qs[((size_t)tn_cur * nd + row) * ldb + col0 + n], qv.x, qv.y);
                    }
                    const double rsum = rs8(v, lr);
                    const int qn = lr >> 1;              // 0: kinetic energy, 1..3: bath qn-1
                    if (qn <= nb) sm[P.off_curp + (((qn == 0) ? nb : qn - 1) * TW2 + w) * N + n + (lr & 1)] = rsum;
                    if (nb > 3) {
                        c3a = colsum8(c3a); c3b = colsum8(c3b);
                        if (lr == 0) st2(&sm[P.off_curp + (3 * TW2 + w) * N + n], c3a, c3b);
                    }
                }
            }
            TICK2(1)
            cta_barrier();                                                      // ---- S2 (all warps)
            TICK2(5)
            nbuf ^= 1;
            {
                { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1/2)
                // ============ round B: second evaluation at time t+1, head p(t+1/2) ============
                {
                    double aDq[NT][2], hB[NT][2], nzr[NT][2];
                    load_direct(tn1, nzr);
                    zero_acc<NT>(aDq);
                    zero_acc<NT>(hB);
                    zero_acc<NT>(uq);
                    smem_matvec2(sm + xq, aDq, uq);
                    matvec1<NT, true>(aM, MAXKS, sm + xpc, LD, hB, lr, lc);
                    TICK2(2)
                    double p1v[NT][2];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        fpot[j][0] = -aDq[j][0];
                        fpot[j][1] = -aDq[j][1];
                        double f0 = fpot[j][0] + uq[j][0], f1 = fpot[j][1] + uq[j][1];
#pragma unroll
                        for (int b = 0; b < NBX; ++b) {
                            double2 nv = ld2(&sm[nbo[b] + nbuf * nbs[b] + 8 * j]);
                            if (b < nb && P.b[b].direct && ip[b] >= 0) nv = make_double2(nzr[j][0], nzr[j][1]);
                            f0 += nv.x;
                            f1 += nv.y;
                        }
                        fb[j][0] = f0;
                        fb[j][1] = f1;
                        const double2 ph = ld2(&sm[xpc + own + 8 * j]);
                        p1v[j][0] = ph.x + dt * (f0 - hB[j][0]) / 2.0;
                        p1v[j][1] = ph.y + dt * (f1 - hB[j][1]) / 2.0;
                    }
                    if (live) {
#pragma unroll
                        for (int j = 0; j < NT; ++j) st2(&sm[xpn + own + 8 * j], p1v[j][0], p1v[j][1]);
                    }
                }
                TICK2(3)
                row_barrier();                                                  // ---- S3 (row warps)
                TICK2(5)
                { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p1, xpn: p(t+1/2)
                // ============ round C: third evaluation, head p1 -> p(t+1), constraint ============
                {
                    double hC[NT][2];
                    zero_acc<NT>(hC);
                    matvec1<NT, true>(aM, MAXKS, sm + xpc, LD, hC, lr, lc);
                    TICK2(4)
                    double pn[NT][2], qn[NT][2];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double2 ph = ld2(&sm[xpn + own + 8 * j]);
                        pn[j][0] = live ? ph.x + dt * (fb[j][0] - hC[j][0]) / 2.0 : 0.0;
                        pn[j][1] = live ? ph.y + dt * (fb[j][1] - hC[j][1]) / 2.0 : 0.0;
                        if (HASCON) {
                            const double2 q0 = ld2(&sm[xq + own + 8 * j]);
                            qn[j][0] = q0.x; qn[j][1] = q0.y;
                        }
                    }
                    if (HASCON) {
                        const double cv = sm[P.off_conv + row];
                        const double dv = sm[P.off_conv + NDP + row];
                        const double av = HASQ ? sm[P.off_conv + 2 * NDP + row] : 0.0;
                        constexpr int NG = (4 * NT + 7) / 8;
                        double tot[NG], part[NG];
#pragma unroll
                        for (int gg = 0; gg < NG; ++gg) {
                            double v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int vi = 8 * gg + i, j = vi >> 2, kind = vi & 3;       // d0, d1, g0, g1 of n-tile j
                                v[i] = (kind < 2 ? pn[j][kind] : qn[j][kind - 2]) * cv;
                            }
                            part[gg] = rs8(v, lr);
                        }
#pragma unroll
                        for (int gg = 0; gg < NG; ++gg) sm[P.off_conp + ((8 * gg + lr) * RW2 + w) * 4 + lc] = part[gg];
                        TICK2(3)
                        row_barrier();                                          // ---- S4 (row warps)
                        TICK2(5)
#pragma unroll
                        for (int gg = 0; gg < NG; ++gg) {
                            double sum = 0.0;
#pragma unroll
                            for (int ww = 0; ww < RW2; ++ww) sum += sm[P.off_conp + ((8 * gg + lr) * RW2 + ww) * 4 + lc];
                            tot[gg] = sum;
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
                            if (!(fabs(d[2]) * P.cmax < 10e-10)) fpot[j][0] += d[2] * dv;
                            if (!(fabs(d[3]) * P.cmax < 10e-10)) fpot[j][1] += d[3] * dv;
                            if (HASQ) { uq[j][0] -= d[2] * av; uq[j][1] -= d[3] * av; }
                        }
                    }
                    if (live) {
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            st2(&sm[xpn + own + 8 * j], pn[j][0], pn[j][1]);
                            if (HASCON) st2(&sm[xq + own + 8 * j], qn[j][0], qn[j][1]);
                        }
#pragma unroll
                        for (int b = 0; b < NBX; ++b) {
                            if (ip[b] < 0 || P.b[b].direct) continue;
                            const FusedBath& B = P.b[b];
                            if (!B.nonlocal) {
#pragma unroll
                                for (int j = 0; j < NT; ++j) st2(&sm[B.off_hist + ip[b] * LD + 8 * j + 2 * lc], pn[j][0], pn[j][1]);
                            } else if (g + 1 < P.nsteps) {
                                const int hb = B.off_hist + (((s + 1) & (FT - 1)) * B.ncp + ip[b]) * LD + 2 * lc;
                                int slot = (B.slot0 + g + 1) % B.R;
                                double* rp = &B.ring[((size_t)slot * B.ncp + ip[b]) * ldb + col0 + 2 * lc];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    st2(&sm[hb + 8 * j], pn[j][0], pn[j][1]);
                                    st2(rp + 8 * j, pn[j][0], pn[j][1]);
                                }
                            }
                        }
                    }
                }
            }
            TICK2(3)
            cta_barrier();                                                      // ---- S1 (all warps)
            TICK2(5)
            { const int tmp = xpc; xpc = xpn; xpn = tmp; }                      // xpc: p(t+1)
            tn = tn1;
        }
    }
#ifdef PYMD_FUSED_TIMING
    if (P.timing && lane == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&P.timing[i]), (unsigned long long)t2acc[i]);
#endif
    // ------------------------------------------------------------------ epilogue
    if (live) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = 8 * j + 2 * lc;
            const double2 pv = ld2(&sm[xpc + own + 8 * j]);
            const double2 qv = ld2(&sm[xq + own + 8 * j]);
            st2(&P.p[(size_t)row * ldb + col0 + n], pv.x, pv.y);
            st2(&P.q[(size_t)row * ldb + col0 + n], qv.x, qv.y);
            st2(&P.fpotn[(size_t)row * ldb + col0 + n], fpot[j][0], fpot[j][1]);
        }
    }

    } else {
    // ================================================================== BATH WARPS
#if F2_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F2_BATH_REGS));
#endif
    // ---- bath-warp state: the pair (bath pb, m8 tile pmt) and its tap / head fragments
    const int pb = roww ? -1 : P.pair_b[w - RW2], pmt = roww ? 0 : P.pair_mt[w - RW2];
    double kf[FT][4], hf[4];
#pragma unroll
    for (int k = 0; k < FT; ++k)
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) kf[k][kq] = 0.0;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) hf[kq] = 0.0;
    if (pb >= 0) {
        const FusedBath& B = P.b[pb];
        const int MT = B.ncp / 8, KQ = B.ncp / 4;
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {
            if (kq < KQ) {
                if (pb != P.rem) hf[kq] = B.hdf[((size_t)pmt * KQ + kq) * 32 + lane];
                if (B.nonlocal) {
#pragma unroll
                    for (int k = 0; k < FT; ++k)
                        if (k < B.ntap) kf[k][kq] = B.kinf[(((size_t)k * MT + pmt) * KQ + kq) * 32 + lane];
                }
            }
        }
    }

    // L2 prefetch of the noise rows two steps ahead: the 128 bath threads cover the 128-byte lines of all rows
    cta_barrier();
    for (int g0 = 0; g0 < P.nsteps; g0 += FT) {
        TICK2(7)
        if (P.mid) mid_phase<NT, NTH2, false>(P, sm, P.off_xp1, min(FT, P.nsteps - g0), g0, col0);
        if (g0 == 0) cta_barrier();
        TICK2(6)
        const int gend = min(P.nsteps, g0 + FT);
        for (int g = g0; g < gend; ++g) {
            const int s = g - g0;
            const int tn_cur = tn;
            const int tn1 = tn + 1 == P.nmd ? 0 : tn + 1;
            {
                const int tn2 = tn1 + 1 == P.nmd ? 0 : tn1 + 1;
                int e = tid - RW2 * 32;
                for (int b = 0; b < nb; ++b) {
                    constexpr int LPR = (N * 8 + 127) / 128;       // 128-byte lines per row tile
                    const int cnt = P.b[b].nc * LPR;
                    for (; e < cnt; e += (TW2 - RW2) * 32)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.b[b].noise + ((size_t)tn2 * P.b[b].nc + e / LPR) * ldb + col0 + (e % LPR) * 16));
                    e -= cnt;
                }
            }
            if (pb >= 0) {
                // ============ bath warp: noise minus tail at time t+1 and head scalars of (pb, pmt) ============
                const FusedBath& B = P.b[pb];
                const int rr = pmt * 8 + lr;
                const bool valid = rr < B.nc;
                const int KQ = B.ncp / 4;
                double2 xi[NT], pt[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    xi[j] = make_double2(0.0, 0.0);
                    pt[j] = make_double2(0.0, 0.0);
                    if (valid) {
                        xi[j] = ld2(&B.noise[((size_t)tn1 * B.nc + rr) * ldb + col0 + 8 * j + 2 * lc]);
                        if (B.nonlocal) pt[j] = ld2(&B.P[((size_t)s * B.nc + rr) * ldb + col0 + 8 * j + 2 * lc]);
                    }
                }
                double acc[NT][2], acc2[NT][2];
                zero_acc<NT>(acc);
                zero_acc<NT>(acc2);
                if (B.nonlocal) {
                    const int ntap = min(s + 1, B.ntap);
                    const double* H0 = sm + B.off_hist + lc * LD + lr;
#pragma unroll
                    for (int k = 1; k <= FT; ++k) {
                        if (k <= ntap) {
                            const double* H = H0 + (s + 1 - k) * B.ncp * LD;        // p_c(t+1-k)
#pragma unroll
                            for (int kq = 0; kq < 4; ++kq) {
                                if (kq < KQ) {
#pragma unroll
                                    for (int j = 0; j < NT; ++j) {
                                        if (k & 1) dmma884(acc[j][0], acc[j][1], kf[k - 1][kq], H[(4 * kq) * LD + 8 * j]);
                                        else dmma884(acc2[j][0], acc2[j][1], kf[k - 1][kq], H[(4 * kq) * LD + 8 * j]);
                                    }
                                }
                            }
                        }
                    }
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double t0v = pt[j].x + (acc[j][0] + acc2[j][0]), t1v = pt[j].y + (acc[j][1] + acc2[j][1]);
                        st2(&sm[B.off_nb + ((nbuf ^ 1) * B.ncp + rr) * LD + 8 * j + 2 * lc], xi[j].x - t0v, xi[j].y - t1v);
                        if (B.nonlocal && g == P.nsteps - 1)
                            st2(&B.tail_cur[(size_t)rr * ldb + col0 + 8 * j + 2 * lc], t0v, t1v);
                    }
                }
                if (pb != P.rem) {
                    // head of a small bath: only the scalars p_c.g_b (currents) are needed
                    const int hs = B.nonlocal ? s : 0;
                    const int inrem = valid ? reinterpret_cast<const int*>(sm + B.off_inrem)[rr] : 0;
                    double gh[NT][2];
                    zero_acc<NT>(gh);
                    const double* H = sm + B.off_hist + (hs * B.ncp + lc) * LD + lr;
#pragma unroll
                    for (int kq = 0; kq < 4; ++kq) {
                        if (kq < KQ) {
#pragma unroll
                            for (int j = 0; j < NT; ++j) dmma884(gh[j][0], gh[j][1], hf[kq], H[(4 * kq) * LD + 8 * j]);
                        }
                    }
#pragma unroll
                    for (int j2 = 0; j2 < NT; j2 += 2) {
                        double v[8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int j = j2 + h;
                            double s0 = 0.0, s1 = 0.0;
                            if (valid) {
                                const double2 pv = ld2(&sm[B.off_hist + (hs * B.ncp + rr) * LD + 8 * j + 2 * lc]);
                                s0 = pv.x * gh[j][0];
                                s1 = pv.y * gh[j][1];
                            }
                            v[4 * h + 0] = s0; v[4 * h + 1] = s1;
                            v[4 * h + 2] = inrem ? s0 : 0.0; v[4 * h + 3] = inrem ? s1 : 0.0;
                        }
                        const double r = rs8(v, lr);     // lane lr: n-tile j2 + (lr>>2), kind lr&3 (s0, s1, r0, r1)
                        const int j = j2 + (lr >> 2), kind = lr & 3;
                        // this warp's own slots: the head enters the bath's current with a minus sign and, for
                        // dofs shared with the remainder bath, the remainder's with a plus sign
                        if (kind < 2) sm[P.off_curp + (pb * TW2 + w) * N + 8 * j + 2 * lc + kind] = -r;
                        else if (P.rem >= 0) sm[P.off_curp + (P.rem * TW2 + w) * N + 8 * j + 2 * lc + kind - 2] = r;
                    }
                }
            }
            TICK2(0)
            cta_barrier();                                                      // ---- S2 (all warps)
            TICK2(5)
            nbuf ^= 1;
            {
                // currents and kinetic energy of time t: fixed-order sum of the per-warp slots
                for (int e = tid - RW2 * 32; e < (nb + 1) * N; e += (TW2 - RW2) * 32) {
                    const int b = e / N, n = e % N;
                    double sum = 0.0;
#pragma unroll
                    for (int ww = 0; ww < TW2; ++ww) sum += sm[P.off_curp + (b * TW2 + ww) * N + n];
                    if (b < nb) P.b[b].cur[(size_t)tn_cur * ldb + col0 + n] = sum;
                    else P.etot[(size_t)tn_cur * ldb + col0 + n] = 0.5 * sum;
                }
            }
            TICK2(1)
            cta_barrier();                                                      // ---- S1 (all warps)
            TICK2(5)
            tn = tn1;
        }
    }
#ifdef PYMD_FUSED_TIMING
    if (P.timing && lane == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&P.timing[16 + i]), (unsigned long long)t2acc[i]);
#endif
    }
    for (int n = tid; n < N; n += NTH2) P.samef[col0 + n] = 1;
}

// shared-memory plan of fused2; returns bytes, or a huge value when the configuration does not qualify
size_t plan2(const pymd_ctx* c, FusedParams& fp, bool hasq, int NT) {
    const int N = 8 * NT, LD = N + 4;
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 1) / 2 * 2; return o; };
    fp.off_xp0 = take(NDP * LD);
    fp.off_xp1 = take(NDP * LD);
    fp.off_xq = take(NDP * LD);
    fp.off_opD = take(8 * MAXKS * 32);
    fp.off_opQ = hasq ? take(8 * MAXKS * 32) : fp.off_opD;
    fp.off_zero = take(LD);
    fp.off_curp = take((c->nbath + 1) * TW2 * N);
    fp.off_conp = take(((4 * NT + 7) / 8) * 8 * RW2 * 4);
    fp.off_conv = take(3 * NDP);
    int npair = 0;
    for (int i = 0; i < 4; ++i) fp.pair_b[i] = fp.pair_mt[i] = -1;
    for (int b = 0; b < c->nbath; ++b) {
        const BathHost& H = c->bath[b];
        FusedBath& B = fp.b[b];
        B.nc = H.nc; B.ncp = H.ncp; B.ml = H.ml; B.nonlocal = H.ml > 1;
        B.direct = (b == fp.rem && H.ml == 1) ? 1 : 0;
        B.ldk = H.ncp + 4;
        B.ntap = H.ml > 1 ? std::min(FT, H.ml - 1) : 0;
        B.off_cids = B.off_inrem = B.off_hd = B.off_nb = B.off_kin = B.off_hist = 0;
        if (B.direct) continue;
        if (H.ncp > 16) return (size_t)1 << 40;
        for (int mt = 0; mt < H.ncp / 8; ++mt) {
            if (npair >= 4) return (size_t)1 << 40;
            fp.pair_b[npair] = b; fp.pair_mt[npair] = mt; ++npair;
        }
        B.off_inrem = take(H.ncp / 2 + 1);
        B.off_nb = take(2 * H.ncp * LD);
        B.off_hist = take((H.ml > 1 ? FT : 1) * H.ncp * LD);
    }
    fp.smem_doubles = off;
    return (size_t)off * sizeof(double);
}

template <int NT, bool Q, bool C>
int launch_f2(const FusedParams& fp, size_t smem, int grid, cudaStream_t st) {
    static bool cfg = false;
    if (!cfg) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(fused2_kernel<NT, Q, C, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(fused2_kernel<NT, Q, C, FB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cfg = true;
    }
    if (fp.nbath <= 3) fused2_kernel<NT, Q, C, 3><<<grid, NTH2, smem, st>>>(fp);
    else fused2_kernel<NT, Q, C, FB_MAX><<<grid, NTH2, smem, st>>>(fp);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

struct Layout {
    FusedParams fp;
    size_t smem_bytes;
};

// shared-memory plan for NT n-tiles; returns bytes (0 if the work-item lists would overflow)
size_t plan(const pymd_ctx* c, int NT, FusedParams& fp) {
    const int N = 8 * NT, LD = N + 4;
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 1) / 2 * 2; return o; };
    fp.off_xp0 = take(NDP * LD);
    fp.off_xp1 = take(NDP * LD);
    fp.off_xq = take(NDP * LD);
    fp.off_zero = take(LD);
    fp.off_curp = take((c->nbath + 1) * FW * N);
    fp.off_conp = take(((4 * NT + 7) / 8) * 8 * FW * 4);
    fp.off_conv = take(3 * NDP);
    int nhead = 0, ntap = 0;
    for (int b = 0; b < c->nbath; ++b) {
        const BathHost& H = c->bath[b];
        FusedBath& B = fp.b[b];
        B.nc = H.nc; B.ncp = H.ncp; B.ml = H.ml; B.nonlocal = H.ml > 1;
        B.direct = (b == fp.rem && H.ml == 1) ? 1 : 0;
        B.ldk = H.ncp + 4;
        B.ntap = H.ml > 1 ? std::min(FT, H.ml - 1) : 0;
        B.off_cids = B.off_inrem = B.off_hd = B.off_nb = B.off_kin = B.off_hist = 0;
        if (B.direct) continue;
        ntap += H.ncp / 8;
        B.off_cids = take(H.ncp / 2 + 1);
        B.off_inrem = take(H.ncp / 2 + 1);
        if (b != fp.rem) { B.off_hd = take(H.ncp * B.ldk); nhead += H.ncp / 8; }
        B.off_nb = take(2 * H.ncp * LD);
        B.off_hist = take((H.ml > 1 ? FT : 1) * H.ncp * LD);
        if (H.ml > 1) B.off_kin = take(B.ntap * H.ncp * B.ldk);
    }
    fp.smem_doubles = off;
    if (nhead + ntap > FW * MAXIT) return (size_t)1 << 40;     // does not fit the item lists
    return (size_t)off * sizeof(double);
}

int pick_nt(const pymd_ctx* c, FusedParams& fp) {
    // trajectory tile 8*NT per CTA: trade SM fill (one CTA per SM) against per-CTA efficiency
    // (larger tiles amortise barriers and give the DMMA pipe more independent chains)
    int best = 0;
    double best_score = -1.0;
    const char* force = getenv("PYMD_FUSED_NT");      // testing knob: force the trajectory tile
    for (int NT : {4, 2, 1}) {
        if (force && atoi(force) != NT) continue;
        if (c->ldb % (8 * NT)) continue;
        if (plan(c, NT, fp) > 227 * 1024) continue;
        const long long ctas = c->ldb / (8 * NT);
        const long long waves = (ctas + 147) / 148;
        const double fill = (double)ctas / (double)(waves * 148);
        const double eff = NT == 4 ? 1.0 : (NT == 2 ? 0.9 : 0.5);      // measured: 16-wide tiles run the 12-warp kernel at 0.95 of the 32-wide rate
        if (fill * eff > best_score) { best_score = fill * eff; best = NT; }
    }
    if (best) plan(c, best, fp);
    return best;
}

void choose_roles(const pymd_ctx* c, FusedParams& fp) {
    fp.rem = -1;
    fp.aq = -1;
    int best = -1;
    for (int b = 0; b < c->nbath; ++b) {
        if (c->bath[b].nc > best) { best = c->bath[b].nc; fp.rem = b; }
        if (c->bath[b].has_aq) fp.aq = b;
    }
}

template <int NT>
int launch_nt(const FusedParams& fp, size_t smem, int grid, bool hasq, bool hascon, cudaStream_t st) {
#define PYMD_FUSED_CASE(Q, C, F)                                                                             \
    {                                                                                                      \
        static bool cfg = false;                                                                           \
        if (!cfg) {                                                                                        \
            PYMD_CUDA_CHECK(cudaFuncSetAttribute(fused_kernel<NT, Q, C, F>,                                \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            cfg = true;                                                                                    \
        }                                                                                                  \
        fused_kernel<NT, Q, C, F><<<grid, FW * 32, smem, st>>>(fp);                                        \
    }
    // operators with more than 12 of the 16 k-steps run all 16 branch-free
    const bool fullk = fp.ks > 12;
    if (fullk) {
        if (hasq && hascon) PYMD_FUSED_CASE(true, true, true)
        else if (hasq) PYMD_FUSED_CASE(true, false, true)
        else if (hascon) PYMD_FUSED_CASE(false, true, true)
        else PYMD_FUSED_CASE(false, false, true)
    } else {
        if (hasq && hascon) PYMD_FUSED_CASE(true, true, false)
        else if (hasq) PYMD_FUSED_CASE(true, false, false)
        else if (hascon) PYMD_FUSED_CASE(false, true, false)
        else PYMD_FUSED_CASE(false, false, false)
    }
#undef PYMD_FUSED_CASE
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace

// Operands shared by the on-chip step kernels (fused.cu, local.cu), built once per finalize():
// the q-coupled electron operator scattered to nd x nd, and c, D c, AqT c for every constraint vector
// (-D q(t+1) and AqT q(t+1) follow from the projection coefficients by linearity, without a mat-vec).
int prepare_scattered_operands(pymd_ctx* c, int aq_bath) {
    const bool hasq = aq_bath >= 0;
    if (hasq && !c->d_aqT) {
        const BathHost& H = c->bath[aq_bath];
        std::vector<double> aqT((size_t)c->nd * c->ldn, 0.0);
        for (int i = 0; i < H.nc; ++i)
            for (int j = 0; j < H.nc; ++j) aqT[(size_t)H.cids[i] * c->ldn + H.cids[j]] = H.aq[(size_t)i * H.nc + j];
        PYMD_CUDA_CHECK(cudaMalloc(&c->d_aqT, aqT.size() * sizeof(double)));
        PYMD_CUDA_CHECK(cudaMemcpy(c->d_aqT, aqT.data(), aqT.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (c->ncon > 0 && !c->d_con3) {
        const int nd = c->nd, nc3 = c->ncon;
        std::vector<double> c3((size_t)4 * nc3 * nd, 0.0);      // c, D c, AqT c, Mp^T c
        for (int k = 0; k < nc3; ++k)
            for (int i = 0; i < nd; ++i) {
                c3[(size_t)k * nd + i] = c->con[(size_t)k * nd + i];
                double sd = 0.0;
                for (int j = 0; j < nd; ++j) sd += c->dyn[(size_t)i * nd + j] * c->con[(size_t)k * nd + j];
                c3[(size_t)(nc3 + k) * nd + i] = sd;
            }
        // Mp^T c (fused3: c.Mp x = (Mp^T c).x lets the constraint dots be taken one round early)
        for (int k = 0; k < nc3; ++k)
            for (int j = 0; j < nd; ++j) {
                double sw = 0.0;
                for (int i = 0; i < nd; ++i) sw += c->mp_host[(size_t)i * c->ldn + j] * c->con[(size_t)k * nd + i];
                c3[(size_t)(3 * nc3 + k) * nd + j] = sw;
            }
        if (hasq) {
            const BathHost& H = c->bath[aq_bath];
            for (int k = 0; k < nc3; ++k)
                for (int i = 0; i < H.nc; ++i) {
                    double sa = 0.0;
                    for (int j = 0; j < H.nc; ++j) sa += H.aq[(size_t)i * H.nc + j] * c->con[(size_t)k * nd + H.cids[j]];
                    c3[(size_t)(2 * nc3 + k) * nd + H.cids[i]] = sa;
                }
        }
        PYMD_CUDA_CHECK(cudaMalloc(&c->d_con3, c3.size() * sizeof(double)));
        PYMD_CUDA_CHECK(cudaMemcpy(c->d_con3, c3.data(), c3.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return PYMD_OK;
}

int fused_supported(const pymd_ctx* c) {
    if (c->nd > NDP || c->ncon > 1 || c->nbath > FB_MAX) return 0;
    int naq = 0;
    for (int b = 0; b < c->nbath; ++b) {
        naq += c->bath[b].has_aq;
        if (c->sd.b[b].fhis) return 0;
        if (c->bath[b].ncp > NDP) return 0;
    }
    if (naq > 1) return 0;
    FusedParams fp{};
    choose_roles(c, fp);
    return pick_nt(c, fp) > 0;
}

int fused_step(pymd_ctx* c, long long t0, int nsteps, cudaStream_t st) {
    StepDev& d = c->sd;
    FusedParams fp{};
    choose_roles(c, fp);
    const int NT = pick_nt(c, fp);
    PYMD_REQUIRE(NT > 0, "fused: configuration does not fit in shared memory");
    const size_t smem = plan(c, NT, fp);
    fp.nd = c->nd; fp.ks = (c->nd + 3) / 4; fp.ldb = c->ldb; fp.ldn = c->ldn; fp.nbath = c->nbath;
    fp.nmd = c->nmd; fp.ncon = c->ncon; fp.dt = c->dt;
    fp.p = d.p; fp.q = d.q; fp.ps = d.ps; fp.qs = d.qs; fp.etot = d.etot; fp.fpotn = d.fpotn; fp.samef = d.samef;
    fp.dyn = c->d_dyn; fp.mp = c->d_mp;
    bool hasq = fp.aq >= 0;
    {
        int rc0 = prepare_scattered_operands(c, fp.aq);
        if (rc0) return rc0;
    }
    fp.aqT = c->d_aqT;
    fp.con3 = c->d_con3;
    fp.cmax = 0.0;
    for (int i = 0; i < (c->ncon > 0 ? c->nd : 0); ++i) fp.cmax = std::max(fp.cmax, std::fabs(c->con[i]));
    for (int b = 0; b < c->nbath; ++b) {
        BathHost& H = c->bath[b];
        FusedBath& B = fp.b[b];
        B.noise = d.b[b].noise; B.cur = d.b[b].cur; B.ring = d.b[b].ring; B.R = H.R > 0 ? H.R : 1;
        B.tail_cur = d.b[b].tail_cur; B.P = d.b[b].tail_next;
        B.head_g = H.d_head; B.lda = H.lda; B.kin_g = H.d_kintra; B.cids_g = H.d_cids; B.inv_g = H.d_inv;
        B.kinf = H.d_kinf; B.hdf = H.d_hdf;
        B.nld = d.b[b].noise_ld ? d.b[b].noise_ld : (long long)H.nc * c->ldb;
    }
    fp.rec_ld = d.rec_ld ? d.rec_ld : (long long)c->nd * c->ldb;
    fp.wrap_row = d.noise_wrap;
    // warp-specialised variant (12 warps) when the trajectory tile is 32 and the baths qualify
    bool use2 = NT >= 2 && !(getenv("PYMD_FUSED2") && atoi(getenv("PYMD_FUSED2")) == 0);
    size_t smem2 = 0;
    if (use2) {
        FusedParams t2 = fp;
        smem2 = plan2(c, t2, hasq, NT);
        if (smem2 > 227 * 1024) use2 = false;
        else fp = t2;
    }
    // two-unit variant (fused3.cu) where it applies: the default for 16- and 32-trajectory tiles
    // (PYMD_FUSED3 = 0 / 1 / unset: never / wherever it applies / per-tile default)
    const char* e3 = getenv("PYMD_FUSED3");
    bool use3 = use2 && (e3 ? atoi(e3) != 0 : (NT == 4 || d.rec_ld != 0 || d.noise_wrap != 0));
    size_t smem3 = 0;
    if (use3) {
        FusedParams t3 = fp;
        smem3 = plan3(c, t3, hasq, NT / 2);
        if (smem3 > 227 * 1024) use3 = false;
        else fp = t3;
    }
    {
        bool strided = d.rec_ld != 0 || d.noise_wrap != 0;
        for (int b = 0; b < c->nbath; ++b) strided = strided || d.b[b].noise_ld != 0;
        PYMD_REQUIRE(!strided || use3, "fused: strided record / noise layouts need the two-unit kernel (16- or 32-trajectory tiles, "
                                       "small baths of at most 16 dofs)");
    }
    int rc;
    if (!c->tail_valid) {
        for (int b = 0; b < c->nbath; ++b)
            if (c->bath[b].ml > 1 && (rc = launch_tail(c, b, 1, 1, d.b[b].tail_cur, st))) return rc;
        c->tail_valid = true;
    }
    const int grid = c->ldb / (8 * NT);
#ifdef PYMD_FUSED_TIMING
    static long long* d_timing = nullptr;
    if (!d_timing) { cudaMalloc(&d_timing, 32 * sizeof(long long)); }
    cudaMemsetAsync(d_timing, 0, 32 * sizeof(long long), st);
    fp.timing = d_timing;
#endif
    fp.nosplit = (getenv("PYMD_FUSED_SPLIT") && atoi(getenv("PYMD_FUSED_SPLIT")) == 0) ? 1 : 0;
    const bool far_on = !(getenv("PYMD_FAR") && atoi(getenv("PYMD_FAR")) == 0);          // testing knob
    // mid mode: every non-local bath has an FFT far field (far.cu); one launch then advances up to the
    // end of the current super-block of kFarS steps, computing the pre-block tails in the kernel
    bool mid = far_on;
    int nonlocal = 0;
    for (int b = 0; b < c->nbath; ++b)
        if (c->bath[b].ml > 1) { ++nonlocal; mid = mid && c->bath[b].d_ghat != nullptr; }
    mid = mid && nonlocal > 0;
    for (int done = 0; done < nsteps;) {
        int ns = std::min(FT, nsteps - done);
        if (mid) {
            unsigned need = 0;
            long long r0 = 0;
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (H.ml <= 1) continue;
                r0 = H.hcount - H.far_base;
                if (!H.far_valid || r0 < 0 || r0 >= kFarS) need |= 1u << b;
            }
            if (need) {
                // a new super-block starts here: far field of everything older than p(t-1)
                for (int b = 0; b < c->nbath; ++b)
                    if (c->bath[b].ml > 1) need |= 1u << b;
                if ((rc = far_launch(c, need, st))) return rc;
                r0 = 0;
            }
            ns = std::min(nsteps - done, kFarS - (int)r0);
            fp.mid = 1;
            fp.r0 = (int)r0;
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (H.ml <= 1) continue;
                fp.hc0 = H.hcount;
                fp.b[b].slot0 = (int)(H.hcount % H.R);
                fp.b[b].kmid = H.d_kmid;
                fp.b[b].F = H.d_far;
            }
        } else {
            fp.mid = 0;
            for (int b = 0; b < c->nbath; ++b) {
                BathHost& H = c->bath[b];
                if (H.ml > 1) {
                    // pre-block part: history older than this block, read once for ns steps
                    if ((rc = launch_tail(c, b, ns, 2, d.b[b].tail_next, st))) return rc;
                    fp.b[b].slot0 = (int)(H.hcount % H.R);
                }
            }
        }
        fp.t0 = t0 + done;
        fp.nsteps = ns;
        fp.carry_valid = c->fpot_valid ? 1 : 0;
        prof_begin(c, 1, st);
        if (use3) rc = launch_f3(fp, smem3, NT / 2, hasq, c->ncon > 0, st);
        else if (use2) {
            const bool con = c->ncon > 0;
            if (NT == 4)
                rc = hasq ? (con ? launch_f2<4, true, true>(fp, smem2, grid, st) : launch_f2<4, true, false>(fp, smem2, grid, st))
                          : (con ? launch_f2<4, false, true>(fp, smem2, grid, st) : launch_f2<4, false, false>(fp, smem2, grid, st));
            else
                rc = hasq ? (con ? launch_f2<2, true, true>(fp, smem2, grid, st) : launch_f2<2, true, false>(fp, smem2, grid, st))
                          : (con ? launch_f2<2, false, true>(fp, smem2, grid, st) : launch_f2<2, false, false>(fp, smem2, grid, st));
        } else
        switch (NT) {
            case 4: rc = launch_nt<4>(fp, smem, grid, hasq, c->ncon > 0, st); break;
            case 2: rc = launch_nt<2>(fp, smem, grid, hasq, c->ncon > 0, st); break;
            default: rc = launch_nt<1>(fp, smem, grid, hasq, c->ncon > 0, st);
        }
        prof_end(c, st);
        if (rc) return rc;
        c->launches++;
        for (int b = 0; b < c->nbath; ++b)
            if (c->bath[b].ml > 1) c->bath[b].hcount += ns;
        c->fpot_valid = true;
        c->fpotc_valid = false;
        done += ns;
    }
#ifdef PYMD_FUSED_TIMING
    {
        long long h[32];
        cudaMemcpy(h, fp.timing, sizeof(h), cudaMemcpyDeviceToHost);
        if (use3) {
            const double nr = (double)grid * 8 * nsteps, nbw = (double)grid * 4 * nsteps;
            const char* rn[8] = {"A.matvec", "A.elem", "reduce+B.matvec", "B.elem/C.elem", "C.matvec", "barrier-wait", "mid", "pre-mid"};
            const char* bn[8] = {"job", "prefetch", "-", "-", "-", "H-wait", "-", "-"};
            double tr = 0, tb = 0;
            fprintf(stderr, "[fused3 timing] row warps, cycles per warp-step:");
            for (int i = 0; i < 8; ++i) { fprintf(stderr, " %s=%.0f", rn[i], h[i] / nr); tr += h[i] / nr; }
            fprintf(stderr, " total=%.0f\n[fused3 timing] bath warps:", tr);
            for (int i = 0; i < 8; ++i) if (bn[i][0] != '-') { fprintf(stderr, " %s=%.0f", bn[i], h[16 + i] / nbw); tb += h[16 + i] / nbw; }
            fprintf(stderr, " total=%.0f\n", tb);
            return PYMD_OK;
        }
        if (use2) {
            // only the LAST launch of this call (the counters are cleared per call, accumulated over its launches)
            const double nr = (double)grid * RW2 * nsteps, nbw = (double)grid * (TW2 - RW2) * nsteps;
            const char* rn[8] = {"A.matvec", "A.elem", "B.matvec", "B.elem/C.elem", "C.matvec", "barrier-wait", "mid", "pre-mid"};
            const char* bn[8] = {"bath.work", "bath.reduce", "-", "-", "-", "barrier-wait", "mid(parked)", "pre-mid"};
            double tr = 0, tb = 0;
            fprintf(stderr, "[fused2 timing] row warps, cycles per warp-step:");
            for (int i = 0; i < 8; ++i) { fprintf(stderr, " %s=%.0f", rn[i], h[i] / nr); tr += h[i] / nr; }
            fprintf(stderr, " total=%.0f\n[fused2 timing] bath warps:", tr);
            for (int i = 0; i < 8; ++i) if (bn[i][0] != '-') { fprintf(stderr, " %s=%.0f", bn[i], h[16 + i] / nbw); tb += h[16 + i] / nbw; }
            fprintf(stderr, " total=%.0f\n", tb);
            return PYMD_OK;
        }
        const double nw = (double)grid * FW * nsteps;
        const char* nm[10] = {"A.matvec", "A.elementwise", "A.items", "barrier-wait", "B.matvec", "B.elementwise", "C.matvec", "C.elementwise", "mid", "pre-mid"};
        double tot = 0; for (int i = 0; i < 10; ++i) tot += h[i] / nw;
        fprintf(stderr, "[fused timing] cycles per warp-step (NT=%d):", NT);
        for (int i = 0; i < 10; ++i) fprintf(stderr, " %s=%.0f", nm[i], h[i] / nw);
        fprintf(stderr, " total=%.0f\n", tot);
    }
#endif
    return PYMD_OK;
}

}  // namespace pymd

// fused3: two-unit variant of the warp-specialised on-chip step kernel (mode 2, nd <= 64).
//
// Measured on fused2 (profiles/r02_fused2_phase_timing.txt): the operator products run at the rate of the
// DMMA pipe (12.6 k of 31 k cycles per step), but between them every warp of the CTA is in the element-wise
// part of a round (11 k cycles) or in the mid phase (5 k) at the same time, and the pipe idles.  Here the
// trajectory tile of a CTA is split into TWO UNITS that advance independently:
//
//   * unit u = 8 NJ trajectories (NJ = 2: 32 per CTA, NJ = 1: 16 per CTA), stepped by FOUR row warps that own
//     16 operator rows each (two m8 tiles): Mp = sum_b scatter(alpha_b K_b[0]) as register A fragments
//     (2 x KS), dyn and AqT as fragment-ordered shared-memory operands.  Warp w of a unit sits on scheduler
//     w % 4, so every scheduler hosts one row warp of each unit: while one unit is in the element-wise part of
//     a round, behind a barrier or staging ring rows, the other one keeps the DMMA pipe of that scheduler busy.
//   * the three rounds of a step (md.py:315-358) are joined by barriers of the unit's 128 row threads only;
//   * FOUR bath warps serve both units: each owns one (bath, m8 tile) pair with the in-block taps dt K[1..8] and
//     the head alpha K[0] as register fragments; per (unit, step) it produces noise-minus-tail of the next time
//     and the complete heat current of its bath rows (p_c . (xi - tail - K[0] p_c), md.py:342), so the row warps
//     only reduce the kinetic energy and the current of the remainder bath (ONE reduce-scatter per step);
//   * producer/consumer named barriers instead of CTA-wide ones: H(u) "history row and pre-block tails of the
//     next step are in place" (row warps arrive, bath warps wait), NBR(u) "noise-minus-tail and head scalars are
//     in place" (bath warps arrive, row warps wait -- it doubles as the round-A barrier of the unit);
//   * the mid field (taps that meet earlier blocks of the 32-step super-block) is computed
//     per unit by its own row warps (warp wr owns the steps 2 wr and 2 wr + 1 of the block), staged through the
//     unit's columns of the free head buffer, while the other unit goes on stepping.
//
// Same operands, carried state and launch sequence as fused2 (fused.cu): fused_step() picks this kernel when
// the baths qualify (every small bath ncp <= 16, at most four pairs, the q-coupled bath is the remainder bath).
#include <algorithm>
#include <cstdlib>

#include "fused.cuh"

namespace pymd {

namespace {

// Register re-balancing between the warp groups (setmaxnreg works on groups of four warps: the two row-warp
// groups and the bath group): 2 x Z_ROW_REGS + Z_BATH_REGS <= 3 x 168.
#ifndef Z_SETMAXNREG
#define Z_SETMAXNREG 1
#endif
// Measured (us per 32-step launch, 32- / 16-trajectory tiles): 200/104: 428 / 253, 208/88: 423 / 258, 216/72: 420 / 263,
// 224/56: 429 / 273, 232/40: 476 / 287 -- 216/72 for the 32-trajectory tiles, 200/104 for the 16-trajectory ones.
#ifndef Z_ROW_REGS
#define Z_ROW_REGS (NJ == 2 ? 216 : 200)
#endif
#ifndef Z_BATH_REGS
#define Z_BATH_REGS (NJ == 2 ? 72 : 104)
#endif
// Unit 1 starts half a step behind unit 0, so that the element-wise part of one unit's round meets the operator
// products of the other (units that start together stay in lockstep and leave the DMMA pipe idle together).
#ifndef Z_OFFSET
#define Z_OFFSET 1
#endif

constexpr int ZU = 2;            // units per CTA
constexpr int ZRW = 4;           // row warps per unit
constexpr int ZBW = 4;           // bath warps (shared by the units)
constexpr int ZTH = (ZU * ZRW + ZBW) * 32;
constexpr int ZSL = ZRW + ZBW;   // per-column slots of the current / energy sums

__device__ __forceinline__ void zbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void zbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// barrier ids: 0 = whole CTA (prologue), 1 + u = R(u) the unit's row warps, 3 + u = NBR(u), 5 + u = H(u)
__device__ __forceinline__ int bar_R(int u) { return 1 + u; }
__device__ __forceinline__ int bar_NBR(int u) { return 3 + u; }
__device__ __forceinline__ int bar_H(int u) { return 5 + u; }

// Pre-block tails of one block for ONE unit and one bath (see mid_bath in fused.cu):
//   P[s] = F[r0b+s] + dt sum_{a=0}^{W-1} K[s+2+a] p_c(t0-1-a)
// by the unit's four row warps; warp wr owns the steps s = 2 wr and 2 wr + 1.
template <int NJ, int MT, int KQ>
__device__ __forceinline__ void zmid_bath(const FusedParams& P, const FusedBath& B, double* sm, int stage_off, int nsblk, int g,
                                          int col0, int u, int wr, int lane) {
    constexpr int N = 16 * NJ, LD = N + 4, ncp = 8 * MT, nA = MT * KQ, NU = 8 * NJ;
    constexpr int spst = NDP / ncp;                     // ring rows per stage
    constexpr int per_slot = ncp * (NU / 2);            // double2 per ring-row tile of the unit
    constexpr int NPF = (spst * per_slot + ZRW * 32 - 1) / (ZRW * 32);
    const int ut = wr * 32 + lane, lr = lane >> 2, lc = lane & 3, cu = NU * u;
    const size_t ldb = P.ldb;
    const int r0b = P.r0 + g, slotb = (B.slot0 + g) % B.R;
    long long Wl = r0b + 1;
    if (Wl > P.hc0 + g) Wl = P.hc0 + g;
    if (Wl > B.ml - 2) Wl = B.ml - 2;
    const int W = (int)Wl;
    // warp wr owns the ADJACENT steps 2 wr and 2 wr + 1 of the block: item (ring row a, step s) uses tap set s + a
    // (dt K[set + 2], host-swizzled A fragments), so the two sets of row a are sets t = 2 wr + a and t + 1, and row
    // a + 1 re-uses set t + 1: ONE new set per ring row, fetched from L2 two rows (~1100 DMMA cycles) ahead.
    const bool mine0 = 2 * wr < nsblk, mine1 = 2 * wr + 1 < nsblk;
    double acc[2][MT][NJ][2];
#pragma unroll
    for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[st][mt][j][0] = acc[st][mt][j][1] = 0.0;
    double2 pf[NPF];
    auto fetch = [&](int a0) {
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
            const int e = ut + i * ZRW * 32;
            const int sl = e / per_slot, j = (e % per_slot) / (NU / 2), c2 = e % (NU / 2);
            pf[i] = make_double2(0.0, 0.0);
            if (e < spst * per_slot && a0 + sl < W) {
                int slot = slotb - 1 - a0 - sl;
                if (slot < 0) slot += B.R;
                pf[i] = __ldcg(reinterpret_cast<const double2*>(&B.ring[((size_t)slot * ncp + j) * ldb + col0 + cu + 2 * c2]));
            }
        }
    };
    const double* Abase = B.kmid + lane;
    auto loadA = [&](double (&dst)[nA], int set) {
        set = min(set, kMidK - 1);                      // sets past the table are never multiplied
#pragma unroll
        for (int i = 0; i < nA; ++i) dst[i] = __ldg(Abase + ((size_t)set * nA + i) * 32);
    };
    // sets t .. t + 3, t = 2 wr + (ring rows multiplied so far); with 32-trajectory tiles (NJ = 2) the registers only
    // allow three (one row ahead: the row itself takes twice as many DMMA cycles there)
    constexpr bool DEEP = NJ == 1;
    double s0[nA], s1[nA], s2[nA], s3[nA];
    int tset = 2 * wr;
    loadA(s0, tset); loadA(s1, tset + 1); loadA(s2, tset + 2);
    if (DEEP) loadA(s3, tset + 3);
    const int nh = g > 0 ? min(W, FT - 1) : 0;          // rows still in the shared-memory block history
    if (W > nh) fetch(nh);
    auto row_products = [&](const double* H) {
#pragma unroll
        for (int kq = 0; kq < KQ; ++kq) {
            double bf[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = H[(4 * kq) * LD + 8 * j];
            if (mine0) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[0][mt][j][0], acc[0][mt][j][1], s0[mt * KQ + kq], bf[j]);
            }
            if (mine1) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[1][mt][j][0], acc[1][mt][j][1], s1[mt * KQ + kq], bf[j]);
            }
        }
        ++tset;
#pragma unroll
        for (int i = 0; i < nA; ++i) { s0[i] = s1[i]; s1[i] = s2[i]; if (DEEP) s2[i] = s3[i]; }
        if (DEEP) loadA(s3, tset + 3);
        else loadA(s2, tset + 2);
    };
    const double* hist = sm + B.off_hist;
    double* stage = sm + stage_off;
    for (int a = 0; a < nh; ++a) row_products(hist + ((FT - 1 - a) * ncp + lc) * LD + lr + cu);
    for (int a0 = nh; a0 < W; a0 += spst) {
        zbar_sync(bar_R(u), ZRW * 32);                  // previous stage fully consumed
#pragma unroll
        for (int i = 0; i < NPF; ++i) {
            const int e = ut + i * ZRW * 32;
            const int sl = e / per_slot, j = (e % per_slot) / (NU / 2), c2 = e % (NU / 2);
            if (e < spst * per_slot) *reinterpret_cast<double2*>(&stage[(sl * ncp + j) * LD + cu + 2 * c2]) = pf[i];
        }
        zbar_sync(bar_R(u), ZRW * 32);
        if (a0 + spst < W) fetch(a0 + spst);
        const int na = min(spst, W - a0);
        for (int sl = 0; sl < na; ++sl) row_products(stage + (sl * ncp + lc) * LD + lr + cu);
    }
#pragma unroll
    for (int st = 0; st < 2; ++st) {
        const int s = 2 * wr + st;
        if (s < nsblk) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int rr = mt * 8 + lr;
                if (rr < B.nc) {
                    const double* Fr = B.F + ((size_t)(r0b + s) * B.nc + rr) * ldb + col0 + cu + 2 * lc;
                    double* Pr = const_cast<double*>(B.P) + ((size_t)s * B.nc + rr) * ldb + col0 + cu + 2 * lc;
                    double2 f[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) f[j] = __ldcg(reinterpret_cast<const double2*>(Fr + 8 * j));
#pragma unroll
                    for (int j = 0; j < NJ; ++j) st2(Pr + 8 * j, f[j].x + acc[st][mt][j][0], f[j].y + acc[st][mt][j][1]);
                }
            }
        }
    }
}

template <int NJ>
__device__ __forceinline__ void zmid_phase(const FusedParams& P, double* sm, int stage_off, int nsblk, int g, int col0, int u, int wr,
                                           int lane) {
    for (int b = 0; b < P.nbath; ++b) {
        const FusedBath& B = P.b[b];
        if (!B.nonlocal) continue;
        if (B.ncp == 16) zmid_bath<NJ, 2, 4>(P, B, sm, stage_off, nsblk, g, col0, u, wr, lane);
        else zmid_bath<NJ, 1, 2>(P, B, sm, stage_off, nsblk, g, col0, u, wr, lane);
    }
    zbar_sync(bar_R(u), ZRW * 32);                      // staging buffer (a head buffer) free again
}

// NJ: n8 tiles per unit; KS: k-steps of the nd x nd operators (15 when nd <= 60, else 16); SPEC: the per-row bath loop
// unrolled without conditions for the two usual bath layouts -- 1: a direct remainder bath first, then exactly two
// shared-memory baths (electron bath + two leads: AuBZ, Au6_3x3); 2: exactly two shared-memory baths and nothing else,
// the first one being the remainder bath (two leads: tests/test.py); 0: anything else (run-time loop)
template <int NJ, bool HASQ, bool HASCON, int KS, int SPEC>
__global__ void __launch_bounds__(ZTH, 1) fused3_kernel(const FusedParams P) {
    constexpr int N = ZU * 8 * NJ, LD = N + 4, NU = 8 * NJ;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
    const bool roww = w < ZU * ZRW;
    const int col0 = blockIdx.x * N;
    const int nd = P.nd, nb = P.nbath;
    const size_t ldb = P.ldb;
    const double dt = P.dt, dt2 = P.dt * P.dt;

    // ------------------------------------------------------------------ prologue (all warps)
    for (int i = tid; i < P.smem_doubles; i += ZTH) sm[i] = 0.0;
    __syncthreads();
    for (int e = tid; e < 8 * MAXKS * 32; e += ZTH) {
        const int mt = e / (MAXKS * 32), k = (e / 32) % MAXKS, ln = e % 32;
        const int r = 8 * mt + (ln >> 2), cidx = 4 * k + (ln & 3);
        const bool in = r < nd && cidx < nd;
        sm[P.off_opD + e] = in ? P.dyn[(size_t)r * P.ldn + cidx] : 0.0;
        if (HASQ) sm[P.off_opQ + e] = in ? P.aqT[(size_t)r * P.ldn + cidx] : 0.0;
    }
    const int tn0 = (int)(P.t0 % P.nmd);
    for (int b = 0; b < nb; ++b) {
        const FusedBath& B = P.b[b];
        if (B.direct) continue;
        int* INREM = reinterpret_cast<int*>(sm + B.off_inrem);
        for (int i = tid; i < B.ncp; i += ZTH)
            INREM[i] = (P.rem >= 0 && b != P.rem && i < B.nc) ? (P.b[P.rem].inv_g[B.cids_g[i]] >= 0) : 0;
        // noise minus tail at time t0 (buffer 0)
        double* NB = sm + P.off_nb3 + B.nbrow0 * LD;
        for (int e = tid; e < B.nc * N; e += ZTH) {
            const int r = e / N, n = e % N;
            double v = B.noise[(size_t)tn0 * B.nld + (size_t)r * ldb + col0 + n];
            if (B.nonlocal) v -= B.tail_cur[(size_t)r * ldb + col0 + n];
            NB[r * LD + n] = v;
        }
    }
    if (HASCON)
        for (int e = tid; e < 4 * P.ncon * nd; e += ZTH) sm[P.off_conv + (e / nd) * NDP + e % nd] = P.con3[e];    // c, D c, AqT c, Mp^T c

    const int xq = P.off_xq;
#ifdef PYMD_FUSED_TIMING
    long long t3acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t3last = clock64();
#define TICK3(i) { const long long _n = clock64(); t3acc[i] += _n - t3last; t3last = _n; }
#else
#define TICK3(i)
#endif
    if (roww) {
        // ================================================================== ROW WARPS
#if Z_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Z_ROW_REGS));      // the bath warp group hands registers over
#endif
        const int u = w >> 2, wr = w & 3, cu = NU * u;
        int row[2], own[2];
        bool live[2];
        // per-thread packed rows of the small baths' noise-minus-tail entries: kept in shared memory (three uses per
        // step; as registers they were spilled to local memory and reloaded in front of the last barrier of the step)
        unsigned* NBT = reinterpret_cast<unsigned*>(sm + P.off_nbt);
        unsigned nbpk[2];
        double mrem[2];
        const double* dnz[2];
        size_t dnz_stride = 0;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            row[m] = 16 * wr + 8 * m + lr;
            live[m] = row[m] < nd;
            own[m] = row[m] * LD + cu + 2 * lc;
            nbpk[m] = 0;
            mrem[m] = 0.0;
            dnz[m] = nullptr;
#pragma unroll
            for (int i = 0; i < FB_MAX; ++i) {
                unsigned r = (unsigned)P.nbzero;
                if (i < P.nsm && live[m]) {
                    const int ip = P.b[P.smb[i]].inv_g[row[m]];
                    if (ip >= 0) r = (unsigned)(P.b[P.smb[i]].nbrow0 + ip);
                }
                nbpk[m] |= r << (8 * i);
            }
            NBT[m * (ZU * ZRW * 32) + tid] = nbpk[m];
            if (P.rem >= 0 && live[m]) {
                const int ip = P.b[P.rem].inv_g[row[m]];
                if (ip >= 0) {
                    mrem[m] = 1.0;
                    if (P.b[P.rem].direct) {
                        dnz[m] = P.b[P.rem].noise + (size_t)ip * ldb + col0 + cu + 2 * lc;
                        dnz_stride = (size_t)P.b[P.rem].nld;
                    }
                }
            }
        }
        int xpc = P.off_xp0, xpn = P.off_xp1;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (!live[m]) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int n = cu + 8 * j + 2 * lc;
                const double2 pv = ld2(&P.p[(size_t)row[m] * ldb + col0 + n]);
                const double2 qv = ld2(&P.q[(size_t)row[m] * ldb + col0 + n]);
                st2(&sm[xpc + own[m] + 8 * j], pv.x, pv.y);
                st2(&sm[xq + own[m] + 8 * j], qv.x, qv.y);
                // p_c(t0) opens the block history and is appended to the HBM ring (functions.rpadleft)
                for (int b = 0; b < nb; ++b) {
                    const FusedBath& B = P.b[b];
                    if (B.direct) continue;
                    const int ip = B.inv_g[row[m]];
                    if (ip < 0) continue;
                    st2(&sm[B.off_hist + ip * LD + n], pv.x, pv.y);
                    if (B.nonlocal) st2(&B.ring[((size_t)B.slot0 * B.ncp + ip) * ldb + col0 + n], pv.x, pv.y);
                }
            }
        }
        zbar_sync(0, ZTH);

        double aM[2][KS];
        auto load_mp = [&]() {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int k = 0; k < KS; ++k)
                    aM[m][k] = (live[m] && 4 * k + lc < nd) ? __ldg(&P.mp[(size_t)row[m] * P.ldn + 4 * k + lc]) : 0.0;
        };
        // acc[m][j] += Mp(rows of this warp) . X(unit's columns)
        auto mp_matvec = [&](const double* X, double (&acc)[2][NJ][2]) {
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const double* xr = X + (4 * k + lc) * LD + lr + cu;
                double bv[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bv[j] = xr[8 * j];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[m][j][0], acc[m][j][1], aM[m][k], bv[j]);
            }
        };
        // a0 += dyn . X, a1 += AqT . X  (fragment-ordered shared-memory operators)
        auto dq_matvec = [&](const double* X, double (&a0)[2][NJ][2], double (&a1)[2][NJ][2]) {
            const double* Ad = sm + P.off_opD + (2 * wr * MAXKS) * 32 + lane;
            const double* Aq = sm + P.off_opQ + (2 * wr * MAXKS) * 32 + lane;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const double* xr = X + (4 * k + lc) * LD + lr + cu;
                double bv[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bv[j] = xr[8 * j];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const double ad = Ad[(m * MAXKS + k) * 32];
                    const double aq = HASQ ? Aq[(m * MAXKS + k) * 32] : 0.0;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884(a0[m][j][0], a0[m][j][1], ad, bv[j]);
                        if (HASQ) dmma884(a1[m][j][0], a1[m][j][1], aq, bv[j]);
                    }
                }
            }
        };
        auto zero3 = [](double (&a)[2][NJ][2]) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < NJ; ++j) a[m][j][0] = a[m][j][1] = 0.0;
        };
        // noise row of the direct (memory-less remainder) bath at time row trow, own entries
        auto load_direct = [&](int trow, double (&z)[2][NJ][2]) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    z[m][j][0] = z[m][j][1] = 0.0;
                    if (dnz[m]) {
                        const double2 v = ld2(dnz[m] + (size_t)trow * dnz_stride + 8 * j);
                        z[m][j][0] = v.x; z[m][j][1] = v.y;
                    }
                }
        };
        // noise-minus-tail rows of the shared-memory baths in buffer `buf` for entry (m, j): loads only (all entries
        // of a round are loaded before any of them is used or stored, so that the loads overlap)
        constexpr int NSL = SPEC ? 2 : FB_MAX;
        constexpr int BOFF = SPEC == 1 ? 1 : 0;      // bath index of the first shared-memory bath in the specialised layouts
        auto nb_load = [&](int buf, int m, int j, double2 (&v)[NSL]) {
            const double* base = sm + P.off_nb3 + buf * P.nbstride + cu + 8 * j + 2 * lc;
            const unsigned pk = NBT[m * (ZU * ZRW * 32) + tid];
#pragma unroll
            for (int i = 0; i < NSL; ++i) {
                v[i] = make_double2(0.0, 0.0);
                if (SPEC || i < P.nsm) v[i] = ld2(base + ((pk >> (8 * i)) & 255u) * LD);
            }
        };
        // f + sum over the baths, in bath order (the order fused2 and the 8-warp kernel add them in): `nz` for the
        // direct remainder bath at its position; the remainder bath's own share (a shared-memory bath) is returned too
        auto nb_add = [&](double2 f, const double2 (&v)[NSL], double nz0, double nz1, double2& rem_part) {
            rem_part = make_double2(0.0, 0.0);
            if (SPEC == 1) {
                f.x += nz0; f.y += nz1;
                f.x += v[0].x; f.y += v[0].y;
                f.x += v[1].x; f.y += v[1].y;
                return f;
            }
            if (SPEC == 2) {
                rem_part = v[0];
                f.x += v[0].x; f.y += v[0].y;
                f.x += v[1].x; f.y += v[1].y;
                return f;
            }
#pragma unroll
            for (int i = 0; i <= FB_MAX; ++i) {
                if (i == P.dpos) { f.x += nz0; f.y += nz1; }
                if (i < FB_MAX && i < P.nsm) {
                    f.x += v[i < NSL ? i : 0].x; f.y += v[i < NSL ? i : 0].y;
                    if (i == P.remi) rem_part = v[i < NSL ? i : 0];
                }
            }
            return f;
        };

        // ring slot that receives p_c(t+1) of the current step, per shared-memory bath (counted, not divided)
        int rslot[NSL], rmod[NSL];
#pragma unroll
        for (int i = 0; i < NSL; ++i) {
            const FusedBath& B = P.b[SPEC ? i + BOFF : P.smb[i < P.nsm ? i : 0]];
            rmod[i] = B.R;
            rslot[i] = (B.slot0 + 1) % B.R;
        }
        double fpot[2][NJ][2], uq[2][NJ][2], fb[2][NJ][2];
        zero3(fpot); zero3(uq); zero3(fb);
        int nbuf = 0;
        int tn = tn0;
        for (int g0 = 0; g0 < P.nsteps; g0 += FT) {
            const int gend = min(P.nsteps, g0 + FT);
            TICK3(7)
            if (P.mid) zmid_phase<NJ>(P, sm, P.off_xp1, gend - g0, g0, col0, u, wr, lane);   // block starts find the head buffers in their initial roles
            load_mp();
            TICK3(6)
            if (g0 == 0) {
                // forces at the initial q
                dq_matvec(sm + xq, fpot, uq);
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) { fpot[m][j][0] = -fpot[m][j][0]; fpot[m][j][1] = -fpot[m][j][1]; }
                if (HASCON && P.carry_valid) {
                    // constrained runs carry md.potforce's cache (md.py:456-457) across launches
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (!live[m]) continue;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const int n = col0 + cu + 8 * j + 2 * lc;
                            if (P.samef[n]) fpot[m][j][0] = P.fpotn[(size_t)row[m] * ldb + n];
                            if (P.samef[n + 1]) fpot[m][j][1] = P.fpotn[(size_t)row[m] * ldb + n + 1];
                        }
                    }
                }
                zbar_sync(bar_R(u), ZRW * 32);          // every warp of the unit has read q before round A overwrites it
            }
            zbar_arrive(bar_H(u), (ZRW + ZBW) * 32);    // history row and pre-block tails of step g0 are in place
#if Z_OFFSET
            if (g0 == 0 && u == 1 && P.nsteps > 1) zbar_sync(7, 2 * ZRW * 32);
#endif
            for (int g = g0; g < gend; ++g) {
                const int s = g - g0;
                const int tn_cur = tn;
                const int tn1 = tn + 1 == P.nmd ? P.wrap_row : tn + 1;
                {
                    // ============ round A: first force evaluation at time t, head p(t) ============
                    double hA[2][NJ][2], nzr[2][NJ][2];
                    load_direct(tn_cur, nzr);
                    zero3(hA);
                    mp_matvec(sm + xpc, hA);
                    TICK3(0)
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.0;
                    // stage 1: every shared-memory load; stage 2: arithmetic; stage 3: the stores (a store between
                    // two entries would order the loads of the second behind it)
                    double2 pvv[2][NJ], qvv[2][NJ], nbv[2][NJ][NSL];
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            pvv[m][j] = live[m] ? ld2(&sm[xpc + own[m] + 8 * j]) : make_double2(0.0, 0.0);
                            qvv[m][j] = ld2(&sm[xq + own[m] + 8 * j]);
                            nb_load(nbuf, m, j, nbv[m][j]);
                        }
                    double2 phv[2][NJ], qnv[2][NJ];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const double2 pv = pvv[m][j], qv = qvv[m][j];
                            double2 cr;
                            const double2 fs = nb_add(make_double2(fpot[m][j][0] + uq[m][j][0] - hA[m][j][0],
                                                                   fpot[m][j][1] + uq[m][j][1] - hA[m][j][1]),
                                                      nbv[m][j], nzr[m][j][0], nzr[m][j][1], cr);
                            const double f0 = fs.x, f1 = fs.y;
                            // force of the remainder bath on this row: its noise (minus tail), minus ALL heads (the
                            // heads of the small baths on shared dofs are added back by the bath warps), plus AqT q
                            double c0 = cr.x + nzr[m][j][0] - mrem[m] * hA[m][j][0];
                            double c1 = cr.y + nzr[m][j][1] - mrem[m] * hA[m][j][1];
                            if (HASQ) { c0 += mrem[m] * uq[m][j][0]; c1 += mrem[m] * uq[m][j][1]; }
                            phv[m][j] = make_double2(pv.x + f0 * dt / 2.0, pv.y + f1 * dt / 2.0);
                            qnv[m][j] = make_double2(qv.x + pv.x * dt + f0 * dt2 / 2.0, qv.y + pv.y * dt + f1 * dt2 / 2.0);
                            v[4 * j + 0] += pv.x * pv.x;
                            v[4 * j + 1] += pv.y * pv.y;
                            v[4 * j + 2] += pv.x * c0;
                            v[4 * j + 3] += pv.y * c1;
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (!live[m]) continue;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const int n = cu + 8 * j + 2 * lc;
                            st2(&sm[xpn + own[m] + 8 * j], phv[m][j].x, phv[m][j].y);
                            st2(&sm[xq + own[m] + 8 * j], qnv[m][j].x, qnv[m][j].y);
                            if (P.ps) st2(&P.ps[(size_t)tn_cur * P.rec_ld + (size_t)row[m] * ldb + col0 + n], pvv[m][j].x, pvv[m][j].y);
                            if (P.qs) st2(&P.qs[(size_t)tn_cur * P.rec_ld + (size_t)row[m] * ldb + col0 + n], qvv[m][j].x, qvv[m][j].y);
                        }
                    }
                    // column sums over this warp's 16 rows: lane lr ends with value lr = (n-tile lr>>2, kind lr&3)
                    const double rsum = rs8(v, lr);
                    const int j = lr >> 2, kind = lr & 3;
                    if (j < NJ) {
                        const int bsel = kind < 2 ? nb : P.rem;
                        if (bsel >= 0) sm[P.off_curp + (bsel * ZSL + wr) * N + cu + 8 * j + 2 * lc + (kind & 1)] = rsum;
                    }
                }
                TICK3(1)
                zbar_sync(bar_NBR(u), (ZRW + ZBW) * 32);                        // ---- S2: unit's row warps + bath warps' arrival
                TICK3(5)
                nbuf ^= 1;
                { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p(t+1/2)
                {
                    // ============ round B: second evaluation at time t+1, head p(t+1/2) ============
                    double aDq[2][NJ][2], hB[2][NJ][2], nzr[2][NJ][2];
                    load_direct(tn1, nzr);
                    zero3(aDq); zero3(uq);
                    dq_matvec(sm + xq, aDq, uq);
                    // currents and kinetic energy of time t: fixed-order sum of the per-warp slots (placed behind the
                    // first operator product of the round so that its loads and adds run while the DMMAs drain)
                    for (int e = wr * 32 + lane; e < (nb + 1) * NU; e += ZRW * 32) {
                        const int b = e / NU, n = cu + e % NU;
                        double sl[ZSL];
#pragma unroll
                        for (int ww = 0; ww < ZSL; ++ww) sl[ww] = sm[P.off_curp + (b * ZSL + ww) * N + n];
                        double sum = 0.0;
#pragma unroll
                        for (int ww = 0; ww < ZSL; ++ww) sum += sl[ww];
                        if (b < nb) P.b[b].cur[(size_t)tn_cur * ldb + col0 + n] = sum;
                        else P.etot[(size_t)tn_cur * ldb + col0 + n] = 0.5 * sum;
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) { fpot[m][j][0] = -aDq[m][j][0]; fpot[m][j][1] = -aDq[m][j][1]; }
                    zero3(hB);
                    mp_matvec(sm + xpc, hB);
                    TICK3(2)
                    double cv[2], wv[2];
                    double vB[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) vB[i] = 0.0;
                    if (HASCON) {
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            cv[m] = sm[P.off_conv + row[m]];
                            wv[m] = sm[P.off_conv + 3 * NDP + row[m]];
                        }
                    }
                    double2 phB[2][NJ], q1B[2][NJ], nbB[2][NJ][NSL], p1B[2][NJ];
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            phB[m][j] = ld2(&sm[xpc + own[m] + 8 * j]);
                            q1B[m][j] = HASCON ? ld2(&sm[xq + own[m] + 8 * j]) : make_double2(0.0, 0.0);
                            nb_load(nbuf, m, j, nbB[m][j]);
                        }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            double2 cr;
                            const double2 fs = nb_add(make_double2(fpot[m][j][0] + uq[m][j][0], fpot[m][j][1] + uq[m][j][1]),
                                                      nbB[m][j], nzr[m][j][0], nzr[m][j][1], cr);
                            const double f0 = fs.x, f1 = fs.y;
                            fb[m][j][0] = f0;
                            fb[m][j][1] = f1;
                            const double2 ph = phB[m][j];
                            const double p1x = ph.x + dt * (f0 - hB[m][j][0]) / 2.0, p1y = ph.y + dt * (f1 - hB[m][j][1]) / 2.0;
                            p1B[m][j] = make_double2(p1x, p1y);
                            if (HASCON && live[m]) {
                                // md.ApplyConstraint (md.py:739-762), one vector: c.p(t+1) before projection is
                                // c.(p1/2 + dt/2 f) - dt/2 (Mp^T c).p1 by linearity, and c.q' is known since round A:
                                // both dots are reduced here, so round C needs no barrier of its own
                                const double2 q1 = q1B[m][j];
                                vB[4 * j + 0] += cv[m] * (ph.x + dt * f0 / 2.0) - dt * (wv[m] * p1x) / 2.0;
                                vB[4 * j + 1] += cv[m] * (ph.y + dt * f1 / 2.0) - dt * (wv[m] * p1y) / 2.0;
                                vB[4 * j + 2] += cv[m] * q1.x;
                                vB[4 * j + 3] += cv[m] * q1.y;
                            }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (!live[m]) continue;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) st2(&sm[xpn + own[m] + 8 * j], p1B[m][j].x, p1B[m][j].y);
                    }
                    if (HASCON) {
                        const double part = rs8(vB, lr);      // lane lr: value lr = (n-tile lr>>2, kind lr&3: d0, d1, g0, g1)
                        sm[P.off_conp + ((u * 8 + lr) * ZRW + wr) * 4 + lc] = part;
                    }
                }
                TICK3(3)
                zbar_sync(bar_R(u), ZRW * 32);                                  // ---- S3
                TICK3(5)
#if Z_OFFSET
                if (g == 0 && u == 0 && P.nsteps > 1) zbar_arrive(7, 2 * ZRW * 32);
#endif
                { const int tmp = xpc; xpc = xpn; xpn = tmp; }      // xpc: p1, xpn: p(t+1/2)
                {
                    // ============ round C: third evaluation, head p1 -> p(t+1), constraint ============
                    // history rows of this thread's two operator rows in the shared-memory baths (-1: not a bath dof),
                    // extracted before the product so that nothing is fetched between it and the last barrier
                    int hrow[2][NSL];
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int i = 0; i < NSL; ++i) {
                            const int r = (int)((NBT[m * (ZU * ZRW * 32) + tid] >> (8 * i)) & 255u);
                            hrow[m][i] = ((SPEC || i < P.nsm) && r != P.nbzero) ? r - P.b[SPEC ? i + BOFF : P.smb[i]].nbrow0 : -1;
                        }
                    double hC[2][NJ][2];
                    zero3(hC);
                    mp_matvec(sm + xpc, hC);
                    TICK3(4)
                    double pn[2][NJ][2], qn[2][NJ][2];
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const double2 ph = ld2(&sm[xpn + own[m] + 8 * j]);
                            pn[m][j][0] = live[m] ? ph.x + dt * (fb[m][j][0] - hC[m][j][0]) / 2.0 : 0.0;
                            pn[m][j][1] = live[m] ? ph.y + dt * (fb[m][j][1] - hC[m][j][1]) / 2.0 : 0.0;
                            qn[m][j][0] = qn[m][j][1] = 0.0;
                            if (HASCON) {
                                const double2 q0 = ld2(&sm[xq + own[m] + 8 * j]);
                                qn[m][j][0] = q0.x; qn[m][j][1] = q0.y;
                            }
                        }
                    if (HASCON) {
                        double cv[2], dv[2], av[2];
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            cv[m] = sm[P.off_conv + row[m]];
                            dv[m] = sm[P.off_conv + NDP + row[m]];
                            av[m] = HASQ ? sm[P.off_conv + 2 * NDP + row[m]] : 0.0;
                        }
                        double tot = 0.0;
#pragma unroll
                        for (int ww = 0; ww < ZRW; ++ww) tot += sm[P.off_conp + ((u * 8 + lr) * ZRW + ww) * 4 + lc];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            double d[4];
#pragma unroll
                            for (int kind = 0; kind < 4; ++kind) d[kind] = __shfl_sync(0xffffffffu, tot, ((4 * j + kind) << 2) | lc);
#pragma unroll
                            for (int m = 0; m < 2; ++m) {
                                pn[m][j][0] -= d[0] * cv[m];
                                pn[m][j][1] -= d[1] * cv[m];
                                qn[m][j][0] -= d[2] * cv[m];
                                qn[m][j][1] -= d[3] * cv[m];
                                // md.sameq (md.py:724-736): max_i |q(t+1)_i - q'_i| = |dq| max_i |c_i|; when it reaches 1e-9
                                // the cached force is dropped: -D q(t+1) = -D q' + dq (D c) by linearity
                                if (!(fabs(d[2]) * P.cmax < 10e-10)) fpot[m][j][0] += d[2] * dv[m];
                                if (!(fabs(d[3]) * P.cmax < 10e-10)) fpot[m][j][1] += d[3] * dv[m];
                                if (HASQ) { uq[m][j][0] -= d[2] * av[m]; uq[m][j][1] -= d[3] * av[m]; }      // AqT q(t+1)
                            }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        if (!live[m]) continue;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            st2(&sm[xpn + own[m] + 8 * j], pn[m][j][0], pn[m][j][1]);
                            if (HASCON) st2(&sm[xq + own[m] + 8 * j], qn[m][j][0], qn[m][j][1]);
                        }
                        // p_c(t+1) enters the block history / the HBM ring
#pragma unroll
                        for (int i = 0; i < NSL; ++i) {
                            const int ip = hrow[m][i];
                            if (ip < 0) continue;
                            const FusedBath& B = P.b[SPEC ? i + BOFF : P.smb[i]];
                            if (!B.nonlocal) {
#pragma unroll
                                for (int j = 0; j < NJ; ++j) st2(&sm[B.off_hist + ip * LD + cu + 8 * j + 2 * lc], pn[m][j][0], pn[m][j][1]);
                            } else if (g + 1 < P.nsteps) {
                                const int hb = B.off_hist + (((s + 1) & (FT - 1)) * B.ncp + ip) * LD + cu + 2 * lc;
                                double* rp = &B.ring[((size_t)rslot[i] * B.ncp + ip) * ldb + col0 + cu + 2 * lc];
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
                                    st2(&sm[hb + 8 * j], pn[m][j][0], pn[m][j][1]);
                                    st2(rp + 8 * j, pn[m][j][0], pn[m][j][1]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NSL; ++i)
                        if (++rslot[i] >= rmod[i]) rslot[i] = 0;
                }
                if (g + 1 < gend) zbar_arrive(bar_H(u), (ZRW + ZBW) * 32);      // history row of the next step is in place
                TICK3(3)
                zbar_sync(bar_R(u), ZRW * 32);                                  // ---- S1
                TICK3(5)
                { const int tmp = xpc; xpc = xpn; xpn = tmp; }                  // xpc: p(t+1)
                tn = tn1;
            }
        }
#ifdef PYMD_FUSED_TIMING
        if (P.timing && lane == 0)
            for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&P.timing[i]), (unsigned long long)t3acc[i]);
#endif
        // ------------------------------------------------------------------ epilogue
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (!live[m]) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int n = cu + 8 * j + 2 * lc;
                const double2 pv = ld2(&sm[xpc + own[m] + 8 * j]);
                const double2 qv = ld2(&sm[xq + own[m] + 8 * j]);
                st2(&P.p[(size_t)row[m] * ldb + col0 + n], pv.x, pv.y);
                st2(&P.q[(size_t)row[m] * ldb + col0 + n], qv.x, qv.y);
                st2(&P.fpotn[(size_t)row[m] * ldb + col0 + n], fpot[m][j][0], fpot[m][j][1]);
            }
        }
    } else {
        // ================================================================== BATH WARPS
#if Z_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Z_BATH_REGS));
#endif
        const int bw = w - ZU * ZRW;
        const int pb = P.pair_b[bw], pmt = P.pair_mt[bw];
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
        zbar_sync(0, ZTH);
        int tn = tn0;
        for (int g = 0; g < P.nsteps; ++g) {
            const int s = g & (FT - 1);
            const int nbuf = g & 1;
            const int tn1 = tn + 1 == P.nmd ? P.wrap_row : tn + 1;
            {
                // L2 prefetch of the noise rows two steps ahead: the 128 bath threads cover the 128-byte lines of all rows
                const int tn2 = tn1 + 1 >= P.nmd ? P.wrap_row : tn1 + 1;
                int e = tid - ZU * ZRW * 32;
                for (int b = 0; b < nb; ++b) {
                    constexpr int LPR = (N * 8 + 127) / 128;       // 128-byte lines per row tile
                    const int cnt = P.b[b].nc * LPR;
                    for (; e < cnt; e += ZBW * 32)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.b[b].noise + (size_t)tn2 * P.b[b].nld + (size_t)(e / LPR) * ldb + col0 + (e % LPR) * 16));
                    e -= cnt;
                }
            }
#pragma unroll 1
            for (int u = 0; u < ZU; ++u) {
                const int cu = NU * u;
                double2 xi[NJ], pt[NJ];
                const bool valid = pb >= 0 && pmt * 8 + lr < P.b[pb >= 0 ? pb : 0].nc;
                const FusedBath& B = P.b[pb >= 0 ? pb : 0];
                const int rr = pmt * 8 + lr;
                // the rows of the NEXT time are read before waiting for the unit: they do not depend on it
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    xi[j] = make_double2(0.0, 0.0);
                    if (valid) xi[j] = ld2(&B.noise[(size_t)tn1 * B.nld + (size_t)rr * ldb + col0 + cu + 8 * j + 2 * lc]);
                }
                // ... and so are the pre-block tails P[s] except at a block start, where the unit's mid phase writes
                // them just before it signals H
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    pt[j] = make_double2(0.0, 0.0);
                    if (valid && B.nonlocal && s > 0) pt[j] = __ldcg(reinterpret_cast<const double2*>(&B.P[((size_t)s * B.nc + rr) * ldb + col0 + cu + 8 * j + 2 * lc]));
                }
                TICK3(1)
                zbar_sync(bar_H(u), (ZRW + ZBW) * 32);  // p_c(t) of the unit is in the block history, P[s] is written
                TICK3(5)
                if (pb >= 0) {
                    const int KQ = B.ncp / 4;
                    if (s == 0) {
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
                            if (valid && B.nonlocal) pt[j] = __ldcg(reinterpret_cast<const double2*>(&B.P[((size_t)s * B.nc + rr) * ldb + col0 + cu + 8 * j + 2 * lc]));
                    }
                    const int hs = B.nonlocal ? s : 0;
                    if (pb != P.rem) {
                        // heat current of this pair's rows at time t: p_c . (xi - tail - K[0] p_c); the head on dofs shared
                        // with the remainder bath is handed back to the remainder's current
                        const int inrem = valid ? reinterpret_cast<const int*>(sm + B.off_inrem)[rr] : 0;
                        double gh[NJ][2];
                        zero_acc<NJ>(gh);
                        const double* H = sm + B.off_hist + (hs * B.ncp + lc) * LD + lr + cu;
#pragma unroll
                        for (int kq = 0; kq < 4; ++kq) {
                            if (kq < KQ) {
#pragma unroll
                                for (int j = 0; j < NJ; ++j) dmma884(gh[j][0], gh[j][1], hf[kq], H[(4 * kq) * LD + 8 * j]);
                            }
                        }
                        double v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = 0.0;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (valid) {
                                const double2 pv = ld2(&sm[B.off_hist + (hs * B.ncp + rr) * LD + cu + 8 * j + 2 * lc]);
                                const double2 nv = ld2(&sm[P.off_nb3 + nbuf * P.nbstride + (B.nbrow0 + rr) * LD + cu + 8 * j + 2 * lc]);
                                v[4 * j + 0] = pv.x * (nv.x - gh[j][0]);
                                v[4 * j + 1] = pv.y * (nv.y - gh[j][1]);
                                v[4 * j + 2] = inrem ? pv.x * gh[j][0] : 0.0;
                                v[4 * j + 3] = inrem ? pv.y * gh[j][1] : 0.0;
                            }
                        }
                        const double r = rs8(v, lr);     // lane lr: n-tile lr>>2, kind lr&3 (own current x2, remainder share x2)
                        const int j = lr >> 2, kind = lr & 3;
                        if (j < NJ) {
                            if (kind < 2) sm[P.off_curp + (pb * ZSL + ZRW + bw) * N + cu + 8 * j + 2 * lc + kind] = r;
                            else if (P.rem >= 0) sm[P.off_curp + (P.rem * ZSL + ZRW + bw) * N + cu + 8 * j + 2 * lc + kind - 2] = r;
                        }
                    }
                    // noise minus memory-kernel tail at time t+1: xi(t+1) - P[s] - sum_{k=1}^{s+1} dt K[k] p_c(t+1-k)
                    double acc[NJ][2], acc2[NJ][2];
                    zero_acc<NJ>(acc);
                    zero_acc<NJ>(acc2);
                    if (B.nonlocal) {
                        const int ntap = min(s + 1, B.ntap);
                        const double* H0 = sm + B.off_hist + lc * LD + lr + cu;
#pragma unroll
                        for (int k = 1; k <= FT; ++k) {
                            if (k <= ntap) {
                                const double* H = H0 + (s + 1 - k) * B.ncp * LD;        // p_c(t+1-k)
#pragma unroll
                                for (int kq = 0; kq < 4; ++kq) {
                                    if (kq < KQ) {
#pragma unroll
                                        for (int j = 0; j < NJ; ++j) {
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
                        for (int j = 0; j < NJ; ++j) {
                            const double t0v = pt[j].x + (acc[j][0] + acc2[j][0]), t1v = pt[j].y + (acc[j][1] + acc2[j][1]);
                            st2(&sm[P.off_nb3 + (nbuf ^ 1) * P.nbstride + (B.nbrow0 + rr) * LD + cu + 8 * j + 2 * lc], xi[j].x - t0v, xi[j].y - t1v);
                            if (B.nonlocal && g == P.nsteps - 1)
                                st2(&B.tail_cur[(size_t)rr * ldb + col0 + cu + 8 * j + 2 * lc], t0v, t1v);
                        }
                    }
                }
                zbar_arrive(bar_NBR(u), (ZRW + ZBW) * 32);
                TICK3(0)
            }
            tn = tn1;
        }
#ifdef PYMD_FUSED_TIMING
        if (P.timing && lane == 0)
            for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&P.timing[16 + i]), (unsigned long long)t3acc[i]);
#endif
    }
    for (int n = tid; n < N; n += ZTH) P.samef[col0 + n] = 1;     // the stored force is the force at q(t)
}

}  // namespace

// Shared-memory plan of fused3 for NJ n-tiles per unit; fills the offsets, the pair table and the list of
// shared-memory ("small") baths.  Returns bytes, or a huge value when the configuration does not qualify.
size_t plan3(const pymd_ctx* c, FusedParams& fp, bool hasq, int NJ) {
    const int N = ZU * 8 * NJ, LD = N + 4;
    const size_t no = (size_t)1 << 40;
    if (fp.aq >= 0 && fp.aq != fp.rem) return no;      // the q-coupled bath must be the remainder bath
    if (c->ncon > 1) return no;
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 1) / 2 * 2; return o; };
    fp.off_xp0 = take(NDP * LD);
    fp.off_xp1 = take(NDP * LD);
    fp.off_xq = take(NDP * LD);
    fp.off_opD = take(8 * MAXKS * 32);
    fp.off_opQ = hasq ? take(8 * MAXKS * 32) : fp.off_opD;
    fp.off_zero = 0;
    fp.off_curp = take((c->nbath + 1) * ZSL * N);
    fp.off_conp = take(ZU * 8 * ZRW * 4);
    fp.off_conv = take(4 * NDP);
    fp.off_nbt = take(ZU * ZRW * 32);            // 2 x 256 unsigned
    int npair = 0, nrows = 0;
    fp.nsm = 0;
    fp.remi = -1;
    fp.dpos = -1;
    for (int i = 0; i < 4; ++i) fp.pair_b[i] = fp.pair_mt[i] = -1;
    for (int i = 0; i < FB_MAX; ++i) fp.smb[i] = 0;
    for (int b = 0; b < c->nbath; ++b) {
        const BathHost& H = c->bath[b];
        FusedBath& B = fp.b[b];
        B.nc = H.nc; B.ncp = H.ncp; B.ml = H.ml; B.nonlocal = H.ml > 1;
        B.direct = (b == fp.rem && H.ml == 1) ? 1 : 0;
        B.ldk = H.ncp + 4;
        B.ntap = H.ml > 1 ? std::min(FT, H.ml - 1) : 0;
        B.off_cids = B.off_inrem = B.off_hd = B.off_nb = B.off_kin = B.off_hist = 0;
        B.nbrow0 = 0;
        if (B.direct) { fp.dpos = fp.nsm; continue; }
        if (H.ncp > 16) return no;
        for (int mt = 0; mt < H.ncp / 8; ++mt) {
            if (npair >= ZBW) return no;
            fp.pair_b[npair] = b; fp.pair_mt[npair] = mt; ++npair;
        }
        if (b == fp.rem) fp.remi = fp.nsm;
        fp.smb[fp.nsm++] = b;
        B.nbrow0 = nrows;
        nrows += H.ncp;
        B.off_inrem = take(H.ncp / 2 + 1);
        B.off_hist = take((H.ml > 1 ? FT : 1) * H.ncp * LD);
    }
    fp.nbzero = nrows;                                  // the zero row follows the rows of the small baths
    fp.nbstride = (nrows + 1) * LD;
    fp.off_nb3 = take(2 * fp.nbstride);
    fp.smem_doubles = off;
    return (size_t)off * sizeof(double);
}

namespace {
template <int NJ, bool Q, bool C, int KS, int SPEC>
int launch_f3_t(const FusedParams& fp, size_t smem, int grid, cudaStream_t st) {
    static bool cfg = false;
    if (!cfg) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(fused3_kernel<NJ, Q, C, KS, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cfg = true;
    }
    fused3_kernel<NJ, Q, C, KS, SPEC><<<grid, ZTH, smem, st>>>(fp);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}
template <int NJ, int KS, int SPEC>
int launch_f3_k(const FusedParams& fp, size_t smem, int grid, bool hasq, bool hascon, cudaStream_t st) {
    if (hasq) return hascon ? launch_f3_t<NJ, true, true, KS, SPEC>(fp, smem, grid, st) : launch_f3_t<NJ, true, false, KS, SPEC>(fp, smem, grid, st);
    return hascon ? launch_f3_t<NJ, false, true, KS, SPEC>(fp, smem, grid, st) : launch_f3_t<NJ, false, false, KS, SPEC>(fp, smem, grid, st);
}
}  // namespace

int launch_f3(const FusedParams& fp, size_t smem, int NJ, bool hasq, bool hascon, cudaStream_t st) {
    const int grid = fp.ldb / (ZU * 8 * NJ);
    const bool k15 = fp.nd <= 60;
    // the usual layouts (see the kernel's SPEC parameter)
    int spec = 0;
    if (fp.nbath == 3 && fp.nsm == 2 && fp.dpos == 0 && fp.remi < 0) spec = 1;
    else if (fp.nbath == 2 && fp.nsm == 2 && fp.dpos < 0 && fp.remi == 0 && !hasq) spec = 2;
#define PYMD_F3_CASE(NJV, KSV)                                                                                         \
    {                                                                                                                  \
        if (spec == 1) return launch_f3_k<NJV, KSV, 1>(fp, smem, grid, hasq, hascon, st);                              \
        if (spec == 2) return hascon ? launch_f3_t<NJV, false, true, KSV, 2>(fp, smem, grid, st)                       \
                                     : launch_f3_t<NJV, false, false, KSV, 2>(fp, smem, grid, st);                     \
        return launch_f3_k<NJV, KSV, 0>(fp, smem, grid, hasq, hascon, st);                                             \
    }
    if (NJ == 2) {
        if (k15) PYMD_F3_CASE(2, 15)
        PYMD_F3_CASE(2, 16)
    }
    if (k15) PYMD_F3_CASE(1, 15)
    PYMD_F3_CASE(1, 16)
#undef PYMD_F3_CASE
}

}  // namespace pymd

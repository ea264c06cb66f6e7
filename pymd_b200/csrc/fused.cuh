// Parameter blocks and small device helpers shared by the on-chip step kernels of fused.cu
// (fused_kernel, fused2_kernel) and fused3.cu (two-unit kernel).
#pragma once
#include "engine.cuh"

namespace pymd {

constexpr int FW = 8;        // warps per CTA
constexpr int NDP = 64;      // padded rows (8 per warp)
constexpr int FT = 8;        // steps per block
constexpr int MAXKS = 16;    // k-steps of the nd x nd operators (nd <= 64)
constexpr int FB_MAX = 4;    // baths supported by the fused path
constexpr int MAXIT = 4;     // work items of each kind per warp

struct FusedBath {
    int nc, ncp, ml, nonlocal, R, slot0, ntap, ldk, direct;       // direct: local remainder bath, noise read straight from HBM
    int off_hd, off_kin, off_hist, off_nb, off_cids, off_inrem;   // smem offsets (doubles)
    int lda;
    int nbrow0;               // fused3: first row of this bath in the combined noise-minus-tail buffer
    long long nld;            // fused3: doubles between consecutive noise rows (nc * ldb unless the records are overlaid)
    const double* noise;
    const double* P;          // [FT][nc][ldb] pre-block tail
    double* tail_cur;         // [nc][ldb] tail(t) carried between launches
    double* ring;
    double* cur;
    const double* head_g;     // [nc][lda]
    const double* kin_g;      // [kBlockT][nc][nc]
    const int* cids_g;
    const int* inv_g;
    // in-kernel mid field (mid == 1): A fragments of dt K[k], k = 2..2+kMidK-1, and the far field
    const double* kmid;       // [kMidK][ncp/8][ncp/4][32]
    const double* F;          // [kFarS][nc][ldb] rows of the current super-block
    const double* kinf;       // fused2: dt K[1..8] as A fragments [8][ncp/8][ncp/4][32]
    const double* hdf;        // fused2: alpha K[0] as A fragments [ncp/8][ncp/4][32]
};

struct FusedParams {
    int nd, ks, ldb, ldn, nbath, nmd, ncon, nsteps, rem, aq, carry_valid;
    int nosplit;              // testing knob (PYMD_FUSED_SPLIT=0): never split work items into n-tile halves
    int mid;                  // 1: several blocks per launch, pre-block tails computed in the kernel
    int r0;                   // steps since the start of the super-block at launch
    long long hc0;            // history rows appended before p(t0)
    long long t0;
    double dt;
    double *p, *q, *ps, *qs, *etot, *fpotn;
    int* samef;
    const double *dyn, *mp, *aqT, *con3;      // con3: [3][ncon][nd] = c, D c, AqT c
    int off_xp0, off_xp1, off_xq, off_curp, off_conp, off_zero, off_conv;
    int off_opD, off_opQ;     // fused2: dyn / AqT as fragment-ordered shared-memory operands
    int off_nbt;              // fused3: per-thread packed noise-minus-tail rows [2][256] (unsigned)
    int off_nb3, nbstride, nbzero;   // fused3: combined noise-minus-tail buffer [2][nbzero + 1][LD], last row zero
    long long rec_ld;         // fused3: doubles between consecutive rows of ps / qs (nd * ldb unless overlaid)
    int wrap_row;             // fused3: noise row read as "row nmd" (0: the noise is periodic in place, phbath.py:196)
    int dpos;                 // fused3: how many shared-memory baths precede the direct remainder bath (-1: none is direct)
    int nsm, smb[4], remi;    // fused3: the shared-memory (non-direct) baths, and the remainder bath's position among them (-1: direct)
    int pair_b[4], pair_mt[4];    // fused2: (bath, m8 tile) owned by each bath warp, -1 = none
    double cmax;              // max_i |c_i| of the (single) constraint vector
    int smem_doubles;
#ifdef PYMD_FUSED_TIMING
    long long* timing;        // [32] accumulated clock64 deltas per phase (tools/build_timing.sh builds only)
#endif
    FusedBath b[FB_MAX];
};

namespace {

__device__ __forceinline__ double colsum8(double v) {     // sum over the 8 lanes sharing lane%4
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}
// Reduce-scatter over the 8 lanes sharing lane%4 (lr = lane>>2): lane lr returns sum_lanes v[lr].
// 7 shuffles for 8 column sums instead of 24.
__device__ __forceinline__ double rs8(const double (&v)[8], int lr) {
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
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

template <int NT>
__device__ __forceinline__ void zero_acc(double (&a)[NT][2]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) a[j][0] = a[j][1] = 0.0;
}


}  // namespace
}  // namespace pymd

// Row-pointer FP64 tensor-core GEMM:  Y[yrow(m)][n] (+)= alpha * sum_k A(m,k) * X[xrow(k)][n]
//
//   * n runs over trajectories (contiguous in memory for X and Y),
//   * every A tile row and every X tile row is a contiguous run in global memory, staged
//     into padded shared memory by ONE TMA bulk copy per row (cp.async.bulk + mbarrier),
//     which gives gather (xrow = cids), scatter (yrow) and the block-Toeplitz memory
//     kernel (A rows slide over a reversed kernel table) for free,
//   * math is DMMA.8x8x4 from shared memory, FP64 accumulate.
#pragma once
#include "common.cuh"

namespace pymd {

struct GemmArgs {
    // A(m, k) = a_base[(m % a_mod) * a_ld + (m / a_mod) * a_ls + seg_aoff[s] + k]  for k in segment s.
    const double* a_base;
    long long a_ld, a_ls;
    int a_mod;
    // X row k lives at x_base + (x_idx ? x_idx[k] : k) * x_ld ; rows outside every segment are 0.
    const double* x_base;
    const int* x_idx;
    long long x_ld;
    // Y row m lives at y_base + (y_idx ? y_idx[m] : m) * y_ld.
    double* y_base;
    const int* y_idx;
    long long y_ld;
    int M, N;
    // K is the union of up to two half-open row ranges (the ring buffer wraps once).
    int nseg;
    int seg_lo[2], seg_hi[2];
    long long seg_aoff[2];
    double alpha;
    int accumulate;
    // batch of independent products (grid.z): operand bases advance by these strides (doubles) per batch index
    int nbatch = 1;
    long long a_bs = 0, x_bs = 0, y_bs = 0;
};

// several independent products in one launch
struct GemmMulti {
    static constexpr int kMax = 4;
    int n = 0;
    GemmArgs a[kMax];
};

// Launch on `stream`; returns a PYMD_* status.
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
int launch_gemm_multi(const GemmMulti& m, cudaStream_t stream);

}  // namespace pymd

// See fft.cuh.  Replaces numpy.fft in myfft.Fourier1D / iFourier1D (reference myfft.py:15-52).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "fft.cuh"

namespace pymd {

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

template <int R>
__device__ __forceinline__ void butterfly(double2* v);

template <>
__device__ __forceinline__ void butterfly<2>(double2* v) {
    const double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(double2* v) {
    const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    const double2 t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<8>(double2* v) {
    const double h = 0.70710678118654752440;
    double2 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        a[r] = cadd(v[r], v[r + 4]);
        b[r] = csub(v[r], v[r + 4]);
    }
    b[1] = make_double2(h * (b[1].x + b[1].y), h * (b[1].y - b[1].x));     // * (1-i)/sqrt2
    b[2] = mul_mi(b[2]);                                                   // * (-i)
    b[3] = make_double2(h * (b[3].y - b[3].x), -h * (b[3].x + b[3].y));    // * (-1-i)/sqrt2
    butterfly<4>(a);
    butterfly<4>(b);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[2 * q] = a[q];
        v[2 * q + 1] = b[q];
    }
}

template <int R>
__global__ void __launch_bounds__(128) fft_pass_kernel(const void* __restrict__ src, void* __restrict__ dst,
                                                       int n, int Ns, long long cols, int conj_in, int conj_out,
                                                       const double2* __restrict__ tw, int twN, int loadmode,
                                                       int storemode, long long src_ld, long long dst_ld,
                                                       double scale, long long dst_cw, long long dst_cs) {
    const int j = blockIdx.y;                   // blockIdx.x: column tile (neighbouring CTAs share DRAM pages)
    const int k = j % Ns;
    const int base = (j / Ns) * Ns * R + k;
    const int nr = n / R;
    double2 w[R];
    if (Ns > 1) {
        const int step = twN / (Ns * R);
#pragma unroll
        for (int r = 1; r < R; ++r) w[r] = tw[(size_t)r * k * step];
    }
    for (long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x; col < cols;
         col += (long long)gridDim.x * blockDim.x) {
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long idx = j + (long long)r * nr;
            if (loadmode == LOAD_C) {
                v[r] = static_cast<const double2*>(src)[idx * src_ld + col];
            } else {
                const double* xr = static_cast<const double*>(src);
                v[r] = make_double2(xr[(2 * idx) * src_ld + col], xr[(2 * idx + 1) * src_ld + col]);
            }
            if (conj_in) v[r].y = -v[r].y;
        }
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
        }
        butterfly<R>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long o = base + (long long)r * Ns;
            double2 out = make_double2(scale * v[r].x, (conj_out ? -scale : scale) * v[r].y);
            if (storemode == STORE_C) {
                static_cast<double2*>(dst)[o * dst_ld + col] = out;
            } else {
                double* dr = static_cast<double*>(dst);
                const long long dc = dst_cw ? (col / dst_cw) * dst_cs + col % dst_cw : col;
                dr[(2 * o) * dst_ld + dc] = out.x;
                dr[(2 * o + 1) * dst_ld + dc] = out.y;
            }
        }
    }
}

// Fused radix R*R pass (R = 8: 64, R = 4: 16).  blockDim = (32 columns, R): thread (x, r1) first
// transforms inputs r1 + R r2 (r2 < R), the R x R intermediate goes through shared memory with
// the inner twiddles W_{R^2}^{r1 q2} applied, thread (x, q2) then transforms over r1 and stores
// outputs R q1 + q2.  Same Stockham indexing as fft_pass_kernel with radix R^2.  The loads of the
// next column tile are in flight while the current one is transformed; per-thread base pointers
// keep the address arithmetic (and with it the register count: two CTAs per SM) small.
template <int R, int LM, int SM>
__global__ void __launch_bounds__(32 * R, 512 / (32 * R)) fft_pass2_kernel(const void* __restrict__ src, void* __restrict__ dst,
                                                                          int n, int Ns, long long cols, int conj_in, int conj_out,
                                                                          const double2* __restrict__ tw, int twN,
                                                                          long long src_ld, long long dst_ld,
                                                                          double scale, long long dst_cw, long long dst_cs) {
    constexpr int RR = R * R;
    __shared__ double2 ex[RR][32];
    // blockIdx.x runs over column tiles so that the CTAs resident at the same time read neighbouring
    // 512-byte pieces of the same rows (DRAM page locality); blockIdx.y is the butterfly index
    const int j = blockIdx.y;
    const int k = j % Ns;
    const int nr = n / RR;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tw0 = ty * k * (twN / (Ns * RR)), twd = R * k * (twN / (Ns * RR));   // outer twiddle index: tw0 + r2 twd
    const int twi = ty * (twN / RR);                                                // inner twiddle index: q2 twi
    const long long cstride = (long long)gridDim.x * 32;
    // element (row r = ty + R r2 of this butterfly, column c): sbase + r2 * sstep + c
    const long long srow = LM == LOAD_C ? 1 : 2;            // real-pair sources hold two rows per complex row
    const long long sbase = (long long)(j + (long long)ty * nr) * srow * src_ld + tx;
    const long long sstep = (long long)R * nr * srow * src_ld;
    const long long drow = SM == STORE_C ? 1 : 2;
    const long long dbase = ((long long)(j / Ns) * Ns * RR + k + (long long)ty * Ns) * drow * dst_ld;
    const long long dstep = (long long)R * Ns * drow * dst_ld;
    auto issue = [&](double2 (&d)[R], long long c0) {
        const bool ok = c0 + tx < cols;
#pragma unroll
        for (int r2 = 0; r2 < R; ++r2) {
            d[r2] = make_double2(0.0, 0.0);
            if (ok) {
                if (LM == LOAD_C) {
                    d[r2] = __ldcs(static_cast<const double2*>(src) + sbase + r2 * sstep + c0);
                } else {
                    const double* xr = static_cast<const double*>(src) + sbase + r2 * sstep + c0;
                    d[r2] = make_double2(__ldcs(xr), __ldcs(xr + src_ld));
                }
            }
        }
    };
    double2 v[R], nx[R];
    long long c0 = (long long)blockIdx.x * 32;
    issue(nx, c0);
    for (; c0 < cols; c0 += cstride) {
        const long long col = c0 + tx;
        const bool ok = col < cols;
#pragma unroll
        for (int r2 = 0; r2 < R; ++r2) v[r2] = nx[r2];
        if (c0 + cstride < cols) issue(nx, c0 + cstride);
        // stage 1: r1 = ty, r2 = 0..R-1
#pragma unroll
        for (int r2 = 0; r2 < R; ++r2) {
            if (conj_in) v[r2].y = -v[r2].y;
            if (Ns > 1 && (r2 > 0 || ty > 0)) v[r2] = cmul(v[r2], tw[tw0 + r2 * twd]);
        }
        butterfly<R>(v);                                // v[q2] = sum_r2 ... exp(-2 pi i r2 q2 / R)
        __syncthreads();                                // previous tile consumed
#pragma unroll
        for (int q2 = 0; q2 < R; ++q2) ex[ty * R + q2][tx] = (q2 > 0 && ty > 0) ? cmul(v[q2], tw[q2 * twi]) : v[q2];
        __syncthreads();
        // stage 2: q2 = ty, over r1
#pragma unroll
        for (int r1 = 0; r1 < R; ++r1) v[r1] = ex[r1 * R + ty][tx];
        butterfly<R>(v);                                // v[q1]: output R q1 + q2
        if (ok) {
            const double si = conj_out ? -scale : scale;
            if (SM == STORE_C) {
                double2* dp = static_cast<double2*>(dst) + dbase + col;
#pragma unroll
                for (int q1 = 0; q1 < R; ++q1) __stcs(dp + q1 * dstep, make_double2(scale * v[q1].x, si * v[q1].y));
            } else {
                const long long dc = dst_cw ? (col / dst_cw) * dst_cs + col % dst_cw : col;
                double* dp = static_cast<double*>(dst) + dbase + dc;
#pragma unroll
                for (int q1 = 0; q1 < R; ++q1) {
                    dp[q1 * dstep] = scale * v[q1].x;
                    dp[q1 * dstep + dst_ld] = si * v[q1].y;
                }
            }
        }
    }
}

// Fused radix-128 pass: 128 = 16 (r1, the thread row) x 8 (r2, in registers).  blockDim = (32 columns, 16):
// thread (x, r1) transforms inputs r1 + 16 r2 (r2 < 8) with one radix-8 butterfly; the 16 x 8 intermediate goes
// through shared memory with the inner twiddles W_128^{r1 q2}; the length-16 transforms over r1 are then split
// over two threads each (radix-2 decimation in frequency: thread row ty = q2 + 8 h forms x[r] +- x[r+8], times
// W_16^r for h = 1, and runs a radix-8 butterfly), so all 16 warps stay busy and thread row ty stores outputs
// ty + 16 m.  With it 2^14 = 128 x 128 is two sweeps over HBM instead of three (64 x 64 x 4), 2^13 = 128 x 64.
template <int LM, int SM>
__global__ void __launch_bounds__(512, 1) fft_pass128_kernel(const void* __restrict__ src, void* __restrict__ dst, int n,
                                                             int Ns, long long cols, int conj_in, int conj_out,
                                                             const double2* __restrict__ tw, int twN, long long src_ld,
                                                             long long dst_ld, double scale, long long dst_cw,
                                                             long long dst_cs) {
    constexpr int R1 = 16, R2 = 8, RR = 128;
    extern __shared__ __align__(16) double2 ex128[];        // [r1 * 8 + q2][32]
    const int j = blockIdx.y;
    const int k = j % Ns;
    const int nr = n / RR;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tw0 = ty * k * (twN / (Ns * RR)), twd = R1 * k * (twN / (Ns * RR));
    const int twi = ty * (twN / RR);
    const long long cstride = (long long)gridDim.x * 32;
    const long long srow = LM == LOAD_C ? 1 : 2;
    const long long sbase = (long long)(j + (long long)ty * nr) * srow * src_ld + tx;
    const long long sstep = (long long)R1 * nr * srow * src_ld;
    const long long drow = SM == STORE_C ? 1 : 2;
    const long long dbase = ((long long)(j / Ns) * Ns * RR + k + (long long)ty * Ns) * drow * dst_ld;   // output ty + 16 m
    const long long dstep = (long long)R1 * Ns * drow * dst_ld;
    const int q2 = ty & 7, hodd = ty >> 3;
    auto issue = [&](double2 (&d)[R2], long long c0) {
        const bool ok = c0 + tx < cols;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) {
            d[r2] = make_double2(0.0, 0.0);
            if (ok) {
                if (LM == LOAD_C) {
                    d[r2] = __ldcs(static_cast<const double2*>(src) + sbase + r2 * sstep + c0);
                } else {
                    const double* xr = static_cast<const double*>(src) + sbase + r2 * sstep + c0;
                    d[r2] = make_double2(__ldcs(xr), __ldcs(xr + src_ld));
                }
            }
        }
    };
    double2 v[R2], nx[R2];
    long long c0 = (long long)blockIdx.x * 32;
    issue(nx, c0);
    for (; c0 < cols; c0 += cstride) {
        const long long col = c0 + tx;
        const bool ok = col < cols;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) v[r2] = nx[r2];
        if (c0 + cstride < cols) issue(nx, c0 + cstride);
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) {
            if (conj_in) v[r2].y = -v[r2].y;
            if (Ns > 1 && (r2 > 0 || ty > 0)) v[r2] = cmul(v[r2], tw[tw0 + r2 * twd]);
        }
        butterfly<8>(v);                                // v[q2'] over r2
        __syncthreads();                                // previous tile consumed
#pragma unroll
        for (int q = 0; q < R2; ++q)
            ex128[(ty * R2 + q) * 32 + tx] = (q > 0 && ty > 0) ? cmul(v[q], tw[q * twi]) : v[q];
        __syncthreads();
        {
            const double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 lo = ex128[(r * R2 + q2) * 32 + tx], hi = ex128[((r + 8) * R2 + q2) * 32 + tx];
                v[r] = hodd ? csub(lo, hi) : cadd(lo, hi);
            }
            if (hodd) {                                  // times W_16^r
                v[1] = cmul(v[1], make_double2(c1, -s1));
                v[2] = make_double2(h * (v[2].x + v[2].y), h * (v[2].y - v[2].x));
                v[3] = cmul(v[3], make_double2(s1, -c1));
                v[4] = mul_mi(v[4]);
                v[5] = cmul(v[5], make_double2(-s1, -c1));
                v[6] = make_double2(h * (v[6].y - v[6].x), -h * (v[6].x + v[6].y));
                v[7] = cmul(v[7], make_double2(-c1, -s1));
            }
            butterfly<8>(v);                            // v[m]: output (q2 + 8 hodd) + 16 m = ty + 16 m
        }
        if (ok) {
            const double si = conj_out ? -scale : scale;
            if (SM == STORE_C) {
                double2* dp = static_cast<double2*>(dst) + dbase + col;
#pragma unroll
                for (int m = 0; m < 8; ++m) __stcs(dp + m * dstep, make_double2(scale * v[m].x, si * v[m].y));
            } else {
                const long long dc = dst_cw ? (col / dst_cw) * dst_cs + col % dst_cw : col;
                double* dp = static_cast<double*>(dst) + dbase + dc;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    dp[m * dstep] = scale * v[m].x;
                    dp[m * dstep + dst_ld] = si * v[m].y;
                }
            }
        }
    }
}



// cp.async staging for the radix-128 passes: thread t keeps its 8 inputs of a column tile in PRIVATE shared-memory
// slots [stage][r2][t] (no barrier: a thread only reads what it fetched itself), two tiles in flight per CTA while a
// third is transformed, and no prefetch registers.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// fft_pass128_kernel with the staged loads (same arithmetic, same results)
template <int LM, int SM>
__global__ void __launch_bounds__(512, 1) fft_pass128s_kernel(const void* __restrict__ src, void* __restrict__ dst, int n,
                                                              int Ns, long long cols, int conj_in, int conj_out,
                                                              const double2* __restrict__ tw, int twN, long long src_ld,
                                                              long long dst_ld, double scale, long long dst_cw,
                                                              long long dst_cs) {
    constexpr int R1 = 16, R2 = 8, RR = 128;
    extern __shared__ __align__(16) double2 ex128[];        // [r1 * 8 + q2][32], then stages [2][8][512]
    double2* stage = ex128 + RR * 32;
    const int j = blockIdx.y;
    const int k = j % Ns;
    const int nr = n / RR;
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 32 + tx;
    const int tw0 = ty * k * (twN / (Ns * RR)), twd = R1 * k * (twN / (Ns * RR));
    const int twi = ty * (twN / RR);
    const long long cstride = (long long)gridDim.x * 32;
    const long long srow = LM == LOAD_C ? 1 : 2;
    const long long sbase = (long long)(j + (long long)ty * nr) * srow * src_ld + tx;
    const long long sstep = (long long)R1 * nr * srow * src_ld;
    const long long drow = SM == STORE_C ? 1 : 2;
    const long long dbase = ((long long)(j / Ns) * Ns * RR + k + (long long)ty * Ns) * drow * dst_ld;
    const long long dstep = (long long)R1 * Ns * drow * dst_ld;
    const int q2 = ty & 7, hodd = ty >> 3;
    auto issue = [&](int st, long long c0) {
        double2* slot = stage + (size_t)st * R2 * 512 + t;
        if (c0 + tx < cols) {
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) {
                const uint32_t d = smem_u32(slot + r2 * 512);
                if (LM == LOAD_C) {
                    cp_async16(d, static_cast<const double2*>(src) + sbase + r2 * sstep + c0);
                } else {
                    const double* xr = static_cast<const double*>(src) + sbase + r2 * sstep + c0;
                    cp_async8(d, xr);
                    cp_async8(d + 8, xr + src_ld);
                }
            }
        } else {
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) slot[r2 * 512] = make_double2(0.0, 0.0);
        }
    };
    double2 v[R2];
    long long c0 = (long long)blockIdx.x * 32;
    issue(0, c0);
    cp_async_commit();
    if (c0 + cstride < cols) issue(1, c0 + cstride);
    cp_async_commit();
    for (int it = 0; c0 < cols; c0 += cstride, ++it) {
        const long long col = c0 + tx;
        const bool ok = col < cols;
        cp_async_wait<1>();
        {
            const double2* slot = stage + (size_t)(it & 1) * R2 * 512 + t;
#pragma unroll
            for (int r2 = 0; r2 < R2; ++r2) v[r2] = slot[r2 * 512];
        }
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) {
            if (conj_in) v[r2].y = -v[r2].y;
            if (Ns > 1 && (r2 > 0 || ty > 0)) v[r2] = cmul(v[r2], tw[tw0 + r2 * twd]);
        }
        butterfly<8>(v);                                // v[q2'] over r2
        // the slots of this stage have been consumed (their values fed the butterfly): refill with tile it + 2
        if (c0 + 2 * cstride < cols) issue(it & 1, c0 + 2 * cstride);
        cp_async_commit();
        __syncthreads();                                // previous tile consumed
#pragma unroll
        for (int q = 0; q < R2; ++q)
            ex128[(ty * R2 + q) * 32 + tx] = (q > 0 && ty > 0) ? cmul(v[q], tw[q * twi]) : v[q];
        __syncthreads();
        {
            const double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 lo = ex128[(r * R2 + q2) * 32 + tx], hi = ex128[((r + 8) * R2 + q2) * 32 + tx];
                v[r] = hodd ? csub(lo, hi) : cadd(lo, hi);
            }
            if (hodd) {                                  // times W_16^r
                v[1] = cmul(v[1], make_double2(c1, -s1));
                v[2] = make_double2(h * (v[2].x + v[2].y), h * (v[2].y - v[2].x));
                v[3] = cmul(v[3], make_double2(s1, -c1));
                v[4] = mul_mi(v[4]);
                v[5] = cmul(v[5], make_double2(-s1, -c1));
                v[6] = make_double2(h * (v[6].y - v[6].x), -h * (v[6].x + v[6].y));
                v[7] = cmul(v[7], make_double2(-c1, -s1));
            }
            butterfly<8>(v);                            // v[m]: output ty + 16 m
        }
        if (ok) {
            const double si = conj_out ? -scale : scale;
            if (SM == STORE_C) {
                double2* dp = static_cast<double2*>(dst) + dbase + col;
#pragma unroll
                for (int m = 0; m < 8; ++m) __stcs(dp + m * dstep, make_double2(scale * v[m].x, si * v[m].y));
            } else {
                const long long dc = dst_cw ? (col / dst_cw) * dst_cs + col % dst_cw : col;
                double* dp = static_cast<double*>(dst) + dbase + dc;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    dp[m * dstep] = scale * v[m].x;
                    dp[m * dstep + dst_ld] = si * v[m].y;
                }
            }
        }
    }
    cp_async_wait<0>();
}

// Last radix-128 pass of a length-n complex transform (Ns = n / 128 sub-transforms done) fused with the real-FFT
// untangling, |X_k|^2 and the column sums of the power spectrum (spectrum.cu).  The outputs of butterfly index j are
// Z[j + Ns r]; their partners Z[n - k] belong to index Ns - j, so one CTA runs BOTH transforms on a column tile, parks
// the first in shared memory and forms |X_k|^2, |X_{n-k}|^2 from the pair without the spectrum ever reaching HBM
// (one sweep instead of three).  Per-thread sums over the CTA's column tiles, then a warp sum; partial[row][cb0 +
// blockIdx.x], rows 0 .. n.  tw: table of n, tw2: table of 2n.
__global__ void __launch_bounds__(512, 1) fft_pass128_power_kernel(const double2* __restrict__ src, int n, long long cols,
                                                                   const double2* __restrict__ tw,
                                                                   const double2* __restrict__ tw2, int ldb, int nvalid,
                                                                   long long c0g, double* __restrict__ partial,
                                                                   int ncb_total, int cb0) {
    constexpr int R1 = 16, R2 = 8, RR = 128;
    extern __shared__ __align__(16) double2 ex128[];        // exchange [128][32], then the parked transform [128][32]
    double2* res = ex128 + RR * 32;
    __shared__ double2 wk[RR];                              // W_2n^{k} of the B outputs, per output r
    const int Ns = n / RR;
    const int ja = blockIdx.y, jb = (Ns - ja) % Ns;
    const bool self = ja == jb;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int q2 = ty & 7, hodd = ty >> 3;
    const int twi = ty * (n / RR);
    const long long cstride = (long long)gridDim.x * 32;
    const long long sstep = (long long)R1 * Ns * cols;
    if (tx == 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) wk[ty + 16 * m] = tw2[jb + Ns * (ty + 16 * m)];
    }
    auto issue = [&](double2 (&d)[R2], int j, long long c0) {
        const bool ok = c0 + tx < cols;
        const double2* sp = src + (long long)(j + (long long)ty * Ns) * cols + tx + c0;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) d[r2] = ok ? __ldcs(sp + r2 * sstep) : make_double2(0.0, 0.0);
    };
    auto transform = [&](double2 (&v)[R2], int j) {
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2)
            if (r2 > 0 || ty > 0) v[r2] = cmul(v[r2], tw[(ty + R1 * r2) * j]);
        butterfly<8>(v);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < R2; ++q)
            ex128[(ty * R2 + q) * 32 + tx] = (q > 0 && ty > 0) ? cmul(v[q], tw[q * twi]) : v[q];
        __syncthreads();
        const double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 lo = ex128[(r * R2 + q2) * 32 + tx], hi = ex128[((r + 8) * R2 + q2) * 32 + tx];
            v[r] = hodd ? csub(lo, hi) : cadd(lo, hi);
        }
        if (hodd) {
            v[1] = cmul(v[1], make_double2(c1, -s1));
            v[2] = make_double2(h * (v[2].x + v[2].y), h * (v[2].y - v[2].x));
            v[3] = cmul(v[3], make_double2(s1, -c1));
            v[4] = mul_mi(v[4]);
            v[5] = cmul(v[5], make_double2(-s1, -c1));
            v[6] = make_double2(h * (v[6].y - v[6].x), -h * (v[6].x + v[6].y));
            v[7] = cmul(v[7], make_double2(-c1, -s1));
        }
        butterfly<8>(v);                                // v[m]: output ty + 16 m
    };
    // |X_k|^2 of the real transform from Z_k = zk, Z_{n-k} = zm and w = W_2n^k
    auto power = [](double2 zk, double2 zm, double2 w) {
        const double ar = zk.x + zm.x, ai = zk.y - zm.y;
        const double br = zk.x - zm.x, bi = zk.y + zm.y;
        const double wr = br * w.x - bi * w.y, wi = br * w.y + bi * w.x;
        const double xr = 0.5 * (ar + wi), xi = 0.5 * (ai - wr);
        return xr * xr + xi * xi;
    };
    double accb[8], acca[8], accn = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) accb[m] = acca[m] = 0.0;
    double2 v[R2], nx[R2];
    long long c0 = (long long)blockIdx.x * 32;
    issue(nx, ja, c0);
    for (; c0 < cols; c0 += cstride) {
        const long long col = c0 + tx;
        const double mask = (col < cols && ((c0g + col) % ldb) < nvalid) ? 1.0 : 0.0;
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) v[r2] = nx[r2];
        issue(nx, jb, c0);
        transform(v, ja);
#pragma unroll
        for (int m = 0; m < 8; ++m) res[(ty + 16 * m) * 32 + tx] = v[m];    // read after transform B's barriers
#pragma unroll
        for (int r2 = 0; r2 < R2; ++r2) v[r2] = nx[r2];
        if (c0 + cstride < cols) issue(nx, ja, c0 + cstride);
        transform(v, jb);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int rb = ty + 16 * m;
            const int ra = ja == 0 ? ((RR - rb) & (RR - 1)) : RR - 1 - rb;
            const double2 za = res[ra * 32 + tx], zb = v[m];
            const double2 w = wk[rb];
            accb[m] += mask * power(zb, za, w);
            // k_a = n - k_b: W_2n^{n - k} = -conj(W_2n^k)
            if (!self) acca[m] += mask * power(za, zb, make_double2(-w.x, w.y));
            if (ja == 0 && rb == 0) accn += mask * power(zb, zb, make_double2(-1.0, 0.0));   // k = n (Nyquist of 2n)
        }
        // the next tile's transform A starts with a barrier after its first butterfly: res is rewritten only after
        // every thread has passed two more barriers
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int rb = ty + 16 * m;
        const int ra = ja == 0 ? ((RR - rb) & (RR - 1)) : RR - 1 - rb;
        const double sb = warp_sum(accb[m]), sa = warp_sum(acca[m]);
        if (tx == 0) {
            partial[(size_t)(jb + Ns * rb) * ncb_total + cb0 + blockIdx.x] = sb;
            if (!self) partial[(size_t)(ja + Ns * ra) * ncb_total + cb0 + blockIdx.x] = sa;
        }
    }
    if (ja == 0 && ty == 0) {
        const double sn = warp_sum(accn);
        if (tx == 0) partial[(size_t)n * ncb_total + cb0 + blockIdx.x] = sn;
    }
}

struct TwKey {
    int dev, N;
    bool operator<(const TwKey& o) const { return dev < o.dev || (dev == o.dev && N < o.N); }
};
std::map<TwKey, double2*> g_tw;
std::mutex g_tw_mu;

}  // namespace

// Radices of the passes, fused ones (128, 64, 16) first.  Fewest sweeps over HBM wins; among plans with the same
// number of sweeps the 64/16 passes are preferred.
int fft_plan(int n, int* radices) {
    static const int cand[6] = {64, 16, 8, 4, 2, 128};
    int best[32], pick[32];                       // indexed by log2 of the remaining length
    best[0] = 0;
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    for (int l = 1; l <= lg; ++l) {
        best[l] = 1 << 20;
        for (int c = 0; c < 6; ++c) {
            int lr = 0;
            while ((1 << lr) < cand[c]) ++lr;
            if (lr <= l && best[l - lr] + 1 < best[l]) { best[l] = best[l - lr] + 1; pick[l] = cand[c]; }
        }
    }
    int np = 0, l = lg;
    while (l > 0) {
        radices[np++] = pick[l];
        int lr = 0;
        while ((1 << lr) < pick[l]) ++lr;
        l -= lr;
    }
    std::sort(radices, radices + np, [](int a, int b) { return a > b; });
    return np;
}

int fft_num_passes(int n) {
    int r[32];
    return fft_plan(n, r);
}

const double2* twiddle_table(int N) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto it = g_tw.find({dev, N});
    if (it != g_tw.end()) return it->second;
    std::vector<double2> h(N);
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int t = 0; t < N; ++t) {
        const long double a = -two_pi * (long double)t / (long double)N;
        h[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    double2* d = nullptr;
    if (cudaMalloc(&d, sizeof(double2) * N) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), sizeof(double2) * N, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    g_tw[{dev, N}] = d;
    return d;
}

int fft_columns(const void* src, void* dst, double2* bufA, double2* bufB, const FftPlan& p,
                cudaStream_t stream, int* npasses) {
    PYMD_REQUIRE(p.n >= 2 && (p.n & (p.n - 1)) == 0, "fft: length %d is not a power of two >= 2", p.n);
    PYMD_REQUIRE(p.cols > 0, "fft: no columns");
    const double2* tw = twiddle_table(p.n);
    PYMD_REQUIRE(tw, "fft: twiddle table allocation failed");
    int radices[32];
    const int np = fft_plan(p.n, radices);
    PYMD_REQUIRE(np == 1 || (bufA && (np == 2 || bufB)), "fft: work buffers required");
    int Ns = 1;
    const void* cur = src;
    double2* pp[2] = {bufA, bufB};
    bool fuse_power = false;
    if (p.power) {
        p.power->fused = false;
        const char* e = getenv("PYMD_POWER_FUSED");               // 0: separate last pass and reduction (tests)
        if (np >= 2 && radices[0] == 128 && p.sign < 0 && p.store == STORE_C && !(e && atoi(e) == 0)) {
            std::rotate(radices, radices + 1, radices + np);     // one radix-128 pass goes last
            fuse_power = true;
        }
    }
    for (int i = 0; i < np; ++i) {
        const int R = radices[i];
        const bool first = i == 0, last = i == np - 1;
        if (last && fuse_power) {
            FftPowerTail& t = *p.power;
            long long xb = (p.cols + 31) / 32;
            if (xb > 74) xb = 74;
            if (xb > t.cb_per) xb = t.cb_per;
            constexpr int smem = 2 * 128 * 32 * sizeof(double2);
            static bool cfg = false;
            if (!cfg) {
                PYMD_CUDA_CHECK(cudaFuncSetAttribute(fft_pass128_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                cfg = true;
            }
            fft_pass128_power_kernel<<<dim3((unsigned)xb, Ns / 2 + 1), dim3(32, 16), smem, stream>>>(
                static_cast<const double2*>(cur), p.n, p.cols, tw, t.tw2, t.ldb, t.nvalid, t.c0g, t.partial, t.ncb_total, t.cb0);
            PYMD_CUDA_CHECK(cudaGetLastError());
            t.fused = true;
            break;
        }
        void* out = last ? dst : static_cast<void*>(pp[i & 1]);
        const int lm = first ? p.load : LOAD_C, sm = last ? p.store : STORE_C;
        const long long sld = first ? p.src_ld : p.cols, dld = last ? p.dst_ld : p.cols;
        const double sc = last ? p.scale : 1.0;
        const long long cw = last ? p.dst_cw : 0, cs = last ? p.dst_cs : 0;
        // ifft(x) = conj(fft(conj(x))): conjugate on the first load and on the last store
        const int ci = (p.sign > 0 && first) ? 1 : 0, co = (p.sign > 0 && last) ? 1 : 0;
        if (R == 128) {
            long long xb = (p.cols + 31) / 32;
            if (xb > 74) xb = 74;                       // one CTA per SM: two butterfly indices per wave, twice the tiles
                                                        // per CTA for the register prefetch (6.9 vs 7.2 ms for 2^14 x 21120)
            PYMD_REQUIRE(p.n / R <= 65535, "fft: length %d too long for the fused pass", p.n);
            dim3 grid((unsigned)xb, p.n / R);
            constexpr int smem128 = 128 * 32 * sizeof(double2);
#define PYMD_FFT128(LMODE, SMODE)                                                                                     \
    do {                                                                                                              \
        static bool cfg = false;                                                                                      \
        if (!cfg) {                                                                                                   \
            PYMD_CUDA_CHECK(cudaFuncSetAttribute(fft_pass128_kernel<LMODE, SMODE>,                                    \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem128));              \
            cfg = true;                                                                                               \
        }                                                                                                             \
        fft_pass128_kernel<LMODE, SMODE><<<grid, dim3(32, 16), smem128, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, \
                                                                                 p.n, sld, dld, sc, cw, cs);          \
    } while (0)
            constexpr int smem128s = 3 * 128 * 32 * sizeof(double2);
#define PYMD_FFT128S(LMODE, SMODE)                                                                                    \
    do {                                                                                                              \
        static bool cfg = false;                                                                                      \
        if (!cfg) {                                                                                                   \
            PYMD_CUDA_CHECK(cudaFuncSetAttribute(fft_pass128s_kernel<LMODE, SMODE>,                                   \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem128s));             \
            cfg = true;                                                                                               \
        }                                                                                                             \
        fft_pass128s_kernel<LMODE, SMODE><<<grid, dim3(32, 16), smem128s, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, \
                                                                                   p.n, sld, dld, sc, cw, cs);        \
    } while (0)
            const char* es = getenv("PYMD_FFT_STAGED");     // read per call: the tests compare both versions
            const int staged = es ? atoi(es) : 1;
            // 8-byte cp.async needs 8-byte aligned real rows, 16-byte ones 16-byte aligned complex rows: always true
            // for the buffers of this library (256-byte aligned allocations, even leading dimensions)
            const bool al = lm == LOAD_C ? (reinterpret_cast<uintptr_t>(cur) % 16 == 0)
                                         : (reinterpret_cast<uintptr_t>(cur) % 8 == 0);
            // measured (2^14 x 61440): complex loads 4.35 -> 4.85 TB/s per pass with the staging; real row pairs (two
            // 8-byte copies per element) gain nothing, so they keep the register prefetch unless PYMD_FFT_STAGED=2
            if (al && (staged >= 2 || (staged == 1 && lm == LOAD_C))) {
                if (lm == LOAD_C && sm == STORE_C) PYMD_FFT128S(LOAD_C, STORE_C);
                else if (lm == LOAD_C) PYMD_FFT128S(LOAD_C, STORE_REAL_ROWS);
                else if (sm == STORE_C) PYMD_FFT128S(LOAD_REAL_PAIRS, STORE_C);
                else PYMD_FFT128S(LOAD_REAL_PAIRS, STORE_REAL_ROWS);
            } else {
                if (lm == LOAD_C && sm == STORE_C) PYMD_FFT128(LOAD_C, STORE_C);
                else if (lm == LOAD_C) PYMD_FFT128(LOAD_C, STORE_REAL_ROWS);
                else if (sm == STORE_C) PYMD_FFT128(LOAD_REAL_PAIRS, STORE_C);
                else PYMD_FFT128(LOAD_REAL_PAIRS, STORE_REAL_ROWS);
            }
#undef PYMD_FFT128S
#undef PYMD_FFT128
        } else if (R >= 16) {
            long long xb = (p.cols + 31) / 32;
            if (xb > 148) xb = 148;                     // half a wave of column tiles per butterfly index
            PYMD_REQUIRE(p.n / R <= 65535, "fft: length %d too long for the fused pass", p.n);
            dim3 grid((unsigned)xb, p.n / R);
#define PYMD_FFT2(RAD, LMODE, SMODE)                                                                       \
    fft_pass2_kernel<RAD, LMODE, SMODE><<<grid, dim3(32, RAD), 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, \
                                                                            p.n, sld, dld, sc, cw, cs)
            if (R == 64) {
                if (lm == LOAD_C && sm == STORE_C) PYMD_FFT2(8, LOAD_C, STORE_C);
                else if (lm == LOAD_C) PYMD_FFT2(8, LOAD_C, STORE_REAL_ROWS);
                else if (sm == STORE_C) PYMD_FFT2(8, LOAD_REAL_PAIRS, STORE_C);
                else PYMD_FFT2(8, LOAD_REAL_PAIRS, STORE_REAL_ROWS);
            } else {
                if (lm == LOAD_C && sm == STORE_C) PYMD_FFT2(4, LOAD_C, STORE_C);
                else if (lm == LOAD_C) PYMD_FFT2(4, LOAD_C, STORE_REAL_ROWS);
                else if (sm == STORE_C) PYMD_FFT2(4, LOAD_REAL_PAIRS, STORE_C);
                else PYMD_FFT2(4, LOAD_REAL_PAIRS, STORE_REAL_ROWS);
            }
#undef PYMD_FFT2
        } else {
            long long xb = (p.cols + 127) / 128;
            if (xb > 592) xb = 592;
            PYMD_REQUIRE(p.n / R <= 65535, "fft: length %d too long for a radix-%d pass", p.n, R);
            dim3 grid((unsigned)xb, p.n / R);
            switch (R) {
                case 8:
                    fft_pass_kernel<8><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc, cw, cs);
                    break;
                case 4:
                    fft_pass_kernel<4><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc, cw, cs);
                    break;
                default:
                    fft_pass_kernel<2><<<grid, 128, 0, stream>>>(cur, out, p.n, Ns, p.cols, ci, co, tw, p.n, lm, sm, sld, dld, sc, cw, cs);
            }
        }
        PYMD_CUDA_CHECK(cudaGetLastError());
        cur = out;
        Ns *= R;
    }
    if (npasses) *npasses = np;
    return PYMD_OK;
}

}  // namespace pymd

// See gemm.cuh.  Producer warps (TMA bulk row copies) + DMMA warps per CTA, mbarrier pipeline.
#include "gemm.cuh"

namespace pymd {

namespace {

constexpr int BK = 32;         // K rows per pipeline stage
constexpr int LDA = BK + 4;    // padded so the A-fragment LDS is bank-conflict free (ld % 16 == 4)

// CTA tile BM x BN, compute warps arranged WM (M) x WN (N), NPROD producer warps.
// cp.async.bulk is a uniform-datapath instruction: a warp whose lanes copy different rows issues
// them one lane at a time (~65 clk each, measured), so the copies of a stage are spread over
// NPROD producer warps to keep the issue time below the DMMA time of the stage.
template <int BM, int BN, int STAGES>
struct __align__(16) Smem {
    double A[STAGES][BM][LDA];
    double B[STAGES][BK][BN + 4];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};

template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
__device__ __forceinline__ void gemm_body(const GemmArgs& g, unsigned char* smem_raw);

template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
__global__ void __launch_bounds__((WM * WN + NPROD) * 32) gemm_kernel(GemmArgs g) {
    g.a_base += (long long)blockIdx.z * g.a_bs;
    g.x_base += (long long)blockIdx.z * g.x_bs;
    g.y_base += (long long)blockIdx.z * g.y_bs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    gemm_body<BM, BN, STAGES, WM, WN, NPROD>(g, smem_raw);
}

// several independent products of similar shape in ONE launch (blockIdx.z selects the problem): the small per-step
// products of the per-step pipeline fill the SMs together instead of one after the other
template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
__global__ void __launch_bounds__((WM * WN + NPROD) * 32) gemm_multi_kernel(const GemmMulti m) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GemmArgs& g = m.a[blockIdx.z];
    if ((int)blockIdx.y * BM >= g.M || (int)blockIdx.x * BN >= g.N) return;     // this problem has fewer tiles
    gemm_body<BM, BN, STAGES, WM, WN, NPROD>(g, smem_raw);
}

template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
__device__ __forceinline__ void gemm_body(const GemmArgs& g, unsigned char* smem_raw) {
    Smem<BM, BN, STAGES>& s = *reinterpret_cast<Smem<BM, BN, STAGES>*>(smem_raw);
    constexpr int LDB = BN + 4;
    constexpr int NCOMPUTE = WM * WN;
    constexpr int WTM = BM / WM, WTN = BN / WN, MT = WTM / 8, NT = WTN / 8;
    static_assert(MT >= 1 && NT >= 1 && WTM % 8 == 0 && WTN % 8 == 0, "bad warp tile");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&s.full[i], NPROD);
            mbar_init(&s.empty[i], NCOMPUTE);
        }
        mbar_fence_init();
    }
    __syncthreads();

    int ntiles = 0;
    for (int sg = 0; sg < g.nseg; ++sg) ntiles += (g.seg_hi[sg] - g.seg_lo[sg] + BK - 1) / BK;

    if (warp >= NCOMPUTE) {
        // ------------------------------ producers ------------------------------
        const int pw = warp - NCOMPUTE;
        const int ncols = min(BN, g.N - n0);
        const uint32_t xbytes = static_cast<uint32_t>(ncols) * 8u;
        int it = 0;
        for (int sg = 0; sg < g.nseg; ++sg) {
            for (int k0 = g.seg_lo[sg]; k0 < g.seg_hi[sg]; k0 += BK, ++it) {
                const int st = it % STAGES;
                if (it >= STAGES) mbar_wait(&s.empty[st], ((it / STAGES) - 1) & 1);
                const int kvalid = min(BK, g.seg_hi[sg] - k0);
                uint32_t mybytes = 0;
                // rows of the stage: [0, BK) are X rows, [BK, BK + BM) are A rows; dealt to the
                // producer lanes round-robin
                for (int e = pw * 32 + lane; e < BK + BM; e += NPROD * 32) {
                    if (e < BK) {
                        if (e < kvalid) {
                            const int k = k0 + e;
                            const long long row = g.x_idx ? g.x_idx[k] : k;
                            bulk_g2s(&s.B[st][e][0], g.x_base + row * g.x_ld + n0, xbytes, &s.full[st]);
                            mybytes += xbytes;
                        } else {
                            for (int c = 0; c < BN; ++c) s.B[st][e][c] = 0.0;   // rows past the K range are zero
                        }
                    } else {
                        const int r = e - BK, m = m0 + r;
                        if (m < g.M) {
                            const double* src = g.a_base + (long long)(m % g.a_mod) * g.a_ld +
                                                (long long)(m / g.a_mod) * g.a_ls + g.seg_aoff[sg] + k0;
                            bulk_g2s(&s.A[st][r][0], src, BK * 8, &s.full[st]);
                            mybytes += BK * 8;
                        }
                    }
                }
                uint32_t total = mybytes;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                __syncwarp();   // zero-fill stores ordered before the arrive below
                if (lane == 0) mbar_arrive_expect_tx(&s.full[st], total);
            }
        }
        return;
    }

    // ------------------------------ DMMA warps ------------------------------
    const int wm = warp / WN, wn = warp % WN;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int it = 0; it < ntiles; ++it) {
        const int st = it % STAGES;
        mbar_wait(&s.full[st], (it / STAGES) & 1);
        const double* As = &s.A[st][wm * WTM + lr][lc];
        const double* Bs = &s.B[st][lc][wn * WTN + lr];
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = As[i * 8 * LDA + kk];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = Bs[kk * LDB + j * 8];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[st]);
    }

    // ------------------------------ epilogue ------------------------------
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int m = m0 + wm * WTM + i * 8 + lr;
        if (m >= g.M) continue;
        const long long row = g.y_idx ? g.y_idx[m] : m;
        double* yrow = g.y_base + row * g.y_ld;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int n = n0 + wn * WTN + j * 8 + 2 * lc;
            if (n >= g.N) continue;
            double2 v = make_double2(g.alpha * acc[i][j][0], g.alpha * acc[i][j][1]);
            double2* dst = reinterpret_cast<double2*>(yrow + n);
            if (g.accumulate) {
                double2 o = *dst;
                v.x += o.x;
                v.y += o.y;
            }

            *dst = v;
        }
    }
}

template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
int launch_multi_t(const GemmMulti& m, cudaStream_t stream) {
    static bool configured = false;
    const int smem = static_cast<int>(sizeof(Smem<BM, BN, STAGES>));
    if (!configured) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(gemm_multi_kernel<BM, BN, STAGES, WM, WN, NPROD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int gx = 0, gy = 0;
    for (int i = 0; i < m.n; ++i) {
        gx = max(gx, (m.a[i].N + BN - 1) / BN);
        gy = max(gy, (m.a[i].M + BM - 1) / BM);
    }
    gemm_multi_kernel<BM, BN, STAGES, WM, WN, NPROD><<<dim3(gx, gy, m.n), (WM * WN + NPROD) * 32, smem, stream>>>(m);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

template <int BM, int BN, int STAGES, int WM, int WN, int NPROD>
int launch_t(const GemmArgs& a, cudaStream_t stream) {
    static bool configured = false;
    const int smem = static_cast<int>(sizeof(Smem<BM, BN, STAGES>));
    if (!configured) {
        PYMD_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BM, BN, STAGES, WM, WN, NPROD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.nbatch > 1 ? a.nbatch : 1);
    gemm_kernel<BM, BN, STAGES, WM, WN, NPROD><<<grid, (WM * WN + NPROD) * 32, smem, stream>>>(a);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

}  // namespace

static int gemm_validate(const GemmArgs& a) {
    PYMD_REQUIRE(a.M > 0 && a.N > 0 && (a.N % 2) == 0, "gemm: bad shape M=%d N=%d", a.M, a.N);
    PYMD_REQUIRE(a.nseg >= 0 && a.nseg <= 2, "gemm: nseg=%d", a.nseg);
    for (int sg = 0; sg < a.nseg; ++sg)
        PYMD_REQUIRE(((a.seg_aoff[sg] + a.seg_lo[sg]) % 2) == 0 && (a.a_ld % 2) == 0 && (a.a_ls % 2) == 0,
                     "gemm: A rows must be 16-byte aligned");
    PYMD_REQUIRE((a.x_ld % 2) == 0 && (a.y_ld % 2) == 0, "gemm: X/Y leading dims must be even");
    return PYMD_OK;
}

// Tile choice by estimated efficiency = SM fill x per-tile DMMA efficiency (measured: a 64x64 CTA
// tile with 32x32 warp tiles needs 1 LDS per 2 DMMA and sustains ~0.9 of the pipe; 32x32 ~0.45).
// Returns 64, 32 or 16 (the BM of the configuration); `nprob` products of this shape share the launch.
static int gemm_tile(const GemmArgs& a, long long nprob) {
    auto fill = [](long long ctas) {          // CTAs on one SM share its DMMA pipe
        const long long per_sm = (ctas + 147) / 148;
        return (double)ctas / (double)(per_sm * 148);
    };
    const long long c64 = nprob * ((a.N + 63) / 64) * ((a.M + 63) / 64);
    const long long c32 = nprob * ((a.N + 31) / 32) * ((a.M + 31) / 32);
    const double pad64 = (double)a.M / (double)(((a.M + 63) / 64) * 64);
    const double pad32 = (double)a.M / (double)(((a.M + 31) / 32) * 32);
    const double e64 = fill(c64) * 0.9 * pad64, e32 = fill(c32) * 0.45 * pad32;
    if (a.M > 16 && e64 >= e32) return 64;
    if (a.M > 16) return 32;
    return 16;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    int rc = gemm_validate(a);
    if (rc) return rc;
    const long long nbz = a.nbatch > 1 ? a.nbatch : 1;
    PYMD_REQUIRE(nbz <= 65535 && (a.a_bs % 2) == 0 && (a.x_bs % 2) == 0 && (a.y_bs % 2) == 0, "gemm: bad batch nbatch=%d", a.nbatch);
    switch (gemm_tile(a, nbz)) {
        case 64: return launch_t<64, 64, 5, 2, 4, 4>(a, stream);   // one CTA per SM: 8 DMMA warps + 4 producer warps
        case 32: return launch_t<32, 32, 3, 2, 2, 1>(a, stream);
        default: return launch_t<16, 32, 3, 2, 2, 1>(a, stream);
    }
}

int launch_gemm_multi(const GemmMulti& m, cudaStream_t stream) {
    PYMD_REQUIRE(m.n >= 1 && m.n <= GemmMulti::kMax, "gemm_multi: n=%d", m.n);
    if (m.n == 1) return launch_gemm(m.a[0], stream);
    int big = 0;
    for (int i = 0; i < m.n; ++i) {
        int rc = gemm_validate(m.a[i]);
        if (rc) return rc;
        PYMD_REQUIRE(m.a[i].nbatch <= 1, "gemm_multi: batched problems are launched alone");
        if ((long long)m.a[i].M * m.a[i].N > (long long)m.a[big].M * m.a[big].N) big = i;
    }
    switch (gemm_tile(m.a[big], m.n)) {
        case 64: return launch_multi_t<64, 64, 5, 2, 4, 4>(m, stream);
        case 32: return launch_multi_t<32, 32, 3, 2, 2, 1>(m, stream);
        default: return launch_multi_t<16, 32, 3, 2, 2, 1>(m, stream);
    }
}

}  // namespace pymd

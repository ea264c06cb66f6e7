// Power spectra of saved trajectories, summed over dofs and over the ensemble
// (reference spectrum.py:8-45 powerspec / powerspec2, md.GetPower md.py:303-313).
//
// Real trajectories are transformed two time-rows at a time: z_m = x_{2m} + i x_{2m+1},
// one half-length complex FFT, then X_k = E_k + W_N^k O_k recovered from Z_k, Z_{N/2-k}.
// |X_k|^2 is reduced over columns in a fixed order (deterministic, no atomics).
#include "../../include/pymd_b200.h"
#include "fft.cuh"

namespace pymd {

int fft_num_passes(int n);

constexpr int PB = 256;

// One thread per (frequency pair k / n-k, column): Z_k and Z_{n-k} are read once and give both
// |X_k|^2 and |X_{n-k}|^2 (X_{n-k} uses the same two values with the roles exchanged).
__global__ void __launch_bounds__(PB) power_partial_kernel(const double2* __restrict__ Z, int n, long long cols,
                                                           int ldb, int nvalid, long long c0,
                                                           const double2* __restrict__ twN,
                                                           double* __restrict__ partial, int ncb_total, int cb0) {
    __shared__ double red[2][PB / 32];
    const int k = blockIdx.x;                       // 0 .. n/2; partner m = n - k
    const int m = n - k;
    const long long col = blockIdx.y * (long long)PB + threadIdx.x;
    double vk = 0.0, vm = 0.0;
    if (col < cols && ((c0 + col) % ldb) < nvalid) {
        const double2 zk = Z[(size_t)(k % n) * cols + col];
        const double2 zm = Z[(size_t)(m % n) * cols + col];
        {
            const double ar = zk.x + zm.x, ai = zk.y - zm.y;      // Zk + conj(Zm)
            const double br = zk.x - zm.x, bi = zk.y + zm.y;      // Zk - conj(Zm)
            const double2 w = twN[k];
            const double wr = br * w.x - bi * w.y, wi = br * w.y + bi * w.x;   // W^k (Zk - conj Zm)
            // X = (A - i*Wb)/2 ;  -i*(wr + i wi) = wi - i wr
            const double xr = 0.5 * (ar + wi), xi = 0.5 * (ai - wr);
            vk = xr * xr + xi * xi;
        }
        {
            const double ar = zm.x + zk.x, ai = zm.y - zk.y;      // Zm + conj(Zk)
            const double br = zm.x - zk.x, bi = zm.y + zk.y;      // Zm - conj(Zk)
            const double2 w = twN[m];
            const double wr = br * w.x - bi * w.y, wi = br * w.y + bi * w.x;
            const double xr = 0.5 * (ar + wi), xi = 0.5 * (ai - wr);
            vm = xr * xr + xi * xi;
        }
    }
    vk = warp_sum(vk);
    vm = warp_sum(vm);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = vk; red[1][threadIdx.x >> 5] = vm; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < PB / 32; ++i) s += red[threadIdx.x][i];
        const int row = threadIdx.x == 0 ? k : m;
        if (threadIdx.x == 0 || m != k) partial[(size_t)row * ncb_total + cb0 + blockIdx.y] = s;
    }
}

__global__ void power_final_kernel(const double* __restrict__ partial, int ncb_total, int N, double dt, int wsq,
                                   double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int k = i <= N / 2 ? i : N - i;
    double s = 0.0;
    for (int c = 0; c < ncb_total; ++c) s += partial[(size_t)k * ncb_total + c];
    const double dw = 2.0 * 3.14159265358979323846 / dt / N;
    // |Fourier1D|^2 = dt^2 |DFT|^2  (myfft.py:29-31), then /dt/nmd (spectrum.py:24,44)
    double val = s * dt * dt / dt / N;
    if (wsq) val = (dw * i) * (dw * i) * val;
    out[i] = val;
}

}  // namespace pymd

using namespace pymd;

// column sums of a [rows][ldb] array in two deterministic stages: CS_CHUNKS row chunks, then their fixed-order sum
namespace pymd {
namespace {
constexpr int CS_CHUNKS = 64;
__global__ void colsum_partial_kernel(const double* a, long long rows, int ldb, double* partial) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ldb) return;
    const long long per = (rows + CS_CHUNKS - 1) / CS_CHUNKS, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
        s0 += a[r * ldb + n]; s1 += a[(r + 1) * ldb + n]; s2 += a[(r + 2) * ldb + n]; s3 += a[(r + 3) * ldb + n];
    }
    for (; r < r1; ++r) s0 += a[r * ldb + n];
    partial[(size_t)blockIdx.y * ldb + n] = (s0 + s1) + (s2 + s3);
}
__global__ void colsum_final_kernel(const double* partial, int ldb, double scale, double* out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ldb) return;
    double s = 0.0;
    for (int c = 0; c < CS_CHUNKS; ++c) s += partial[(size_t)c * ldb + n];
    out[n] = scale * s;
}
}  // namespace
}  // namespace pymd

extern "C" {

static long long ps_chunk(long long total, long long chunk) {
    if (chunk <= 0 || chunk > total) chunk = total;
    return chunk;
}

size_t pymd_powerspec_work_bytes(int nmd, int nd, int ldb, long long chunk_cols) {
    const long long total = (long long)nd * ldb, chunk = ps_chunk(total, chunk_cols);
    const long long ncb_total = ((total + chunk - 1) / chunk) * ((chunk + PB - 1) / PB);
    return (size_t)(nmd / 2 + 1) * ncb_total * sizeof(double) + 2 * (size_t)(nmd / 2) * chunk * sizeof(double2) + 256;
}

static int powerspec_impl(const double* traj, int nmd, int nd, int ldb, int nvalid, double dt, int wsq, double* out,
                          void* work, long long chunk_cols, void* stream, long long row_ld);

int pymd_powerspec(const double* traj, int nmd, int nd, int ldb, int nvalid, double dt, int wsq, double* out,
                   void* work, long long chunk_cols, void* stream) {
    return powerspec_impl(traj, nmd, nd, ldb, nvalid, dt, wsq, out, work, chunk_cols, stream, 0);
}

int pymd_powerspec_ld(const double* traj, long long row_stride, int nmd, int nd, int ldb, int nvalid, double dt, int wsq,
                      double* out, void* work, long long chunk_cols, void* stream) {
    PYMD_REQUIRE(row_stride >= (long long)nd * ldb && row_stride % 2 == 0, "powerspec: row stride %lld", row_stride);
    return powerspec_impl(traj, nmd, nd, ldb, nvalid, dt, wsq, out, work, chunk_cols, stream, row_stride);
}

static int powerspec_impl(const double* traj, int nmd, int nd, int ldb, int nvalid, double dt, int wsq, double* out,
                          void* work, long long chunk_cols, void* stream, long long row_ld) {
    PYMD_REQUIRE(traj && out && work, "powerspec: null argument");
    PYMD_REQUIRE(nmd >= 4 && (nmd & (nmd - 1)) == 0, "powerspec: nmd=%d must be a power of two >= 4", nmd);
    PYMD_REQUIRE(nvalid > 0 && nvalid <= ldb, "powerspec: nvalid=%d ldb=%d", nvalid, ldb);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int half = nmd / 2;
    const long long total = (long long)nd * ldb, chunk = ps_chunk(total, chunk_cols);
    PYMD_REQUIRE(chunk % 2 == 0, "powerspec: chunk_cols must be even");
    const int nchunks = (int)((total + chunk - 1) / chunk);
    const int cb_per = (int)((chunk + PB - 1) / PB), ncb_total = nchunks * cb_per;
    double* partial = static_cast<double*>(work);
    size_t off = ((size_t)(half + 1) * ncb_total * sizeof(double) + 255) / 256 * 256;
    double2* bufA = reinterpret_cast<double2*>(static_cast<char*>(work) + off);
    double2* bufB = bufA + (size_t)half * chunk;
    const double2* twN = twiddle_table(nmd);
    PYMD_REQUIRE(twN, "powerspec: twiddle table allocation failed");
    PYMD_CUDA_CHECK(cudaMemsetAsync(partial, 0, (size_t)(half + 1) * ncb_total * sizeof(double), st));
    const int np = fft_num_passes(half);
    for (int ch = 0; ch < nchunks; ++ch) {
        const long long c0 = ch * chunk, cols = (c0 + chunk <= total) ? chunk : total - c0;
        FftPlan p{};
        p.n = half; p.cols = cols; p.sign = -1;
        p.load = LOAD_REAL_PAIRS; p.store = STORE_C;
        p.src_ld = row_ld ? row_ld : total; p.dst_ld = cols; p.scale = 1.0;
        // passes write bufA, bufB, bufA, ...; the last one reads buf[(np-2)&1] and must write the other
        double2* dst = (np % 2) ? bufA : bufB;
        double2* a = bufA;
        double2* b = bufB;
        // the last radix-128 pass forms |X_k|^2 and the column sums itself when the plan has one (fft.cu)
        FftPowerTail tail{ldb, nvalid, c0, twN, partial, ncb_total, ch * cb_per, cb_per, false};
        p.power = &tail;
        int rc = fft_columns(traj + c0, dst, a, b, p, st, nullptr);
        if (rc) return rc;
        if (!tail.fused) {
            dim3 grid(half / 2 + 1, (unsigned)((cols + PB - 1) / PB));
            power_partial_kernel<<<grid, PB, 0, st>>>(dst, half, cols, ldb, nvalid, c0, twN, partial, ncb_total, ch * cb_per);
            PYMD_CUDA_CHECK(cudaGetLastError());
        }
    }
    power_final_kernel<<<(nmd + 255) / 256, 256, 0, st>>>(partial, ncb_total, nmd, dt, wsq, out);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}

int pymd_column_sums(const double* a_dev, long long rows, int ldb, double scale, double* out_dev, void* work_dev, void* stream) {
    PYMD_REQUIRE(a_dev && out_dev && work_dev && rows > 0 && ldb > 0, "column_sums: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(work_dev);
    colsum_partial_kernel<<<dim3((ldb + 127) / 128, CS_CHUNKS), 128, 0, st>>>(a_dev, rows, ldb, partial);
    colsum_final_kernel<<<(ldb + 127) / 128, 128, 0, st>>>(partial, ldb, scale, out_dev);
    PYMD_CUDA_CHECK(cudaGetLastError());
    return PYMD_OK;
}
size_t pymd_column_sums_work_bytes(int ldb) { return (size_t)CS_CHUNKS * ldb * sizeof(double); }

size_t pymd_fft_work_bytes(int n, long long cols) { return 2 * (size_t)n * cols * sizeof(double2); }

int pymd_fft_num_passes(int n) {
    if (n < 2 || (n & (n - 1)) != 0) return -1;
    return fft_num_passes(n);
}

int pymd_fft_columns(double* data, int n, long long cols, int sign, void* work, void* stream) {
    PYMD_REQUIRE(data && work, "fft_columns: null argument");
    PYMD_REQUIRE(sign == 1 || sign == -1, "fft_columns: sign must be +1 or -1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double2* A = static_cast<double2*>(work);
    double2* B = A + (size_t)n * cols;
    FftPlan p{};
    p.n = n; p.cols = cols; p.sign = sign;
    p.load = LOAD_C; p.store = STORE_C; p.src_ld = cols; p.dst_ld = cols; p.scale = 1.0;
    // out of place into A/B, then copy back (keeps the pass kernel free of in-place hazards)
    const int np = fft_num_passes(n);
    double2* dst = (np % 2) ? A : B;
    double2* a = A;
    double2* b = B;
    int rc = fft_columns(data, dst, a, b, p, st, nullptr);
    if (rc) return rc;
    PYMD_CUDA_CHECK(cudaMemcpyAsync(data, dst, (size_t)n * cols * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return PYMD_OK;
}

}  // extern "C"

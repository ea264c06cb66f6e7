// Quantum initial condition for the whole ensemble (reference md.initialise, md.py:253-290):
// every eigenmode i of the dynamical matrix gets a random phase 2 pi r_i and amplitude
// am_i = sqrt((n_B(hw_i, T) + 1/2) * 2 / hw_i) (0 for hw_i < 0.01); the constraint projection
// (md.ApplyConstraint, md.py:739-762) is applied after every mode, as the reference does.
#include "engine.cuh"

namespace pymd {

constexpr int IX = 32, IY = 8;

__global__ void __launch_bounds__(IX* IY) init_state_kernel(double* __restrict__ p, double* __restrict__ q,
                                                            const double* __restrict__ U,
                                                            const double* __restrict__ hw,
                                                            const double* __restrict__ am,
                                                            const double* __restrict__ r,
                                                            const double* __restrict__ con, int ncon, int nd,
                                                            int ldb) {
    __shared__ double red[IY][IX];
    const int n = blockIdx.x * IX + threadIdx.x;
    for (int i = threadIdx.y; i < nd; i += IY) {
        p[(size_t)i * ldb + n] = 0.0;
        q[(size_t)i * ldb + n] = 0.0;
    }
    for (int m = 0; m < nd; ++m) {
        const double ang = 2.0 * 3.14159265358979323846 * r[(size_t)m * ldb + n];
        const double c = cos(ang), s = sin(ang);
        const double a = am[m], w = hw[m];
        for (int i = threadIdx.y; i < nd; i += IY) {
            const size_t o = (size_t)i * ldb + n;
            const double u = U[(size_t)i * nd + m];
            q[o] = q[o] + u * a * c;
            p[o] = p[o] - w * u * a * s;
        }
        if (ncon == 0) continue;
        __syncthreads();
        double dq[PYMD_MAX_CONSTRAINTS], dp[PYMD_MAX_CONSTRAINTS];
        for (int cc = 0; cc < ncon; ++cc) {
            double sq = 0.0, sp = 0.0;
            for (int i = threadIdx.y; i < nd; i += IY) {
                const double cv = con[(size_t)cc * nd + i];
                sq += q[(size_t)i * ldb + n] * cv;
                sp += p[(size_t)i * ldb + n] * cv;
            }
            red[threadIdx.y][threadIdx.x] = sq;
            __syncthreads();
            sq = 0.0;
            for (int y = 0; y < IY; ++y) sq += red[y][threadIdx.x];
            __syncthreads();
            red[threadIdx.y][threadIdx.x] = sp;
            __syncthreads();
            sp = 0.0;
            for (int y = 0; y < IY; ++y) sp += red[y][threadIdx.x];
            __syncthreads();
            dq[cc] = sq;
            dp[cc] = sp;
        }
        for (int i = threadIdx.y; i < nd; i += IY) {
            const size_t o = (size_t)i * ldb + n;
            double qq = q[o], pp = p[o];
            for (int cc = 0; cc < ncon; ++cc) {
                const double cv = con[(size_t)cc * nd + i];
                qq -= dq[cc] * cv;
                pp -= dp[cc] * cv;
            }
            q[o] = qq;
            p[o] = pp;
        }
        __syncthreads();
    }
}

}  // namespace pymd

using namespace pymd;

extern "C" int pymd_init_state(pymd_ctx* c, const double* U, const double* hw, const double* am,
                               const double* r, void* stream) {
    PYMD_REQUIRE(c && U && hw && am && r, "init_state: null argument");
    PYMD_REQUIRE(c->sd.p && c->sd.q, "init_state: state not bound");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nd = c->nd;
    double *dU = nullptr, *dv = nullptr, *dcon = nullptr;
    PYMD_CUDA_CHECK(cudaMalloc(&dU, sizeof(double) * nd * nd));
    PYMD_CUDA_CHECK(cudaMalloc(&dv, sizeof(double) * 2 * nd));
    PYMD_CUDA_CHECK(cudaMemcpyAsync(dU, U, sizeof(double) * nd * nd, cudaMemcpyHostToDevice, st));
    PYMD_CUDA_CHECK(cudaMemcpyAsync(dv, hw, sizeof(double) * nd, cudaMemcpyHostToDevice, st));
    PYMD_CUDA_CHECK(cudaMemcpyAsync(dv + nd, am, sizeof(double) * nd, cudaMemcpyHostToDevice, st));
    if (c->ncon > 0) {
        PYMD_CUDA_CHECK(cudaMalloc(&dcon, sizeof(double) * c->ncon * nd));
        PYMD_CUDA_CHECK(cudaMemcpyAsync(dcon, c->con.data(), sizeof(double) * c->ncon * nd, cudaMemcpyHostToDevice, st));
    }
    init_state_kernel<<<c->ldb / IX, dim3(IX, IY), 0, st>>>(c->sd.p, c->sd.q, dU, dv, dv + nd, r, dcon, c->ncon, nd, c->ldb);
    c->launches++;
    PYMD_CUDA_CHECK(cudaGetLastError());
    PYMD_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dU); cudaFree(dv); cudaFree(dcon);
    c->fpot_valid = false;
    return PYMD_OK;
}

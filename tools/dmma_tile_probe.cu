// DMMA.8x8x4 throughput from SHARED-MEMORY operands for several warp-tile shapes (MT x NT m8n8 tiles)
// and warps per SM: finds the inner-loop shape the GEMM / fused kernels should use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_tile_probe tools/dmma_tile_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// each warp owns an (8*MT) x (8*NT) tile; A tile [8*MT][36], B tile [32][8*NT+4] in smem, K loop of 32
template <int MT, int NT, int REGA>
__global__ void probe(double* out, int iters) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lr = lane >> 2, lc = lane & 3;
    constexpr int LDA = 36, LDB = 8 * NT + 4;
    double* As = sm + warp * (8 * MT * LDA + 32 * LDB);
    double* Bs = As + 8 * MT * LDA;
    for (int i = lane; i < 8 * MT * LDA + 32 * LDB; i += 32) As[i] = 1e-3 * (i % 7);
    __syncwarp();
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double areg[MT][8];
    if (REGA) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) areg[i][k] = As[(i * 8 + lr) * LDA + 4 * k + lc];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = REGA ? areg[i][k] : As[(i * 8 + lr) * LDA + 4 * k + lc];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = Bs[(4 * k + lc) * LDB + 8 * j + lr];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MT, int NT, int REGA>
void run(double* out, int sms, int warps) {
    const int iters = 512;
    const size_t smem = (size_t)warps * (8 * MT * 36 + 32 * (8 * NT + 4)) * 8;
    if (smem > 220 * 1024) return;
    cudaFuncSetAttribute(probe<MT, NT, REGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<MT, NT, REGA><<<sms, warps * 32, smem>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MT, NT, REGA><<<sms, warps * 32, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 256 * MT * NT * 8 * iters * (double)warps * sms;
    printf("{\"MT\": %d, \"NT\": %d, \"A_in_regs\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f, \"err\": \"%s\"}\n", MT, NT, REGA,
           warps, flops / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    for (int warps : {4, 8, 16}) {
        run<4, 4, 0>(out, sms, warps);
        run<2, 4, 0>(out, sms, warps);
        run<4, 2, 0>(out, sms, warps);
        run<2, 2, 0>(out, sms, warps);
        run<1, 4, 0>(out, sms, warps);
        run<1, 2, 0>(out, sms, warps);
        run<1, 4, 1>(out, sms, warps);
        run<1, 2, 1>(out, sms, warps);
        run<2, 4, 1>(out, sms, warps);
    }
    return 0;
}

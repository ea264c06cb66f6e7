// Measures the FP64 throughput ceilings of the device: DMMA.8x8x4 (mma.sync f64) and scalar
// DFMA, as a function of warps per SM and independent accumulator chains.  The roofline for the
// tensor-bound kernels of pymd_b200 uses these numbers (MEASURED_PEAKS.json has no FP64 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_kernel(double* out, int iters) {
    double c[ILP][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
    double c[ILP];
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n \"dmma\": [", p.name, sms, p.clockRate);
    const int iters = 4096;
    bool first = true;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int ilp : {1, 2, 4, 8}) {
            float ms = 0;
            auto run = [&](auto k) { ms = time_ms([&] { k<<<sms, warps * 32>>>(out, iters); }); };
            if (ilp == 1) run(dmma_kernel<1>); else if (ilp == 2) run(dmma_kernel<2>);
            else if (ilp == 4) run(dmma_kernel<4>); else run(dmma_kernel<8>);
            const double flops = 2.0 * 256 * (double)iters * ilp * warps * sms;
            printf("%s\n  {\"warps_per_sm\": %d, \"ilp\": %d, \"tflops\": %.2f}", first ? "" : ",", warps, ilp,
                   flops / ms * 1e-9);
            first = false;
        }
    }
    printf("],\n \"dfma\": [");
    first = true;
    for (int warps : {4, 8, 16, 32}) {
        for (int ilp : {1, 4, 8}) {
            float ms = 0;
            auto run = [&](auto k) { ms = time_ms([&] { k<<<sms, warps * 32>>>(out, iters * 8); }); };
            if (ilp == 1) run(dfma_kernel<1>); else if (ilp == 4) run(dfma_kernel<4>); else run(dfma_kernel<8>);
            const double flops = 2.0 * 32 * (double)iters * 8 * ilp * warps * sms;
            printf("%s\n  {\"warps_per_sm\": %d, \"ilp\": %d, \"tflops\": %.2f}", first ? "" : ",", warps, ilp,
                   flops / ms * 1e-9);
            first = false;
        }
    }
    printf("]}\n");
    return 0;
}

// fp64_peak.cu — measured FP64 issue ceilings on the box's B200 (roofline denominators for
// the fp64-bound kernels; MEASURED_PEAKS.json only carries HBM and bf16 numbers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_peak fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void dfma_tput(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __fma_rn(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ddiv_tput(double* out, int iters, double a) {
    double x0 = 1.0 + threadIdx.x * 1e-9, x1 = 2.0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
        x0 = __ddiv_rn(x0, a);
        x1 = __ddiv_rn(x1, a);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1;
}
__global__ void dfma_lat(double* out, int iters, double a, double b, long long* cycles) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b);
        x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b); x = __fma_rn(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void ddiv_lat(double* out, int iters, double a, long long* cycles) {
    double x = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) { x = __ddiv_rn(x, a); x = __ddiv_rn(x, a); x = __ddiv_rn(x, a); x = __ddiv_rn(x, a); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int iters = 20000;
    for (int blocks_per_sm : {1, 2, 4}) {
        for (int threads : {256, 512, 1024}) {
            if (blocks_per_sm * threads > 2048) continue;
            dfma_tput<8><<<sms * blocks_per_sm, threads>>>(out, 100, 1.0000001, 1e-9);
            cudaEventRecord(e0);
            dfma_tput<8><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)sms * blocks_per_sm * threads * iters * 8;
            printf("DFMA tput  blocks/SM=%d threads=%4d : %.2f T DFMA/s  (%.1f per clk per SM at %d MHz)\n", blocks_per_sm,
                   threads, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (p.clockRate * 1e3), p.clockRate / 1000);
        }
    }
    ddiv_tput<<<sms * 2, 1024>>>(out, 100, 1.0000001);
    cudaEventRecord(e0);
    ddiv_tput<<<sms * 2, 1024>>>(out, 2000, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("DDIV tput  : %.3f T div/s\n", (double)sms * 2 * 1024 * 2000 * 2 / ms / 1e9);
    dfma_lat<<<1, 32>>>(out, 1000, 1.0000001, 1e-9, cyc); cudaDeviceSynchronize();
    printf("DFMA dependent latency : %.2f cycles\n", (double)*cyc / (1000 * 8));
    ddiv_lat<<<1, 32>>>(out, 1000, 1.0000001, cyc); cudaDeviceSynchronize();
    printf("DDIV dependent latency : %.2f cycles\n", (double)*cyc / (1000 * 4));
    return 0;
}

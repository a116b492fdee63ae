// subpart_bench.cu — which warps of a CTA share an SM sub-partition (FP64 pipe)?  148 CTAs x 32 warps, only the
// warps selected by (warp % stride == 0 && warp / stride < count) run the pass-2 recurrence.
#include <cuda_runtime.h>
#include <cstdio>
struct Recip { double b, yh, yl; };
__device__ __forceinline__ double div_recip(double a, const Recip& r) {
    double t = __dmul_rn(a, r.yl);
    double q0 = __fma_rn(a, r.yh, t);
    double res = __fma_rn(-r.b, q0, a);
    return __fma_rn(res, r.yh, q0);
}
__global__ void pass2(int stride, int count, double lam, Recip r, Recip z, int steps, double* out) {
    const int w = threadIdx.x >> 5;
    if (w % stride != 0 || w / stride >= count) return;
    double p = 1.0, acc = 0.0, di = 1.0, pn = 0.0;
    for (int n = 0; n < steps; n += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double a_ = __dmul_rn(p, lam);
            pn = div_recip(p, z);
            p = div_recip(a_, r);
            acc = __dadd_rn(acc, __dmul_rn(di, pn));
            di = __dadd_rn(di, 1.0);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + p + pn;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    Recip r{1.0009765625, 0.0, 0.0}; r.yh = 1.0 / r.b; r.yl = (1.0 - r.b * r.yh) / r.b;
    Recip z{3.7, 0.0, 0.0}; z.yh = 1.0 / z.b; z.yl = (1.0 - z.b * z.yh) / z.b;
    const int steps = 40000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cfg[][2] = {{1, 1}, {1, 4}, {1, 8}, {4, 2}, {4, 4}, {4, 8}, {2, 8}, {8, 4}, {1, 32}};
    for (auto& c : cfg) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            pass2<<<148, 1024>>>(c[0], c[1], 1.0, r, z, steps, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        printf("warps {k*%d, k<%d}: %.1f cycles/step\n", c[0], c[1], ms * 1e-3 * 1.965e9 / steps);
    }
    return 0;
}

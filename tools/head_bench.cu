// head_bench.cu — cycles per step of the chain's HEAD phase (one table entry {rate, 1/rate hi, lo, packed}
// per step, 32 B), one warp, for several ways of getting the entry to the step that needs it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o head_bench head_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
struct Recip { double b, yh, yl; };
__device__ __forceinline__ double div_recip(double a, const Recip& r) {
    double t = __dmul_rn(a, r.yl);
    double q0 = __fma_rn(a, r.yh, t);
    double res = __fma_rn(-r.b, q0, a);
    return __fma_rn(res, r.yh, q0);
}
__device__ __forceinline__ void load_recip(const double* tab, int n, Recip& r) {
    const double2 v = *reinterpret_cast<const double2*>(tab + 4 * n);
    r.b = v.x; r.yh = v.y; r.yl = tab[4 * n + 2];
}
__device__ __forceinline__ void load_recip_nc(const double* tab, int n, Recip& r) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(tab + 4 * n));
    r.b = v.x; r.yh = v.y; r.yl = __ldg(tab + 4 * n + 2);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// mode 0: 2-deep register pipeline (as grid_kernel), 1: + prefetch.L1 every step, 2: __ldg loads, 3: __ldg + prefetch,
// 4: 4-deep register pipeline, 5: 8-deep register pipeline
template <int MODE>
__global__ void head(const double* __restrict__ tab, int len, int reps, double lam, double* out, long long* cyc) {
    double p = 1.0, sum = 1.0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const double* tb = tab + (size_t)rep * 4 * len;  // fresh (cold) table every repetition
        if (MODE <= 3) {
            Recip A, B;
            if (MODE >= 2) { load_recip_nc(tb, 0, A); load_recip_nc(tb, 1, B); } else { load_recip(tb, 0, A); load_recip(tb, 1, B); }
#define STEP(R)                                                                           \
    p = div_recip(__dmul_rn(p, lam), R);                                                  \
    if (MODE >= 2) load_recip_nc(tb, n + 2 < len ? n + 2 : len - 1, R);                   \
    else load_recip(tb, n + 2 < len ? n + 2 : len - 1, R);                                \
    if (MODE == 1 || MODE == 3) prefetch_l1(tb + 4 * (n + 14));                           \
    sum = __dadd_rn(sum, p);                                                              \
    ++n;
            for (int n = 0; n + 4 <= len;) { STEP(A) STEP(B) STEP(A) STEP(B) }
#undef STEP
        } else {
            constexpr int D = MODE == 4 ? 4 : 8;
            Recip R[D];
#pragma unroll
            for (int k = 0; k < D; ++k) load_recip(tb, k, R[k]);
            for (int n = 0; n + D <= len; n += D) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    p = div_recip(__dmul_rn(p, lam), R[k]);
                    load_recip(tb, n + k + D < len ? n + k + D : len - 1, R[k]);
                    sum = __dadd_rn(sum, p);
                }
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = sum + p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the warp stages 32 entries at a time in shared memory (coalesced 1 KB load, double buffered)
__global__ void head_smem(const double* __restrict__ tab, int len, int reps, double lam, double* out, long long* cyc) {
    __shared__ double4 buf[2][32];
    double p = 1.0, sum = 1.0;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const double4* tb = reinterpret_cast<const double4*>(tab + (size_t)rep * 4 * len);
        double4 nxt = tb[lane];
        for (int n0 = 0; n0 < len; n0 += 32) {
            const int cur = (n0 >> 5) & 1;
            buf[cur][lane] = nxt;
            __syncwarp();
            if (n0 + 32 < len) nxt = tb[n0 + 32 + lane];
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const double4 e = buf[cur][k];
                Recip R{e.x, e.y, e.z};
                p = div_recip(__dmul_rn(p, lam), R);
                sum = __dadd_rn(sum, p);
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = sum + p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    const int len = 256, reps = 64;
    double *tab, *out; long long* cyc;
    const size_t n = (size_t)len * 4 * reps;
    cudaMalloc(&tab, n * sizeof(double)); cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
    double* h = new double[n];
    for (size_t i = 0; i < n / 4; ++i) { double b = 1.0 + 1e-3 * (i % 7); h[4 * i] = b; h[4 * i + 1] = 1.0 / b; h[4 * i + 2] = (1.0 - b * (1.0 / b)) / b; h[4 * i + 3] = 0; }
    cudaMemcpy(tab, h, n * sizeof(double), cudaMemcpyHostToDevice);
    const char* names[] = {"2-deep regs (as now)", "2-deep + prefetch.L1 14 ahead", "2-deep, ld.global.nc", "ld.global.nc + prefetch", "4-deep regs", "8-deep regs"};
#define RUN(M) for (int w = 0; w < 2; ++w) { head<M><<<1, 32>>>(tab, len, reps, 1.0, out, cyc); cudaDeviceSynchronize(); } \
    printf("%-34s: %.1f cycles/step\n", names[M], (double)cyc[0] / (len * reps));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    for (int w = 0; w < 2; ++w) { head_smem<<<1, 32>>>(tab, len, reps, 1.0, out, cyc); cudaDeviceSynchronize(); }
    printf("%-34s: %.1f cycles/step\n", "smem-staged 32 entries / warp", (double)cyc[0] / (len * reps));
    return 0;
}

// Dependent-issue latency microbenchmarks on one warp (development aid): cycles per dependent instruction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP> __global__ void lat(double* out, long long* cyc, double a, double b, int iters) {
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (OP == 0) x = fma(x, y, a);
            if (OP == 1) x = x * y;
            if (OP == 2) x = x + y;
            if (OP == 3) { double r_; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r_) : "d"(x)); x = r_; }
            if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (OP == 5) { unsigned v = __double2loint(x); v = min(v - 12345u, v + 1000000000u); x = __hiloint2double(__double2hiint(x), (int)v); }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput with N independent chains per thread and W warps per SM sub-partition
template <int CH> __global__ void thr(double* out, long long* cyc, double a, double b, int iters) {
    double x[CH];
    for (int c = 0; c < CH; ++c) x[c] = a + c + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], b, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H", "SHFL(64-bit = 2 SHFL)", "IADD+VIADDMNMX (2 int ops)"};
    long long h;
    const int iters = 2000;
#define RUN(OP) lat<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s %.2f cycles per dependent op\n", names[OP], (double)h / (iters * 16));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    // FP64 throughput per SM sub-partition vs warps x chains: 1 CTA per SM of W*4 warps (W per SMSP)
#define THR(CH, W) thr<CH><<<1, 128 * W>>>(d, c, 1.0000001, 0.9999999, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("warps/SMSP %d chains %d: %.2f cycles per DFMA per SMSP (2.0 = pipe saturated)\n", W, CH, (double)h / (iters * 8.0 * CH * W));
    THR(1, 1) THR(2, 1) THR(4, 1) THR(8, 1) THR(1, 2) THR(2, 2) THR(1, 4) THR(2, 4) THR(3, 4) THR(4, 4) THR(1, 8) THR(2, 8)
    return 0;
}

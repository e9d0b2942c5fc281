// Issue model microbenchmark (development aid): does an FP64 instruction leave its second pipe cycle free for an
// integer instruction of another (or the same) warp?  Body = NF independent DFMAs + NI independent integer add / xor per
// thread, W warps per SM sub-partition; prints cycles per body per warp and SMSP next to the two models
//   overlap : max(2 NF, NF + NI)      (FP64 pipe takes a warp instruction every 2 cycles, issue port 1 per cycle)
//   serial  : 2 NF + NI               (an FP64 instruction blocks the issue port for both cycles)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NF, int NI> __global__ void mix(double* out, long long* cyc, double a, double b, unsigned k, int iters) {
    double x[NF > 0 ? NF : 1];
    unsigned v[NI > 0 ? NI : 1];
    for (int c = 0; c < NF; ++c) x[c] = a + c + threadIdx.x * 1e-9;
    for (int c = 0; c < NI; ++c) v[c] = threadIdx.x + c;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int c = 0; c < (NF > NI ? NF : NI); ++c) {
                if (c < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(b), "d"(a));
                if (c < NI) { if (r & 1) v[c] ^= k; else v[c] += k; }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < NF; ++c) s += x[c];
    for (int c = 0; c < NI; ++c) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&c, 8);
    long long h;
    const int iters = 4000;
#define RUN(NF, NI, W) mix<NF, NI><<<1, 128 * W>>>(d, c, 1.0000001, 0.9999999, 12345u, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("NF %2d NI %2d warps/SMSP %d: %6.2f cycles per body per warp   overlap model %3d   serial model %3d\n", NF, NI, W, (double)h / (iters * 4.0 * W), \
           (2 * NF > NF + NI ? 2 * NF : NF + NI), 2 * NF + NI);
    RUN(8, 0, 4) RUN(0, 8, 4) RUN(0, 16, 4) RUN(8, 4, 4) RUN(8, 8, 4) RUN(8, 16, 4) RUN(4, 8, 4) RUN(4, 16, 4)
    RUN(8, 8, 2) RUN(8, 8, 8) RUN(8, 16, 8)
    return 0;
}

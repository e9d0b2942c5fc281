// FP32 counterpart of rfbw.cu (development aid): issue rate of scalar FFMA and packed FFMA2 (fma.rn.f32x2) with shared
// operands versus three different register sources, 8 independent chains per thread, W warps per SM sub-partition.
//   MODE 0: x = fma(x, b, a)        b, a shared (reuse cache / uniform)
//   MODE 1: x = fma(x, y_c, z_c)    three different registers
//   MODE 2: X = fma2(X, B, A)       packed, B, A shared
//   MODE 3: X = fma2(X, Y_c, Z_c)   packed, three different register pairs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
template <int MODE, int CH> __global__ void rf(float* out, const float* in, long long* cyc, float a, float b, int iters) {
    float x[CH], y[CH], z[CH];
    unsigned long long X[CH], Y[CH], Z[CH], A, B;
    for (int c = 0; c < CH; ++c) {
        x[c] = in[c] + threadIdx.x * 1e-6f; y[c] = in[CH + c]; z[c] = in[2 * CH + c];
        X[c] = ((unsigned long long)__float_as_uint(x[c]) << 32) | __float_as_uint(y[c]);
        Y[c] = ((unsigned long long)__float_as_uint(y[c]) << 32) | __float_as_uint(z[c]);
        Z[c] = ((unsigned long long)__float_as_uint(z[c]) << 32) | __float_as_uint(x[c] + 1.0f);
    }
    A = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(b);
    B = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) x[c] = fmaf(x[c], b, a);
                if (MODE == 1) x[c] = fmaf(x[c], y[c], z[c]);
                if (MODE == 2) X[c] = fma2(X[c], B, A);
                if (MODE == 3) X[c] = fma2(X[c], Y[c], Z[c]);
            }
    }
    long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CH; ++c) s += x[c] + y[c] + z[c] + __uint_as_float((unsigned)X[c]) + __uint_as_float((unsigned)(X[c] >> 32)) + __uint_as_float((unsigned)Y[c]) + __uint_as_float((unsigned)Z[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float *d, *in; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&in, 1024); cudaMalloc(&c, 8);
    float hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0f + 1e-4f * i;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    long long h;
    const int iters = 4000;
    const char* names[] = {"FFMA  x, b, a (shared b, a)", "FFMA  x, y_c, z_c (3 distinct)", "FFMA2 X, B, A (shared B, A)", "FFMA2 X, Y_c, Z_c (3 distinct)"};
#define RUN(M, CH, W) rf<M, CH><<<1, 128 * W>>>(d, in, c, 1.0000001f, 0.9999999f, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-32s chains %d warps/SMSP %d: %.2f cycles per instruction per SMSP\n", names[M], CH, W, (double)h / (iters * 4.0 * CH * W));
    RUN(0, 8, 4) RUN(1, 8, 4) RUN(2, 8, 4) RUN(3, 8, 4)
    RUN(0, 8, 8) RUN(1, 8, 8) RUN(2, 8, 8) RUN(3, 8, 8)
    return 0;
}

// FP64 pipe rate vs operand sources (development aid): does a DFMA whose three operands are three different
// register pairs still issue every 2 cycles per SM sub-partition, or is it limited by register-file reads?
//   MODE 0: x = fma(x, b, a)      b, a shared by all chains (operand reuse cache / same registers)
//   MODE 1: x = fma(x, y_c, z_c)  three distinct register pairs per instruction, different for every chain
//   MODE 2: x = x * y_c           two distinct register pairs
//   MODE 3: x = fma(x, y_c, K)    K an immediate/constant-bank operand
//   MODE 4: x = fma(y_c, z_c, x)  accumulate form (the complex MAC pattern)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE, int CH> __global__ void rf(double* out, const double* in, long long* cyc, double a, double b, int iters) {
    double x[CH], y[CH], z[CH];
    for (int c = 0; c < CH; ++c) {
        x[c] = in[c] + threadIdx.x * 1e-9;
        y[c] = in[CH + c];
        z[c] = in[2 * CH + c];
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) x[c] = fma(x[c], b, a);
                if (MODE == 1) x[c] = fma(x[c], y[c], z[c]);
                if (MODE == 2) x[c] = x[c] * y[c];
                if (MODE == 3) x[c] = fma(x[c], y[c], 0.333333333333);
                if (MODE == 4) x[c] = fma(y[c], z[c], x[c]);
            }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += x[c] + y[c] + z[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *d, *in; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&in, 1024); cudaMalloc(&c, 8);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    long long h;
    const int iters = 4000;
    const char* names[] = {"fma(x, b, a) shared b,a", "fma(x, y_c, z_c) 3 distinct", "x * y_c", "fma(x, y_c, const)", "fma(y_c, z_c, x)"};
#define RUN(M, CH, W) rf<M, CH><<<1, 128 * W>>>(d, in, c, 1.0000001, 0.9999999, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-30s chains %d warps/SMSP %d: %.2f cycles per FP64 instruction per SMSP\n", names[M], CH, W, (double)h / (iters * 4.0 * CH * W));
    RUN(0, 8, 4) RUN(1, 8, 4) RUN(2, 8, 4) RUN(3, 8, 4) RUN(4, 8, 4)
    RUN(0, 8, 2) RUN(1, 8, 2) RUN(2, 8, 2) RUN(3, 8, 2) RUN(4, 8, 2)
    RUN(1, 4, 4) RUN(4, 4, 4) RUN(1, 8, 1) RUN(4, 8, 1)
    return 0;
}

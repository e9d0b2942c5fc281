// Does nvcc 12.9 / sm_100a emit packed FP32 (FFMA2 / FMUL2 / FADD2) and what is its issue rate? (development probe)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k2(float2* out, int n, float2 a, float2 b) {
    float2 x0 = out[threadIdx.x], x1 = a, x2 = b, x3 = make_float2(a.y, b.x);
    for (int i = 0; i < n; ++i) {
        x0 = __ffma2_rn(x0, a, b); x1 = __ffma2_rn(x1, a, b); x2 = __ffma2_rn(x2, a, b); x3 = __ffma2_rn(x3, a, b);
        x0 = __fmul2_rn(x0, b);    x1 = __fadd2_rn(x1, a);    x2 = __fmul2_rn(x2, b);    x3 = __fadd2_rn(x3, a);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(x0.x + x1.x + x2.x + x3.x, x0.y + x1.y + x2.y + x3.y);
}
__global__ void k1(float* out, int n, float a, float b) {
    float x0 = out[threadIdx.x], x1 = a, x2 = b, x3 = a + b, x4 = a - b, x5 = a * b, x6 = b - a, x7 = a * a;
    for (int i = 0; i < n; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        x0 *= b; x1 += a; x2 *= b; x3 += a; x4 *= b; x5 += a; x6 *= b; x7 += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    float2* d2; float* d1;
    const int blocks = 148 * 8, threads = 256, n = 1 << 16;
    cudaMalloc(&d2, blocks * threads * sizeof(float2)); cudaMalloc(&d1, blocks * threads * sizeof(float));
    cudaMemset(d2, 0, blocks * threads * sizeof(float2)); cudaMemset(d1, 0, blocks * threads * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k2<<<blocks, threads>>>(d2, n, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f)); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        // per iteration: 8 packed instructions = 16 lane ops (FMA counted as 1 op here)
        printf("packed : %.3f ms, %.2f T lane-ops/s, %.2f T warp-instr/s\n", ms, 16.0 * n * blocks * threads / ms / 1e9, 8.0 * n * blocks * threads / 32 / ms / 1e9);
        cudaEventRecord(e0); k1<<<blocks, threads>>>(d1, n, 1.0001f, 1e-3f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("scalar : %.3f ms, %.2f T lane-ops/s, %.2f T warp-instr/s\n", ms, 16.0 * n * blocks * threads / ms / 1e9, 16.0 * n * blocks * threads / 32 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Register-file bandwidth model (development aid): FP64 instructions with k distinct register-pair operands next to
// integer instructions with 2 or 3 distinct register operands.  If the vector register file delivers one 64-bit operand
// (or two 32-bit ones, one per bank) per cycle and SM sub-partition, a body of NF three-register DFMAs and NI
// three-register IADD3s costs about 3 NF + 1.5 NI cycles, not max(3 NF, NF + NI).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NF, int NI, int IREGS, int FREGS> __global__ void rfmix(double* out, const double* in, long long* cyc, int iters) {
    double x[NF > 0 ? NF : 1], y[NF > 0 ? NF : 1], z[NF > 0 ? NF : 1];
    unsigned v[NI > 0 ? NI : 1], w[NI > 0 ? NI : 1], u[NI > 0 ? NI : 1];
    for (int c = 0; c < NF; ++c) { x[c] = in[c] + threadIdx.x * 1e-9; y[c] = in[8 + c]; z[c] = in[16 + c]; }
    for (int c = 0; c < NI; ++c) { v[c] = threadIdx.x + c; w[c] = (unsigned)in[24 + c] + c; u[c] = (unsigned)in[40 + c] * 3 + c; }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < (NF > NI ? NF : NI); ++c) {
                if (c < NF) { if (FREGS == 3) x[c] = fma(x[c], y[c], z[c]); else x[c] = x[c] * y[c]; }
                if (c < NI) { if (IREGS == 3) v[c] = v[c] + w[c] + u[c]; else v[c] = v[c] + w[c]; }
                if (c + NF < NI) { const int d = c + NF; if (IREGS == 3) v[d] = v[d] + w[d] + u[d]; else v[d] = v[d] + w[d]; }
            }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < NF; ++c) s += x[c] + y[c] + z[c];
    for (int c = 0; c < NI; ++c) s += v[c] + w[c] + u[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *d, *in; long long* c; cudaMalloc(&d, 1 << 22); cudaMalloc(&in, 1024); cudaMalloc(&c, 8);
    double hin[64]; for (int i = 0; i < 64; ++i) hin[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    long long h;
    const int iters = 4000;
#define RUN(NF, NI, IR, FR, W) rfmix<NF, NI, IR, FR><<<1, 128 * W>>>(d, in, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("NF %d (%d-reg FP64)  NI %2d (%d-reg int)  warps/SMSP %d: %6.2f cycles per body per warp\n", NF, FR, NI, IR, W, (double)h / (iters * 4.0 * W));
    RUN(8, 0, 3, 3, 4) RUN(8, 0, 3, 2, 4) RUN(0, 8, 2, 3, 4) RUN(0, 8, 3, 3, 4) RUN(0, 16, 3, 3, 4)
    RUN(8, 8, 2, 3, 4) RUN(8, 8, 3, 3, 4) RUN(8, 16, 3, 3, 4) RUN(8, 16, 2, 3, 4)
    RUN(8, 8, 3, 2, 4) RUN(8, 16, 3, 2, 4) RUN(8, 16, 2, 2, 4)
    RUN(4, 8, 3, 3, 4) RUN(4, 16, 3, 3, 4)
    return 0;
}

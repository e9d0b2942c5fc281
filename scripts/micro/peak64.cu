// Why does tp3_peak_probe read 34.2 TFLOP/s when a one-register DFMA issues every 2.05 cycles per sub-partition
// (= 36.3 TFLOP/s at 1965 MHz on 148 SMs)?  
// Full-device DFMA / DMUL / mixed chains for a range of chains per thread, CTAs per SM and run lengths; for each run the
// SM clock that actually applied is measured inside the kernel (clock64 against globaltimer), so that a power-capped
// burst shows up as a lower clock rather than as "lost" issue slots.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o _build/peak64 peak64.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// MODE 0: DFMA x = x*a + b (a, b uniform); MODE 1: DMUL x = x*a; MODE 2: alternating DFMA / DMUL; MODE 3: DADD
template <int MODE, int CH> __global__ void probe(double* out, long long* cyc, unsigned long long* ns, double a, double b, int iters) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    const long long t0 = clock64();
    const unsigned long long g0 = gtimer();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 64 / CH; ++r)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) x[c] = fma(x[c], a, b);
                if (MODE == 1) x[c] = x[c] * a;
                if (MODE == 2) x[c] = (c & 1) ? x[c] * a : fma(x[c], a, b);
                if (MODE == 3) x[c] = x[c] + b;
            }
    }
    const long long t1 = clock64();
    const unsigned long long g1 = gtimer();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { *cyc = t1 - t0; *ns = g1 - g0; }
}

template <int MODE, int CH> void run(const char* name, int ctas_per_sm, int threads, int iters, int sms, double* out, long long* cyc, unsigned long long* ns) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; long long hc = 0; unsigned long long hn = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<MODE, CH><<<sms * ctas_per_sm, threads>>>(out, cyc, ns, 0.999999, 1e-6, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) { best = ms; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hn, ns, 8, cudaMemcpyDeviceToHost); }
    }
    const double inst = (double)sms * ctas_per_sm * threads * (double)iters * 64.0;   // lane-instructions
    const double flop = inst * (MODE == 0 ? 2.0 : MODE == 2 ? 1.5 : 1.0);
    const double warps_per_smsp = ctas_per_sm * threads / 32.0 / 4.0;
    const double cyc_per_inst = (double)hc / ((double)iters * 64.0 * warps_per_smsp);
    printf("%-10s chains %2d  %d x %3d threads/SM  %7.2f ms  %6.2f TFLOP/s  %5.3f cycles/instr/SMSP  SM clock during run %7.1f MHz\n", name, CH, ctas_per_sm,
           threads, best, flop / (best * 1e-3) / 1e12, cyc_per_inst, 1e3 * (double)hc / (double)hn);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; long long* cyc; unsigned long long* ns;
    cudaMalloc(&out, (size_t)sms * 16 * 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&ns, 8);
    printf("%s, %d SMs, nominal clock %d MHz\n", p.name, sms, p.clockRate / 1000);
    // the library's probe: 8 chains, 8 CTAs x 256 threads per SM, 4096 iterations
    run<0, 8>("DFMA", 8, 256, 4096, sms, out, cyc, ns);
    run<0, 8>("DFMA", 8, 256, 65536, sms, out, cyc, ns);   // ~16x longer: a sustained rather than a burst figure
    run<0, 4>("DFMA", 8, 256, 4096, sms, out, cyc, ns);
    run<0, 16>("DFMA", 8, 256, 4096, sms, out, cyc, ns);
    run<0, 16>("DFMA", 4, 256, 4096, sms, out, cyc, ns);
    run<0, 8>("DFMA", 4, 128, 8192, sms, out, cyc, ns);
    run<0, 8>("DFMA", 2, 128, 8192, sms, out, cyc, ns);
    run<0, 8>("DFMA", 1, 128, 16384, sms, out, cyc, ns);
    run<1, 8>("DMUL", 8, 256, 4096, sms, out, cyc, ns);
    run<2, 8>("DFMA/DMUL", 8, 256, 4096, sms, out, cyc, ns);
    run<3, 8>("DADD", 8, 256, 4096, sms, out, cyc, ns);
    return 0;
}

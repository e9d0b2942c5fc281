// Hand-written FP64 elementary functions for the fused kernel.
//
// Why not CUDA's libdevice versions: ncu on the first kernel (profiles/r01_v1_fast_f64_ranf.txt)
// showed only 31 % of the issue slots going to the FP64 pipe; log()/sincospi()/sqrt() spend more
// instructions on UMOV pairs (64-bit immediates), integer fix-ups and slow-path branches than on
// DFMAs.  Here the coefficients are kernel parameters (uniform registers, see TP3_COEFF_IMM below), there
// are no special-case branches (the operands of this kernel are known to be finite and normal), and
// the argument reductions use what we know about the inputs:
//   neg_log      x in (0, 2): 128-entry table of (1/c, -log c) in shared memory + degree-6 log1p
//   sincos_256   u in [0, 1): 2 pi u = 2 pi (k + d) / 256, table of 256 directions in shared memory + rotation by
//                the exact remainder d (|d| <= 1/2) with degree-5 / degree-6 polynomials; no quadrant logic
//   sqrt / rsqrt / rcp        MUFU seed (RSQ64H / RCP64H) + Newton steps in FMA form
// Accuracy (tests/test_gpu_parity.py::test_fastmath_accuracy, measured on B200): <= 2 ulp for sqrt / rcp / rsqrt,
// <= 3e-16 (log) and <= 2.5e-16 (sin, cos) absolute, i.e. ~1e-16 relative, four orders of magnitude inside the 1e-13 per-event budget that keeps every
// accumulated quantity within 1e-10 of the reference.
#pragma once

namespace tp3 {

#include "fastmath_tables.inc"

#ifndef TP3_COEFF_IMM
#define TP3_COEFF_IMM 2
#endif
// The polynomial coefficients of -log and sin/cos.  A DFMA whose three sources are three different register pairs issues
// every 3 cycles on B200 instead of every 2 (the vector register file delivers one 64-bit operand per cycle,
// profiles/r01_micro_fp64_operands.txt), so a Horner step wants its coefficient in a UNIFORM register:
//   2  kernel parameters (PhysParams::fc): loaded once with LDCU into uniform registers, like the physics parameters
//   1  literals: ptxas rebuilds them in vector registers with two IMAD.MOV per coefficient and event
//   0  __constant__ arrays: loaded into vector registers with LDC.64 once per event
struct FastCoef {
    double log1p[5], neg_ln2, rot_sin[3], rot_cos[3];
    double u_scale, u_bias, phi_scale, phi_bias;  // RANF word -> uniform (u32_times): 1e-9, -(2^52 1e-9), 256e-9, -(2^52 256e-9)
    double u_scale2, u_scale_sq;                  // 2e-9 (cos_theta = 2u - 1 in one FMA), 1e-18 (r r' from the integers)
};
#define TP3_FAST_COEF_INIT {{TP3_LOG1P_COEFFS}, TP3_NEG_LN2, {TP3_ROT_SIN_COEFFS}, {TP3_ROT_COS_COEFFS}, \
                            1e-9, -4503599627370496.0 * 1e-9, 256e-9, -4503599627370496.0 * 256e-9, 2e-9, 1e-18}
#if TP3_COEFF_IMM == 2
#define TP3_COEFF_ARRAYS                                                                                   \
    const double *const kLog1p = fm.fc->log1p, *const kRotSin = fm.fc->rot_sin, *const kRotCos = fm.fc->rot_cos; \
    const double kNegLn2 = fm.fc->neg_ln2;
#elif TP3_COEFF_IMM == 1
// Polynomial coefficients as literals: ptxas materialises them in uniform registers (UMOV pairs), and a DFMA whose
// third source is a uniform register reads two register pairs from the vector register file instead of three.
// (As __constant__ arrays they were loaded into vector registers with LDC.64 once per event and every Horner step
// became a three-register DFMA: 3 cycles instead of 2 on B200, profiles/r01_micro_fp64_operands.txt.)
#define TP3_COEFF_ARRAYS                                   \
    constexpr double kLog1p[5] = {TP3_LOG1P_COEFFS};       \
    constexpr double kNegLn2 = TP3_NEG_LN2;                \
    constexpr double kRotSin[3] = {TP3_ROT_SIN_COEFFS};    \
    constexpr double kRotCos[3] = {TP3_ROT_COS_COEFFS};
#else
#define TP3_COEFF_ARRAYS
__constant__ double kLog1p[5] = {TP3_LOG1P_COEFFS};
__constant__ double kNegLn2 = TP3_NEG_LN2;
__constant__ double kRotSin[3] = {TP3_ROT_SIN_COEFFS};
__constant__ double kRotCos[3] = {TP3_ROT_COS_COEFFS};
#endif

#ifdef TP3_EXPERIMENT_SMALL_TABLES   /* timing experiment only (wrong results): how much would 20 resident warps per SM buy? */
#define TP3_LOG_MASK 63
#define TP3_SC_MASK 127
#else
#define TP3_LOG_MASK 127
#define TP3_SC_MASK 255
#endif
struct FastMathSmem {
    double2 log_tab[TP3_LOG_MASK + 1];
    double2 sincos_tab[TP3_SC_MASK + 1];  // {sin, cos}(2 pi k / 256)
};
// What the elementary functions need: the CTA's tables and the coefficients (a kernel parameter)
struct FastMath {
    const FastMathSmem* sm;
    const FastCoef* fc;
};

// The tables live in GLOBAL memory (L2-resident, 6 KB) and are copied with coalesced 128-bit loads, all of a lane's loads in flight
// at once.  (Until session 50 they were __constant__: every lane reads a different entry, and a constant-bank load with 32 different
// addresses is served one address at a time -- 768 serialized accesses per CTA, which a one-warp CTA that lives for a single batch
// or for one 2500-event part of the faster-evgen physics kernel pays at every start.)
template <int THREADS> __device__ __forceinline__ void fastmath_load(FastMathSmem* sm) {
#pragma unroll
    for (int i = 0; i <= TP3_LOG_MASK; i += THREADS)
        if (i + (int)threadIdx.x <= TP3_LOG_MASK) sm->log_tab[i + threadIdx.x] = __ldg(&kLogTable[i + threadIdx.x]);
#pragma unroll
    for (int i = 0; i <= TP3_SC_MASK; i += THREADS)
        if (i + (int)threadIdx.x <= TP3_SC_MASK) sm->sincos_tab[i + threadIdx.x] = __ldg(&kSinCosTable[i + threadIdx.x]);
}

__device__ __forceinline__ double mufu_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double mufu_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/x, x finite normal: seed error e0 -> e0^3 with three FMAs
__device__ __forceinline__ double fast_rcp(double x) {
    const double y = mufu_rcp(x);
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// sqrt(x) and 1/sqrt(x) together (Goldschmidt form), x > 0 finite normal
__device__ __forceinline__ void fast_sqrt_rsqrt(double x, double& s, double& rs) {
    const double y = mufu_rsqrt(x);
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    s = g;
    rs = h + h;
}
// 1/sqrt(x) alone, x > 0 finite normal: the h half of the Goldschmidt pair above (the last update of g is not needed)
__device__ __forceinline__ double fast_rsqrt(double x) {
    const double y = mufu_rsqrt(x);
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    h = fma(h, r, h);
    return h + h;
}
// sqrt(x), x >= 0; x is biased by 1e-300 so that 0 gives 1e-150 (0 for every use in this kernel) instead of
// 0 * inf; the nonzero arguments here are >= 4e-9 and unchanged by the bias
#ifndef TP3_SQRT_RESIDUAL
#define TP3_SQRT_RESIDUAL 1
#endif
__device__ __forceinline__ double fast_sqrt(double x) {
    x += 1e-300;
    const double y = mufu_rsqrt(x);
    double g = x * y;
    const double h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);  // ~40 good bits
#if TP3_SQRT_RESIDUAL
    // second step on the residual x - g^2 with the SEED-accurate h = 1/(2 sqrt x): the correction is 2^-40 of the
    // result, so 20 good bits of h are enough, and the refinement of h (one DFMA) is not needed
    return fma(fma(-g, g, x), h, g);
#else
    const double h1 = fma(h, r, h);
    return fma(g, fma(-g, h1, 0.5), g);
#endif
}

// -log(x) for finite normal x > 0
__device__ __forceinline__ double fast_neg_log(double x, const FastMath fm) {
    const FastMathSmem* const sm = fm.sm;
    TP3_COEFF_ARRAYS
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int k = (hi >> 20) - 1023;
    const double2 t = sm->log_tab[(hi >> 13) & TP3_LOG_MASK];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, kLog1p[4], kLog1p[3]);
    p = fma(r, p, kLog1p[2]);
    p = fma(r, p, kLog1p[1]);
    p = fma(r, p, kLog1p[0]);
    p = fma(r * r, p, r);  // log1p(r)
    return fma((double)k, kNegLn2, t.y) - p;
}

// sin(2 pi u), cos(2 pi u) given t = 256 u in [0, 256]: table of 256 directions + rotation by the remainder.
// 2 pi u = 2 pi (k + d) / 256 with k = rint(t), |d| <= 1/2 (exact), so the rotation angle is at most pi/256 and
// degree-5 / degree-6 Taylor polynomials are exact to 1e-17; no quadrant logic, 12 FP64 instructions.
__device__ __forceinline__ void fast_sincos_256(double t, const FastMath fm, double& s, double& c) {
    const FastMathSmem* const sm = fm.sm;
    TP3_COEFF_ARRAYS
    const double kf = rint(t);
    const int k = __double2int_rn(t) & TP3_SC_MASK;
    const double2 sc = sm->sincos_tab[k];
    const double d = t - kf;
    const double d2 = d * d;
    const double sb = d * fma(d2, fma(d2, kRotSin[2], kRotSin[1]), kRotSin[0]);
    const double cb = fma(d2, fma(d2, fma(d2, kRotCos[2], kRotCos[1]), kRotCos[0]), 1.0);
    s = fma(sc.x, cb, sc.y * sb);
    c = fma(sc.y, cb, -(sc.x * sb));
}

// (double)n * scale for 0 <= n < 2^32 with ONE FMA and no I2F: the word n dropped into the low half
// of 2^52 is exactly 2^52 + n, and fma(2^52 + n, scale, -(2^52 * scale)) rounds the exact product
// n * scale once — bit-identical to the conversion followed by a multiply (2^52 * scale is exact).
#ifndef TP3_U32_I2F_LITERAL
#define TP3_U32_I2F_LITERAL 0
#endif
__device__ __forceinline__ double u32_times(uint32_t n, double scale) {
#if TP3_U32_I2F_LITERAL
    return (double)(int)n * scale;
#else
    const double d = __hiloint2double(0x43300000, (int)n);
    return fma(d, scale, -4503599627370496.0 * scale);
#endif
}

// The same with scale and -(2^52 * scale) supplied by the caller (kernel parameters: uniform registers, loaded once,
// instead of four immediates rebuilt for every event).
#ifndef TP3_U32_I2F
#define TP3_U32_I2F 1
#endif
__device__ __forceinline__ double u32_times(uint32_t n, double scale, double neg_bias) {
#if TP3_U32_I2F
    // I2F.F64 (XU pipe: one 32-bit register read) + a two-operand DMUL, literally ranf.rs:99.  The single-FMA form below
    // needs two moves to build the register pair and reads three register pairs (3 cycles of the register file).
    return (double)(int)n * scale;
#else
    return fma(__hiloint2double(0x43300000, (int)n), scale, neg_bias);
#endif
}

}  // namespace tp3

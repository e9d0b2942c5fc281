// Per-event physics of e+e- -> 3 photons, register resident, templated on the run's Float.
//
//   gen_event      RAMBO for 3 massless photons       src/evgen.rs:89-132,174-207
//   keep_event     the four cuts as a predicate        src/evcut.rs:42-96
//   me_literal     spinor products + helicity amps     src/spinor.rs:31-162, src/matelems.rs:53-86
//   me_fast        the same five helicity sums after algebraic reduction:
//                  |s_ij|^2 = 2 p_i.p_j turns the A and B+ sums into real dot products, the mixed
//                  term needs only s_jk^2 and u_k = s_0k s_1k, and no square root is left
//                  (one reciprocal per photon + one for the common denominator).
#pragma once

#include <cstdint>

#include "fastmath.cuh"

namespace tp3 {

template <class F> struct Num;
template <> struct Num<double> {
    static constexpr double MIN_POSITIVE = 2.2250738585072014e-308;
    static constexpr double TWO_PI = 6.283185307179586476925286766559;
    static constexpr double RAC8 = 2.8284271247461900976033774484194;
};
template <> struct Num<float> {
    static constexpr float MIN_POSITIVE = 1.17549435e-38f;
    static constexpr float TWO_PI = 2.0f * 3.14159265358979323846f;
    static constexpr float RAC8 = 2.0f * 1.41421356237309504880f;
};

// Result type of a comparison, and "is this (warp-uniform) parameter positive": bool / the value itself for the
// scalar Floats; the two-events-per-lane type of f32x2.cuh supplies its own.
template <class F> struct MaskOf { using type = bool; };
__device__ __forceinline__ bool uniform_positive(double x) { return x > 0.0; }
__device__ __forceinline__ bool uniform_positive(float x) { return x > 0.0f; }

__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float log_t(float x) { return logf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double abs_t(double x) { return fabs(x); }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }

// Fast-path math: hand-written FP64 (fastmath.cuh). The f32 fast path goes straight to the SFU
// approximations (MUFU.LG2 / SIN / COS / RCP / RSQ, 2-3 ulp): the stated f32 bounds are three
// orders of magnitude looser than that, and the f32 literal kernel keeps the IEEE libdevice functions.
// (The .ftz forms are ONE MUFU instruction each; without the modifier the compiler wraps every call in a rescaling of
// subnormal arguments -- FSETP, FMUL by 2^24, FSEL, the MUFU, an FMUL back -- that no argument here can need: the reciprocals
// are taken of E + Z > MIN_POSITIVE, of norms and of m + E, the logarithm of e + MIN_POSITIVE, the square roots of x + 1e-30.
// For normal arguments and results both forms return the same bits.)
__device__ __forceinline__ float mufu_rcp_f(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsqrt_f(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2_f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double rcp_t(double x) { return fast_rcp(x); }
__device__ __forceinline__ float rcp_t(float x) { return mufu_rcp_f(x); }
__device__ __forceinline__ double neg_log_t(double x, const FastMath sm) { return fast_neg_log(x, sm); }
__device__ __forceinline__ float neg_log_t(float x, const FastMath) { return -(mufu_lg2_f(x) * 0.693147182f); }  // __logf, one MUFU
// The azimuth uniform arrives pre-scaled by an exact power of two: 256 u in f64 (table + rotation), 4 u in f32
// (quarter turns for the SFU path).
template <class F> struct PhiScale;
template <> struct PhiScale<double> { static constexpr double value = 256.0; };
template <> struct PhiScale<float> { static constexpr float value = 4.0f; };
__device__ __forceinline__ void sincos_scaled_t(double t, const FastMath fm, double* s, double* c) { fast_sincos_256(t, fm, *s, *c); }
// sin, cos of t quarter turns, 0 <= t <= 4, on the SFU.
// Shipped form (TP3_F32_SINCOS_DIRECT = 1): the angle is reflected into x = pi - theta = (2 - t) pi/2, |x| <= pi -- the range on
// which sin.approx / cos.approx have their documented absolute error (2^-21.4) -- and sin(theta) = sin(x), cos(theta) = -cos(x):
// one subtraction, one multiplication and the two SFU calls; the sign folds into the consumer's operand modifier.
// A/B form (= 0, shipped until session 47): quadrant split, q = rint(t) from the float adder (t + 1.5 * 2^23 holds rint(t) in its low
// mantissa bits and (t + M) - M is that integer as a float), remainder |x| <= pi/4, swap and signs from q: 7 % more issue slots for
// the whole kernel.  Both forms agree with the f32 oracle to the same per-event figures (profiles/r02_f32_sincos_ab.txt,
// profiles/r02_f32_per_event.txt); over the 1e7 events of the golden run they decide 5 resp. 1 cuts differently from the reference.
#ifndef TP3_F32_SINCOS_DIRECT
#define TP3_F32_SINCOS_DIRECT 1
#endif
constexpr float kRintMagic = 12582912.0f;
constexpr float kHalfPiF = 1.57079632679489661923f;
__device__ __forceinline__ float mufu_sin_f(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos_f(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void sincos_quadrant(float x, int q, float* s, float* c) {
    const float ps = __sinf(x), pc = __cosf(x);
    const bool swap = q & 1;
    const float s0 = swap ? pc : ps, c0 = swap ? ps : pc;
    *s = __int_as_float(__float_as_int(s0) ^ (int)((unsigned)(q & 2) << 30));
    *c = __int_as_float(__float_as_int(c0) ^ (int)((unsigned)((q + 1) & 2) << 30));
}
__device__ __forceinline__ void sincos_scaled_t(float t, const FastMath, float* s, float* c) {
#if TP3_F32_SINCOS_DIRECT
    const float x = (2.0f - t) * kHalfPiF;
    *s = mufu_sin_f(x);
    *c = -mufu_cos_f(x);
#else
    const float tm = __fadd_rn(t, kRintMagic);
    const float qf = __fadd_rn(tm, -kRintMagic);
    sincos_quadrant((t - qf) * kHalfPiF, __float_as_int(tm), s, c);
#endif
}
__device__ __forceinline__ double sqrt_pos_t(double x) { return fast_sqrt(x); }
__device__ __forceinline__ float sqrt_pos_t(float x) { return x * mufu_rsqrt_f(x + 1e-30f); }
__device__ __forceinline__ void sqrt_rsqrt_t(double x, double* s, double* rs) { fast_sqrt_rsqrt(x, *s, *rs); }
__device__ __forceinline__ void sqrt_rsqrt_t(float x, float* s, float* rs) {
    *rs = mufu_rsqrt_f(x);
    *s = x * *rs;
}

// Kernel-side view of tp3_params in the run's Float.
template <class F> struct PhysParams {
    F e_total, acut, bcut, e_min, sincut;
    F g_a, g_beta_p, g_beta_m;
    F sigma_contribs[5];
    // Products of the parameters above that me_fast / keep_event would otherwise recompute for every event in vector
    // FP64 (the compiler does not hoist all of them); same operations, same order, done once on the host in F:
    F k_m0;    // (g_a g_a) 8
    F k_m1;    // (g_beta_p g_beta_p) ((8 e^2) e^2)
    F k_m2;    // (g_beta_m g_beta_m) ((4 e^2) e^2)
    F k_mix;   // (-(g_a g_beta_p)) e^2
    F omb_e, he; // (1 - bcut) / e, e / 2
    FastCoef fc;  // coefficients of the hand-written FP64 functions (fastmath.cuh); unused by the f32 kernels
};

// ------------------------------------------------------------------ event generation
// Conformal transform of RAMBO to the total energy + optional sort (evgen.rs:94-118):
// q[k] = raw (X, Y, Z, E) of photon k  ->  p[k] = (X, Y, Z, E), optionally sorted by decreasing E.
// CONS3: the third photon is what 4-momentum conservation leaves, p_2 = (0,0,0,e_total) - p_0 - p_1, instead of the
// transform of q_2 (3 additions instead of 9 multiply-adds for the E and X the cuts need).  The transform conserves the
// total momentum to rounding error, so the two differ by ~1e-16 e_total; the pair cuts, |s_ij|^2 and the survivor queue of
// the fused kernel already rely on the same identity.
template <class F, bool SORT, bool LITERAL, bool CONS3 = false>
__device__ __forceinline__ void conformal_transform(const F q[3][4], F e_total, F p[3][4]) {
    F r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = (q[0][c] + q[1][c]) + q[2][c];
    const F m2 = r[3] * r[3] - ((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
    F alpha, m, beta;
    if constexpr (LITERAL) {
        alpha = e_total / m2;
        m = sqrt_t(m2);
        beta = (F)1 / (m + r[3]);
    } else {
        F rs;
        sqrt_rsqrt_t(m2, &m, &rs);
        alpha = e_total * (rs * rs);
        beta = rcp_t(m + r[3]);
    }
    const F am = alpha * m;
#pragma unroll
    for (int k = 0; k < (CONS3 ? 2 : 3); ++k) {
        const F rq = (q[k][0] * r[0] + q[k][1] * r[1]) + q[k][2] * r[2];
        p[k][3] = alpha * (r[3] * q[k][3] - rq);
        if constexpr (LITERAL) {
            const F b = beta * rq - q[k][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) p[k][c] = alpha * (m * q[k][c] + b * r[c]);
        } else {
            const F ab = alpha * (beta * rq - q[k][3]);
#pragma unroll
            for (int c = 0; c < 3; ++c) p[k][c] = am * q[k][c] + ab * r[c];
        }
    }
    if constexpr (CONS3) {
#pragma unroll
        for (int c = 0; c < 3; ++c) p[2][c] = -(p[0][c] + p[1][c]);
        p[2][3] = (e_total - p[0][3]) - p[1][3];
    }
    if constexpr (SORT) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = a + 1; b < 3; ++b) {
                const bool sw = p[b][3] > p[a][3];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const F x = p[a][c], y = p[b][c];
                    p[a][c] = sw ? y : x;
                    p[b][c] = sw ? x : y;
                }
            }
    }
}


#ifndef TP3_CONS3
#define TP3_CONS3 1
#endif
// Third photon from momentum conservation: the f64 fast paths only (in f32 the rounding error of the sum is 1e-7 e_total)
template <class F, bool LITERAL> struct Cons3 { static constexpr bool value = false; };
template <> struct Cons3<double, false> { static constexpr bool value = TP3_CONS3 != 0; };

// Hook called between the stages of an event so that independent work of the same warp (the next
// iteration's random numbers) can be interleaved with the FP64 chains; NoTick does nothing.
struct NoTick {
    template <int K> __device__ __forceinline__ void at() const {}
};

// One photon of generate_raw (evgen.rs:182-206): u = (cos_theta, phi, r, r') uniforms -> q = (X, Y, Z, E)
template <class F, bool LITERAL>
__device__ __forceinline__ void raw_photon(const F* u, const FastMath fm, F q[4]) {
    const F c = (F)2 * u[0] - (F)1;
    const F e = u[2] * u[3];
    F sphi, cphi, st, en;
    if constexpr (LITERAL) {
        sincos_t(Num<F>::TWO_PI * u[1], &sphi, &cphi);
        st = sqrt_t((F)1 - c * c);
        en = -log_t(e + Num<F>::MIN_POSITIVE);
    } else {
#ifdef TP3_EXPERIMENT_KO_SINCOS   /* timing experiments only (wrong results): knock one function out */
        sphi = u[1];
        cphi = (F)1 - u[1];
#else
        sincos_scaled_t(u[1], fm, &sphi, &cphi);
#endif
#ifdef TP3_EXPERIMENT_KO_SQRT
        st = (F)1 - c * c;
#else
        st = sqrt_pos_t((F)1 - c * c);
#endif
#ifdef TP3_EXPERIMENT_KO_LOG
        en = e + (F)1;
#else
        en = neg_log_t(e + Num<F>::MIN_POSITIVE, fm);
#endif
    }
    if constexpr (LITERAL) {
        q[0] = en * (st * sphi);
        q[1] = en * (st * cphi);
    } else {  // one multiplication less
        const F es = en * st;
        q[0] = es * sphi;
        q[1] = es * cphi;
    }
    q[2] = en * c;
    q[3] = en;
}

// The same for the f64 RANF fast path, straight from the stream integers d_j = (double)n_j (ranf.rs:99 would first
// scale each to u_j = 1e-9 d_j): cos_theta = 2e-9 d_0 - 1 in one FMA, 256 phi-turns = 256e-9 d_1, and
// r r' = 1e-18 (d_2 d_3) with one multiplication less.  Each differs from the reference's expression by one rounding.
// (rr = d_2 d_3 rounded once: the caller forms it as the exact 64-bit integer product converted once where that fits)
__device__ __forceinline__ void raw_photon_ints(const double* d, double rr, const FastMath fm, double q[4]) {
    const double c = fma(d[0], fm.fc->u_scale2, -1.0);
    const double e = fma(rr, fm.fc->u_scale_sq, Num<double>::MIN_POSITIVE);
    double sphi, cphi;
    fast_sincos_256(d[1] * fm.fc->phi_scale, fm, sphi, cphi);
    const double st = fast_sqrt(fma(-c, c, 1.0));
    const double en = fast_neg_log(e, fm);
    const double es = en * st;
    q[0] = es * sphi;
    q[1] = es * cphi;
    q[2] = en * c;
    q[3] = en;
}
template <bool SORT, class Tick>
__device__ __forceinline__ void gen_event_ints(const double d[12], const double rr[3], double e_total, const FastMath fm, double p[3][4], Tick& tick) {
    double q[3][4];
    raw_photon_ints(d, rr[0], fm, q[0]);
    tick.template at<1>();
    raw_photon_ints(d + 4, rr[1], fm, q[1]);
    tick.template at<2>();
    raw_photon_ints(d + 8, rr[2], fm, q[2]);
    tick.template at<3>();
    conformal_transform<double, SORT, false, Cons3<double, false>::value>(q, e_total, p);
}

// u[12]: uniforms in the reference's draw order (per photon: cos_theta, phi, r, r'; evgen.rs:182-187). In the fast
// variant the phi slot holds PhiScale<F>::value * u (an exact power-of-two scaling) instead of u.
template <class F, bool SORT, bool LITERAL, class Tick>
__device__ __forceinline__ void gen_event(const F u[12], F e_total, const FastMath fm, F p[3][4], Tick& tick) {
    F q[3][4];
    raw_photon<F, LITERAL>(u, fm, q[0]);
    tick.template at<1>();
    raw_photon<F, LITERAL>(u + 4, fm, q[1]);
    tick.template at<2>();
    raw_photon<F, LITERAL>(u + 8, fm, q[2]);
    tick.template at<3>();
    conformal_transform<F, SORT, LITERAL, Cons3<F, LITERAL>::value>(q, e_total, p);
}

// `faster-evgen` raw generation (evgen.rs:143-173): cos_theta and exp(-E) from 9 uniforms, the azimuth
// from a point (x, y) of the unit disc with squared radius r2 (the rejection loop lives in the caller,
// because it decides how many random numbers the event consumes).
template <class F, bool SORT>
__device__ __forceinline__ void gen_event_faster(const F u9[9], const F xy[3][2], const F r2[3], F e_total,
                                                 const FastMath fm, F p[3][4]) {
    F q[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const F c = (F)2 * u9[k] - (F)1;
        const F e = u9[3 + k] * u9[6 + k];
        const F st = sqrt_pos_t((F)1 - c * c);
        const F en = neg_log_t(e + Num<F>::MIN_POSITIVE, fm);
        F s_, n;
        sqrt_rsqrt_t(r2[k], &s_, &n);  // n = 1 / sqrt(r2)
        q[k][0] = en * (st * (xy[k][0] * n));
        q[k][1] = en * (st * (xy[k][1] * n));
        q[k][2] = en * c;
        q[k][3] = en;
    }
    conformal_transform<F, SORT, false, Cons3<F, false>::value>(q, e_total, p);
}

// ------------------------------------------------------------------------------ cuts
// The beam is along X: p(e-) = (-E/2, 0, 0, E/2) (evgen.rs:66-69), so p_gamma . p_e = -X E/2 and
// the common factor E/2 drops out of evcut.rs:52-62 and :80-92.
template <class F, bool SORT, bool LITERAL>
__device__ __forceinline__ typename MaskOf<F>::type keep_event(const F p[3][4], const PhysParams<F>& P) {
    // All comparisons are evaluated (no short circuit: they are independent, a chain of && would serialise them)
    typename MaskOf<F>::type ok;
    if constexpr (SORT) ok = !(p[2][3] < P.e_min);
    else ok = (p[0][3] >= P.e_min) & (p[1][3] >= P.e_min) & (p[2][3] >= P.e_min);  // event.rs:96-105 (min E < e_min rejects)
#pragma unroll
    for (int k = 0; k < 3; ++k) ok = ok & !(abs_t(p[k][0]) > P.acut * p[k][3]);
    if constexpr (LITERAL) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = a + 1; b < 3; ++b) {
                const F num = (p[a][0] * p[b][0] + p[a][1] * p[b][1]) + p[a][2] * p[b][2];
                ok = ok & !(num > P.bcut * (p[a][3] * p[b][3]));
            }
    } else {
        // p_i.p_j (3-vectors) = E_i E_j - (p_i + p_j)^2 / 2 = E_i E_j - e (e - 2 E_k) / 2 by momentum conservation
        // (the transform of evgen.rs:94-106 conserves the total 4-momentum (0,0,0,e) to rounding error), so
        // "cos > bcut" reads (1 - bcut) E_i E_j > e (e/2 - E_k): 3 FP64 instructions per pair instead of 6.
        // divided by e: (1 - bcut)/e E_i E_j + E_k > e/2 — a multiply, one FMA and a compare per pair
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = (k == 0) ? 1 : 0, j = (k == 2) ? 1 : 2;
            ok = ok & !(fma_t(p[i][3] * p[j][3], P.omb_e, p[k][3]) > P.he);
        }
    }
    if (uniform_positive(P.sincut)) {  // |n_x| < sincut |n|  (uniform branch; the default sincut is 0)
        const F nx = p[0][1] * p[1][2] - p[0][2] * p[1][1];
        const F ny = p[0][2] * p[1][0] - p[0][0] * p[1][2];
        const F nz = p[0][0] * p[1][1] - p[0][1] * p[1][0];
        const F nn = sqrt_t((nx * nx + ny * ny) + nz * nz);
        ok = ok & !(abs_t(nx) < P.sincut * nn);
    }
    return ok;
}

// ------------------------------------------------------------------ matrix elements
template <class F> struct Cplx {
    F re, im;
};
template <class F> __device__ __forceinline__ Cplx<F> cmul(Cplx<F> a, Cplx<F> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class F> __device__ __forceinline__ Cplx<F> csqr(Cplx<F> a) { return cmul(a, a); }
// acc + a b with four FMAs (a separate multiply and add would be six instructions)
template <class F> __device__ __forceinline__ Cplx<F> cmac(Cplx<F> acc, Cplx<F> a, Cplx<F> b) {
    return {fma_t(a.re, b.re, fma_t(-a.im, b.im, acc.re)), fma_t(a.re, b.im, fma_t(a.im, b.re, acc.im))};
}
template <class F> __device__ __forceinline__ Cplx<F> cadd(Cplx<F> a, Cplx<F> b) { return {a.re + b.re, a.im + b.im}; }
template <class F> __device__ __forceinline__ Cplx<F> csub(Cplx<F> a, Cplx<F> b) { return {a.re - b.re, a.im - b.im}; }
template <class F> __device__ __forceinline__ Cplx<F> cscale(Cplx<F> a, F s) { return {a.re * s, a.im * s}; }
template <class F> __device__ __forceinline__ Cplx<F> cdiv(Cplx<F> a, Cplx<F> b) {  // num-complex 0.4.4
    const F n = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / n, (a.im * b.re - a.re * b.im) / n};
}
template <class F> __device__ __forceinline__ F cnorm(Cplx<F> a) { return a.re * a.re + a.im * a.im; }

// Operation-for-operation transcription (spinor.rs:31-162, matelems.rs:53-86). m[5] = helicity sums.
template <class F> __device__ void me_literal(const F p[3][4], const PhysParams<F>& P, F m[5]) {
    F ev[5][4];
    const F half = P.e_total / (F)2;
    ev[0][0] = -half; ev[0][1] = 0; ev[0][2] = 0; ev[0][3] = half;
    ev[1][0] = half;  ev[1][1] = 0; ev[1][2] = 0; ev[1][3] = half;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) ev[2 + k][c] = p[k][c];
    F xx[5];
    Cplx<F> fx[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        xx[i] = sqrt_t(ev[i][3] + ev[i][2]);
        if (xx[i] > Num<F>::MIN_POSITIVE) fx[i] = {ev[i][0] / xx[i], ev[i][1] / xx[i]};
        else fx[i] = {sqrt_t((F)2 * ev[i][3]), (F)0};
    }
    Cplx<F> s[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) s[i][j] = csub(cscale(fx[i], xx[j]), cscale(fx[j], xx[i]));
    auto S = [&](int i, int j) { return s[i][j]; };
    auto T = [&](int i, int j) { return Cplx<F>{-s[i][j].re, s[i][j].im}; };
    const F mr8 = -Num<F>::RAC8;
    auto a_ppm = [&](int k1, int k2, int k3) {
        return cdiv(cmul(cscale(S(0, 1), mr8), csqr(S(0, k3))), cmul(cmul(cmul(S(0, k1), S(0, k2)), S(1, k1)), S(1, k2)));
    };
    auto a_pmm = [&](int k1, int k2, int k3) {
        return cdiv(cmul(cscale(T(0, 1), mr8), csqr(T(1, k1))), cmul(cmul(cmul(T(1, k2), T(1, k3)), T(0, k2)), T(0, k3)));
    };
    auto bp_ppm = [&](int k1, int k2, int k3) { return cmul(cscale(T(0, 1), mr8), csqr(cmul(T(k1, k2), S(k3, 0)))); };
    auto bp_pmm = [&](int k1, int k2, int k3) { return cmul(cscale(S(0, 1), mr8), csqr(cmul(T(k1, 1), S(k2, k3)))); };
    auto bm_ppp = [&](int k1, int k2, int k3) {
        return cmul(cscale(S(0, 1), mr8),
                    cadd(cadd(csqr(cmul(T(k1, k2), T(k3, 1))), csqr(cmul(T(k1, k3), T(k2, 1)))), csqr(cmul(T(k2, k3), T(k1, 1)))));
    };
    auto bm_mmm = [&](int k1, int k2, int k3) {
        return cmul(cscale(T(0, 1), mr8),
                    cadd(cadd(csqr(cmul(S(k1, 0), S(k2, k3))), csqr(cmul(S(k2, 0), S(k1, k3)))), csqr(cmul(S(k3, 0), S(k1, k2)))));
    };
    const Cplx<F> zero = {0, 0};
    Cplx<F> a[8] = {zero, a_pmm(4, 2, 3), a_pmm(3, 2, 4), a_ppm(3, 4, 2), a_pmm(2, 3, 4), a_ppm(2, 4, 3), a_ppm(2, 3, 4), zero};
    Cplx<F> bp[8] = {zero, bp_pmm(4, 2, 3), bp_pmm(3, 2, 4), bp_ppm(3, 4, 2), bp_pmm(2, 3, 4), bp_ppm(2, 4, 3), bp_ppm(2, 3, 4), zero};
    Cplx<F> bm[8] = {bm_mmm(2, 3, 4), zero, zero, zero, zero, zero, zero, bm_ppp(2, 3, 4)};
    m[0] = m[1] = m[2] = m[3] = m[4] = 0;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        const Cplx<F> A = cscale(a[h], P.g_a), Bp = cscale(bp[h], P.g_beta_p), Bm = cscale(bm[h], P.g_beta_m);
        const Cplx<F> mix = cmul(cscale(A, (F)2), Cplx<F>{Bp.re, -Bp.im});
        m[0] += cnorm(A);
        m[1] += cnorm(Bp);
        m[2] += cnorm(Bm);
        m[3] += mix.re;
        m[4] += mix.im;
    }
}

// Algebraically reduced form.  With e = e_total, h^2 = e/2 and, per photon k,
//   A_k = E_k + Z_k,  c_k = X_k + i Y_k,  g_k = c_k^2 / A_k  (= fx_k^2),
//   s_0k^2 = h^2 (A_k + 2 c_k + g_k),  s_1k^2 = h^2 (A_k - 2 c_k + g_k),  u_k = s_0k s_1k = -h^2 (A_k - g_k),
//   P_k = |s_0k|^2 = e (E_k + X_k),  Q_k = |s_1k|^2 = e (E_k - X_k),
// per pair  s_ij^2 = g_i A_j - 2 c_i c_j + g_j A_i,  R_ij = |s_ij|^2 = 2 (E_i E_j - p_i.p_j),
// and D = prod_k P_k Q_k = |u_0 u_1 u_2|^2, the helicity sums of matelems.rs:72-86 are
//   m0 = g_a^2  8 e^2 / D * sum_k P_k Q_k (P_k^2 + Q_k^2)
//   m1 = g_b+^2 8 e^2     * sum_k R_ij^2 (P_k^2 + Q_k^2)
//   m2 = g_b-^2 8 e^2     * ( |sum_k s_0k^2 s_ij^2|^2 + |sum_k s_1k^2 s_ij^2|^2 )
//   m3 + i m4 = -16 e^2 g_a g_b+ * sum_k ( P_k^2 W_k + Q_k^2 conj W_k ),  W_k = s_ij^2 u_k conj(U) / D, U = u_0 u_1 u_2
// (ij = the two photons other than k).  The common powers of h^2 are pulled out of the sums.
// photon along -Z (spinor.rs:42-46): xx = 0, fx = sqrt(2E)  =>  A = 0, c = 0, g = 2E.
// Reached in f32 (E + Z rounds to 0 about once per 1e7 events), never observed in f64.
#ifndef TP3_DEGENERATE_VOTE
#define TP3_DEGENERATE_VOTE 0   // 1: the fix-up behind a warp vote + uniform branch (measured 0.5 % slower than the predicated moves)
#endif
template <class F> __device__ __forceinline__ void degenerate_fix(F& A, Cplx<F>& g, F& X, F& Y, F E) {
    if (!(A > Num<F>::MIN_POSITIVE)) {
        A = 0;
        g = {E + E, (F)0};
        X = 0;
        Y = 0;
    }
}

// TP3_DEGENERATE_VOTE: apply the spinor.rs:42-46 fix-up only if some lane of the (converged part of the) warp needs it
// (in f64 none ever has), behind one vote + a uniform branch instead of predicated moves.  Measured slightly slower than
// the predicated form (profiles/r01_ab_variants.txt), so it is off; the packed f32 type always applies the fix per half.
template <class F> __device__ __forceinline__ bool any_degenerate(const F A[3]) {
    return __any_sync(__activemask(), !(A[0] > Num<F>::MIN_POSITIVE) | !(A[1] > Num<F>::MIN_POSITIVE) | !(A[2] > Num<F>::MIN_POSITIVE));
}

template <class F> __device__ __forceinline__ void me_fast(const F p[3][4], const PhysParams<F>& P, F m[5]) {
    const F e = P.e_total;
    F A[3], Ep[3], Em[3];       // A_k, E_k + X_k, E_k - X_k
    Cplx<F> g[3], tk[3], ub[3];  // g_k, t_k = A_k + g_k, A_k - g_k
    F cx[3], cy[3];                     // c_k = X_k + i Y_k
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const F X = p[k][0], Y = p[k][1];
        const F Z = p[k][2], E = p[k][3];
        Ep[k] = E + X;
        Em[k] = E - X;
        A[k] = E + Z;
        const F iA = rcp_t(A[k]);
        g[k].re = (X * X - Y * Y) * iA;
        g[k].im = ((X + X) * Y) * iA;
        cx[k] = X;
        cy[k] = Y;
    }
#if TP3_DEGENERATE_VOTE
    if (any_degenerate(A))
#endif
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) degenerate_fix(A[k], g[k], cx[k], cy[k], p[k][3]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tk[k] = {A[k] + g[k].re, g[k].im};
        ub[k] = {A[k] - g[k].re, -g[k].im};
    }
    // pairs, indexed by the photon k they exclude: (i,j) = (1,2), (0,2), (0,1)
    Cplx<F> s2[3];
    F R[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = (k == 0) ? 1 : 0, j = (k == 2) ? 1 : 2;
        const Cplx<F> cc = cmul(Cplx<F>{cx[i], cy[i]}, Cplx<F>{cx[j], cy[j]});
        s2[k].re = g[i].re * A[j] + g[j].re * A[i] - (cc.re + cc.re);
        s2[k].im = g[i].im * A[j] + g[j].im * A[i] - (cc.im + cc.im);
        // |s_ij|^2 = 2 p_i.p_j = (p_i + p_j)^2 = (p_beams - p_k)^2 = e (e - 2 E_k): momentum conservation,
        // which RAMBO's conformal transform guarantees to rounding error (evgen.rs:94-106)
        R[k] = e * (e - (p[k][3] + p[k][3]));
    }
    // T_k = (E+X)(E-X), PQ2_k = (E+X)^2 + (E-X)^2, DPQ_k = (E+X)^2 - (E-X)^2   (units of e^2 pulled out)
    F T[3], S2[3], D2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        T[k] = Ep[k] * Em[k];
        const F a = Ep[k] * Ep[k], b = Em[k] * Em[k];
        S2[k] = a + b;
        D2[k] = a - b;
    }
    const F Dn = (T[0] * T[1]) * T[2];  // D = e^6 * Dn
    const F iDn = rcp_t(Dn);
    // m0 = g_a^2 * 8 e^2 * e^4 sum(T S2) / (e^6 Dn)
    m[0] = P.k_m0 * ((T[0] * S2[0] + T[1] * S2[1] + T[2] * S2[2]) * iDn);
    // m1 = g_b+^2 * 8 e^2 * e^2 sum(R^2 S2)
    m[1] = P.k_m1 * (R[0] * R[0] * S2[0] + R[1] * R[1] * S2[1] + R[2] * R[2] * S2[2]);
    // m2: s_0k^2 = (e/2)(t_k + 2 c_k), s_1k^2 = (e/2)(t_k - 2 c_k) with t_k = A_k + g_k, and
    // |Sa + Sb|^2 + |Sa - Sb|^2 = 2 (|Sa|^2 + |Sb|^2) for Sa = sum t_k s2_k, Sb = sum 2 c_k s2_k
    Cplx<F> Sa = cmul(tk[0], s2[0]), Sb = cmul(Cplx<F>{cx[0], cy[0]}, s2[0]);
#pragma unroll
    for (int k = 1; k < 3; ++k) {
        Sa = cmac(Sa, tk[k], s2[k]);
        Sb = cmac(Sb, Cplx<F>{cx[k], cy[k]}, s2[k]);
    }
    m[2] = P.k_m2 * (cnorm(Sa) + (F)4 * cnorm(Sb));
    // mixed: u_k = -(e/2) ub_k, U = -(e/2)^3 Ub, W_k = s2_k u_k conj(U) / D = s2_k ub_k conj(Ub) (e/2)^4 / (e^6 Dn)
    const Cplx<F> Ub = cmul(cmul(ub[0], ub[1]), ub[2]);
    const Cplx<F> Uc = {Ub.re, -Ub.im};
    F re = 0, im = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const Cplx<F> w = cmul(cmul(s2[k], ub[k]), Uc);
        re += S2[k] * w.re;
        im += D2[k] * w.im;
    }
    // P_k^2 = e^2 (E+X)^2: total factor -16 e^2 g_a g_b+ * e^2 * (e/2)^4 / e^6 = -g_a g_b+ e^2
    const F cm = P.k_mix * iDn;
    m[3] = cm * re;
    m[4] = cm * im;
}

}  // namespace tp3

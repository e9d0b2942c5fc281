// Two single-precision events per lane: Blackwell's packed FP32 arithmetic (add/sub/mul/fma.f32x2, SASS FADD2 /
// FMUL2 / FFMA2: one issue slot for two lanes' worth of work).
//
// The f32 kernel is issue bound, not pipe bound (profiles/r01_final_fast_f32_xoshiro.txt: 545 issue slots per event,
// 52 % of them scalar FMUL / FFMA / FADD), so halving the arithmetic issue slots is the lever.  `f2` is a number type
// the physics templates (physics.cuh) can be instantiated with: the operators are inline PTX WITHOUT a rounding
// modifier, which lets ptxas contract mul + add/sub into FFMA2 (with operand negation) exactly as it contracts the
// scalar code, so both halves compute what the scalar f32 kernel computes.  Transcendentals go through the SFU one
// half at a time.
#pragma once

#include <cstdint>
#include <cstring>

#include "physics.cuh"

namespace tp3 {

struct f2 {
    unsigned long long v;  // low word = event 0, high word = event 1
    f2() = default;
    __host__ __device__ __forceinline__ f2(float x, float y) {
#ifdef __CUDA_ARCH__
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(y));
#else
        uint32_t a, b;
        std::memcpy(&a, &x, 4);
        std::memcpy(&b, &y, 4);
        v = (unsigned long long)b << 32 | a;
#endif
    }
    __host__ __device__ __forceinline__ f2(float s) : f2(s, s) {}
    __host__ __device__ __forceinline__ f2(double s) : f2((float)s, (float)s) {}
    __host__ __device__ __forceinline__ f2(int s) : f2((float)s, (float)s) {}
    __device__ __forceinline__ float lo() const {
        float x, y;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
        return x;
    }
    __device__ __forceinline__ float hi() const {
        float x, y;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
        return y;
    }
};

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(-a.lo(), -a.hi()); }  // folds into an operand modifier
__device__ __forceinline__ f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
__device__ __forceinline__ f2 operator+(f2 a, float b) { return a + f2(b); }
__device__ __forceinline__ f2 fma_t(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 abs_t(f2 a) { return f2(fabsf(a.lo()), fabsf(a.hi())); }

// comparison results of the two events
struct m2 {
    bool x, y;
};
__device__ __forceinline__ m2 operator&(m2 a, m2 b) { return {a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ m2 operator!(m2 a) { return {!a.x, !a.y}; }
__device__ __forceinline__ m2 operator>(f2 a, f2 b) { return {a.lo() > b.lo(), a.hi() > b.hi()}; }
__device__ __forceinline__ m2 operator<(f2 a, f2 b) { return {a.lo() < b.lo(), a.hi() < b.hi()}; }
__device__ __forceinline__ m2 operator>=(f2 a, f2 b) { return {a.lo() >= b.lo(), a.hi() >= b.hi()}; }
__device__ __forceinline__ f2 select(m2 m, f2 a, f2 b) { return f2(m.x ? a.lo() : b.lo(), m.y ? a.hi() : b.hi()); }

template <> struct MaskOf<f2> { using type = m2; };
template <> struct Num<f2> {
    static constexpr float MIN_POSITIVE = Num<float>::MIN_POSITIVE;
};
template <> struct PhiScale<f2> { static constexpr float value = PhiScale<float>::value; };

__device__ __forceinline__ bool uniform_positive(f2 x) { return x.lo() > 0.0f; }  // both halves hold the same parameter

// SFU functions, one half at a time
__device__ __forceinline__ f2 rcp_t(f2 x) { return f2(rcp_t(x.lo()), rcp_t(x.hi())); }
__device__ __forceinline__ f2 sqrt_t(f2 x) { return f2(sqrt_t(x.lo()), sqrt_t(x.hi())); }
__device__ __forceinline__ f2 sqrt_pos_t(f2 x) { return f2(sqrt_pos_t(x.lo()), sqrt_pos_t(x.hi())); }
__device__ __forceinline__ f2 neg_log_t(f2 x, const FastMath fm) { return f2(neg_log_t(x.lo(), fm), neg_log_t(x.hi(), fm)); }
__device__ __forceinline__ void sqrt_rsqrt_t(f2 x, f2* s, f2* rs) {
    float s0, r0, s1, r1;
    sqrt_rsqrt_t(x.lo(), &s0, &r0);
    sqrt_rsqrt_t(x.hi(), &s1, &r1);
    *s = f2(s0, s1);
    *rs = f2(r0, r1);
}
__device__ __forceinline__ void sincos_scaled_t(f2 t, const FastMath, f2* s, f2* c) {
    // the argument of both halves in packed arithmetic (see the scalar forms in physics.cuh: same operations, same bits);
    // the inline-PTX operators are opaque to the compiler, so (t + M) - M is not simplified
    float s0, c0, s1, c1;
#if TP3_F32_SINCOS_DIRECT
    const f2 x = (f2(2.0f) - t) * f2(kHalfPiF);
    s0 = mufu_sin_f(x.lo());
    s1 = mufu_sin_f(x.hi());
    c0 = mufu_cos_f(x.lo());
    c1 = mufu_cos_f(x.hi());
    *s = f2(s0, s1);
    *c = -f2(c0, c1);
#else
    const f2 magic(kRintMagic);
    const f2 tm = t + magic;
    const f2 qf = tm - magic;
    const f2 x = (t - qf) * f2(kHalfPiF);
    sincos_quadrant(x.lo(), __float_as_int(tm.lo()), &s0, &c0);
    sincos_quadrant(x.hi(), __float_as_int(tm.hi()), &s1, &c1);
    *s = f2(s0, s1);
    *c = f2(c0, c1);
#endif
}

// me_fast's photon-along-(-Z) case (spinor.rs:42-46), per half
__device__ __forceinline__ bool any_degenerate(const f2*) { return true; }  // the fix is applied per half, branch-free
__device__ __forceinline__ void degenerate_fix(f2& A, Cplx<f2>& g, f2& X, f2& Y, f2 E) {
    const m2 ok = A > f2(Num<float>::MIN_POSITIVE);
    if (ok.x && ok.y) return;  // (about one event in 1e7)
    const f2 zero(0.0f);
    A = select(ok, A, zero);
    g.re = select(ok, g.re, E + E);
    g.im = select(ok, g.im, zero);
    X = select(ok, X, zero);
    Y = select(ok, Y, zero);
}

}  // namespace tp3

// `faster-evgen` on the GPU (src/evgen.rs:143-173, 221-249).
//
// Under this feature an event consumes a DATA-DEPENDENT number of random numbers: 9, then 6, then 2
// more for every re-roll of a point that fell outside the unit disc; and RANF serves every request
// from a single 55-number round, discarding the rest of the round when the request does not fit
// (ranf.rs:87-92).  The position of event e in the stream therefore depends on every earlier event of
// the batch, which is exactly why the reference's reproducible multi-threaded mode re-runs the
// rejection loop on its scheduler thread (evgen.rs:257-267, "Bottleneck!" in VALIDATION.md:46-48).
//
// To reproduce the reference's integers exactly, the stream is walked sequentially from known start states:
// ONE THREAD = ONE BATCH (32 consecutive batches per warp) or, when the scan also supplies the states at the 32
// lane boundaries inside every batch (fe_scan.cuh), ONE THREAD = 313 EVENTS; private generator state per thread
// (RANF: 56 words in shared memory, column layout, so a warp's accesses are conflict free; xoshiro:
// registers).  Batch start states come from the scheduler: the host pre-advances its generator batch by
// batch exactly like the reference's scheduler thread does, or, under faster-threading, every batch is
// re-seeded / jump()ed on the device.  The physics is the same gen -> cuts -> matrix-element code as the
// main kernel; the rejection test itself is evaluated without FMA contraction so that every accept /
// re-roll decision is the reference's.
#pragma once

#include "kernels.cuh"

namespace tp3 {

constexpr int kFeThreads = 128;

struct FeArgs {
    uint64_t first_batch;
    uint64_t n_batches;
    uint32_t last_batch_len;
    uint32_t jump_seeding;
    uint32_t split;                // 1: one thread per batch; 32: one warp per batch, lane l owns events [313 l, 313 (l + 1))
    const uint32_t* ranf_states;   // [n_batches * split][57]: numbers[0..55] + index (sequential mode)
    const uint64_t* xo_states;     // [n_batches][4]
    tp3_acc* out;
    int32_t ranf_seed;
};

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// RANF with the reference's request semantics, state in shared memory column `col`
template <class F> struct RanfThread {
    uint32_t* st;  // &numbers[0][thread], stride kFeThreads words
    int index;
    __device__ __forceinline__ uint32_t& n(int i) { return st[i * kFeThreads]; }
    __device__ void reset() {  // ranf.rs:106-119
        for (int i = 1; i < 25; ++i) n(i) = ranf_sub(n(i), n(i + 31));
        for (int i = 25; i < 56; ++i) n(i) = ranf_sub(n(i), n(i - 24));
    }
    __device__ void seed(int32_t s) {  // ranf.rs:36-66
        for (int i = 0; i < 56; ++i) n(i) = 0;
        n(55) = (uint32_t)s;
        int j = s, k = 1;
        for (int i = 1; i < 55; ++i) {
            const int ii = (21 * i) % 55;
            n(ii) = (uint32_t)k;
            const int nk = j - k;
            j = k;
            k = nk < 0 ? nk + (int)kRanfMod : nk;
        }
        for (int r = 0; r < 10; ++r) reset();
        index = 55;
    }
    template <int N> __device__ __forceinline__ void take(F out[N]) {  // ranf.rs:78-102
        if (index < N) {
            reset();
            index = 55;
        }
        index -= N;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t w = n(index + 1 + i);
            out[i] = sizeof(F) == 8 ? (F)((double)(int)w * 1e-9) : (F)((float)(int)w * 1e-9f);
        }
    }
};

template <class F> struct XoThread;
template <> struct XoThread<double> {
    Xoshiro256Lane g;
    template <int N> __device__ __forceinline__ void take(double out[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = to_uniform_xo(g.next());
    }
};
template <> struct XoThread<float> {
    Xoshiro128Lane g;
    template <int N> __device__ __forceinline__ void take(float out[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = to_uniform_xo(g.next());
    }
};

template <class F, class Gen>
__device__ __forceinline__ void fe_simulate(Gen& gen, int n_ev, const PhysParams<F>& P, const FastMathSmem* fm, bool warp_reduce,
                                            tp3_acc* out) {
    LaneAcc<F> acc;
    acc.clear();
    for (int ev = 0; ev < n_ev; ++ev) {
        F u9[9], v6[6];
        gen.template take<9>(u9);
        gen.template take<6>(v6);
        F xy[3][2], r2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // from_iterator fills the 3x2 matrix column-major (evgen.rs:223-225)
            xy[k][0] = (F)2 * v6[k] - (F)1;
            xy[k][1] = (F)2 * v6[3 + k] - (F)1;
            r2[k] = add_rn(mul_rn(xy[k][0], xy[k][0]), mul_rn(xy[k][1], xy[k][1]));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // evgen.rs:231-241 (the `< MIN_POSITIVE^2` test is dead: the constant underflows to 0)
            while (r2[k] > (F)1) {
                F w[2];
                gen.template take<2>(w);
                xy[k][0] = (F)2 * w[0] - (F)1;
                xy[k][1] = (F)2 * w[1] - (F)1;
                r2[k] = add_rn(mul_rn(xy[k][0], xy[k][0]), mul_rn(xy[k][1], xy[k][1]));
            }
        }
        F p[3][4];
        gen_event_faster<F, false>(u9, xy, r2, P.e_total, fm, p);  // the sums do not depend on the photon order
        if (keep_event<F, false, false>(p, P)) {
            F m[5];
            me_fast<F>(p, P, m);
            acc.integrate(m, P.sigma_contribs);
        }
    }
    if (warp_reduce) {  // warp-uniform: the 32 lanes hold the parts of one batch (same tree as simulate_kernel)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                acc.spm2[k] += shfl_xor_t(acc.spm2[k], off);
                acc.vars[k] += shfl_xor_t(acc.vars[k], off);
            }
            acc.sigma += shfl_xor_t(acc.sigma, off);
            acc.variance += shfl_xor_t(acc.variance, off);
            acc.selected += __shfl_xor_sync(0xffffffffu, acc.selected, off);
        }
        if (threadIdx.x & 31) return;
    }
    out->selected_events = acc.selected;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        out->spm2[k] = (double)acc.spm2[k];
        out->vars[k] = (double)acc.vars[k];
    }
    out->sigma = (double)acc.sigma;
    out->variance = (double)acc.variance;
}

template <class F, int RNG>
__global__ void __launch_bounds__(kFeThreads) faster_evgen_kernel(const FeArgs a, const PhysParams<F> P) {
    __shared__ FastMathSmem fm;
    __shared__ uint32_t ranf_state[RNG == RNG_RANF ? 56 * kFeThreads : 1];
    fastmath_load(&fm);
    __syncthreads();
    const uint64_t unit = (uint64_t)blockIdx.x * kFeThreads + threadIdx.x;
    const bool split = a.split == 32;
    const uint64_t slot = split ? unit >> 5 : unit;
    if (slot >= a.n_batches) return;  // whole warps when split
    const int n_batch = (slot + 1 == a.n_batches) ? (int)a.last_batch_len : kBatch;
    const int part = split ? (int)(unit & 31) : 0;
    const int n_ev = split ? max(0, min(n_batch - part * kLaneEvents, kLaneEvents)) : n_batch;
    const uint64_t batch = a.first_batch + slot;
    if (RNG == RNG_RANF) {
        RanfThread<F> gen;
        gen.st = ranf_state + threadIdx.x;
        if (a.jump_seeding) {
            gen.seed((int32_t)((uint32_t)a.ranf_seed + 123456u * (uint32_t)batch));
        } else {
            const uint32_t* s = a.ranf_states + unit * 57;
            for (int i = 0; i < 56; ++i) gen.n(i) = s[i];
            gen.index = (int)s[56];
        }
        fe_simulate<F, RanfThread<F>>(gen, n_ev, P, &fm, split, a.out + slot);
    } else {
        XoThread<F> gen;
        const uint64_t* s = a.xo_states + 4 * slot;
        gen.g.s0 = (decltype(gen.g.s0))s[0];
        gen.g.s1 = (decltype(gen.g.s0))s[1];
        gen.g.s2 = (decltype(gen.g.s0))s[2];
        gen.g.s3 = (decltype(gen.g.s0))s[3];
        fe_simulate<F, XoThread<F>>(gen, n_ev, P, &fm, split, a.out + slot);
    }
}

}  // namespace tp3

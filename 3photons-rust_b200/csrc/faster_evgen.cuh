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
// (RANF: 55 words in shared memory, one row per lane with an odd stride, so a warp's accesses are conflict free;
// xoshiro: registers).  The lanes of a warp issue their requests in lock-step so that refills are shared work.  Batch start states come from the scheduler: the host pre-advances its generator batch by
// batch exactly like the reference's scheduler thread does, or, under faster-threading, every batch is
// re-seeded / jump()ed on the device.  The physics is the same gen -> cuts -> matrix-element code as the
// main kernel; the rejection test itself is evaluated without FMA contraction so that every accept /
// re-roll decision is the reference's.
#pragma once

#include "kernels.cuh"

namespace tp3 {

#ifndef TP3_FE_THREADS
#define TP3_FE_THREADS 128
#endif
constexpr int kFeThreads = TP3_FE_THREADS;

struct FeArgs {
    uint64_t first_batch;
    uint64_t n_batches;
    uint32_t last_batch_len;
    uint32_t jump_seeding;
    uint32_t split;                // 1: one thread per batch; 32: one warp per batch, lane l owns events [313 l, 313 (l + 1))
    const uint32_t* ranf_states;   // [n_batches * split][57]: numbers[0..55] + index (sequential mode)
    const uint64_t* xo_states;     // [n_batches * split][4]
    tp3_acc* out;
    int32_t ranf_seed;
};

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

constexpr int kFeRow = kRanfLag + 2;  // words per private generator: row[k] = numbers[k + 1]; odd stride, so that the
                                      // lanes' private accesses AND a warp's cooperative access to one row are conflict free

// RANF with the reference's request semantics (ranf.rs:78-102), one private generator per lane.
// All requests are issued WARP-SYNCHRONOUSLY (every lane calls take() together, with its own `want`), so that
// a refill (ranf.rs:106-119) is not 55 dependent subtractions by one lane while 31 wait, but one cooperative step
// of the whole warp per lane that needs it: 55 new numbers from <= 4 old ones each (ranf_next_slot).
template <class F> struct RanfLane {
    uint32_t* row;        // this lane's generator
    uint32_t* warp_rows;  // lane 0's
    int index;
    __device__ __forceinline__ uint32_t& n(int i) { return row[i - 1]; }  // numbers[i], i = 1..55
    __device__ void reset_private() {  // ranf.rs:106-119 (seeding only)
        for (int i = 1; i < 25; ++i) n(i) = ranf_sub(n(i), n(i + 31));
        for (int i = 25; i < 56; ++i) n(i) = ranf_sub(n(i), n(i - 24));
    }
    __device__ void seed(int32_t s) {  // ranf.rs:36-66
        for (int i = 1; i < 56; ++i) n(i) = 0;
        n(55) = (uint32_t)s;
        int j = s, k = 1;
        for (int i = 1; i < 55; ++i) {
            const int ii = (21 * i) % 55;
            n(ii) = (uint32_t)k;
            const int nk = j - k;
            j = k;
            k = nk < 0 ? nk + (int)kRanfMod : nk;
        }
        for (int r = 0; r < 10; ++r) reset_private();
        index = 55;
    }
    __device__ __forceinline__ void load(const uint32_t* s) {  // numbers[0..55] + index
        for (int i = 1; i < 56; ++i) n(i) = s[i];
        index = (int)s[56];
    }
    template <int N> __device__ __forceinline__ void take(bool want, F out[N], int lane) {
        const bool refill = want && index < N;
        unsigned m = __ballot_sync(0xffffffffu, refill);
        // Lane l < 24 owns slots l + 1, l + 25 and (l < 7) l + 49 of the row being refilled: numbers[i] -= numbers[i + 31]
        // for i <= 24, then numbers[i] -= NEW numbers[i - 24] — a chain that stays inside the lane (as in RanfWarpStream).
        const int l = lane < 24 ? lane : 23, l7 = lane < 7 ? lane : 6;
        while (m) {  // warp-uniform
            uint32_t* r = warp_rows + (__ffs(m) - 1) * kFeRow;
            m &= m - 1;
            const uint32_t o1 = r[l], o32 = r[l + 31], o25 = r[l + 24], o49 = r[l7 + 48];
            const uint32_t va = ranf_sub(o1, o32), vb = ranf_sub(o25, va), vc = ranf_sub(o49, vb);
            __syncwarp();
            if (lane < 24) {
                r[l] = va;
                r[l + 24] = vb;
            }
            if (lane < 7) r[l7 + 48] = vc;
            __syncwarp();
        }
        if (refill) index = kRanfLag;
        if (want) index -= N;
        const int at = want ? index : 0;  // idle lanes read valid words and ignore them
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t w = row[at + i];
            out[i] = sizeof(F) == 8 ? (F)u32_times(w, 1e-9) : (F)((float)(int)w * 1e-9f);
        }
    }
};

template <class F> struct XoLane;
template <> struct XoLane<double> {
    Xoshiro256Lane g;
    template <int N> __device__ __forceinline__ void take(bool want, double out[N], int) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = 0.5;
        if (want) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = to_uniform_xo(g.next());
        }
    }
};
template <> struct XoLane<float> {
    Xoshiro128Lane g;
    template <int N> __device__ __forceinline__ void take(bool want, float out[N], int) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = 0.5f;
        if (want) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = to_uniform_xo(g.next());
        }
    }
};

// The events of one lane, in lock-step with the other lanes of the warp (which walk other streams): 9 numbers,
// 6 numbers, then one 2-number re-roll per step while any lane still has a point outside the unit disc.
template <class F, class Gen>
__device__ __forceinline__ void fe_simulate(Gen& gen, int n_ev, const PhysParams<F>& P, const FastMath fm, bool warp_reduce,
                                            tp3_acc* out) {
    const int lane = threadIdx.x & 31;
    LaneAcc<F> acc;
    acc.clear();
    const int n_max = __reduce_max_sync(0xffffffffu, n_ev);
    for (int ev = 0; ev < n_max; ++ev) {
        const bool act = ev < n_ev;
        F u9[9], v6[6];
        gen.template take<9>(act, u9, lane);
        gen.template take<6>(act, v6, lane);
        F xy[3][2], r2[3];
        bool pend[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // from_iterator fills the 3x2 matrix column-major (evgen.rs:223-225)
            xy[k][0] = (F)2 * v6[k] - (F)1;
            xy[k][1] = (F)2 * v6[3 + k] - (F)1;
            r2[k] = add_rn(mul_rn(xy[k][0], xy[k][0]), mul_rn(xy[k][1], xy[k][1]));
            pend[k] = act && r2[k] > (F)1;  // evgen.rs:231-241 (the `< MIN_POSITIVE^2` test is dead: the constant underflows to 0)
        }
        while (__any_sync(0xffffffffu, pend[0] | pend[1] | pend[2])) {
            const bool need = pend[0] | pend[1] | pend[2];
            F w[2];
            gen.template take<2>(need, w, lane);
            const F x = (F)2 * w[0] - (F)1, y = (F)2 * w[1] - (F)1;
            const F rr = add_rn(mul_rn(x, x), mul_rn(y, y));
            const bool still = rr > (F)1;
            // the points are settled in order: 0 until it is inside, then 1, then 2
            const bool s0 = pend[0], s1 = !s0 && pend[1], s2 = !s0 && !s1 && pend[2];
            if (s0) { xy[0][0] = x; xy[0][1] = y; r2[0] = rr; pend[0] = still; }
            if (s1) { xy[1][0] = x; xy[1][1] = y; r2[1] = rr; pend[1] = still; }
            if (s2) { xy[2][0] = x; xy[2][1] = y; r2[2] = rr; pend[2] = still; }
        }
        if (act) {
            F p[3][4];
            gen_event_faster<F, false>(u9, xy, r2, P.e_total, fm, p);  // the sums do not depend on the photon order
            if (keep_event<F, false, false>(p, P)) {
                F m[5];
                me_fast<F>(p, P, m);
                acc.integrate(m, P.sigma_contribs);
            }
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
        if (lane) return;
    }
    if (!out) return;  // lanes past the end of the launch only keep the warp in step
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
__global__ void __launch_bounds__(kFeThreads, 512 / kFeThreads) faster_evgen_kernel(const FeArgs a, const PhysParams<F> P) {
    __shared__ FastMathSmem fm;
    __shared__ uint32_t ranf_state[RNG == RNG_RANF ? kFeRow * kFeThreads : 1];
    fastmath_load<kFeThreads>(&fm);
    __syncthreads();
    const uint64_t unit = (uint64_t)blockIdx.x * kFeThreads + threadIdx.x;
    const bool split = a.split == 32;
    const uint64_t slot = split ? unit >> 5 : unit;
    // whole warps leave together; otherwise every lane stays (the requests are warp-synchronous), idle ones with no events
    if ((split ? slot : unit - (threadIdx.x & 31)) >= a.n_batches) return;
    const bool live = slot < a.n_batches;
    const int n_batch = !live ? 0 : (slot + 1 == a.n_batches) ? (int)a.last_batch_len : kBatch;
    const int part = split ? (int)(unit & 31) : 0;
    const int n_ev = split ? max(0, min(n_batch - part * kLaneEvents, kLaneEvents)) : n_batch;
    const uint64_t batch = a.first_batch + slot;
    tp3_acc* out = live ? a.out + slot : nullptr;
    if (RNG == RNG_RANF) {
        RanfLane<F> gen;
        gen.warp_rows = ranf_state + (threadIdx.x & ~31) * kFeRow;
        gen.row = ranf_state + threadIdx.x * kFeRow;
        gen.index = kRanfLag;
        if (!live) {
            for (int i = 1; i < 56; ++i) gen.n(i) = 0;
        } else if (a.jump_seeding) {
            gen.seed((int32_t)((uint32_t)a.ranf_seed + 123456u * (uint32_t)batch));
        } else {
            gen.load(a.ranf_states + unit * 57);
        }
        __syncwarp();
        fe_simulate<F, RanfLane<F>>(gen, n_ev, P, FastMath{&fm, &P.fc}, split, out);
    } else {
        XoLane<F> gen;
        const uint64_t* s = a.xo_states + 4 * (live ? (split ? unit : slot) : 0);  // split: [n_batches * 32][4], batch-major
        gen.g.s0 = (decltype(gen.g.s0))s[0];
        gen.g.s1 = (decltype(gen.g.s0))s[1];
        gen.g.s2 = (decltype(gen.g.s0))s[2];
        gen.g.s3 = (decltype(gen.g.s0))s[3];
        fe_simulate<F, XoLane<F>>(gen, n_ev, P, FastMath{&fm, &P.fc}, split, out);
    }
}

}  // namespace tp3

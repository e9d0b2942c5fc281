// `faster-evgen` with RANF, second design: WALK the stream once, PHYSICS on event records.
//
// Under this feature the position of an event in the random stream depends on every earlier event (9 numbers, 6 numbers,
// 2 more per re-roll of a point outside the unit disc, and RANF discards the rest of a 55-number round when a request does
// not fit: evgen.rs:143-173,221-249, ranf.rs:78-102).  The reference's reproducible scheduler therefore walks the stream
// sequentially (evgen.rs:257-267).  Here:
//
//   1. fe_walk_kernel     ONE LANE = ONE SEGMENT of consecutive rounds, private generator per lane (started by jump-ahead).
//        A lane starts `warm` rounds BEFORE its segment with all nine possible consumer states (fe_scan.cuh) and follows
//        only the states that are still possible; every request that does not fit puts every walk back at a round start,
//        so the nine walks coincide after 2.2 rounds on average (never more than 17 in 4e5 trials; `warm` = 24 leaves
//        ~1e-8 per segment, and a segment that has not collapsed is simply redone from its predecessor's exit state).
//        From its segment start on the lane's state is the TRUE one: it serves the requests exactly as the reference does
//        and writes, for every event that starts in its segment, a 64-byte RECORD with the 15 integers the event keeps:
//        the nine numbers of random_array::<9>() and the (x, y) numbers of the three accepted points.
//        The accept / re-roll decision is EXACT: with X = 2n - 1e9 the reference's floating-point test r^2 > 1 can only
//        differ from the integer test X^2 + Y^2 > 1e18 when the two sides are within the rounding error of the
//        floating-point expression (|X^2 + Y^2 - 1e18| <= 2665 in f64, 1.9e12 in f32, see FeTol); outside that band the
//        integer comparison decides, inside it the reference's own expression is evaluated (no FMA contraction).
//   2. (host)             prefix sum of the per-segment event counts: the absolute index of every segment's first event.
//   3. fe_physics_kernel  ONE LANE = ONE EVENT: records are fetched with cp.async one warp iteration ahead, then the same
//        generation -> cuts -> survivor compaction -> matrix elements as the fused default kernel.  A batch is cut into
//        four fixed parts of 2500 events (one warp each, partial sums folded in part order), so that a pass of a few
//        thousand batches still balances over the 2368 resident warps; the result does not depend on the launch shape.
//
// Cost (profiles/r02_fe_*): the walk is integer work (generator + ~3.8 disc tests per event), the physics kernel has no
// generator and no rejection loop left.  HBM carries 64 B per event each way, far below the bandwidth the FP64-bound
// physics kernel leaves idle.
#pragma once

#include <cstddef>

#include "fe_scan.cuh"

namespace tp3 {

struct alignas(32) FeRecord {
    // w[0..8]: random_array::<9>() in the reference's order (evgen.rs:149-153: column-major 3x3 -> cos_theta of the three
    // photons, then the two factors of each exp(-E)); w[9 + 2k], w[10 + 2k]: the numbers behind (x, y) of accepted point k.
    uint32_t w[16];
};

constexpr int kFeParts = 4;                       // fixed parts of a batch in the physics kernel
constexpr int kFePartLen = kBatch / kFeParts;     // 2500 events
constexpr int kFeWarm = 24;                       // warm-up rounds before a segment (see the head of this file)
constexpr int kFeWarmFirst = 256;                 // ... before the first segment of a pass, which has no predecessor to redo it from
constexpr int kFeSlotsPerRound = 4;               // an event takes >= 15 of the 55 numbers of a round: at most 4 start in one

// The floating-point test of the reference may disagree with the exact one only for |X^2 + Y^2 - 1e18| <= 24 eps 1e18 = 2665
// in f64 and 32 eps 1e18 = 1.9e12 in f32.  The top bits of S = X^2 + Y^2 settle every case outside a
// band that contains that interval with a factor >= 2 to spare:
//   f64: S >> 32 against 1e18 >> 32: undecided only if the high words are equal, |S - 1e18| < 2^31.4 (probability 2e-9)
//   f32: S >> 43 against (1e18 -+ 2^42) >> 43 = 113686 / 113687 (probability 9e-6)
// and only then is the reference's own floating-point expression evaluated (out of line: it must not cost the common path).
template <class F> struct FeBand;
template <> struct FeBand<double> { static constexpr int shift = 32; static constexpr uint32_t lo = 0x0DE0B6B3u, hi = 0x0DE0B6B3u; };
template <> struct FeBand<float> { static constexpr int shift = 43; static constexpr uint32_t lo = 113686u, hi = 113687u; };
static_assert((1000000000000000000ull >> 32) == 0x0DE0B6B3ull, "high word of 1e18");
static_assert(((1000000000000000000ull - (1ull << 42)) >> 43) == 113686ull && ((1000000000000000000ull + (1ull << 42)) >> 43) == 113687ull, "f32 band");

template <class F> __device__ __noinline__ bool fe_outside_reference(uint32_t a, uint32_t b) { return fe_outside<F>(a, b); }

template <class F> __device__ __forceinline__ bool fe_outside_exact(uint32_t a, uint32_t b) {
    const int x = (int)(a + a) - 1000000000, y = (int)(b + b) - 1000000000;  // 2n - 1e9, in (-1e9, 1e9)
    // high words of the two squares: their sum is the high word of S = x^2 + y^2 or one less (the carry of the low words)
    const uint32_t top = ((uint32_t)__mulhi(x, x) + (uint32_t)__mulhi(y, y)) >> (FeBand<F>::shift - 32);
    if (__builtin_expect(top + 1u >= FeBand<F>::lo && top <= FeBand<F>::hi, 0)) return fe_outside_reference<F>(a, b);
    return top > FeBand<F>::hi;
}

// Branch-free form for the hot loops: the decision where it is certain, `amb` raised where the reference's expression has to
// decide (the caller then redoes the test with fe_outside_exact, once, out of the common path).
template <class F> __device__ __forceinline__ bool fe_outside_fast(uint32_t a, uint32_t b, bool& amb) {
    constexpr uint32_t width = FeBand<F>::hi - FeBand<F>::lo + 1u;
    const int x = (int)(a + a) - 1000000000, y = (int)(b + b) - 1000000000;
    const uint32_t top = ((uint32_t)__mulhi(x, x) + (uint32_t)__mulhi(y, y)) >> (FeBand<F>::shift - 32);
    const uint32_t d = top - (FeBand<F>::lo - 1u);  // wraps (negative) below the band: inside; above `width`: outside
    amb |= d <= width;
    return (int)d > (int)width;
}

// State after the six-number request (three points, flags "outside") and after an accepted re-roll: the tables of
// fe_scan.cuh (fe_walk_entry), one nibble per case.
__device__ __forceinline__ int fe_after_six(bool n0, bool n1, bool n2) {
    constexpr uint32_t kAfterSix = 0x0u | (2u << 4) | (6u << 8) | (4u << 12) | (8u << 16) | (3u << 20) | (7u << 24) | (5u << 28);
    return (int)((kAfterSix >> (4 * ((int)n0 | (int)n1 << 1 | (int)n2 << 2))) & 15u);
}
__device__ __forceinline__ int fe_after_roll(int s) {  // s in 2..8: the first point still outside is now inside
    constexpr unsigned long long kAfterRoll = (0ull << 8) | (8ull << 12) | (6ull << 16) | (7ull << 20) | (0ull << 24) | (8ull << 28) | (0ull << 32);
    return (int)((kAfterRoll >> (4 * s)) & 15ull);
}

// numbers[i + 1] of this lane's private generator (see FeWalkSmem: lane l owns bank l)
#define FE_ROW(i) row[(i) << 5]

// One round from entry state s with the exact test, nothing recorded (warm-up): exit state, events started.
template <class F> __device__ __forceinline__ int fe_walk_round_exact(const uint32_t* row, int s, int& count) {
    int idx = kRanfLag;
    count = 0;
    for (;;) {
        if (s == 0) {
            if (idx < 9) break;
            idx -= 9;
            ++count;
            s = 1;
        } else if (s == 1) {
            if (idx < 6) break;
            idx -= 6;
            s = fe_after_six(fe_outside_exact<F>(FE_ROW(idx), FE_ROW(idx + 3)), fe_outside_exact<F>(FE_ROW(idx + 1), FE_ROW(idx + 4)),
                             fe_outside_exact<F>(FE_ROW(idx + 2), FE_ROW(idx + 5)));
        } else {
            if (idx < 2) break;
            idx -= 2;
            if (!fe_outside_exact<F>(FE_ROW(idx), FE_ROW(idx + 1))) s = fe_after_roll(s);
        }
    }
    return s;
}

#ifndef TP3_FE_GEN_REGS
#define TP3_FE_GEN_REGS 0   // 1: generator state in registers (fewer shared-memory accesses, but 96 registers: 20 warps per SM instead of 28: slower)
#endif

struct FeWalkArgs {
    const uint32_t* jump_table;   // [kRanfDigits][256][55], the seeded round 0 behind it (api.cu)
    uint64_t first_round;         // absolute index of the round where segment 0 starts
    uint32_t n_seg, seg_rounds;
    uint32_t warm, warm_first;
    const uint32_t* seg_list;     // redo mode: the segments to walk (else null: all of them) ...
    const uint8_t* seg_entry;     // ... and their true entry states
    uint32_t n_list;
    uint32_t* seg_count;          // [n_seg] events started in the segment
    uint8_t* seg_exit;            // [n_seg] consumer state at the end of the segment (0xff: the walks never coincided)
    uint8_t* seg_fail;            // [n_seg] 1: the nine walks had not coincided at the segment start -> redo
    FeRecord* records;            // [n_seg][kFeSlotsPerRound * seg_rounds], or null: count only
};

struct FeWalkSmem {
    uint32_t win[2 * kRanfLag + 2];
    // one private generator per lane, numbers[k + 1] of lane l at tile[k][l]: every lane owns a bank, so the walk's
    // data-dependent indices are as conflict free as the generator's fixed ones (rows with an odd stride made the walk's
    // loads 3-4 wavefronts each: the MIO queue was 20 % of the stalls, profiles/r02_fe_walk_*.txt)
    uint32_t tile[kRanfLag][32];
};

// ranf.rs:106-119 in place on a lane's private row.
__device__ __forceinline__ void fe_next_round(uint32_t* row) {
#pragma unroll
    for (int k = 0; k < 24; ++k) FE_ROW(k) = ranf_sub(FE_ROW(k), FE_ROW(k + 31));
#pragma unroll
    for (int k = 24; k < kRanfLag; ++k) FE_ROW(k) = ranf_sub(FE_ROW(k), FE_ROW(k - 24));
}

// ONE WARP = ONE CTA (7.5 KB of shared memory, 56 registers): a block fits wherever one CTA of the physics kernel retires,
// which is what lets the walk of the next pass run next to the physics of this one (api.cu: fe_stream_simulate).
constexpr int kFeWalkThreads = 32;
template <class F>
__global__ void __launch_bounds__(kFeWalkThreads, TP3_FE_GEN_REGS ? 20 : 28) fe_walk_kernel(const FeWalkArgs a) {
    __shared__ FeWalkSmem w;
    const int lane = threadIdx.x & 31;
    const bool redo = a.seg_list != nullptr;
    const uint32_t n_items = redo ? a.n_list : a.n_seg;
    const uint32_t item0 = blockIdx.x * 32;
    if (item0 >= n_items) return;
    const uint32_t* base = a.jump_table + (size_t)kRanfDigits * 256 * kRanfLag;  // the seeded round 0

    // where every lane's walk starts, and its generator there (jump-ahead is a warp-cooperative operation)
    uint32_t my_seg = 0;
    uint64_t my_start = 0;
    int my_warm = 0;
    uint64_t prev_start = 0;
    for (int k = 0; k < 32; ++k) {
        if (item0 + k >= n_items) break;  // warp-uniform
        const uint32_t g = redo ? a.seg_list[item0 + k] : item0 + k;
        const uint64_t s_round = a.first_round + (uint64_t)g * a.seg_rounds;
        const uint64_t want_warm = redo ? 0 : (g == 0 ? a.warm_first : a.warm);
        const uint64_t start = s_round >= want_warm ? s_round - want_warm : 0;
        if (k == 0 || redo) ranf_jump_to_round(w.win, base, start, a.jump_table, lane);
        else ranf_jump_to_round(w.win, w.win, start - prev_start, a.jump_table, lane);  // segments of a warp are consecutive
        prev_start = start;
        for (int i = lane; i < kRanfLag; i += 32) w.tile[i][k] = w.win[i];
        if (lane == k) {
            my_seg = g;
            my_start = start;
            my_warm = (int)(s_round - start);
        }
        __syncwarp();
    }
    const bool live = item0 + lane < n_items;
    uint32_t* const row = &w.tile[0][lane];
    FeRecord* const rec_base = (a.records && live) ? a.records + (size_t)my_seg * kFeSlotsPerRound * a.seg_rounds : nullptr;

    // consumer state: where each of the nine possible entry states is now (warm-up), then the single true state
    uint64_t cur = 0x876543210ull;
    bool single = false;
    int s = 0;
    if (redo) {
        single = true;
        s = live ? a.seg_entry[item0 + lane] : 0;
    } else if (my_start == 0) {  // the very beginning of the stream: a new event is about to start
        single = true;
        s = 0;
    }
    uint32_t count = 0;
    bool done = !live, rec = false, in_event = false, failed = false;
    uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0, u4 = 0, u5 = 0, u6 = 0, u7 = 0, u8 = 0, p0a = 0, p0b = 0, p1a = 0, p1b = 0, p2a = 0, p2b = 0;
    const int seg_end = my_warm + (int)a.seg_rounds;

#if TP3_FE_GEN_REGS
    // The generator lives in REGISTERS (55 of them) and a round is copied to shared memory only for the walk's data-dependent
    // indices: 55 shared-memory stores per round instead of 110 loads + 55 stores.
    uint32_t g[kRanfLag];
#pragma unroll
    for (int k = 0; k < kRanfLag; ++k) g[k] = FE_ROW(k);
#endif
    for (int t = 0; !__all_sync(0xffffffffu, done); ++t) {
#if TP3_FE_GEN_REGS
        if (t > 0) {  // ranf.rs:106-119
#pragma unroll
            for (int k = 0; k < 24; ++k) g[k] = ranf_sub(g[k], g[k + 31]);
#pragma unroll
            for (int k = 24; k < kRanfLag; ++k) g[k] = ranf_sub(g[k], g[k - 24]);
#pragma unroll
            for (int k = 0; k < kRanfLag; ++k) FE_ROW(k) = g[k];
        }
#else
        if (t > 0 && !done) fe_next_round(row);
#endif
        if (!done && t == my_warm && !single) {  // the segment starts and its state is still ambiguous: the host redoes it
            failed = true;
            done = true;
        }
        // ---- warm-up, while several entry states are still possible (2.2 rounds on average): every candidate walked, slowly
        const bool many = !done && !single;
        if (many) {
            uint32_t img = 0;
#pragma unroll
            for (int q = 0; q < 9; ++q) img |= 1u << ((cur >> (4 * q)) & 15u);
            uint64_t m = 0;
            while (img) {
                const int q = __ffs(img) - 1;
                img &= img - 1;
                int c;
                m |= (uint64_t)fe_walk_round_exact<F>(row, q, c) << (4 * q);
            }
            uint64_t nxt = 0;
#pragma unroll
            for (int q = 0; q < 9; ++q) nxt |= ((m >> (4 * ((cur >> (4 * q)) & 15u))) & 15ull) << (4 * q);
            cur = nxt;
            if (cur == (cur & 15u) * kFeNine4) {
                single = true;
                s = (int)(cur & 15u);
            }
        }  // (no `continue`: the other lanes of the warp may be further on, and the votes below are warp-wide)
        if (!done && t == seg_end) {
            if (a.seg_exit) a.seg_exit[my_seg] = (uint8_t)s;  // state at the end of the segment = entry state of the next one
            if (!in_event) done = true;                        // no event in progress: nothing left to do
        }
        // ---- one state: the stream is served exactly as the reference serves it.  Before the segment (the rest of the warm-up)
        // nothing is counted or recorded; in the segment every event that starts is counted and its record written; past the
        // segment's end only the event in progress is finished.
        const bool may_start = t < seg_end, counting = t >= my_warm;
        bool active = !done && !many;
        int idx = kRanfLag;
        // One round, request by request, the lanes of the warp in step: up to four times [9 numbers][6 numbers + three tests]
        // [re-rolls]; a lane that enters the round with requests pending skips what is already served in its first turn (it then
        // has room for at most three event starts: 2 + 3 x 15 + 9 > 55), a lane whose next request does not fit is finished.
        for (int e = 0; e < 6; ++e) {
            if (active && s == 0) {
                if (idx >= 9 && may_start) {  // a new event starts here (random_array::<9>, evgen.rs:149)
                    idx -= 9;
                    u0 = FE_ROW(idx); u1 = FE_ROW(idx + 1); u2 = FE_ROW(idx + 2); u3 = FE_ROW(idx + 3); u4 = FE_ROW(idx + 4);
                    u5 = FE_ROW(idx + 5); u6 = FE_ROW(idx + 6); u7 = FE_ROW(idx + 7); u8 = FE_ROW(idx + 8);
                    in_event = true;
                    rec = counting;
                    count += counting ? 1u : 0u;
                    s = 1;
                } else {
                    active = false;
                }
            }
            if (active && s == 1) {
                if (idx >= 6) {  // the three points, column-major 3x2 (evgen.rs:223-225): point k = (v[k], v[3 + k])
                    idx -= 6;
                    const uint32_t a0 = FE_ROW(idx), a1 = FE_ROW(idx + 1), a2 = FE_ROW(idx + 2), a3 = FE_ROW(idx + 3), a4 = FE_ROW(idx + 4), a5 = FE_ROW(idx + 5);
                    bool amb = false;
                    bool n0 = fe_outside_fast<F>(a0, a3, amb), n1 = fe_outside_fast<F>(a1, a4, amb), n2 = fe_outside_fast<F>(a2, a5, amb);
                    if (__builtin_expect(amb, 0)) {
                        n0 = fe_outside_exact<F>(a0, a3);
                        n1 = fe_outside_exact<F>(a1, a4);
                        n2 = fe_outside_exact<F>(a2, a5);
                    }
                    p0a = a0; p0b = a3; p1a = a1; p1b = a4; p2a = a2; p2b = a5;  // re-rolled points are overwritten below
                    s = fe_after_six(n0, n1, n2);
                } else {
                    active = false;
                }
            }
            while (active && s >= 2) {  // (a plain divergent loop: the lanes re-converge behind it)
                if (idx >= 2) {  // one re-roll of the first point that is still outside (evgen.rs:231-241)
                    idx -= 2;
                    const uint32_t x = FE_ROW(idx), y = FE_ROW(idx + 1);
                    if (s < 6) { p0a = x; p0b = y; }
                    else if (s < 8) { p1a = x; p1b = y; }
                    else { p2a = x; p2b = y; }
                    bool amb = false;
                    bool out = fe_outside_fast<F>(x, y, amb);
                    if (__builtin_expect(amb, 0)) out = fe_outside_exact<F>(x, y);
                    if (!out) s = fe_after_roll(s);
                } else {
                    active = false;
                }
            }
            if (active && s == 0 && in_event) {  // the event is complete: its record
                in_event = false;
                if (rec && rec_base) {
                    // two 256-bit stores (STG.E.256): every lane writes to another line, and what bounds this kernel is the
                    // number of (lane, store instruction) pairs the L1TEX tag stage sees, not the bytes
                    FeRecord* o = rec_base + (count - 1);
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(u0), "r"(u1), "r"(u2), "r"(u3), "r"(u4),
                                 "r"(u5), "r"(u6), "r"(u7)
                                 : "memory");
                    asm volatile("st.global.v8.b32 [%0+32], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(u8), "r"(p0a), "r"(p0b), "r"(p1a),
                                 "r"(p1b), "r"(p2a), "r"(p2b), "r"(0u)
                                 : "memory");
                }
                rec = false;
                if (!may_start) {  // that was the event straddling the end of the segment
                    active = false;
                    done = true;
                }
            }
            if (!__any_sync(0xffffffffu, active)) break;
        }
    }
    if (live) {
        a.seg_count[my_seg] = count;
        a.seg_fail[my_seg] = failed ? 1 : 0;
        if (failed && a.seg_exit) a.seg_exit[my_seg] = 0xff;
    }
}

// ---------------------------------------------------------------------------------------------------- physics
// Event from its record (evgen.rs:143-173): the f64 path works straight from the integers like gen_event_ints.
template <class F> __device__ __forceinline__ void fe_event_from_record(const uint32_t w[15], F e_total, const FastMath fm, F p[3][4]);
template <> __device__ __forceinline__ void fe_event_from_record<double>(const uint32_t w[15], double e_total, const FastMath fm, double p[3][4]) {
    double q[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double c = fma((double)(int)w[k], fm.fc->u_scale2, -1.0);
        // (the product of the two integers is exact in 64 bits: one conversion rounds it as the multiplication of the two doubles would)
        const double e = fma((double)((unsigned long long)w[3 + k] * (unsigned long long)w[6 + k]), fm.fc->u_scale_sq, Num<double>::MIN_POSITIVE);
        const double x = fma((double)(int)w[9 + 2 * k], fm.fc->u_scale2, -1.0), y = fma((double)(int)w[10 + 2 * k], fm.fc->u_scale2, -1.0);
        const double en = fast_neg_log(e, fm);
        // sin(theta) / |(x, y)| = sqrt(a) / sqrt(r2) = a / sqrt(a r2), a = 1 - c^2: ONE reciprocal square root instead of a
        // square root and a reciprocal square root (7 FP64 instructions less per photon).  The bias keeps a = 0 (c = -1: the
        // draw 0) away from 0 x inf: the factor is then exactly 0, as in the reference; a r2 >= 1.6e-26 otherwise.
        const double a = fma(-c, c, 1.0);
        const double esn = (en * a) * fast_rsqrt(fma(a, fma(x, x, y * y), 1e-290));
        q[k][0] = esn * x;
        q[k][1] = esn * y;
        q[k][2] = en * c;
        q[k][3] = en;
    }
    conformal_transform<double, false, false, Cons3<double, false>::value>(q, e_total, p);
}
template <> __device__ __forceinline__ void fe_event_from_record<float>(const uint32_t w[15], float e_total, const FastMath fm, float p[3][4]) {
    float u9[9], xy[3][2], r2[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) u9[i] = (float)(int)w[i] * 1e-9f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        xy[k][0] = 2.0f * ((float)(int)w[9 + 2 * k] * 1e-9f) - 1.0f;
        xy[k][1] = 2.0f * ((float)(int)w[10 + 2 * k] * 1e-9f) - 1.0f;
        r2[k] = xy[k][0] * xy[k][0] + xy[k][1] * xy[k][1];
    }
    gen_event_faster<float, false>(u9, xy, r2, e_total, fm, p);
}

struct FePhysArgs {
    const FeRecord* records;
    uint32_t slots_per_seg;        // kFeSlotsPerRound * seg_rounds
    const uint64_t* seg_events;    // [n_seg + 1] absolute index of the first event that starts in each segment of the pass
    const uint32_t* unit_seg;      // [n_units] segment of the pass holding the unit's first event
    uint64_t first_event;          // absolute index of the first event of unit 0 (a batch start)
    uint64_t end_event;            // absolute index one past the last event wanted (a short last batch)
    uint32_t n_units;              // kFeParts per batch
    uint32_t n_warps;
    tp3_acc* out_parts;            // [n_units]
    // per-event observables (tp3.h): device histograms, or null
    uint32_t hist_bins;
    unsigned long long* hist_counts;   // [TP3_HIST_OBSERVABLES][hist_bins]
    double* hist_weights;              // [kHistReplicas][TP3_HIST_OBSERVABLES][hist_bins]
};

template <class F> struct FePhysSmem {
    double2 log_tab[128];                                                   // FastMathSmem::log_tab (no sin / cos in this event generator)
    alignas(16) typename Pair<F>::type queue[2 * kQueuePhotons][kQueue];  // survivor queue (kernels.cuh)
    alignas(16) uint4 stage[2][4][32];                                      // records of this / the next warp iteration
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// ONE WARP = ONE CTA, like the fused default kernel; warp w takes units [n_units w / W, n_units (w + 1) / W).
// HIST: 0 = none; 1 / 2 = the per-event observable epilogue of the default kernel (kernels.cuh: hist_fill) with the photons in
// generation order / sorted by decreasing energy.
template <class F, int HIST = 0> __global__ void __launch_bounds__(32, 16) fe_physics_kernel(const FePhysArgs a, const PhysParams<F> P) {
    __shared__ FePhysSmem<F> sm;
    extern __shared__ __align__(16) unsigned char fe_hist_raw[];
    const int lane = threadIdx.x;
    const int hist_n = HIST ? TP3_HIST_OBSERVABLES * (int)a.hist_bins : 0;
    uint32_t* const hist_c = reinterpret_cast<uint32_t*>(fe_hist_raw);
    double* const hist_w = HIST ? a.hist_weights + (size_t)(blockIdx.x % kHistReplicas) * hist_n : nullptr;
    if (HIST)
        for (int i = lane; i < hist_n; i += 32) hist_c[i] = 0u;
#pragma unroll
    for (int i = 0; i < 128; i += 32) sm.log_tab[i + lane] = __ldg(&kLogTable[i + lane]);  // coalesced 128-bit loads from global memory (fastmath.cuh)
    __syncwarp();
    static_assert(offsetof(FastMathSmem, log_tab) == 0, "log_tab must lead FastMathSmem");
    const FastMath fm{reinterpret_cast<const FastMathSmem*>(sm.log_tab), &P.fc};  // only log_tab is read (fast_neg_log)
    const uint64_t u_lo = (uint64_t)a.n_units * blockIdx.x / a.n_warps, u_hi = (uint64_t)a.n_units * (blockIdx.x + 1) / a.n_warps;
    for (uint64_t u = u_lo; u < u_hi; ++u) {
        const uint64_t e0 = a.first_event + (u / kFeParts) * kBatch + (u % kFeParts) * kFePartLen;
        const uint64_t e1 = min(e0 + kFePartLen, a.end_event);
        const int n_ev = e1 > e0 ? (int)(e1 - e0) : 0;
        uint32_t g = a.unit_seg[u];
        auto fetch = [&](int it) {  // this lane's record of warp iteration `it` -> stage buffer (asynchronous)
            const uint64_t ev = min(e0 + (uint64_t)it * 32 + lane, e1 - 1);  // lanes past the end re-read the last record
            while (ev >= __ldg(a.seg_events + g + 1)) ++g;                    // events of a unit are in stream order
            const uint4* src = reinterpret_cast<const uint4*>(a.records + (size_t)g * a.slots_per_seg + (ev - __ldg(a.seg_events + g)));
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async16(&sm.stage[it & 1][c][lane], src + c);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        LaneAcc<F> acc;
        acc.clear();
        int q_head = 0, q_count = 0;
        const int n_it = (n_ev + 31) / 32;
        if (n_it) fetch(0);
        for (int it = 0; it < n_it; ++it) {
            if (it + 1 < n_it) {
                fetch(it + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            uint32_t w[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 v = sm.stage[it & 1][c][lane];
                w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
            }
            F p[3][4];
            fe_event_from_record<F>(w, P.e_total, fm, p);
            const bool keep = it * 32 + lane < n_ev && keep_event<F, false, false>(p, P);
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) queue_push<F>(sm.queue, (q_head + q_count + __popc(mask & ((1u << lane) - 1u))) & (kQueue - 1), p);
            q_count += __popc(mask);
            __syncwarp();
            if (q_count >= 32) {
                F e[3][4];
                queue_pop<F>(sm.queue, (q_head + lane) & (kQueue - 1), P.e_total, e);
                __syncwarp();
                q_head = (q_head + 32) & (kQueue - 1);
                q_count -= 32;
                F m[5];
                me_fast<F>(e, P, m);
                acc.integrate(m, P.sigma_contribs);
                if (HIST) hist_fill<F, HIST == 2>(e, m, P, hist_c, hist_w, (int)a.hist_bins);
            }
        }
        if (lane < q_count) {  // drain
            F e[3][4];
            queue_pop<F>(sm.queue, (q_head + lane) & (kQueue - 1), P.e_total, e);
            F m[5];
            me_fast<F>(e, P, m);
            acc.integrate(m, P.sigma_contribs);
            if (HIST) hist_fill<F, HIST == 2>(e, m, P, hist_c, hist_w, (int)a.hist_bins);
        }
        __syncwarp();
        F v[12];
        acc.fields(v);
        uint32_t n = acc.selected;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] += shfl_xor_t(v[k], off);
            n += __shfl_xor_sync(0xffffffffu, n, off);
        }
        if (lane == 0) {
            tp3_acc* o = a.out_parts + u;
            o->selected_events = n;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                o->spm2[k] = (double)v[k];
                o->vars[k] = (double)v[5 + k];
            }
            o->sigma = (double)v[10];
            o->variance = (double)v[11];
        }
    }
    if (HIST) {  // CTA histograms -> device histograms
        __syncwarp();
        for (int i = lane; i < hist_n; i += 32) {
            const uint32_t n = hist_c[i];
            if (n) atomicAdd(a.hist_counts + i, (unsigned long long)n);
        }
    }
}

// Batch accumulator = its part accumulators merged in part order, in the run's Float (resacc.rs:133-139).
template <class F>
__global__ void fe_combine_parts_kernel(const tp3_acc* __restrict__ parts, uint64_t n_batches, tp3_acc* __restrict__ out, int n_parts = kFeParts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b = i / 13;
    const int f = (int)(i % 13);
    if (b >= n_batches) return;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(parts + b * n_parts);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(out + b);
    if (f == 0) {
        unsigned long long n = 0;
        for (int k = 0; k < n_parts; ++k) n += src[k * 13];
        dst[0] = n;
    } else {
        F s = (F)__longlong_as_double((long long)src[f]);
        for (int k = 1; k < n_parts; ++k) s += (F)__longlong_as_double((long long)src[k * 13 + f]);
        dst[f] = (unsigned long long)__double_as_longlong((double)s);
    }
}

}  // namespace tp3

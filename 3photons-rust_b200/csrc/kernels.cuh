// The fused per-batch kernel (replaces the closure of src/main.rs:103-128 and everything it
// calls) plus the small parity/seeding kernels.
//
// Work decomposition: ONE WARP = ONE BATCH of <= 10 000 events (src/scheduling/mod.rs:21) = ONE CTA for the plain fast
// kernel (16 one-warp CTAs per SM; the histogram / literal / f32-xoshiro kernels keep four warps per CTA).  A warp
// regenerates its batch's random stream from one jump-ahead (rng.cuh), runs generation + cuts on 32 events per iteration,
// compacts the survivors (70.8 % at the default cuts) into a shared-memory queue and evaluates the matrix elements only
// on full warps, then folds its 13 sums with shuffles and writes one 104-byte accumulator.  There is no CTA-level
// barrier after the prologue, so warps never wait for each other.  What bounds it (the vector register file) and what
// follows from that is in DESIGN.md section 4d.
//
// Scheduling (multi_threading.rs:46-70 hands out batches to its workers one by one).  The launch's batches are cut into UNITS
// of consecutive batches, in batch order; the sequential RANF stream is re-positioned once per unit and simply continues
// inside it.  Two schedules (profiles/r02_schedule_ab.txt):
//   * dynamic (shipped): one warp per unit, dispatched by the hardware in unit order.  Big units first (16 batches for the f64
//     kernel, 8 for f32, fewer in short launches), their number a multiple of 4 x SMs; then half a wave of half units, half a wave
//     of quarter units and two waves of single batches, so that the device drains within one batch time whatever the launch
//     size (api.cu: fill_schedule; DESIGN.md section 4e has the trace of every unit this was designed from; round 1 used 1-8
//     equal batches per CTA: 8.8 waves and a 3 % tail at 125 000 batches per GPU).
//   * static: exactly as many warps as the device holds, each walking the same number of units (+- one batch).  No tail
//     at all, but 11 % SLOWER: warps that start together stay in step, so their integer phases (stream generation)
//     and FP64 phases collide instead of overlapping; warps that start at staggered times do not.
// Units complete roughly in batch order either way, which is what the ordered fold below needs.
//
// Ordered fold (ResultsAccumulator::merge in batch order, sequential.rs:24-36 / multi_threading.rs:107-126) inside the
// kernel: a warp that finishes a unit publishes it and then TRIES to take the fold lock; the holder adds every unit
// that is complete, in unit (= batch) order, to the running accumulator.  Nobody waits: a warp that finds the lock taken
// goes on simulating, and the holder re-checks after releasing, so a unit published meanwhile is never left behind.
// The sum is the strict left fold over the batches whichever warps did the adding.
#pragma once

#include <cstdint>

#include "../../include/tp3.h"
#include "physics.cuh"
#include "f32x2.cuh"
#include "rng.cuh"

namespace tp3 {

#ifndef TP3_MIN_CTAS
#define TP3_MIN_CTAS 4
#endif
#ifndef TP3_TICK_SPREAD
#define TP3_TICK_SPREAD 0   // 1: one round between each physics stage; 0: whole chain at stage TP3_TICK_AT
#endif
#ifndef TP3_TICK_AT
#define TP3_TICK_AT 1
#endif
// f32 two-events-per-lane kernel: 20 warps per SM (96 registers with a few spills beat 128 registers at 16 warps), as
// one-warp CTAs with RANF (+2.7 %) and four-warp CTAs with xoshiro (one-warp CTAs: -0.5 %); profiles/r01_ab_variants.txt
__host__ __device__ constexpr int x2_warps(int rng) { return rng == 0 ? 1 : 4; }
#ifndef TP3_WARPS
#define TP3_WARPS 4         // warps (batches) per CTA of the kernels that share per-CTA state: histogram epilogue, literal, f32 x2, dump
#endif
#ifndef TP3_FAST_WARPS
#define TP3_FAST_WARPS 1    // warps per CTA of the plain fast kernel: one-warp CTAs free their SM slot as soon as their batches are
#endif                      // done instead of waiting for the slowest of four warps (+2 %, profiles/r01_ab_variants.txt)
#ifndef TP3_FAST_MIN_CTAS
#define TP3_FAST_MIN_CTAS (16 / TP3_FAST_WARPS)   // 16 warps per SM at 128 registers
#endif
#ifndef TP3_GEN_FROM_INTS
#define TP3_GEN_FROM_INTS 1 // f64 RANF fast kernel: cos_theta, phi and r r' straight from the stream integers (6 FP64 fewer per event)
#endif
#ifndef TP3_ACC_SMEM
#define TP3_ACC_SMEM 0      // 1: the lanes' 12 partial sums live in shared memory (frees 24 registers in the event loop; slower)
#endif
#ifndef TP3_P3_CONS
#define TP3_P3_CONS 1       // 1: the survivor queue holds two photons, the third is rebuilt from 4-momentum conservation
#endif
constexpr int kWarps = TP3_WARPS;        // batches per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kBatch = TP3_EVENT_BATCH_SIZE;
constexpr int kLaneEvents = (kBatch + 31) / 32;  // xoshiro: contiguous events per lane (313)
constexpr int kQueue = 64;               // survivor queue slots per warp (< 32 pending + <= 32 new)
constexpr int kQueuePhotons = TP3_P3_CONS ? 2 : 3;

enum RngKind { RNG_RANF = 0, RNG_XOSHIRO = 1 };

struct SimArgs {
    uint64_t first_batch;      // index of the first batch of this launch within the run
    uint64_t n_batches;        // batches in this launch
    uint32_t last_batch_len;   // events in the last batch of the launch
    uint32_t jump_seeding;     // TP3_FASTER_THREADING: batch b starts after b rng.jump()s
    // A batch can be cut into `batch_parts` equal PARTS of `batch_events` events (sequential RANF stream only): the launch then
    // runs over part indices -- first_batch, n_batches and `out` count parts, consecutive parts are consecutive in the stream
    // exactly as consecutive batches are -- and a combine kernel adds the parts of a batch in part order (api.cu).
    uint32_t batch_events;     // events per slot of the launch: TP3_EVENT_BATCH_SIZE / batch_parts
    uint32_t batch_parts;      // 1: a slot is a batch
    // static schedule (see the head of this file)
    uint32_t n_warps;          // W: warps in the grid
    uint32_t unit_batches;     // consecutive batches per unit in the full rounds (the RANF stream simply continues inside a unit)
    uint32_t full_rounds;      // rounds of W full units; the rest is split evenly in one more round
    uint32_t dynamic;          // >= 1: one unit per warp, n_warps units dispatched by the hardware in unit order: dynamic - 1 RAMP units
                               //    of 1, 2, .., 8, 1, 2, .. batches (an A/B option: unequal first units to take the first wave out of
                               //    step at once; it does not pay), then `full_rounds` units of unit_batches batches, then single
                               //    batches (the last waves are short, so the tail is < 1 batch)
    uint32_t taper;            // dynamic schedule: between the big units and the single batches, `taper` units of unit_batches / 2 and
                               //    `taper` units of unit_batches / 4 batches: the warps leave the big units at different times, and
                               //    every stage absorbs the spread of the one before it
    uint32_t epoch;            // value that marks a unit of THIS launch as done in unit_done
    uint32_t* unit_done;       // [(full_rounds + 1) * W], or null: no in-kernel fold
    struct FoldState* fold;    // running accumulator of the ordered fold, or null
    const uint32_t* ranf_table;        // [kRanfDigits][256][55]
    const uint64_t* xo_batch_states;   // [n_batches][4] from the seeding kernel
    const uint64_t* xo_lane_polys;     // [32][4] jump polynomial of each lane's offset in the batch
    tp3_acc* out;                      // [n_batches]
    uint32_t ranf_base[kRanfLag];      // seeded round 0 (ranf.rs:28-66), slot order
    int32_t ranf_seed;
    // per-event observables (tp3.h): device histograms, or null
    uint32_t hist_bins;
    unsigned long long* hist_counts;   // [TP3_HIST_OBSERVABLES][hist_bins]
    double* hist_weights;              // [kHistReplicas][TP3_HIST_OBSERVABLES][hist_bins]
#ifdef TP3_TRACE_UNITS                 /* diagnostic build only (scripts/unit_trace.py): when and where every unit ran */
    unsigned long long* trace;         // [units][3]: globaltimer at the unit's start and end, smid | warpid << 16 | batches << 32
#endif
};

// Ordered fold of a launch (one per device slot, reset by the host before the launch).
struct FoldState {
    uint32_t lock;                 // 0 free, 1 held
    uint32_t pad;
    unsigned long long next_unit;  // first unit that has not been added yet
    tp3_acc running;               // fold of the batches of units [0, next_unit)
};

struct DumpArgs {
    uint32_t n_events;   // events of the batch to dump
    uint64_t* words;     // [n_events][12] raw stream words (may be null)
    double* momenta;     // [n_events][3][4]             (may be null)
    int32_t* kept;       // [n_events]
    double* m2;          // [n_events][5]
};

template <class F, int RNG> struct RawWord;
template <class F> struct RawWord<F, RNG_RANF> { using type = uint32_t; };
template <> struct RawWord<double, RNG_XOSHIRO> { using type = uint64_t; };
template <> struct RawWord<float, RNG_XOSHIRO> { using type = uint32_t; };

// raw stream word -> uniform in [0,1): ranf.rs:99 / rand 0.8.5 Standard distribution (Appendix B.3)
__device__ __forceinline__ double to_uniform_xo(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ float to_uniform_xo(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

template <class F> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

template <class F, int Q = kQueue, int NP = kQueuePhotons> struct WarpSmem {
    RanfWarpSmem ranf;
    // survivors' momenta as (X,Y) and (Z,E) pairs per photon, pair-major: consecutive slots are
    // consecutive 16-byte (f64) words, so the 128-bit accesses of a warp are conflict free
    alignas(16) typename Pair<F>::type queue[2 * NP][Q];
};
template <class F, int W = kWarps> struct BlockSmem {
    FastMathSmem fm;
    WarpSmem<F> w[W];
#if TP3_ACC_SMEM
    F acc[W][12][32];  // field-major: a warp's access to one field is one conflict-free wavefront (two for f64)
#endif
};
// Warps per CTA / minimum resident CTAs of simulate_kernel: the plain fast kernel runs as one-warp CTAs.
__host__ __device__ constexpr int sim_warps(bool literal, bool hist) { return (literal || hist) ? kWarps : TP3_FAST_WARPS; }
__host__ __device__ constexpr int sim_min_ctas(bool literal, bool hist) { return literal ? 2 : hist ? TP3_MIN_CTAS : TP3_FAST_MIN_CTAS; }

// Per-warp random source: hands each lane the 12 raw words of "its" event of iteration `it`.
template <class F, int RNG> struct WarpRng;

// RANF: lanes interleave (event = 32 it + lane); the warp regenerates the global stream cooperatively.
template <class F> struct WarpRng<F, RNG_RANF> {
    RanfWarpStream s;
    int n;
    template <class WS> __device__ void init(const SimArgs& a, WS* sm, uint64_t batch, uint64_t /*slot*/, int n_ev, int lane) {
        n = n_ev;
        if (a.jump_seeding) {
            // rng.jump() = reseed with seed + 123456 per batch (ranf.rs:136-140), i32 wrapping
            uint32_t* y = reinterpret_cast<uint32_t*>(&sm->queue[0][0]);  // the queue is idle during seeding
            ranf_seed_warp(y, y + 64, (int32_t)((uint32_t)a.ranf_seed + 123456u * (uint32_t)batch), lane);
            s.init(&sm->ranf, y, 0, a.ranf_table, lane);
        } else {
            s.init(&sm->ranf, a.ranf_base, (uint64_t)kDrawsPerEvent * a.batch_events * batch, a.ranf_table, lane);
        }
    }
    __device__ int iterations() const { return (n + 31) / 32; }
    __device__ int event_of(int it, int lane) const {
        const int e = 32 * it + lane;
        return e < n ? e : -1;
    }
    __device__ __forceinline__ void draws(int, int lane, uint32_t w[12]) { s.draws(lane, w); }
    // Pipelined refill for the next warp iteration (see RanfWarpStream::begin_next)
    __device__ __forceinline__ void begin_next(int lane) { s.begin_next(lane); }
    template <int K> __device__ __forceinline__ void tick(int lane) { s.template tick<K>(lane); }
    // The next batch of the sequential stream starts right after this batch's last draw: step past
    // the (possibly partial) last warp iteration instead of jumping again.
    __device__ bool next_batch(const SimArgs& a, int n_next, int lane) {
        if (a.jump_seeding) return false;
        if (n_next == 0 || n == 0) {  // the empty trailing parts of a short last batch (batch_parts): no draw is read any more
            n = 0;
            return n_next == 0;
        }
        // (advance() needs the step to leave the first buffered round: true for whole batches, 192 draws into the last warp
        // iteration, and for the part sizes api.cu allows: 5000, 2000 and 1000 events end 96, 192 and 96 draws into theirs)
        s.advance(kDrawsPerEvent * (n - 32 * (iterations() - 1)), lane);
        n = n_next;
        return true;
    }
    // ranf.rs:99: (n as Float) * 1e-9; `phi` asks for PhiScale<F>::value * u instead (an exact scaling)
    __device__ static F uniform(uint32_t w, bool phi, const FastCoef& fc) {
        if (sizeof(F) == 8) return (F)(phi ? u32_times(w, fc.phi_scale, fc.phi_bias) : u32_times(w, fc.u_scale, fc.u_bias));
        const float u = (float)(int)w * 1e-9f;
        return (F)(phi ? 4.0f * u : u);
    }
};

// xoshiro: per-lane state in registers, each lane owns a contiguous run of kLaneEvents events.
template <class F, class Lane> struct XoshiroWarpRng {
    Lane g;
    int lo, hi, n;
    template <class WS> __device__ void init(const SimArgs& a, WS*, uint64_t, uint64_t slot, int n_ev, int lane) {
        n = n_ev;
        lo = lane * kLaneEvents;
        hi = min(n_ev, lo + kLaneEvents);
        const uint64_t* st = a.xo_batch_states + 4 * slot;
        g.s0 = (decltype(g.s0))st[0]; g.s1 = (decltype(g.s0))st[1]; g.s2 = (decltype(g.s0))st[2]; g.s3 = (decltype(g.s0))st[3];
        if (lane && lo < hi) g.apply(a.xo_lane_polys + 4 * lane);
    }
    __device__ int iterations() const { return min(n, kLaneEvents); }
    __device__ int event_of(int it, int) const { return lo + it < hi ? lo + it : -1; }
    __device__ bool next_batch(const SimArgs&, int, int) { return false; }  // re-seeded from the batch state
    template <class W> __device__ __forceinline__ void draws(int, int, W w[12]) {
#pragma unroll
        for (int j = 0; j < 12; ++j) w[j] = g.next();
    }
    __device__ __forceinline__ void begin_next(int) {}
    template <int K> __device__ __forceinline__ void tick(int) {}
};
template <> struct WarpRng<double, RNG_XOSHIRO> : XoshiroWarpRng<double, Xoshiro256Lane> {
    __device__ static double uniform(uint64_t w, bool phi, const FastCoef&) { return (phi ? 256.0 : 1.0) * to_uniform_xo(w); }
};
template <> struct WarpRng<float, RNG_XOSHIRO> : XoshiroWarpRng<float, Xoshiro128Lane> {
    __device__ static float uniform(uint32_t w, bool phi, const FastCoef&) { return (phi ? 4.0f : 1.0f) * to_uniform_xo(w); }
};

template <class F> struct LaneAcc {
    F spm2[5], vars[5], sigma, variance;
    uint32_t selected;
    __device__ void clear() {
#pragma unroll
        for (int k = 0; k < 5; ++k) spm2[k] = vars[k] = 0;
        sigma = variance = 0;
        selected = 0;
    }
    // resacc.rs:121-129
    __device__ __forceinline__ void integrate(const F m[5], const F sc[5]) {
        selected += 1;
        F w = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            spm2[k] += m[k];
            vars[k] += m[k] * m[k];
            w += m[k] * sc[k];
        }
        sigma += w;
        variance += w * w;
    }
    __device__ __forceinline__ void fields(F v[12]) const {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            v[k] = spm2[k];
            v[5 + k] = vars[k];
        }
        v[10] = sigma;
        v[11] = variance;
    }
};

// The same sums kept in the warp's shared memory: 12 load-add-store per matrix-element step instead of 24 live registers.
template <class F> struct SmemLaneAcc {
    F (*a)[32];
    int lane;
    uint32_t selected;
    __device__ void clear() {
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k][lane] = 0;
        selected = 0;
    }
    __device__ __forceinline__ void integrate(const F m[5], const F sc[5]) {
        selected += 1;
        F w = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            a[k][lane] += m[k];
            a[5 + k][lane] += m[k] * m[k];
            w += m[k] * sc[k];
        }
        a[10][lane] += w;
        a[11][lane] += w * w;
    }
    __device__ __forceinline__ void fields(F v[12]) const {
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = a[k][lane];
    }
};

// Tick hook handed to gen_event: step K of the pipelined refill, if there is a next iteration.
template <class F, int RNG> struct RngTick {
    WarpRng<F, RNG>& rng;
    int lane;
    bool more;
    template <int K> __device__ __forceinline__ void at() {
#if TP3_TICK_SPREAD
        if (more) rng.template tick<K>(lane);
#else
        if (more && K == TP3_TICK_AT) {  // whole chain in one place: no fences inside, the compiler interleaves it
            rng.template tick<1>(lane); rng.template tick<2>(lane); rng.template tick<3>(lane); rng.template tick<4>(lane);
            rng.template tick<5>(lane); rng.template tick<6>(lane); rng.template tick<7>(lane);
        }
#endif
    }
};

template <class F> __device__ __forceinline__ F shfl_xor_t(F v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Per-event observables (tp3.h).  Event counts: u32 histograms of the CTA in dynamic shared memory (native shared
// atomics; a CTA sees < 2^32 events), added to the device histograms when the CTA ends.  Weight sums: f64 reductions
// straight to one of kHistReplicas device copies in L2 (fire-and-forget RED.ADD.F64; a shared-memory f64 add would be a
// compare-and-swap loop), the copies are summed when the histograms are fetched.
constexpr int kHistReplicas = 64;
__device__ __forceinline__ int hist_bin(double t, int nb) {  // t in [0, 1] up to rounding
    const int b = (int)(t * (double)nb);
    return min(max(b, 0), nb - 1);
}
template <class F, bool SORT>
__device__ __forceinline__ void hist_fill(const F e[3][4], const F m[5], const PhysParams<F>& P, uint32_t* hc, double* hw, int nb) {
    // hc: the CTA's counts (shared), hw: this CTA's replica of the device weights (global)
    F w = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) w += m[k] * P.sigma_contribs[k];
    F en[3] = {e[0][3], e[1][3], e[2][3]}, px[3] = {e[0][0], e[1][0], e[2][0]};
    if (SORT) {  // evgen.rs:109-118
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = a + 1; b < 3; ++b) {
                const bool sw = en[b] > en[a];
                const F ea = en[a], eb = en[b], xa = px[a], xb = px[b];
                en[a] = sw ? eb : ea; en[b] = sw ? ea : eb;
                px[a] = sw ? xb : xa; px[b] = sw ? xa : xb;
            }
    }
    const double inv = 2.0 / (double)P.e_total;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int bx = hist_bin((double)en[k] * inv, nb);
        const int bc = hist_bin(0.5 + 0.5 * (double)(px[k] * rcp_t(en[k])), nb);  // reciprocal: <= 1 ulp from the quotient
        atomicAdd(hc + k * nb + bx, 1u);
        atomicAdd(hw + k * nb + bx, (double)w);
        atomicAdd(hc + (3 + k) * nb + bc, 1u);
        atomicAdd(hw + (3 + k) * nb + bc, (double)w);
    }
}

// Survivor queue of a warp: momenta as (X,Y) and (Z,E) pairs per photon.  With TP3_P3_CONS only photons 0 and 1 are
// stored; photon 2 is what 4-momentum conservation leaves (the transform of evgen.rs:94-106 maps the photons' total
// momentum to (0,0,0,e_total) up to rounding error, the same identity the pair cuts and |s_ij|^2 already rely on).
template <class F, class Q> __device__ __forceinline__ void queue_push(Q queue, int slot, const F p[3][4]) {
#pragma unroll
    for (int k = 0; k < kQueuePhotons; ++k) {
        queue[2 * k][slot] = {p[k][0], p[k][1]};
        queue[2 * k + 1][slot] = {p[k][2], p[k][3]};
    }
}
template <class F, class Q> __device__ __forceinline__ void queue_pop(Q queue, int slot, F e_total, F e[3][4]) {
#pragma unroll
    for (int k = 0; k < kQueuePhotons; ++k) {
        const typename Pair<F>::type xy = queue[2 * k][slot], ze = queue[2 * k + 1][slot];
        e[k][0] = xy.x; e[k][1] = xy.y; e[k][2] = ze.x; e[k][3] = ze.y;
    }
    if (kQueuePhotons == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) e[2][c] = -(e[0][c] + e[1][c]);
        e[2][3] = (e_total - e[0][3]) - e[1][3];
    }
}

// Number of events of batch `slot` of the launch.
__device__ __forceinline__ int batch_len(const SimArgs& a, uint64_t slot) {
    if (a.batch_parts <= 1) return (slot + 1 == a.n_batches) ? (int)a.last_batch_len : kBatch;
    if (slot + a.batch_parts < a.n_batches) return (int)a.batch_events;  // a part of a full batch
    // the parts of the launch's last batch share its last_batch_len events, in order
    const int before = (int)((slot + a.batch_parts - a.n_batches) * a.batch_events);
    return max(0, min((int)a.batch_events, (int)a.last_batch_len - before));
}

// Batches [lo, hi) of the launch that make up unit u.
// (dynamic schedule: the first argument has its top bit set and carries the number of ramp units, a multiple of 8, in its
// other bits; `full_rounds` is the number of big units; the rest are single batches)
constexpr uint32_t kSchedDynamic = 0x80000000u;
__host__ __device__ __forceinline__ void unit_range_dynamic(uint64_t ramp, uint64_t big, uint64_t unit_batches, uint64_t taper, uint64_t u,
                                                            uint64_t& lo, uint64_t& hi) {
    const uint64_t ramp_batches = (ramp >> 3) * 36;  // 1 + 2 + .. + 8 per group of eight ramp units
    if (u < ramp) {
        const uint64_t j = u & 7;
        lo = (u >> 3) * 36 + j * (j + 1) / 2;
        hi = lo + j + 1;
        return;
    }
    u -= ramp;
    uint64_t base = ramp_batches, size = unit_batches;
    if (u >= big) {  // taper: units of unit_batches / 2, then of unit_batches / 4, then single batches
        u -= big;
        base += big * size;
        size >>= 1;
        if (u >= taper) {
            u -= taper;
            base += taper * size;
            size >>= 1;
            if (u >= taper) {
                u -= taper;
                base += taper * size;
                size = 1;
            }
        }
    }
    lo = base + u * size;
    hi = lo + size;
}
__device__ __forceinline__ void unit_range(uint32_t n_warps, uint32_t full_rounds, uint32_t unit_batches, uint32_t taper, uint64_t n_batches,
                                           uint64_t u, uint64_t& lo, uint64_t& hi) {
    if (n_warps & kSchedDynamic) {
        unit_range_dynamic(n_warps & ~kSchedDynamic, full_rounds, unit_batches, taper, u, lo, hi);
        return;
    }
    const uint64_t W = n_warps, full = (uint64_t)full_rounds * W;
    if (u < full) {
        lo = u * unit_batches;
        hi = lo + unit_batches;
    } else {
        const uint64_t base = full * unit_batches, rem = n_batches - base, w = u - full;
        lo = base + rem * w / W;
        hi = base + rem * (w + 1) / W;
    }
}
__device__ __forceinline__ void unit_range(const SimArgs& a, uint64_t u, uint64_t& lo, uint64_t& hi) {
    unit_range(a.dynamic ? ((a.dynamic - 1) | kSchedDynamic) : a.n_warps, a.full_rounds, a.unit_batches, a.taper, a.n_batches, u, lo, hi);
}
// Number of units of a launch.
__device__ __forceinline__ uint64_t unit_count(uint32_t n_warps, uint32_t full_rounds, uint32_t unit_batches, uint32_t taper,
                                               uint64_t n_batches) {
    if (n_warps & kSchedDynamic) {
        const uint64_t ramp = n_warps & ~kSchedDynamic;
        const uint64_t taper_batches = (uint64_t)taper * ((unit_batches >> 1) + (unit_batches >> 2));
        return ramp + full_rounds + 2ull * taper + (n_batches - (ramp >> 3) * 36 - (uint64_t)full_rounds * unit_batches - taper_batches);
    }
    return ((uint64_t)full_rounds + 1) * n_warps;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_sc() { asm volatile("fence.sc.gpu;" ::: "memory"); }

// Publish unit u (its accumulators are in a.out) and fold whatever is ready.  Called by all lanes of the warp.
// F is the run's Float: under f32 the merge adds in f32 (resacc.rs:133-139 on f32 fields), like tp3_merge.
// (The schedule is passed by value: a reference to the kernel's parameter block would force a local-memory copy of it.)
template <class F>
__device__ __noinline__ void fold_publish(FoldState* fs, uint32_t* unit_done, const tp3_acc* out, uint32_t n_warps, uint32_t full_rounds,
                                          uint32_t unit_batches, uint32_t taper, uint32_t epoch, uint64_t n_batches, uint64_t u,
                                          int lane) {
    const uint64_t n_units = unit_count(n_warps, full_rounds, unit_batches, taper, n_batches);
    __syncwarp();
    if (lane == 0) {
        __threadfence();  // this unit's accumulators (written by lane 0) before the flag
        *reinterpret_cast<volatile uint32_t*>(unit_done + u) = epoch;
    }
    for (;;) {
        uint32_t got = 0;
        if (lane == 0) {
            fence_sc();  // flag store (or lock release below) before the lock read: the other side does the mirror image
            got = atomicCAS(&fs->lock, 0u, 1u) == 0u;
            if (got) __threadfence();
        }
        got = __shfl_sync(0xffffffffu, got, 0);
        if (!got) return;  // the holder re-checks after it releases
        uint64_t nu = *reinterpret_cast<volatile unsigned long long*>(&fs->next_unit);
        // lane k < 12 carries field k (spm2[5], vars[5], sigma, variance), lane 12 the event count
        const unsigned long long* run = reinterpret_cast<const unsigned long long*>(&fs->running);
        F acc = 0;
        uint64_t cnt = 0;
        if (nu > 0) {
            if (lane < 12) acc = (F)__longlong_as_double((long long)__ldcg(run + 1 + lane));
            else if (lane == 12) cnt = __ldcg(run);
        }
        const int word = lane < 12 ? 1 + lane : 0;
        for (;;) {
            // how many of the next 32 units are complete: one flag per lane, one round trip
            const bool ready = nu + lane < n_units && ld_acquire_u32(unit_done + nu + lane) == epoch;
            const unsigned not_ready = ~__ballot_sync(0xffffffffu, ready);
            const unsigned m = not_ready ? (unsigned)__ffs(not_ready) - 1u : 32u;  // leading run of ready units
            if (m == 0) break;
            uint64_t lo, hi, lo2;
            unit_range(n_warps, full_rounds, unit_batches, taper, n_batches, nu, lo, hi);
            unit_range(n_warps, full_rounds, unit_batches, taper, n_batches, nu + m - 1, lo2, hi);  // units are consecutive batch ranges
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(out) + word;
            for (uint64_t b = lo; b < hi; b += 16) {  // 16 loads in flight, then the dependent chain of additions
                unsigned long long v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (b + j < hi && lane < 13) ? __ldcg(src + (b + j) * 13) : 0ull;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (b + j < hi) {
                        if (lane < 12) acc += (F)__longlong_as_double((long long)v[j]);
                        else cnt += v[j];
                    }
                }
            }
            nu += m;
            if (m < 32) break;
        }
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&fs->running);
        if (lane < 12) __stcg(dst + 1 + lane, (unsigned long long)__double_as_longlong((double)acc));
        else if (lane == 12) __stcg(dst, (unsigned long long)cnt);
        __syncwarp();
        bool again = false;
        if (lane == 0) {
            *reinterpret_cast<volatile unsigned long long*>(&fs->next_unit) = nu;
            __threadfence();
            atomicExch(&fs->lock, 0u);
            fence_sc();
            again = nu < n_units && ld_acquire_u32(unit_done + nu) == epoch;  // published while we held the lock?
        }
        if (!__shfl_sync(0xffffffffu, (uint32_t)again, 0)) return;
    }
}

template <class F, int RNG, bool SORT, bool LITERAL, bool HIST = false>
__global__ void __launch_bounds__(32 * sim_warps(LITERAL, HIST), sim_min_ctas(LITERAL, HIST)) simulate_kernel(const SimArgs a, const PhysParams<F> P) {
    constexpr int kWarps = sim_warps(LITERAL, HIST), kThreads = 32 * kWarps;  // this kernel's CTA shape
    __shared__ BlockSmem<F, kWarps> sm;
    extern __shared__ __align__(16) unsigned char hist_raw[];
    using Word = typename RawWord<F, RNG>::type;
    const int warp = kWarps == 1 ? 0 : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    fastmath_load<kThreads>(&sm.fm);
    const int hist_n = HIST ? TP3_HIST_OBSERVABLES * (int)a.hist_bins : 0;
    uint32_t* const hist_c = reinterpret_cast<uint32_t*>(hist_raw);
    double* const hist_w = HIST ? a.hist_weights + (size_t)(blockIdx.x % kHistReplicas) * hist_n : nullptr;
    if (HIST) {
        for (int i = threadIdx.x; i < hist_n; i += kThreads) hist_c[i] = 0u;
    }
    __syncthreads();
    typename Pair<F>::type(*queue)[kQueue] = sm.w[warp].queue;

    // The fast kernel never needs the sorted order: the energy cut is min(E_1,E_2,E_3) either way, the plane
    // normal p_a x p_b has the same direction for any photon pair (momenta sum to zero), and the helicity sums
    // are symmetric under photon permutations (the reference's no-photon-sorting golden is a symlink to the
    // default one). The sort (evgen.rs:109-118) stays in the literal kernel and in the per-event dump.
    constexpr bool kSort = SORT && LITERAL;

    WarpRng<F, RNG> rng;
    const uint64_t wid = (uint64_t)blockIdx.x * kWarps + warp;
  const uint32_t last_round = a.dynamic ? 0u : a.full_rounds;  // dynamic schedule: this warp's one unit is unit `wid`
  for (uint32_t round = 0; round <= last_round; ++round) {
    const uint64_t unit = (uint64_t)round * a.n_warps + wid;
    uint64_t unit_lo, unit_hi;
    unit_range(a, unit, unit_lo, unit_hi);
    if (unit_lo >= a.n_batches) break;  // (warps past the last unit of a dynamic grid rounded up to whole CTAs)
#ifdef TP3_TRACE_UNITS
    if (a.trace && lane == 0) {
        unsigned long long t;
        unsigned smid, wid_hw;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid_hw));
        a.trace[3 * unit] = t;
        a.trace[3 * unit + 2] = smid | (wid_hw << 16) | (unsigned long long)(unit_hi - unit_lo) << 32;
    }
#endif
  for (uint64_t slot = unit_lo; slot < unit_hi; ++slot) {
    const int n_ev = batch_len(a, slot);
    if (slot == unit_lo || !rng.next_batch(a, n_ev, lane)) rng.init(a, &sm.w[warp], a.first_batch + slot, slot, n_ev, lane);

#if TP3_ACC_SMEM
    SmemLaneAcc<F> acc{sm.acc[warp], lane, 0};
#else
    LaneAcc<F> acc;
#endif
    acc.clear();
    int q_head = 0, q_count = 0;  // survivor queue (warp-uniform)
    const int n_it = rng.iterations();
    for (int it = 0; it < n_it; ++it) {
        Word w[12];
        rng.draws(it, lane, w);
        // The random numbers of the NEXT iteration are regenerated in seven steps interleaved with this
        // iteration's physics (each step is a short serial chain; alone it would stall the warp).
#ifdef TP3_EXPERIMENT_FAKE_RNG   /* timing experiment only: reuse the same draws, results are wrong */
        const bool more = false;
#else
        const bool more = it + 1 < n_it;
#endif
        if (more) rng.begin_next(lane);
        RngTick<F, RNG> tick{rng, lane, more};
        F p[3][4];  // lanes past the end compute on valid but unused draws
        if constexpr (TP3_GEN_FROM_INTS && sizeof(F) == 8 && !LITERAL) {
            // straight from the stream integers (the scales are kernel parameters: 1e-9-based for RANF, powers of two for
            // xoshiro256+, where (x >> 11) 2^-53 makes this form EXACTLY the reference's arithmetic with four multiplications less)
            double d[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) d[j] = RNG == RNG_RANF ? (double)(int)w[j] : (double)((unsigned long long)w[j] >> 11);
            double rr[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)  // RANF: n n' < 1e18 is exact in 64 bits: one IMAD.WIDE + one conversion instead of two + a DMUL
                rr[k] = RNG == RNG_RANF ? (double)((unsigned long long)w[4 * k + 2] * (unsigned long long)w[4 * k + 3]) : d[4 * k + 2] * d[4 * k + 3];
            gen_event_ints<kSort>(d, rr, P.e_total, FastMath{&sm.fm, &P.fc}, p, tick);
        } else {
            F u[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) u[j] = WarpRng<F, RNG>::uniform(w[j], !LITERAL && (j & 3) == 1, P.fc);
            gen_event<F, kSort, LITERAL>(u, P.e_total, FastMath{&sm.fm, &P.fc}, p, tick);
        }
        tick.template at<4>();
        const bool keep = rng.event_of(it, lane) >= 0 && keep_event<F, kSort, LITERAL>(p, P);
        tick.template at<5>();
        if (LITERAL) {
            tick.template at<6>();
            tick.template at<7>();
            if (keep) {
                F m[5];
                me_literal<F>(p, P, m);
                acc.integrate(m, P.sigma_contribs);
            }
            continue;
        }
        // evcut.rs:42-96 as a predicate mask; survivors are compacted so that the matrix elements
        // (75 % of the FP64 work) always run on full warps.
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int slot = (q_head + q_count + __popc(mask & ((1u << lane) - 1u))) & (kQueue - 1);
            queue_push<F>(queue, slot, p);
        }
        q_count += __popc(mask);
        __syncwarp();
        tick.template at<6>();
        tick.template at<7>();
        if (q_count >= 32) {
            const int slot = (q_head + lane) & (kQueue - 1);
            F e[3][4];
            queue_pop<F>(queue, slot, P.e_total, e);
            __syncwarp();
            q_head = (q_head + 32) & (kQueue - 1);
            q_count -= 32;
#ifndef TP3_EXPERIMENT_NO_ME
            F m[5];
            me_fast<F>(e, P, m);
            acc.integrate(m, P.sigma_contribs);
            if (HIST) hist_fill<F, SORT>(e, m, P, hist_c, hist_w, (int)a.hist_bins);
#else
            acc.spm2[0] += e[0][0] + e[1][1] + e[2][2];
#endif
        }
    }
    if (!LITERAL && lane < q_count) {  // drain
        const int slot = (q_head + lane) & (kQueue - 1);
        F e[3][4];
        queue_pop<F>(queue, slot, P.e_total, e);
        F m[5];
        me_fast<F>(e, P, m);
        acc.integrate(m, P.sigma_contribs);
        if (HIST) hist_fill<F, SORT>(e, m, P, hist_c, hist_w, (int)a.hist_bins);
    }

    // ResultsAccumulator of the batch: xor-shuffle tree over the 32 lane-partials (deterministic)
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
        tp3_acc* o = a.out + slot;
        o->selected_events = n;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            o->spm2[k] = (double)v[k];
            o->vars[k] = (double)v[5 + k];
        }
        o->sigma = (double)v[10];
        o->variance = (double)v[11];
    }
    __syncwarp();
  }
#ifdef TP3_TRACE_UNITS
    if (a.trace && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[3 * unit + 1] = t;
    }
#endif
    if (a.fold) fold_publish<F>(a.fold, a.unit_done, a.out, a.dynamic ? ((a.dynamic - 1) | kSchedDynamic) : a.n_warps, a.full_rounds, a.unit_batches, a.taper, a.epoch, a.n_batches, unit, lane);
  }
    if (HIST) {  // CTA histograms -> device histograms
        __syncthreads();
        for (int i = threadIdx.x; i < hist_n; i += kThreads) {
            const uint32_t n = hist_c[i];
            if (n) atomicAdd(a.hist_counts + i, (unsigned long long)n);
        }
    }
}

// ---- f32, two events per lane (f32x2.cuh) --------------------------------------------------------------------------
// Same streams, same event physics and the same batch sums as simulate_kernel<float, RNG, ., false> (up to the order
// of the additions), but every lane carries the events of two warp iterations through the generation, the cuts and the
// matrix elements in packed FP32 arithmetic.
constexpr int kQueue2 = 128;  // < 64 pending + <= 64 new survivors

// The uniforms of the two events of a lane, scaled TOGETHER (one packed multiply for two raw words; xoshiro: the azimuth's
// factor 4 is part of the power of two).  Same bits as WarpRng<float, RNG>::uniform on each half: every step is either exact or
// the same single rounding.
template <int RNG> __device__ __forceinline__ f2 uniform_x2(uint32_t a, uint32_t b, bool phi);
template <> __device__ __forceinline__ f2 uniform_x2<RNG_XOSHIRO>(uint32_t a, uint32_t b, bool phi) {
    return f2((float)(a >> 8), (float)(b >> 8)) * f2(phi ? 1.0f / 4194304.0f : 1.0f / 16777216.0f);
}
template <> __device__ __forceinline__ f2 uniform_x2<RNG_RANF>(uint32_t a, uint32_t b, bool phi) {
    const f2 u = f2((float)(int)a, (float)(int)b) * f2(1e-9f);
    return phi ? u * f2(4.0f) : u;
}

template <int RNG>
__global__ void __launch_bounds__(32 * x2_warps(RNG), 20 / x2_warps(RNG)) simulate_kernel_x2(const SimArgs a, const PhysParams<f2> P) {
    constexpr int kWarps = x2_warps(RNG);  // this kernel's CTA shape
    using F = float;
    using Word = typename RawWord<F, RNG>::type;
    __shared__ WarpSmem<F, kQueue2, 3> smw[kWarps];
    const int warp = kWarps == 1 ? 0 : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    // Survivor queue, component-major: row 4 k + c holds component c (X, Y, Z, E) of photon k.  Queue slot s lives at word
    // (s & 64) | (s & 31) << 1 | (s >> 5 & 1) of its row, so that the two events a lane takes when 64 survivors are popped
    // (slots q_head + lane and q_head + 32 + lane, q_head = 0 or 64) are ADJACENT words: one 64-bit load per component
    // delivers the packed pair as the arithmetic wants it (event a in the low half, event b in the high half), and a
    // survivor is pushed with 32-bit stores straight from the half it sits in.  (Photon-major float2 pairs cost 81
    // register moves per pop to transpose, profiles/r02_f32x2_xoshiro_kernel.txt.)
    float(*qf)[kQueue2] = reinterpret_cast<float(*)[kQueue2]>(&smw[warp].queue[0][0]);
    static_assert(sizeof(smw[0].queue) == 12 * kQueue2 * sizeof(float) && kQueue2 == 128, "12 component rows of 128 slots");
    auto qword = [](int s) { return (s & 64) | ((s & 31) << 1) | ((s >> 5) & 1); };
    const FastMath fm{nullptr, nullptr};  // the f32 elementary functions are SFU instructions, no tables

    WarpRng<F, RNG> rng;
    const uint64_t wid = (uint64_t)blockIdx.x * kWarps + warp;
  const uint32_t last_round = a.dynamic ? 0u : a.full_rounds;  // dynamic schedule: this warp's one unit is unit `wid`
  for (uint32_t round = 0; round <= last_round; ++round) {
    const uint64_t unit = (uint64_t)round * a.n_warps + wid;
    uint64_t unit_lo, unit_hi;
    unit_range(a, unit, unit_lo, unit_hi);
    if (unit_lo >= a.n_batches) break;  // (warps past the last unit of a dynamic grid rounded up to whole CTAs)
    for (uint64_t slot = unit_lo; slot < unit_hi; ++slot) {
        const int n_ev = batch_len(a, slot);
        if (slot == unit_lo || !rng.next_batch(a, n_ev, lane)) rng.init(a, &smw[warp], a.first_batch + slot, slot, n_ev, lane);

        f2 spm2[5], vars[5], sigma(0.0f), variance(0.0f);
#pragma unroll
        for (int k = 0; k < 5; ++k) spm2[k] = vars[k] = f2(0.0f);
        uint32_t selected = 0;
        // resacc.rs:121-129 for the two events of a lane; `valid` masks the drain's last, half-filled step
        auto integrate = [&](const f2 (&e)[3][4], m2 valid) {
            f2 m[5];
            me_fast<f2>(e, P, m);
            f2 w(0.0f);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                m[k] = select(valid, m[k], f2(0.0f));
                spm2[k] += m[k];
                vars[k] += m[k] * m[k];
                w += m[k] * P.sigma_contribs[k];
            }
            sigma += w;
            variance += w * w;
            selected += (uint32_t)valid.x + (uint32_t)valid.y;
        };
        auto pop = [&](int head, f2 (&e)[3][4]) {  // slots head + lane (low halves) and head + 32 + lane (high halves)
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) e[k][c].v = *reinterpret_cast<const unsigned long long*>(&qf[4 * k + c][head + 2 * lane]);
        };

        int q_head = 0, q_count = 0;  // survivor queue (warp-uniform)
        const int n_it = rng.iterations();
        for (int it = 0; it < n_it; it += 2) {
            Word w0[12], w1[12];
            rng.draws(it, lane, w0);
            const bool have_b = it + 1 < n_it;
            if (have_b) {  // same call sequence as the one-event kernel: refill, then the next iteration's draws
                rng.begin_next(lane);
                rng.template tick<1>(lane); rng.template tick<2>(lane); rng.template tick<3>(lane); rng.template tick<4>(lane);
                rng.template tick<5>(lane); rng.template tick<6>(lane); rng.template tick<7>(lane);
                rng.draws(it + 1, lane, w1);
            } else {
#pragma unroll
                for (int j = 0; j < 12; ++j) w1[j] = w0[j];
            }
            const bool more = it + 2 < n_it;
            if (more) rng.begin_next(lane);
            RngTick<F, RNG> tick{rng, lane, more};
            f2 u[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) u[j] = uniform_x2<RNG>(w0[j], w1[j], (j & 3) == 1);
            f2 p[3][4];
            gen_event<f2, false, false>(u, P.e_total, fm, p, tick);
            const m2 ok = keep_event<f2, false, false>(p, P);
            const bool keep_a = ok.x && rng.event_of(it, lane) >= 0;
            const bool keep_b = ok.y && have_b && rng.event_of(it + 1, lane) >= 0;
            // evcut.rs:42-96 as predicate masks; the survivors of both iterations are compacted into the queue
            const unsigned mask_a = __ballot_sync(0xffffffffu, keep_a), mask_b = __ballot_sync(0xffffffffu, keep_b);
            const unsigned below = (1u << lane) - 1u;
            if (keep_a) {
                const int w = qword((q_head + q_count + __popc(mask_a & below)) & (kQueue2 - 1));
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) qf[4 * k + c][w] = p[k][c].lo();
            }
            q_count += __popc(mask_a);
            if (keep_b) {
                const int w = qword((q_head + q_count + __popc(mask_b & below)) & (kQueue2 - 1));
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) qf[4 * k + c][w] = p[k][c].hi();
            }
            q_count += __popc(mask_b);
            __syncwarp();
            if (q_count >= 64) {
                f2 e[3][4];
                pop(q_head, e);
                __syncwarp();
                q_head = (q_head + 64) & (kQueue2 - 1);
                q_count -= 64;
                integrate(e, m2{true, true});
            }
        }
        if (q_count > 0) {  // drain (warp-uniform condition; fewer than 64 events)
            f2 e[3][4];
            pop(q_head, e);
            // lanes without an event still hold finite momenta of earlier events or zeros: give them a harmless event
            const m2 valid{lane < q_count, 32 + lane < q_count};
            const f2 one(1.0f), zero(0.0f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                e[k][0] = select(valid, e[k][0], k == 0 ? one : zero);
                e[k][1] = select(valid, e[k][1], k == 1 ? one : zero);
                e[k][2] = select(valid, e[k][2], k == 2 ? one : zero);
                e[k][3] = select(valid, e[k][3], one + one);
            }
            integrate(e, valid);
        }

        // ResultsAccumulator of the batch: the two halves, then the xor-shuffle tree over the 32 lane-partials
        float v[12];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            v[k] = spm2[k].lo() + spm2[k].hi();
            v[5 + k] = vars[k].lo() + vars[k].hi();
        }
        v[10] = sigma.lo() + sigma.hi();
        v[11] = variance.lo() + variance.hi();
        uint32_t n = selected;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] += shfl_xor_t(v[k], off);
            n += __shfl_xor_sync(0xffffffffu, n, off);
        }
        if (lane == 0) {
            tp3_acc* o = a.out + slot;
            o->selected_events = n;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                o->spm2[k] = (double)v[k];
                o->vars[k] = (double)v[5 + k];
            }
            o->sigma = (double)v[10];
            o->variance = (double)v[11];
        }
        __syncwarp();
    }
    if (a.fold) fold_publish<float>(a.fold, a.unit_done, a.out, a.dynamic ? ((a.dynamic - 1) | kSchedDynamic) : a.n_warps, a.full_rounds, a.unit_batches, a.taper, a.epoch, a.n_batches, unit, lane);
  }
}

// Parity hook: same streams, same event -> lane mapping, per-event outputs instead of sums.
template <class F, int RNG, bool SORT, bool LITERAL>
__global__ void __launch_bounds__(kThreads) dump_kernel(const SimArgs a, const PhysParams<F> P, const DumpArgs d) {
    __shared__ BlockSmem<F> sm;
    using Word = typename RawWord<F, RNG>::type;
    const int lane = threadIdx.x & 31;
    fastmath_load<kThreads>(&sm.fm);
    __syncthreads();
    if (threadIdx.x >= 32) return;  // one warp = one batch; the dump is for a single batch
    WarpRng<F, RNG> rng;
    rng.init(a, &sm.w[0], a.first_batch, 0, batch_len(a, 0), lane);
    const int n_it = rng.iterations();
    for (int it = 0; it < n_it; ++it) {
        Word w[12];
        rng.draws(it, lane, w);
        if (it + 1 < n_it) {
            rng.begin_next(lane);
            rng.template tick<1>(lane); rng.template tick<2>(lane); rng.template tick<3>(lane); rng.template tick<4>(lane);
            rng.template tick<5>(lane); rng.template tick<6>(lane); rng.template tick<7>(lane);
        }
        const int e = rng.event_of(it, lane);
        if (e < 0 || e >= (int)d.n_events) continue;
        if (d.words)
            for (int j = 0; j < 12; ++j) d.words[(size_t)e * 12 + j] = (uint64_t)w[j];
        if (!d.momenta) continue;
        F p[3][4];
        NoTick no_tick;
        if constexpr (TP3_GEN_FROM_INTS && sizeof(F) == 8 && !LITERAL) {  // as in simulate_kernel
            double d[12];
            for (int j = 0; j < 12; ++j) d[j] = RNG == RNG_RANF ? (double)(int)w[j] : (double)((unsigned long long)w[j] >> 11);
            double rr[3];
            for (int k = 0; k < 3; ++k)
                rr[k] = RNG == RNG_RANF ? (double)((unsigned long long)w[4 * k + 2] * (unsigned long long)w[4 * k + 3]) : d[4 * k + 2] * d[4 * k + 3];
            gen_event_ints<SORT>(d, rr, P.e_total, FastMath{&sm.fm, &P.fc}, p, no_tick);
        } else {
            F u[12];
            for (int j = 0; j < 12; ++j) u[j] = WarpRng<F, RNG>::uniform(w[j], !LITERAL && (j & 3) == 1, P.fc);
            gen_event<F, SORT, LITERAL>(u, P.e_total, FastMath{&sm.fm, &P.fc}, p, no_tick);
        }
        const bool k = keep_event<F, SORT, LITERAL>(p, P);
        F m[5] = {0, 0, 0, 0, 0};
        if (k) {
            if (LITERAL) me_literal<F>(p, P, m);
            else me_fast<F>(p, P, m);
        }
        for (int q = 0; q < 3; ++q)
            for (int c = 0; c < 4; ++c) d.momenta[((size_t)e * 3 + q) * 4 + c] = (double)p[q][c];
        d.kept[e] = k;
        for (int c = 0; c < 5; ++c) d.m2[(size_t)e * 5 + c] = (double)m[c];
    }
}

// Parity hook for the shipped f32 kernel: the per-event outputs of simulate_kernel_x2's PACKED physics (two events per
// lane through gen_event<f2>, keep_event<f2>, me_fast<f2>), same streams and the same pairing of warp iterations.
template <int RNG>
__global__ void __launch_bounds__(32) dump_kernel_x2(const SimArgs a, const PhysParams<f2> P, const DumpArgs d) {
    using F = float;
    using Word = typename RawWord<F, RNG>::type;
    __shared__ WarpSmem<F, kQueue2, 3> smw;
    const int lane = threadIdx.x & 31;
    const FastMath fm{nullptr, nullptr};
    WarpRng<F, RNG> rng;
    rng.init(a, &smw, a.first_batch, 0, batch_len(a, 0), lane);
    const int n_it = rng.iterations();
    for (int it = 0; it < n_it; it += 2) {
        Word w0[12], w1[12];
        rng.draws(it, lane, w0);
        const bool have_b = it + 1 < n_it;
        if (have_b) {
            rng.begin_next(lane);
            rng.template tick<1>(lane); rng.template tick<2>(lane); rng.template tick<3>(lane); rng.template tick<4>(lane);
            rng.template tick<5>(lane); rng.template tick<6>(lane); rng.template tick<7>(lane);
            rng.draws(it + 1, lane, w1);
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) w1[j] = w0[j];
        }
        if (it + 2 < n_it) {
            rng.begin_next(lane);
            rng.template tick<1>(lane); rng.template tick<2>(lane); rng.template tick<3>(lane); rng.template tick<4>(lane);
            rng.template tick<5>(lane); rng.template tick<6>(lane); rng.template tick<7>(lane);
        }
        f2 u[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) u[j] = uniform_x2<RNG>(w0[j], w1[j], (j & 3) == 1);
        f2 p[3][4];
        NoTick no_tick;
        gen_event<f2, false, false>(u, P.e_total, fm, p, no_tick);
        const m2 ok = keep_event<f2, false, false>(p, P);
        f2 m[5];
        me_fast<f2>(p, P, m);
        const int ev[2] = {rng.event_of(it, lane), have_b ? rng.event_of(it + 1, lane) : -1};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = ev[h];
            if (e < 0 || e >= (int)d.n_events) continue;
            if (d.words)
                for (int j = 0; j < 12; ++j) d.words[(size_t)e * 12 + j] = (uint64_t)(h ? w1[j] : w0[j]);
            if (!d.momenta) continue;
            const bool k = h ? ok.y : ok.x;
            for (int q = 0; q < 3; ++q)
                for (int c = 0; c < 4; ++c) d.momenta[((size_t)e * 3 + q) * 4 + c] = (double)(h ? p[q][c].hi() : p[q][c].lo());
            d.kept[e] = k;
            for (int c = 0; c < 5; ++c) d.m2[(size_t)e * 5 + c] = k ? (double)(h ? m[c].hi() : m[c].lo()) : 0.0;
        }
    }
}

// xoshiro seeding kernel: state at the start of each batch of the launch.
//   sequential stream: base advanced by 120000 * batch outputs; faster-threading: batch jump()s.
// digit_polys[k][d] = G^(d * 256^k), G = x^120000 or the jump() polynomial; words per poly = 4.
template <class Lane>
__global__ void xoshiro_seed_kernel(uint64_t first_batch, uint64_t n_batches, const uint64_t* __restrict__ digit_polys,
                                    int n_digits, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint64_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_batches) return;
    const uint64_t batch = first_batch + i;
    Lane g;
    g.s0 = (decltype(g.s0))b0; g.s1 = (decltype(g.s0))b1; g.s2 = (decltype(g.s0))b2; g.s3 = (decltype(g.s0))b3;
    for (int k = 0; k < n_digits; ++k) {
        const unsigned dgt = (unsigned)((batch >> (8 * k)) & 0xffu);
        if (dgt) g.apply(digit_polys + ((size_t)k * 256 + dgt) * 4);
    }
    out[4 * i + 0] = g.s0; out[4 * i + 1] = g.s1; out[4 * i + 2] = g.s2; out[4 * i + 3] = g.s3;
}

// ResultsAccumulator::merge (resacc.rs:133-139) as a strict left fold in batch order (bit-identical to
// the host fold). The only serial part is the chain of additions (one per batch and field); everything
// else is arranged around it: warps 1..7 stage tiles of 256 accumulators in shared memory (coalesced,
// double buffered) while lane k < 12 of warp 0 adds field k of the current tile sequentially and lane 12
// sums the event counts.
constexpr int kMergeThreads = 256, kMergeTile = 224;  // 2 x 224 x 104 B = 46.6 KB of static shared memory
template <class F> __global__ void __launch_bounds__(kMergeThreads) merge_kernel(const tp3_acc* __restrict__ in, uint64_t n, tp3_acc* out,
                                                                                 bool first_chunk) {
    constexpr int kWords = 13;
    __shared__ uint64_t tile[2][kMergeTile * kWords];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t* src = reinterpret_cast<const uint64_t*>(in);
    const uint64_t total = n * kWords;
    auto stage = [&](int buf, uint64_t t, int first, int step) {
        const uint64_t base = t * kMergeTile * kWords;
        for (int i = first; i < kMergeTile * kWords; i += step) {
            const uint64_t idx = base + i;
            tile[buf][i] = idx < total ? src[idx] : 0ull;
        }
    };
    // a later chunk continues the fold from the running accumulator left in `out` by the previous chunk
    uint64_t* dst = reinterpret_cast<uint64_t*>(out);
    F acc = 0;
    uint64_t cnt = 0;
    if (!first_chunk) {
        if (tid < 12) acc = (F)__longlong_as_double((long long)dst[1 + tid]);
        else if (tid == 12) cnt = dst[0];
    }
    const uint64_t n_tiles = (n + kMergeTile - 1) / kMergeTile;
    stage(0, 0, tid, kMergeThreads);
    __syncthreads();
    for (uint64_t t = 0; t < n_tiles; ++t) {
        const int cur = (int)(t & 1);
        if (tid >= 32) {
            if (t + 1 < n_tiles) stage(cur ^ 1, t + 1, tid - 32, kMergeThreads - 32);
        } else {
            const int m = (int)min((uint64_t)kMergeTile, n - t * kMergeTile);
            if (lane < 12) {
                const uint64_t* col = &tile[cur][1 + lane];
                int b = 0;
                if (t == 0 && first_chunk) {  // the fold starts FROM the first accumulator (sequential.rs:24-26)
                    acc = (F)__longlong_as_double((long long)col[0]);
                    b = 1;
                }
                if (b == 0 && m == kMergeTile) {
#pragma unroll 32
                    for (int i = 0; i < kMergeTile; ++i) acc += (F)__longlong_as_double((long long)col[i * kWords]);
                } else {
                    for (; b < m; ++b) acc += (F)__longlong_as_double((long long)col[b * kWords]);
                }
            } else if (lane == 12) {
                for (int b = 0; b < m; ++b) cnt += tile[cur][b * kWords];
            }
        }
        __syncthreads();
    }
    if (tid < 12) dst[1 + tid] = (uint64_t)__double_as_longlong((double)acc);
    else if (tid == 12) dst[0] = cnt;
}

// Parity hook for the hand-written FP64 functions (fastmath.cuh): out[i] = f_which(in[i]).
__global__ void fastmath_probe_kernel(int which, uint32_t n, const double* __restrict__ in, double* __restrict__ out, const FastCoef fc) {
    __shared__ FastMathSmem fms;
    fastmath_load<256>(&fms);  // (launched with 256 threads per CTA, api.cu)
    const FastMath fm{&fms, &fc};
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = in[i];
        double a = 0, b = 0;
        switch (which) {
            case 0: a = fast_neg_log(x, fm); break;
            case 1: fast_sincos_256(256.0 * x, fm, a, b); break;
            case 2: fast_sincos_256(256.0 * x, fm, b, a); break;
            case 3: a = fast_sqrt(x); break;
            case 4: a = fast_rcp(x); break;
            case 5: fast_sqrt_rsqrt(x, b, a); break;
            case 6: fast_sqrt_rsqrt(x, a, b); break;
            case 7: a = mufu_rcp(x); break;
            case 8: a = mufu_rsqrt(x); break;
            case 9: a = u32_times((uint32_t)x, 1e-9); break;
            case 10: a = u32_times((uint32_t)x, 256e-9); break;
        }
        out[i] = a;
    }
}

// Peak probes: 8 independent FMA chains per thread, x = x * a + 1 with the addend an IMMEDIATE.  Round 1 passed both a
// and b in registers: DFMA R, R, Ra.reuse, Rb.reuse depends on the operand reuse cache, which does not survive a switch to
// another warp, so part of the instructions read three register pairs (3 cycles instead of 2) and
// the probe read 34.2 TFLOP/s.  With two register sources the FP64 pipe issues every 2.0 cycles per sub-partition:
// 37.1 TFLOP/s at 1965 MHz, the nominal 148 SMs x 64 lanes x 2 flop (scripts/micro/peak64.cu, profiles/r02_peak64_probe.txt).
template <class F> __global__ void fma_probe_kernel(F* out, int iters, F a) {
    F x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma_t(x0, a, (F)1); x1 = fma_t(x1, a, (F)1); x2 = fma_t(x2, a, (F)1); x3 = fma_t(x3, a, (F)1);
            x4 = fma_t(x4, a, (F)1); x5 = fma_t(x5, a, (F)1); x6 = fma_t(x6, a, (F)1); x7 = fma_t(x7, a, (F)1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace tp3

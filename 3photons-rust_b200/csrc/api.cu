// C ABI (include/tp3.h) over the sm_100a kernels: context, seeding tables, launches.
// No CPU fallback lives here: without a usable device every compute entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jump_tables.hpp"
#include "kernels.cuh"
#include "faster_evgen.cuh"
#include "fe_scan.cuh"
#include "fe_scan_xo.cuh"
#include "fe_stream.cuh"

using namespace tp3;

namespace {

thread_local std::string g_create_error;

struct DeviceSlot {
    int dev = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint32_t* d_ranf_table = nullptr;
    tp3_acc* d_batch_parts = nullptr;  // [n * parts] part accumulators of a launch with batch_parts > 1
    size_t batch_parts_cap = 0;
    bool fold_in_kernel = false;       // the last launch folded inside the simulation kernel (else: merge_kernel behind it)
    uint64_t* d_xo_digit_polys = nullptr;
    uint64_t* d_xo_lane_polys = nullptr;
    uint64_t* d_xo_states = nullptr;
    size_t xo_states_cap = 0;
    tp3_acc* d_out = nullptr;
    size_t out_cap = 0;
    FoldState* d_fold = nullptr;           // in-kernel ordered fold of the last launch (kernels.cuh)
    uint32_t* d_unit_done = nullptr;       // per-unit completion marks, compared with `epoch`
    size_t unit_done_cap = 0;
    uint32_t epoch = 0;
    unsigned long long* d_hist_counts = nullptr;  // per-event observables: [TP3_HIST_OBSERVABLES][hist_bins]
    double* d_hist_weights = nullptr;
    uint32_t* d_fe_ranf_states = nullptr;  // faster-evgen: [n][57] batch start states from the host scheduler
    size_t fe_states_cap = 0;
    // faster-evgen scan workspace (fe_scan.cuh), kept between calls: grows to the largest pass, freed with the context
    void* d_fe_bnd = nullptr;
    size_t fe_bnd_cap = 0;
    uint64_t* d_fe_maps = nullptr;
    size_t fe_maps_cap = 0;
    uint8_t *d_fe_seg_exit = nullptr, *d_fe_seg_state = nullptr;
    uint32_t* d_fe_seg_count = nullptr;
    uint64_t* d_fe_seg_events = nullptr;
    size_t fe_seg_cap = 0;
    // faster-evgen + xoshiro scan workspace (fe_scan_xo.cuh)
    uint32_t* d_xo_scan = nullptr;  // exit A | exit B | count | mismatch flag
    size_t xo_scan_cap = 0;
    void* d_xo_bnd = nullptr;
    size_t xo_bnd_cap = 0;
    // faster-evgen stream pipeline (fe_stream.cuh): two sets of grow-only workspaces, so that the walk of pass k + 1
    // (integer work) runs next to the physics of pass k (FP64 work) on a second stream
    struct FsBuf {
        FeRecord* d_records = nullptr;
        size_t records_cap = 0;
        uint32_t* d_count = nullptr;     // per segment: events started
        uint8_t *d_exit = nullptr, *d_fail = nullptr;
        uint64_t* d_seg_events = nullptr;
        uint32_t* d_redo_list = nullptr;
        uint8_t* d_redo_entry = nullptr;
        size_t seg_cap = 0;
        uint32_t* d_unit_seg = nullptr;
        tp3_acc* d_parts = nullptr;
        size_t unit_cap = 0;
        cudaEvent_t phys_done = nullptr; // the physics kernel that last read this buffer's records
        bool phys_pending = false;
        // host side of the pass in flight on this buffer
        std::vector<uint32_t> h_count, h_unit_seg;
        std::vector<uint8_t> h_exit, h_fail;
        std::vector<uint64_t> h_seg_events;
        uint64_t first_round = 0, first_events = 0, n_seg = 0;
        uint32_t seg_rounds = 0;
        bool count_only = false;
    } fs[2];
    cudaStream_t fs_stream2 = nullptr;
    cudaEvent_t fs_event = nullptr;
    // last launch
    uint64_t last_first = 0, last_n = 0;
    SimArgs last_args;                       // its schedule (fused default kernel)
    cudaStream_t copy_stream = nullptr;      // tp3_simulate_batches: accumulators go to the host while the kernel runs
    cudaEvent_t copy_event = nullptr, copy_event2 = nullptr;
    unsigned long long* h_progress = nullptr;  // pinned
};

}  // namespace

struct tp3_ctx {
    tp3_params params;
    std::vector<DeviceSlot> devs;
    std::string err;
    uint64_t launches = 0;
    uint32_t hist_bins = 0;  // per-event observables on (tp3_histograms_enable)
    // tp3_set_option: test / A-B switches, read once here instead of from the environment on every launch
    int64_t opt_unit_batches = 0;    // consecutive batches per scheduling unit (0 = by launch size)
    int64_t opt_grid_warps = 0;      // warps in the grid (0 = what the device holds at once)
    int64_t opt_batch_parts = 1;     // parts per batch (1, 2, 5, 10; 0 = as many of those as still fit one wave of warps): small runs
                                     //    then fill the device; the per-batch sums depend on it in their last bits (tp3.h)
    int64_t opt_taper_units = 0;     // dynamic schedule: units per taper stage (-1 = no taper, 0 = one wave)
    int64_t opt_tail_singles = 0;    // dynamic schedule: single-batch units at the end of the launch (0 = auto)
    int64_t opt_align_units = 1;     // dynamic schedule: the big units end on a multiple of 4 x SMs units
    int64_t opt_ramp_units = 0;      // dynamic schedule: ramp units of 1, 2, .., 8 batches at the head of the launch (A/B only)
    int64_t opt_sched_dynamic = 1;   // 1 (default): one unit per warp, dispatched by the hardware in unit order; 0: static balanced schedule (kernels.cuh)
    int64_t opt_f32_scalar = 0;      // f32: one event per lane instead of the packed two-events-per-lane kernel
    int64_t opt_fe_split = 0;        // faster-evgen: 1 = one thread per batch, 32 = one lane per 313 events, 0 = by launch size
    int64_t opt_fe_host_scan = 0;    // faster-evgen: batch start states from the host walk (the reference's method; cross-check)
    int64_t opt_fe_xo_seg_units = 0; // faster-evgen + xoshiro scan: segment length in units of 2048 outputs (0 = by launch size)
    int64_t opt_fe_timing = 0;       // print the scan phases of every call to stderr
    int64_t stat_fe_xo_pass_b = 0;   // tp3_get_stat: pass-B repetitions of the last xoshiro scan
    int64_t opt_fe_legacy = 0;       // faster-evgen + RANF: the round-1 pipeline (scan over transition maps + one lane per 313 events)
    int64_t opt_fe_pass_segments = 0;// faster-evgen stream pipeline: segments per pass (0 = one full wave of lanes)
    int64_t opt_fe_seg_rounds = 0;   // ... rounds per segment (0 = by pass size, <= 1024)
    int64_t opt_fe_warm = 0;         // ... warm-up rounds before a segment (0 = kFeWarm); small values exercise the redo path
    int64_t opt_fe_serial = 0;       // ... 1: passes one after the other on the slot's stream (no walk / physics overlap)
    int64_t stat_fe_passes = 0, stat_fe_redone = 0;  // passes and redone segments of the last call
    // faster-evgen stream pipeline: a round (segment boundary of an earlier pass) whose absolute event index is known
    uint64_t fs_round = 0, fs_events = 0;
    // host copies of the seeding data
    uint32_t ranf_base[kRanfLag];
    std::vector<uint32_t> ranf_table;
    std::vector<uint64_t> xo_digit_polys;   // [n_digits][256][4]
    std::vector<uint64_t> xo_lane_polys;    // [32][4]
    uint64_t xo_base[4];
    int xo_digits = 0;
    // faster-evgen, sequential stream: the host scheduler's generator, positioned at the start of batch fe_pos
    uint64_t fe_pos = 0;
    uint32_t fe_ranf[56];
    int fe_ranf_index = 55;
    uint64_t fe_xo[4];
    bool fe_ready = false;
    // faster-evgen, RANF, sequential stream: where the device scan stands (always at a round start)
    uint64_t scan_round = 0, scan_events = 0;
    int scan_state = 0;
};

namespace {

#define TP3_CUDA(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return TP3_E_CUDA;                                                                           \
        }                                                                                                \
    } while (0)


template <class F> struct ScalarOf { using type = F; };
template <> struct ScalarOf<f2> { using type = float; };

template <class F> PhysParams<F> phys_params(const tp3_params& p) {
    using S = typename ScalarOf<F>::type;  // host arithmetic in the run's Float (this file is built with -ffp-contract=off)
    PhysParams<F> q;
    q.e_total = (F)p.e_total;
    q.acut = (F)p.acut;
    q.bcut = (F)p.bcut;
    q.e_min = (F)p.e_min;
    q.sincut = (F)p.sincut;
    q.g_a = (F)p.g_a;
    q.g_beta_p = (F)p.g_beta_p;
    q.g_beta_m = (F)p.g_beta_m;
    for (int k = 0; k < 5; ++k) q.sigma_contribs[k] = (F)p.sigma_contribs[k];
    {
        const S e = (S)p.e_total, ga = (S)p.g_a, gp = (S)p.g_beta_p, gm = (S)p.g_beta_m, e2 = e * e;
        q.k_m0 = (F)((ga * ga) * (S)8);
        q.k_m1 = (F)((gp * gp) * (((S)8 * e2) * e2));
        q.k_m2 = (F)((gm * gm) * (((S)4 * e2) * e2));
        q.k_mix = (F)((-(ga * gp)) * e2);
        q.omb_e = (F)(((S)1 - (S)p.bcut) / e);
        q.he = (F)((S)0.5 * e);
    }
    q.fc = FastCoef TP3_FAST_COEF_INIT;
    if (p.flags & TP3_STANDARD_RANDOM) {
        // xoshiro256+: the stream integer is d = x >> 11 and the uniform d 2^-53 (rand 0.8.5 Standard): cos_theta = 2u - 1 =
        // fma(d, 2^-52, -1), 256 u = d 2^-45, r r' = (d d') 2^-106 -- all exact rescalings, so gen_event_ints computes the
        // reference's own values with one multiplication per uniform less
        q.fc.u_scale2 = 0x1p-52;
        q.fc.u_scale_sq = 0x1p-106;
        q.fc.phi_scale = 0x1p-45;
    }
    return q;
}

// What a launcher needs to lay out the static schedule of kernels.cuh.
struct Sched {
    int sm_count;
    int64_t unit_batches;  // 0 = auto
    int64_t grid_warps;    // 0 = auto
    bool stream_continues; // sequential RANF: a warp's next batch continues the stream, so units of several batches pay
    int64_t dynamic;       // 1: one unit per warp, dispatched by the hardware (kernels.cuh)
    SimArgs* filled;       // out (may be null): the schedule the launcher chose
    int64_t ramp_units;    // 0 = none
    int64_t taper_units;   // -1 = none, 0 = auto (one wave per stage), > 0: units per taper stage
    int64_t tail_singles;  // 0 = auto, > 0: single-batch units at the end of the launch (at least)
    bool align_units;      // the big units end on a multiple of 4 x SMs units (see fill_schedule)
};

// Fill the schedule fields of `a` for a kernel that runs `warps`-warp CTAs, `ctas_per_sm` of them per SM.
cudaError_t fill_schedule(SimArgs& a, int warps, int ctas_per_sm, const Sched& sc, uint64_t max_unit = 8) {
    if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
    uint64_t W = (uint64_t)sc.sm_count * ctas_per_sm * warps;  // what the device holds at once
    if (sc.grid_warps > 0) W = ((uint64_t)sc.grid_warps + warps - 1) / warps * warps;
    if (sc.dynamic) {
        // Units in batch order, one warp each: big units of `unit` batches, then the tail described below.
        // Big units: the largest power of two <= max_unit that keeps a big unit shorter than half of what the launch takes (fixed
        // units of 8 made a launch of 12 000 batches take 4.3 instead of 3.4 ms).  max_unit is 16 for the f64 kernel -- every unit pays
        // one jump-ahead and one CTA turnover, 10-20 us: 0.15 % at 1e6 batches -- and 8 for the f32 kernels, whose launches from
        // 125 000 batches up came out 0.2-0.3 ms LONGER with units of 16 (profiles/r02_tail_scan.txt).
        uint64_t unit = 1;
        if (sc.stream_continues)
            while (unit < max_unit && 2 * unit * 2 * W <= a.n_batches) unit *= 2;
        if (sc.unit_batches > 0) unit = (uint64_t)sc.unit_batches;
        // ramp: the first wave's warps start together; unequal first units (1, 2, .., 8 batches) take them out of step at once
        // (measured: it does not pay, profiles/r02_schedule_ab.txt: the first wave drifts apart quickly enough on its own; off by default)
        uint64_t ramp = sc.ramp_units > 0 ? (uint64_t)sc.ramp_units / 8 * 8 : 0;
        if (ramp / 8 * 36 > a.n_batches) ramp = 0;
        const uint64_t ramp_batches = ramp / 8 * 36;
        // Tail (profiles/r02_tail_scan.txt, r02_unit_trace.txt).  The warps leave the big units at different times -- the starts of
        // 2368 consecutive units are spread over 4 of the 5 ms a big unit takes -- and what follows has to absorb that spread and
        // end with single batches, so that the device drains within one batch time.  With room for it: half a wave of half units,
        // half a wave of quarter units, then two waves of single batches (3 W units for 5 W batches; four waves of single batches,
        // the first version, are 4 W units for 4 W batches and measured 0.06-0.15 ms slower at the 125 000 batches one GPU gets of
        // the 1e10-event run on eight).  Short launches: single batches for the last four waves.
        const uint64_t left = a.n_batches - ramp_batches;
        uint64_t taper = 0, singles = std::min<uint64_t>(left, 4 * W);
        if (unit >= 4 && unit % 4 == 0 && sc.taper_units >= 0) {
            const uint64_t t = sc.taper_units > 0 ? (uint64_t)sc.taper_units : W / 2;
            const uint64_t s1 = sc.tail_singles > 0 ? (uint64_t)sc.tail_singles : 2 * W;
            const uint64_t need = t * (unit / 2 + unit / 4) + s1;
            if (left >= 2 * need) {
                taper = t;
                singles = s1;
            }
        } else if (sc.tail_singles > 0) {
            singles = std::min<uint64_t>(left, (uint64_t)sc.tail_singles);
        }
        const uint64_t taper_batches = taper * (unit / 2 + unit / 4);
        uint64_t big = unit > 1 ? (left - taper_batches - singles) / unit : 0;
        // ALIGNMENT (profiles/r02_unit_trace.txt).  The four one-warp CTAs that share a sub-partition start a launch together,
        // the scheduler serves them strictly in order, and from then on they stay 1.2 ms apart: in steady state 4 x SMs consecutive
        // units -- one per sub-partition of the device -- start within 0.3 ms of each other, four such bands per 5 ms wave.  If the
        // big units end inside a band, half of an SM's sub-partitions hold a big unit more than the others, nothing evens that
        // out (a new CTA goes where a slot frees), and the launch takes 0.1-0.2 ms longer: the run time as a function of
        // the launch size alternates with a period of 4 x SMs big units (profiles/r02_tail_scan.txt).  So the big units end on a
        // band boundary and the batches this frees (< 4 x SMs units' worth) run as single batches.
        if (warps == 1 && sc.align_units && ramp == 0) {
            const uint64_t M = 4ull * (uint64_t)sc.sm_count;
            if (big >= 2 * M) big -= big % M;
        }
        const uint64_t units = ramp + big + 2 * taper + (left - taper_batches - big * unit);
        a.taper = (uint32_t)taper;
        a.dynamic = 1 + (uint32_t)ramp;
        a.unit_batches = (uint32_t)unit;
        a.full_rounds = (uint32_t)big;
        a.n_warps = (uint32_t)((units + warps - 1) / warps * warps);
        if (sc.filled) *sc.filled = a;
        return cudaSuccess;
    }
    if (a.n_batches < W) W = (a.n_batches + warps - 1) / warps * warps;  // one batch per warp, no full round
    uint64_t unit = 1;
    if (sc.stream_continues) {
        unit = a.n_batches / (W * 4);  // at least ~4 rounds, so that the ordered fold overlaps the simulation
        unit = unit < 1 ? 1 : unit > 8 ? 8 : unit;
    }
    if (sc.unit_batches > 0) unit = (uint64_t)sc.unit_batches;
    a.n_warps = (uint32_t)W;
    a.unit_batches = (uint32_t)unit;
    a.full_rounds = (uint32_t)(a.n_batches / (W * unit));
    if (sc.filled) *sc.filled = a;
    return cudaSuccess;
}
// Resident CTAs per SM of a kernel (asked once per kernel and dynamic shared memory size).
template <class K> int ctas_per_sm(K kernel, int threads, size_t dyn) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn) != cudaSuccess) return 0;
    return n;
}

template <class F, int RNG, bool SORT, bool LITERAL>
cudaError_t launch_sim(SimArgs a, const tp3_params& p, cudaStream_t st, const Sched& sc) {
    constexpr int warps = sim_warps(LITERAL, false);
    auto kernel = simulate_kernel<F, RNG, SORT, LITERAL>;
    static const int occ = ctas_per_sm(kernel, 32 * warps, 0);
    if (cudaError_t e = fill_schedule(a, warps, occ, sc, sizeof(F) == 8 && !LITERAL ? 16 : 8)) return e;
    kernel<<<a.n_warps / warps, 32 * warps, 0, st>>>(a, phys_params<F>(p));
    return cudaGetLastError();
}
// Fast kernel with the per-event observable epilogue: the CTA's event counts in dynamic shared memory (4 bytes per bin).
template <class F, int RNG, bool SORT>
cudaError_t launch_sim_hist(SimArgs a, const tp3_params& p, cudaStream_t st, const Sched& sc) {
    const size_t dyn = (size_t)TP3_HIST_OBSERVABLES * a.hist_bins * sizeof(uint32_t);
    auto kernel = simulate_kernel<F, RNG, SORT, false, true>;
    if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)) return e;
    if (cudaError_t e = fill_schedule(a, kWarps, ctas_per_sm(kernel, kThreads, dyn), sc)) return e;
    kernel<<<a.n_warps / kWarps, kThreads, dyn, st>>>(a, phys_params<F>(p));
    return cudaGetLastError();
}
template <class F, int RNG, bool SORT, bool LITERAL>
void launch_dump(const SimArgs& a, const tp3_params& p, const DumpArgs& d, cudaStream_t st) {
    dump_kernel<F, RNG, SORT, LITERAL><<<1, kThreads, 0, st>>>(a, phys_params<F>(p), d);
}

using SimFn = cudaError_t (*)(SimArgs, const tp3_params&, cudaStream_t, const Sched&);
using DumpFn = void (*)(const SimArgs&, const tp3_params&, const DumpArgs&, cudaStream_t);

template <class F, int RNG> SimFn pick_sim2(bool sort, bool literal) {
    if (sort) return literal ? launch_sim<F, RNG, true, true> : launch_sim<F, RNG, true, false>;
    return literal ? launch_sim<F, RNG, false, true> : launch_sim<F, RNG, false, false>;
}
template <class F, int RNG> DumpFn pick_dump2(bool sort, bool literal) {
    if (sort) return literal ? launch_dump<F, RNG, true, true> : launch_dump<F, RNG, true, false>;
    return literal ? launch_dump<F, RNG, false, true> : launch_dump<F, RNG, false, false>;
}
// f32 fast kernel, two events per lane in packed FP32 arithmetic (f32x2.cuh)
template <int RNG> cudaError_t launch_sim_x2(SimArgs a, const tp3_params& p, cudaStream_t st, const Sched& sc) {
    constexpr int warps = x2_warps(RNG);
    auto kernel = simulate_kernel_x2<RNG>;
    static const int occ = ctas_per_sm(kernel, 32 * warps, 0);
    if (cudaError_t e = fill_schedule(a, warps, occ, sc)) return e;
    kernel<<<a.n_warps / warps, 32 * warps, 0, st>>>(a, phys_params<f2>(p));
    return cudaGetLastError();
}
template <class F, int RNG> SimFn pick_sim_hist2(bool sort) { return sort ? launch_sim_hist<F, RNG, true> : launch_sim_hist<F, RNG, false>; }
SimFn pick_sim(const tp3_params& p, bool hist, bool f32_scalar) {
    const bool f32 = p.flags & TP3_F32, xo = p.flags & TP3_STANDARD_RANDOM;
    const bool sort = !(p.flags & TP3_NO_PHOTON_SORTING), lit = p.kernel == TP3_KERNEL_LITERAL;
    if (hist) {
        if (f32) return xo ? pick_sim_hist2<float, RNG_XOSHIRO>(sort) : pick_sim_hist2<float, RNG_RANF>(sort);
        return xo ? pick_sim_hist2<double, RNG_XOSHIRO>(sort) : pick_sim_hist2<double, RNG_RANF>(sort);
    }
    if (f32 && !lit && !f32_scalar)  // (the one-event-per-lane f32 kernel stays for A/B runs and tests: tp3_set_option)
        return xo ? launch_sim_x2<RNG_XOSHIRO> : launch_sim_x2<RNG_RANF>;
    if (f32) return xo ? pick_sim2<float, RNG_XOSHIRO>(sort, lit) : pick_sim2<float, RNG_RANF>(sort, lit);
    return xo ? pick_sim2<double, RNG_XOSHIRO>(sort, lit) : pick_sim2<double, RNG_RANF>(sort, lit);
}
template <int RNG> void launch_dump_x2(const SimArgs& a, const tp3_params& p, const DumpArgs& d, cudaStream_t st) {
    dump_kernel_x2<RNG><<<1, 32, 0, st>>>(a, phys_params<f2>(p), d);
}
DumpFn pick_dump(const tp3_params& p, bool f32_scalar) {
    const bool f32 = p.flags & TP3_F32, xo = p.flags & TP3_STANDARD_RANDOM;
    const bool sort = !(p.flags & TP3_NO_PHOTON_SORTING), lit = p.kernel == TP3_KERNEL_LITERAL;
    // the shipped f32 kernel is the packed one: its per-event dump goes through the same packed physics (unsorted photons)
    if (f32 && !lit && !f32_scalar) return xo ? launch_dump_x2<RNG_XOSHIRO> : launch_dump_x2<RNG_RANF>;
    if (f32) return xo ? pick_dump2<float, RNG_XOSHIRO>(sort, lit) : pick_dump2<float, RNG_RANF>(sort, lit);
    return xo ? pick_dump2<double, RNG_XOSHIRO>(sort, lit) : pick_dump2<double, RNG_RANF>(sort, lit);
}

// Seeding tables for the xoshiro streams (host, once per context).
void build_xoshiro_tables(tp3_ctx* c) {
    const bool f32 = c->params.flags & TP3_F32;
    const bool jump = c->params.flags & TP3_FASTER_THREADING;
    const int deg = f32 ? 128 : 256;
    Gf2Mod mod;
    if (f32) {
        Xoshiro128 g = xoshiro128_seed(12345);  // standard.rs:23
        for (int i = 0; i < 4; ++i) c->xo_base[i] = g.s[i];
        mod = xoshiro_min_poly(g, 128);
    } else {
        Xoshiro256 g = xoshiro256_seed(12345);
        for (int i = 0; i < 4; ++i) c->xo_base[i] = g.s[i];
        mod = xoshiro_min_poly(g, 256);
    }
    // per-batch generator: G = x^(12 * 10000) (sequential stream) or jump() = x^(2^(deg/2)); under sequential faster-evgen
    // batches have no fixed stride and the tables step by kXoSegUnit outputs for the scan of fe_scan_xo.cuh instead
    const bool seq_faster = (c->params.flags & TP3_FASTER_EVGEN) && !jump;
    Gf2Poly G = jump ? gf2_x_pow(1, deg / 2, mod)
                     : gf2_x_pow(seq_faster ? (uint64_t)kXoSegUnit : (uint64_t)kDrawsPerEvent * kBatch, 0, mod);
    c->xo_digits = 5;  // 2^40 batches
    c->xo_digit_polys.assign((size_t)c->xo_digits * 256 * 4, 0);
    Gf2Poly unit = G;
    for (int k = 0; k < c->xo_digits; ++k) {
        Gf2Poly cur;
        cur.w[0] = 1;
        for (int d = 0; d < 256; ++d) {
            std::memcpy(&c->xo_digit_polys[((size_t)k * 256 + d) * 4], cur.w, 32);
            cur = gf2_mul_mod(cur, unit, mod);
        }
        unit = cur;
    }
    c->xo_lane_polys.assign((size_t)32 * 4, 0);
    for (int l = 0; l < 32; ++l) {  // lane l starts kLaneEvents * l events into the batch
        Gf2Poly p = gf2_x_pow((uint64_t)kDrawsPerEvent * kLaneEvents * l, 0, mod);
        std::memcpy(&c->xo_lane_polys[(size_t)l * 4], p.w, 32);
    }
}


// ---- faster-evgen: the scheduler's pre-advance of the master generator (evgen.rs:257-267) -------------
// The reference's own method, kept on the host as the CROSS-CHECK of the GPU scans (fe_scan.cuh for RANF,
// fe_scan_xo.cuh for xoshiro), which supply the start states in every run: this code is reached only with the test
// option `fe_host_scan` set.  It is what the reference's scheduler thread does between spawning batch tasks
// (multi_threading.rs:59-64); the accept / re-roll test is evaluated exactly as the reference does (no FMA, run's Float).
template <class F> struct FeHostRanf {
    uint32_t* n;  // numbers[0..55]
    int& index;
    void reset() {
        for (int i = 1; i < 25; ++i) { int32_t v = (int32_t)n[i] - (int32_t)n[i + 31]; n[i] = (uint32_t)(v < 0 ? v + (int32_t)RANF_MOD : v); }
        for (int i = 25; i < 56; ++i) { int32_t v = (int32_t)n[i] - (int32_t)n[i - 24]; n[i] = (uint32_t)(v < 0 ? v + (int32_t)RANF_MOD : v); }
    }
    template <int N> void take(F* out) {
        if (index < N) { reset(); index = 55; }
        index -= N;
        for (int i = 0; i < N; ++i) out[i] = (F)(int32_t)n[index + 1 + i] * (F)1e-9;
    }
};
template <class F> struct FeHostXo;
template <> struct FeHostXo<double> {
    Xoshiro256 g;
    template <int N> void take(double* out) {
        for (int i = 0; i < N; ++i) { const uint64_t r = g.s[0] + g.s[3]; g.step(); out[i] = (double)(r >> 11) * (1.0 / 9007199254740992.0); }
    }
};
template <> struct FeHostXo<float> {
    Xoshiro128 g;
    template <int N> void take(float* out) {
        for (int i = 0; i < N; ++i) { const uint32_t r = g.s[0] + g.s[3]; g.step(); out[i] = (float)(r >> 8) * (1.0f / 16777216.0f); }
    }
};
template <class F, class Gen> void fe_skip_events(Gen& gen, uint64_t n_events) {
    for (uint64_t e = 0; e < n_events; ++e) {
        F u9[9], v6[6];
        gen.template take<9>(u9);
        gen.template take<6>(v6);
        for (int k = 0; k < 3; ++k) {
            volatile F x = (F)2 * v6[k] - (F)1, y = (F)2 * v6[3 + k] - (F)1;
            volatile F xx = x * x, yy = y * y;
            F r2 = xx + yy;
            while (r2 > (F)1) {
                F w[2];
                gen.template take<2>(w);
                x = (F)2 * w[0] - (F)1;
                y = (F)2 * w[1] - (F)1;
                xx = x * x;
                yy = y * y;
                r2 = xx + yy;
            }
        }
    }
}

// Start states of batches [first, first + n) of the sequential stream -> host arrays.
void fe_host_states(tp3_ctx* c, uint64_t first, uint64_t n, std::vector<uint32_t>& ranf_out, std::vector<uint64_t>& xo_out) {
    const bool f32 = c->params.flags & TP3_F32, xo = c->params.flags & TP3_STANDARD_RANDOM;
    if (!c->fe_ready || c->fe_pos > first) {  // (re)start from the seeded generator
        if (xo) {
            for (int i = 0; i < 4; ++i) c->fe_xo[i] = c->xo_base[i];
        } else {
            c->fe_ranf[0] = 0;
            ranf_seed_state(RANF_DEFAULT_SEED, c->fe_ranf + 1);
            c->fe_ranf_index = 55;
        }
        c->fe_pos = 0;
        c->fe_ready = true;
    }
    auto skip = [&](uint64_t events) {
        if (xo) {
            if (f32) { FeHostXo<float> g; for (int i = 0; i < 4; ++i) g.g.s[i] = (uint32_t)c->fe_xo[i]; fe_skip_events<float>(g, events); for (int i = 0; i < 4; ++i) c->fe_xo[i] = g.g.s[i]; }
            else { FeHostXo<double> g; for (int i = 0; i < 4; ++i) g.g.s[i] = c->fe_xo[i]; fe_skip_events<double>(g, events); for (int i = 0; i < 4; ++i) c->fe_xo[i] = g.g.s[i]; }
        } else {
            if (f32) { FeHostRanf<float> g{c->fe_ranf, c->fe_ranf_index}; fe_skip_events<float>(g, events); }
            else { FeHostRanf<double> g{c->fe_ranf, c->fe_ranf_index}; fe_skip_events<double>(g, events); }
        }
    };
    while (c->fe_pos < first) { skip(TP3_EVENT_BATCH_SIZE); ++c->fe_pos; }
    if (xo) xo_out.resize(n * 4); else ranf_out.resize(n * 57);
    for (uint64_t b = 0; b < n; ++b) {
        if (xo) std::memcpy(&xo_out[b * 4], c->fe_xo, 32);
        else { std::memcpy(&ranf_out[b * 57], c->fe_ranf, 56 * 4); ranf_out[b * 57 + 56] = (uint32_t)c->fe_ranf_index; }
        skip(TP3_EVENT_BATCH_SIZE);  // every batch but the last of a run is full; a short last batch ends the run
        ++c->fe_pos;
    }
}


// ---- faster-evgen + RANF: start states by a scan over per-round transition maps (fe_scan.cuh) --------------------
// Fills s.d_fe_ranf_states for batches [first, first + n) of the sequential stream, on the device: the host only
// chains one small map per segment (64 to 2048 rounds) between two kernels of a pass.
// `split` boundaries per batch (1, or 32 with part_len 313): n * split generator states in batch-major order.
int fe_device_states_ranf(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t split) {
    const bool f32 = c->params.flags & TP3_F32;
    const uint32_t part_len = split == 1 ? (uint32_t)TP3_EVENT_BATCH_SIZE : (uint32_t)kLaneEvents;
    const uint64_t n_bnd = n * split;
    const bool timing = c->opt_fe_timing != 0;
    auto now = []() { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    double t_maps = 0, t_chain = 0, t_bnd = 0, t_states = 0;
    int passes = 0;
    if (first * (uint64_t)TP3_EVENT_BATCH_SIZE < c->scan_events || c->scan_round == 0) {
        c->scan_round = 0;
        c->scan_events = 0;
        c->scan_state = 0;
    }
    if (s.fe_states_cap < n_bnd) {
        if (s.d_fe_ranf_states) TP3_CUDA(c, cudaFree(s.d_fe_ranf_states));
        s.d_fe_ranf_states = nullptr;
        s.fe_states_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_fe_ranf_states, n_bnd * 57 * sizeof(uint32_t)));
        s.fe_states_cap = n_bnd;
    }
    if (s.fe_bnd_cap < n_bnd) {
        cudaFree(s.d_fe_bnd);
        s.d_fe_bnd = nullptr;
        s.fe_bnd_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_fe_bnd, n_bnd * sizeof(FeBoundary)));
        s.fe_bnd_cap = n_bnd;
    }
    FeBoundary* d_bnd = static_cast<FeBoundary*>(s.d_fe_bnd);
    // event index of the last boundary wanted, and of the first one a following call for the next batches would want
    const uint64_t last_target = (first + n - 1) * (uint64_t)TP3_EVENT_BATCH_SIZE + (uint64_t)(split - 1) * part_len;
    const uint64_t next_first = (first + n) * (uint64_t)TP3_EVENT_BATCH_SIZE;
    int rc = TP3_OK;
    auto check = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == TP3_OK) {
            c->err = std::string(what) + ": " + cudaGetErrorString(e);
            rc = TP3_E_CUDA;
        }
        return e == cudaSuccess;
    };
    // a boundary is found when the running event count passes it: scan until the count exceeds last_target
    while (rc == TP3_OK) {
        const uint64_t remaining = last_target + 1 - c->scan_events;
        uint64_t n_rounds = (uint64_t)((double)remaining * 0.325) + 8192;  // 0.3206 rounds per event (16.64 numbers + 3.1 discarded per round)
        if (n_rounds > (1ull << 28)) n_rounds = 1ull << 28;                 // 2 GB of maps per pass
        // one lane per segment: enough segments to fill the device (~8 warps per scheduler), as long as possible otherwise
        uint32_t seg_rounds = 64;
        while (seg_rounds < (uint32_t)kFeMaxSegRounds && n_rounds / (2 * seg_rounds) >= (uint64_t)s.sm_count * 4 * 8 * 32) seg_rounds *= 2;
        const uint64_t n_seg = (n_rounds + seg_rounds - 1) / seg_rounds;
        if (s.fe_maps_cap < n_rounds) {
            cudaFree(s.d_fe_maps);
            s.d_fe_maps = nullptr;
            s.fe_maps_cap = 0;
            if (!check(cudaMalloc(&s.d_fe_maps, n_rounds * 8), "cudaMalloc maps")) break;
            s.fe_maps_cap = n_rounds;
        }
        if (s.fe_seg_cap < n_seg) {
            cudaFree(s.d_fe_seg_exit); cudaFree(s.d_fe_seg_state); cudaFree(s.d_fe_seg_count); cudaFree(s.d_fe_seg_events);
            s.d_fe_seg_exit = s.d_fe_seg_state = nullptr; s.d_fe_seg_count = nullptr; s.d_fe_seg_events = nullptr;
            s.fe_seg_cap = 0;
            if (!check(cudaMalloc(&s.d_fe_seg_exit, n_seg * 9), "cudaMalloc")) break;
            if (!check(cudaMalloc(&s.d_fe_seg_state, n_seg), "cudaMalloc")) break;
            if (!check(cudaMalloc(&s.d_fe_seg_count, n_seg * 9 * 4), "cudaMalloc")) break;
            if (!check(cudaMalloc(&s.d_fe_seg_events, n_seg * 8), "cudaMalloc")) break;
            s.fe_seg_cap = n_seg;
        }
        uint64_t* const d_maps = s.d_fe_maps;
        uint8_t *const d_seg_exit = s.d_fe_seg_exit, *const d_seg_state = s.d_fe_seg_state;
        uint32_t* const d_seg_count = s.d_fe_seg_count;
        uint64_t* const d_seg_events = s.d_fe_seg_events;
        const unsigned map_blocks = (unsigned)((n_seg + 127) / 128), seg_blocks = (unsigned)((n_seg + 127) / 128);
        if (f32) fe_round_maps_kernel<float><<<map_blocks, 128, 0, s.stream>>>(s.d_ranf_table, c->scan_round, n_rounds, seg_rounds, d_maps);
        else fe_round_maps_kernel<double><<<map_blocks, 128, 0, s.stream>>>(s.d_ranf_table, c->scan_round, n_rounds, seg_rounds, d_maps);
        fe_segment_kernel<<<seg_blocks, 128, 0, s.stream>>>(d_maps, n_rounds, seg_rounds, d_seg_exit, d_seg_count);
        c->launches += 2;
        ++passes;
        auto t0 = now();
        std::vector<uint8_t> seg_exit(n_seg * 9), seg_state(n_seg);
        std::vector<uint32_t> seg_count(n_seg * 9);
        std::vector<uint64_t> seg_events(n_seg);
        if (!check(cudaMemcpyAsync(seg_exit.data(), d_seg_exit, n_seg * 9, cudaMemcpyDeviceToHost, s.stream), "D2H")) break;
        if (!check(cudaMemcpyAsync(seg_count.data(), d_seg_count, n_seg * 9 * 4, cudaMemcpyDeviceToHost, s.stream), "D2H")) break;
        if (!check(cudaStreamSynchronize(s.stream), "fe scan maps")) break;
        auto t1 = now();
        int state = c->scan_state;
        uint64_t events = c->scan_events;
        for (uint64_t g = 0; g < n_seg; ++g) {  // chain the segment maps (multi_threading.rs:59-64 does this event by event)
            seg_state[g] = (uint8_t)state;
            seg_events[g] = events;
            events += seg_count[g * 9 + state];
            state = seg_exit[g * 9 + state];
        }
        if (!check(cudaMemcpyAsync(d_seg_state, seg_state.data(), n_seg, cudaMemcpyHostToDevice, s.stream), "H2D")) break;
        if (!check(cudaMemcpyAsync(d_seg_events, seg_events.data(), n_seg * 8, cudaMemcpyHostToDevice, s.stream), "H2D")) break;
        auto t2 = now();
        fe_boundaries_kernel<<<seg_blocks, 128, 0, s.stream>>>(d_maps, c->scan_round, n_rounds, seg_rounds, d_seg_state, d_seg_events, split,
                                                               part_len, first * split, n_bnd, d_bnd);
        ++c->launches;
        if (!check(cudaStreamSynchronize(s.stream), "fe boundaries")) break;  // the host vectors die with this iteration
        t_maps += ms(t0, t1);
        t_chain += ms(t1, t2);
        t_bnd += ms(t2, now());
        if (events > last_target) {
            // Done. Leave the scan at the last segment start that a call for the following batches can resume from.
            uint64_t g = n_seg - 1;
            while (g > 0 && seg_events[g] > next_first) --g;
            c->scan_round += g * seg_rounds;
            c->scan_events = seg_events[g];
            c->scan_state = seg_state[g];
            break;
        }
        c->scan_round += n_rounds;
        c->scan_events = events;
        c->scan_state = state;
    }
    if (rc == TP3_OK) {
        auto t0 = now();
        const uint32_t chain = split == 1 ? 4 : split;
        const unsigned blocks = (unsigned)(((n_bnd + chain - 1) / chain + 3) / 4);
        if (f32) fe_batch_states_kernel<float><<<blocks, 128, 0, s.stream>>>(s.d_ranf_table, d_bnd, n_bnd, chain, s.d_fe_ranf_states);
        else fe_batch_states_kernel<double><<<blocks, 128, 0, s.stream>>>(s.d_ranf_table, d_bnd, n_bnd, chain, s.d_fe_ranf_states);
        ++c->launches;
        check(cudaGetLastError(), "fe batch states");
        check(cudaStreamSynchronize(s.stream), "fe batch states");
        t_states = ms(t0, now());
    }
    if (timing)
        std::fprintf(stderr, "[tp3 fe scan] batches %llu split %u passes %d: maps+segments %.2f ms, host chain %.2f ms, boundaries %.2f ms, states %.2f ms\n",
                     (unsigned long long)n, split, passes, t_maps, t_chain, t_bnd, t_states);
    return rc;
}

// ---- faster-evgen + xoshiro: batch start states by coalescing segment walks (fe_scan_xo.cuh) -----------------------
// Fills s.d_xo_states[0 .. n) for batches [first, first + n) of the sequential stream and leaves the context's generator
// (c->fe_xo, the same bookkeeping as the host walk) at the start of batch first + n, so that consecutive calls continue.
// `split` = 32 also locates the 32 lane starts inside every batch (events 10000 b + 313 l), batch-major, so that a warp
// can share a batch as in the RANF path; pass C then walks from a segment entry once per LANE start, which only pays
// with short segments (4096 outputs = 246 events on average): used for runs too small to fill the GPU with one thread
// per batch.
int fe_device_states_xo(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t split) {
    const bool f32 = c->params.flags & TP3_F32;
    if (!c->fe_ready || c->fe_pos > first) {  // (re)start from the seeded generator
        for (int i = 0; i < 4; ++i) c->fe_xo[i] = c->xo_base[i];
        c->fe_pos = 0;
        c->fe_ready = true;
    }
    const uint64_t lead = first - c->fe_pos;                                   // batches to pass over before `first`
    const uint64_t e_last = (lead + n) * (uint64_t)TP3_EVENT_BATCH_SIZE;       // event index (from the base) where batch first + n starts
    const uint64_t n_bnd = n * split + 1;                                      // lane / batch starts of batches first .. first + n - 1, then the start of batch first + n
    auto target = [&](uint64_t j) {  // event index (from the base) of boundary j
        return j == n * split ? e_last : (lead + j / split) * (uint64_t)TP3_EVENT_BATCH_SIZE + (j % split) * (uint64_t)kLaneEvents;
    };
    if (s.xo_states_cap < n_bnd) {
        if (s.d_xo_states) TP3_CUDA(c, cudaFree(s.d_xo_states));
        s.d_xo_states = nullptr;
        s.xo_states_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_xo_states, n_bnd * 4 * sizeof(uint64_t)));
        s.xo_states_cap = n_bnd;
    }
    if (s.xo_bnd_cap < n_bnd) {
        cudaFree(s.d_xo_bnd);
        s.d_xo_bnd = nullptr;
        s.xo_bnd_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_xo_bnd, n_bnd * sizeof(XoBoundary)));
        s.xo_bnd_cap = n_bnd;
    }
    const uint64_t b0 = c->fe_xo[0], b1 = c->fe_xo[1], b2 = c->fe_xo[2], b3 = c->fe_xo[3];
    double margin = 1.004;  // 16.64 outputs per event on average; the count is checked and the scan extended if short
    for (int attempt = 0; attempt < 8; ++attempt, margin *= 1.02) {
        const uint64_t outputs = (uint64_t)((double)(e_last + 1) * 16.65 * margin) + 65536;
        // one lane per segment: as long as possible (the entry walk and the jump-ahead are per segment) while the device stays full
        uint32_t seg_units = 64;
        while (seg_units > 2 && outputs / ((uint64_t)seg_units * kXoSegUnit) < (uint64_t)s.sm_count * 4 * 128 * 2) seg_units /= 2;
        // Lane starts: pass C walks from a segment entry once per 313 events, so segments must be short, but a 4096-output
        // segment fails to coalesce about once in 800 (event lengths are odd, which slows the merging down: measured 3.7 %
        // per 2048 outputs) and pass B would run 2-3 times; 8192 outputs (2e-6) is the better trade.
        if (split > 1) seg_units = 4;
        if (c->opt_fe_xo_seg_units >= 1 && c->opt_fe_xo_seg_units <= 64)  // test hook: short segments make pass B repeat
            seg_units = (uint32_t)c->opt_fe_xo_seg_units;
        const uint64_t seg_len = (uint64_t)seg_units * kXoSegUnit;
        const uint64_t n_seg = (outputs + seg_len - 1) / seg_len + 1;
        if (n_seg * seg_units >= (1ull << (8 * c->xo_digits))) {
            c->err = "faster-evgen scan beyond the reach of the xoshiro jump tables";
            return TP3_E_INVALID;
        }
        if (s.xo_scan_cap < 3 * n_seg + 1) {
            cudaFree(s.d_xo_scan);
            s.d_xo_scan = nullptr;
            s.xo_scan_cap = 0;
            TP3_CUDA(c, cudaMalloc(&s.d_xo_scan, (3 * n_seg + 1) * sizeof(uint32_t)));
            s.xo_scan_cap = 3 * n_seg + 1;
        }
        uint32_t *exit_a = s.d_xo_scan, *exit_b = s.d_xo_scan + n_seg, *count = s.d_xo_scan + 2 * n_seg, *flag = s.d_xo_scan + 3 * n_seg;
        const unsigned blocks = (unsigned)((n_seg + 127) / 128);
        auto walk = [&](const uint32_t* prev, const uint32_t* check, uint32_t* out_exit) {
            if (f32) xo_walk_kernel<float><<<blocks, 128, 0, s.stream>>>(b0, b1, b2, b3, n_seg, seg_units, s.d_xo_digit_polys, c->xo_digits, prev, check, out_exit, count, flag);
            else xo_walk_kernel<double><<<blocks, 128, 0, s.stream>>>(b0, b1, b2, b3, n_seg, seg_units, s.d_xo_digit_polys, c->xo_digits, prev, check, out_exit, count, flag);
            ++c->launches;
        };
        walk(nullptr, nullptr, exit_a);  // pass A
        bool settled = false;
        int passes_b = 0;
        for (int pass = 0; pass < 64 && !settled; ++pass, ++passes_b) {  // pass B until the exits reproduce themselves
            TP3_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(uint32_t), s.stream));
            walk(exit_a, exit_a, exit_b);
            uint32_t mismatch = 1;
            TP3_CUDA(c, cudaMemcpyAsync(&mismatch, flag, sizeof mismatch, cudaMemcpyDeviceToHost, s.stream));
            TP3_CUDA(c, cudaStreamSynchronize(s.stream));
            settled = mismatch == 0;
            if (!settled) std::swap(exit_a, exit_b);
        }
        if (!settled) {
            c->err = "faster-evgen xoshiro scan did not settle";
            return TP3_E_CUDA;
        }
        c->stat_fe_xo_pass_b = passes_b;
        if (c->opt_fe_timing)
            std::fprintf(stderr, "[tp3 fe xo scan] batches %llu split %u: %llu segments of %llu outputs, pass B x %d\n", (unsigned long long)n, split,
                         (unsigned long long)n_seg, (unsigned long long)seg_len, passes_b);
        std::vector<uint32_t> h_count(n_seg);
        TP3_CUDA(c, cudaMemcpyAsync(h_count.data(), count, n_seg * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));
        // the segment of every wanted event index (multi_threading.rs:59-64 finds these by walking event by event)
        std::vector<XoBoundary> bnd(n_bnd);
        uint64_t cum = 0, j = 0;
        for (uint64_t g = 0; g < n_seg && j < n_bnd; ++g) {
            const uint64_t next = cum + h_count[g];
            while (j < n_bnd && target(j) < next) {  // targets increase with j
                bnd[j].seg = g;
                bnd[j].skip = (uint32_t)(target(j) - cum);
                bnd[j].pad = 0;
                ++j;
            }
            cum = next;
        }
        if (j < n_bnd) continue;  // the estimate was short: scan further
        TP3_CUDA(c, cudaMemcpyAsync(s.d_xo_bnd, bnd.data(), n_bnd * sizeof(XoBoundary), cudaMemcpyHostToDevice, s.stream));
        const unsigned bblocks = (unsigned)((n_bnd + 127) / 128);
        const XoBoundary* d_bnd = static_cast<const XoBoundary*>(s.d_xo_bnd);
        if (f32) xo_boundary_states_kernel<float><<<bblocks, 128, 0, s.stream>>>(b0, b1, b2, b3, seg_units, s.d_xo_digit_polys, c->xo_digits, exit_a, d_bnd, n_bnd, s.d_xo_states);
        else xo_boundary_states_kernel<double><<<bblocks, 128, 0, s.stream>>>(b0, b1, b2, b3, seg_units, s.d_xo_digit_polys, c->xo_digits, exit_a, d_bnd, n_bnd, s.d_xo_states);
        ++c->launches;
        TP3_CUDA(c, cudaGetLastError());
        TP3_CUDA(c, cudaMemcpyAsync(c->fe_xo, s.d_xo_states + 4 * (n_bnd - 1), 32, cudaMemcpyDeviceToHost, s.stream));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));  // bnd dies with this scope; fe_xo is read by the next call
        c->fe_pos = first + n;
        return TP3_OK;
    }
    c->err = "faster-evgen xoshiro scan: event count estimate failed";
    return TP3_E_CUDA;
}

// ---- faster-evgen + RANF, sequential stream: walk -> event records -> physics (fe_stream.cuh) -----------------------
// Fills s.d_out[0 .. n) with the accumulators of batches [first, first + n).  The stream is processed in passes of
// consecutive rounds; the context remembers a round whose absolute event index is known, so consecutive calls continue.
template <class T> int fs_grow(tp3_ctx* c, T*& ptr, size_t& cap, size_t need) {
    if (cap >= need) return TP3_OK;
    if (ptr) TP3_CUDA(c, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    TP3_CUDA(c, cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return TP3_OK;
}

// A TILE of the stream (tp3_fe_tile_device): every event that STARTS in rounds [first_round, first_round + n_rounds), at most
// max_events of them, counted from 0 at the tile start.  Tiles of consecutive round ranges partition the events of the
// run exactly, so G ranks can each take one without knowing how many events precede it (the reference's scheduler needs
// exactly that knowledge, evgen.rs:257-267, which is why its reproducible mode does not scale).
struct FeTile {
    uint64_t first_round, n_rounds, max_events;  // n_rounds = 0: no round limit; max_events = 0: no event limit
    uint64_t events_done, batches_done;          // out
};

int fe_stream_simulate(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t last_len, FeTile* tile = nullptr) {
    using FsBuf = DeviceSlot::FsBuf;
    const bool f32 = c->params.flags & TP3_F32;
    const uint64_t B = TP3_EVENT_BATCH_SIZE;
    const uint64_t kNoLimit = ~0ull;
    // events wanted: [e_lo, e_hi); in a tile limited by rounds e_hi is only known once the walk has reached the tile's end
    const uint64_t e_lo = first * B;
    uint64_t e_hi = tile ? (tile->max_events ? tile->max_events : kNoLimit) : (first + n - 1) * B + last_len;
    const uint64_t stop_round = (tile && tile->n_rounds) ? tile->first_round + tile->n_rounds : kNoLimit;
    uint64_t tile_round = tile ? tile->first_round : 0, tile_events = 0;
    uint64_t& pos_round = tile ? tile_round : c->fs_round;    // a round (segment boundary) whose event index is known ...
    uint64_t& pos_events = tile ? tile_events : c->fs_events;  // ... and that index
    if (!tile && pos_events > e_lo) pos_round = pos_events = 0;  // the known point is past the start: from the seed again
    if (tile) n = e_hi == kNoLimit ? kNoLimit : (e_hi + B - 1) / B;
    c->stat_fe_passes = c->stat_fe_redone = 0;
    const double kRoundsPerEvent = 0.3252;  // 3.076 events start per round on average (measured); only sizes the passes
    const uint64_t lanes_per_wave = (uint64_t)s.sm_count * 26 * 32;  // the walk kernel alone holds 26 one-warp CTAs per SM (shared memory)
    const uint32_t warm = c->opt_fe_warm > 0 ? (uint32_t)c->opt_fe_warm : (uint32_t)kFeWarm;
    const bool overlap = !c->opt_fe_serial;
    if (!s.fs_stream2) {
        // The walks run on a HIGH-PRIORITY stream of their own, the physics kernels on the slot's stream: the blocks of the walk
        // of pass k + 1 (integer work) are placed as soon as CTAs of the physics kernel of pass k (FP64 work) retire.
        int prio_lo = 0, prio_hi = 0;
        TP3_CUDA(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        TP3_CUDA(c, cudaStreamCreateWithPriority(&s.fs_stream2, cudaStreamNonBlocking, prio_hi));
        TP3_CUDA(c, cudaEventCreateWithFlags(&s.fs_event, cudaEventDisableTiming));
        for (auto& b : s.fs) TP3_CUDA(c, cudaEventCreateWithFlags(&b.phys_done, cudaEventDisableTiming));
    }
    cudaStream_t const sw = overlap ? s.fs_stream2 : s.stream, sp = s.stream;  // walks / physics
    for (auto& b : s.fs) b.phys_pending = false;
    {
        // Both kernels ask for the largest shared-memory carve-out: CTAs of two kernels only share an SM if they agree on the
        // L1 / shared split, and the walk of pass k + 1 is meant to run next to the physics of pass k.
        static bool once = [] {
            cudaFuncSetAttribute(fe_walk_kernel<double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(fe_walk_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(fe_physics_kernel<double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(fe_physics_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            return true;
        }();
        (void)once;
    }
    // the walk stream starts after whatever the caller has queued on the slot's stream
    TP3_CUDA(c, cudaEventRecord(s.fs_event, s.stream));
    TP3_CUDA(c, cudaStreamWaitEvent(s.fs_stream2, s.fs_event, 0));

    uint64_t done_batches = 0;  // batches [first, first + done_batches) have their physics launched
    auto launch_walk = [&](FsBuf& b, const FeWalkArgs& a, uint32_t items) {
        const unsigned blocks = (items + 31) / 32;
        if (f32) fe_walk_kernel<float><<<blocks, kFeWalkThreads, 0, sw>>>(a);
        else fe_walk_kernel<double><<<blocks, kFeWalkThreads, 0, sw>>>(a);
        ++c->launches;
    };
    auto walk_args = [&](FsBuf& b) {
        FeWalkArgs w;
        std::memset(&w, 0, sizeof w);
        w.jump_table = s.d_ranf_table;
        w.first_round = b.first_round;
        w.n_seg = (uint32_t)b.n_seg;
        w.seg_rounds = b.seg_rounds;
        w.warm = warm;
        w.warm_first = kFeWarmFirst;
        w.seg_count = b.d_count;
        w.seg_exit = b.d_exit;
        w.seg_fail = b.d_fail;
        w.records = b.count_only ? nullptr : b.d_records;
        return w;
    };
    auto fetch_counts = [&](FsBuf& b) -> int {
        TP3_CUDA(c, cudaMemcpyAsync(b.h_count.data(), b.d_count, b.n_seg * 4, cudaMemcpyDeviceToHost, sw));
        TP3_CUDA(c, cudaMemcpyAsync(b.h_exit.data(), b.d_exit, b.n_seg, cudaMemcpyDeviceToHost, sw));
        TP3_CUDA(c, cudaMemcpyAsync(b.h_fail.data(), b.d_fail, b.n_seg, cudaMemcpyDeviceToHost, sw));
        return TP3_OK;
    };
    // Size pass `b` from the known point (round, events) towards the first batch that still has to be served, and start its walk.
    auto start_pass = [&](FsBuf& b, uint64_t round, uint64_t events, uint64_t next_batch) -> int {
        const uint64_t next_event = (first + next_batch) * B;
        // far from the first wanted event: count-only passes (no records) until about 800 batches before it
        b.count_only = (double)(next_event - events) * kRoundsPerEvent > 8.0 * (double)lanes_per_wave;
        const uint64_t want_events = b.count_only ? next_event - events : e_hi - events;
        uint64_t want_rounds = e_hi == kNoLimit ? kNoLimit : (uint64_t)((double)want_events * kRoundsPerEvent * 1.002) + 4096;
        if (stop_round != kNoLimit && want_rounds > stop_round - round) want_rounds = stop_round - round;  // a tile ends at its last round
        const uint64_t max_seg = c->opt_fe_pass_segments > 0 ? (uint64_t)c->opt_fe_pass_segments : lanes_per_wave;
        uint32_t seg_rounds = 64;
        if (c->opt_fe_seg_rounds > 0) seg_rounds = (uint32_t)c->opt_fe_seg_rounds;
        else
            while (seg_rounds < 512 && want_rounds / seg_rounds > lanes_per_wave) seg_rounds *= 2;  // fill the device first, then lengthen
        if (stop_round != kNoLimit)
            while ((stop_round - round) % seg_rounds) seg_rounds /= 2;  // a tile ends on a segment boundary
        uint64_t n_seg = (want_rounds + seg_rounds - 1) / seg_rounds;
        if (n_seg > max_seg) n_seg = max_seg;
        if (b.count_only) {  // stop a little before the first wanted event, at a segment boundary
            const uint64_t lim = (uint64_t)((double)want_events * kRoundsPerEvent * 0.99) / seg_rounds;
            if (n_seg > lim) n_seg = lim;
            if (n_seg == 0) n_seg = 1;
        }
        b.first_round = round;
        b.first_events = events;
        b.n_seg = n_seg;
        b.seg_rounds = seg_rounds;
        int rc = TP3_OK;
        if (!b.count_only && (rc = fs_grow(c, b.d_records, b.records_cap, (size_t)n_seg * kFeSlotsPerRound * seg_rounds))) return rc;
        if (b.seg_cap < n_seg + 1) {
            size_t cap = 0;
            if ((rc = fs_grow(c, b.d_count, cap, n_seg + 1))) return rc;
            cap = 0; if ((rc = fs_grow(c, b.d_exit, cap, n_seg + 1))) return rc;
            cap = 0; if ((rc = fs_grow(c, b.d_fail, cap, n_seg + 1))) return rc;
            cap = 0; if ((rc = fs_grow(c, b.d_seg_events, cap, n_seg + 1))) return rc;
            cap = 0; if ((rc = fs_grow(c, b.d_redo_list, cap, n_seg + 1))) return rc;
            cap = 0; if ((rc = fs_grow(c, b.d_redo_entry, cap, n_seg + 1))) return rc;
            b.seg_cap = n_seg + 1;
        }
        b.h_count.resize(n_seg);
        b.h_exit.resize(n_seg);
        b.h_fail.resize(n_seg);
        if (b.phys_pending) {  // the records of this buffer are still being read by the physics kernel of two passes ago
            TP3_CUDA(c, cudaStreamWaitEvent(sw, b.phys_done, 0));
            b.phys_pending = false;
        }
        launch_walk(b, walk_args(b), (uint32_t)n_seg);
        TP3_CUDA(c, cudaGetLastError());
        return fetch_counts(b);
    };

    int cur = 0;
    int rc = start_pass(s.fs[0], pos_round, pos_events, 0);
    if (rc) return rc;
    while (true) {
        FsBuf& b = s.fs[cur];
        TP3_CUDA(c, cudaStreamSynchronize(sw));  // the walk of this pass is over (the physics of the previous pass may still run)
        ++c->stat_fe_passes;
        const uint64_t n_seg = b.n_seg;
        // ---- segments whose nine walks had not coincided at their start: redo them from the predecessor's exit state
        for (int round = 0;; ++round) {
            std::vector<uint32_t> list;
            std::vector<uint8_t> entry;
            for (uint64_t g = 0; g < n_seg; ++g) {
                if (!b.h_fail[g]) continue;
                if (g == 0) {
                    c->err = "faster-evgen walk: the warm-up of a pass did not settle";
                    return TP3_E_CUDA;
                }
                if (b.h_fail[g - 1] || b.h_exit[g - 1] == 0xff) continue;  // its predecessor is redone first
                list.push_back((uint32_t)g);
                entry.push_back(b.h_exit[g - 1]);
            }
            if (list.empty()) break;
            if (round == 63) {
                c->err = "faster-evgen walk: too many redo rounds (fe_warm too small)";
                return TP3_E_INVALID;
            }
            c->stat_fe_redone += (int64_t)list.size();
            TP3_CUDA(c, cudaMemcpyAsync(b.d_redo_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, sw));
            TP3_CUDA(c, cudaMemcpyAsync(b.d_redo_entry, entry.data(), entry.size(), cudaMemcpyHostToDevice, sw));
            FeWalkArgs r = walk_args(b);
            r.seg_list = b.d_redo_list;
            r.seg_entry = b.d_redo_entry;
            r.n_list = (uint32_t)list.size();
            launch_walk(b, r, r.n_list);
            TP3_CUDA(c, cudaGetLastError());
            if ((rc = fetch_counts(b))) return rc;
            TP3_CUDA(c, cudaStreamSynchronize(sw));  // (also: `list` and `entry` die with this iteration)
        }
        // ---- 2. absolute index of every segment's first event (multi_threading.rs:59-64 finds these event by event)
        b.h_seg_events.resize(n_seg + 1);
        b.h_seg_events[0] = b.first_events;
        for (uint64_t g = 0; g < n_seg; ++g) b.h_seg_events[g + 1] = b.h_seg_events[g] + b.h_count[g];
        const uint64_t pass_end = b.h_seg_events[n_seg];
        if (b.first_round + n_seg * b.seg_rounds == stop_round && e_hi > pass_end) {  // the tile ends here: now its size is known
            e_hi = pass_end;
            n = (e_hi + B - 1) / B;
        }
        // ---- batches whose events all start in this pass
        uint64_t b_hi = done_batches;
        if (!b.count_only) {
            while (b_hi < n) {
                const uint64_t end = std::min((first + b_hi + 1) * B, e_hi);
                if (end > pass_end) break;
                ++b_hi;
            }
        }
        // ---- where the next pass (or call) starts: the last segment boundary at or before the next batch's first event
        const uint64_t next = (first + b_hi) * B;
        uint64_t g_next = n_seg;
        while (g_next > 0 && b.h_seg_events[g_next] > next) --g_next;
        if (g_next == 0 && b_hi == done_batches && b_hi < n) {  // the pass holds less than the one batch it was sized for
            c->err = "faster-evgen stream pipeline: a pass must hold at least one batch (fe_pass_segments x fe_seg_rounds too small)";
            return TP3_E_INVALID;
        }
        pos_round = b.first_round + g_next * b.seg_rounds;
        pos_events = b.h_seg_events[g_next];
        const uint64_t lo_batches = done_batches;
        done_batches = b_hi;
        const bool more = done_batches < n;
        // ---- 3. physics on the records of this pass ...
        if (b_hi > lo_batches) {
            const uint64_t nb = b_hi - lo_batches, n_units = nb * kFeParts;
            if (b.unit_cap < n_units) {
                size_t cap = 0;
                if ((rc = fs_grow(c, b.d_unit_seg, cap, n_units))) return rc;
                cap = 0; if ((rc = fs_grow(c, b.d_parts, cap, n_units))) return rc;
                b.unit_cap = n_units;
            }
            b.h_unit_seg.resize(n_units);
            uint64_t g = 0;
            for (uint64_t u = 0; u < n_units; ++u) {
                const uint64_t ev = (first + lo_batches + u / kFeParts) * B + (u % kFeParts) * (uint64_t)kFePartLen;
                while (g + 1 < n_seg && b.h_seg_events[g + 1] <= ev) ++g;
                b.h_unit_seg[u] = (uint32_t)g;
            }
            TP3_CUDA(c, cudaMemcpyAsync(b.d_seg_events, b.h_seg_events.data(), (n_seg + 1) * 8, cudaMemcpyHostToDevice, sp));
            TP3_CUDA(c, cudaMemcpyAsync(b.d_unit_seg, b.h_unit_seg.data(), n_units * 4, cudaMemcpyHostToDevice, sp));
            FePhysArgs ph;
            std::memset(&ph, 0, sizeof ph);
            ph.records = b.d_records;
            ph.slots_per_seg = (uint32_t)(kFeSlotsPerRound * b.seg_rounds);
            ph.seg_events = b.d_seg_events;
            ph.unit_seg = b.d_unit_seg;
            ph.first_event = (first + lo_batches) * B;
            ph.end_event = e_hi;
            ph.n_units = (uint32_t)n_units;
            ph.out_parts = b.d_parts;
            // One CTA per unit (2500 events), dispatched by the hardware: warps that start at staggered times do not run in
            // lock step (the static split, as many CTAs as the device holds, was 11 % slower for the default kernel), and as
            // they retire the blocks of the next pass's walk, queued on the other stream, move in next to them.
            uint64_t W = n_units;
            if (c->opt_grid_warps > 0 && (uint64_t)c->opt_grid_warps < W) W = (uint64_t)c->opt_grid_warps;
            if (c->hist_bins) {
                // per-event observables: as many CTAs as the device holds, each flushes its shared-memory counts once
                W = std::min<uint64_t>(W, (uint64_t)s.sm_count * 16);
                ph.hist_bins = c->hist_bins;
                ph.hist_counts = s.d_hist_counts;
                ph.hist_weights = s.d_hist_weights;
                ph.n_warps = (uint32_t)W;
                const size_t dyn = (size_t)TP3_HIST_OBSERVABLES * c->hist_bins * sizeof(uint32_t);
                const bool sorted = !(c->params.flags & TP3_NO_PHOTON_SORTING);
                if (f32 && sorted) fe_physics_kernel<float, 2><<<(unsigned)W, 32, dyn, sp>>>(ph, phys_params<float>(c->params));
                else if (f32) fe_physics_kernel<float, 1><<<(unsigned)W, 32, dyn, sp>>>(ph, phys_params<float>(c->params));
                else if (sorted) fe_physics_kernel<double, 2><<<(unsigned)W, 32, dyn, sp>>>(ph, phys_params<double>(c->params));
                else fe_physics_kernel<double, 1><<<(unsigned)W, 32, dyn, sp>>>(ph, phys_params<double>(c->params));
            } else {
                ph.n_warps = (uint32_t)W;
                if (f32) fe_physics_kernel<float><<<(unsigned)W, 32, 0, sp>>>(ph, phys_params<float>(c->params));
                else fe_physics_kernel<double><<<(unsigned)W, 32, 0, sp>>>(ph, phys_params<double>(c->params));
            }
            ++c->launches;
            TP3_CUDA(c, cudaGetLastError());
            const unsigned cb = (unsigned)((nb * 13 + 255) / 256);
            if (f32) fe_combine_parts_kernel<float><<<cb, 256, 0, sp>>>(b.d_parts, nb, s.d_out + lo_batches);
            else fe_combine_parts_kernel<double><<<cb, 256, 0, sp>>>(b.d_parts, nb, s.d_out + lo_batches);
            ++c->launches;
            TP3_CUDA(c, cudaGetLastError());
            TP3_CUDA(c, cudaEventRecord(b.phys_done, sp));
            b.phys_pending = true;
        }
        if (!more) break;
        // ---- ... while the next pass is walked on the other stream (queued behind the physics that last used that buffer)
        cur ^= 1;
        if ((rc = start_pass(s.fs[cur], pos_round, pos_events, done_batches))) return rc;
    }
    if (tile) {
        tile->events_done = e_hi;
        tile->batches_done = done_batches;
    }
    // (every physics kernel ran on the slot's stream, behind the walk it depends on: the slot's stream owns the result)
    return TP3_OK;
}

template <class F, int RNG> void launch_fe(const FeArgs& a, const tp3_params& p, cudaStream_t st) {
    const uint64_t units = a.n_batches * a.split;
    faster_evgen_kernel<F, RNG><<<(unsigned)((units + kFeThreads - 1) / kFeThreads), kFeThreads, 0, st>>>(a, phys_params<F>(p));
}

int ensure_out(tp3_ctx* c, DeviceSlot& s, uint64_t n) {
    if (!s.d_fold) TP3_CUDA(c, cudaMalloc(&s.d_fold, sizeof(FoldState)));
    if (s.out_cap < n) {
        if (s.d_out) TP3_CUDA(c, cudaFree(s.d_out));
        s.d_out = nullptr;
        s.out_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_out, n * sizeof(tp3_acc)));
        s.out_cap = n;
    }
    if ((c->params.flags & TP3_STANDARD_RANDOM) && s.xo_states_cap < n) {
        if (s.d_xo_states) TP3_CUDA(c, cudaFree(s.d_xo_states));
        s.d_xo_states = nullptr;
        s.xo_states_cap = 0;
        TP3_CUDA(c, cudaMalloc(&s.d_xo_states, n * 4 * sizeof(uint64_t)));
        s.xo_states_cap = n;
    }
    return TP3_OK;
}

#ifdef TP3_TRACE_UNITS
unsigned long long* g_trace = nullptr;  // diagnostic build only: device buffer [units][3], set by tp3_debug_set_trace
#endif
SimArgs make_args(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t last_len) {
    SimArgs a;
    std::memset(&a, 0, sizeof a);
    a.first_batch = first;
    a.n_batches = n;
    a.last_batch_len = last_len;
    a.batch_events = TP3_EVENT_BATCH_SIZE;
    a.batch_parts = 1;
    a.jump_seeding = (c->params.flags & TP3_FASTER_THREADING) ? 1u : 0u;
#ifdef TP3_TRACE_UNITS
    a.trace = g_trace;
#endif
    a.ranf_table = s.d_ranf_table;
    a.xo_batch_states = s.d_xo_states;
    a.xo_lane_polys = s.d_xo_lane_polys;
    a.out = s.d_out;
    std::memcpy(a.ranf_base, c->ranf_base, sizeof a.ranf_base);
    a.ranf_seed = RANF_DEFAULT_SEED;
    return a;
}

// Enqueue seeding (xoshiro) + the fused kernel for [first, first+n) on one device slot.
// `fold`: also left-fold the accumulators in batch order into s.d_fold->running (in the kernel; merge_kernel under faster-evgen).
int enqueue_range(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t last_len, bool fold = false) {
    TP3_CUDA(c, cudaSetDevice(s.dev));
    int rc = ensure_out(c, s, n);
    if (rc) return rc;
    if (n > 0x7fffffffull) {
        c->err = "too many batches in one launch";
        return TP3_E_INVALID;
    }
    if (!(c->params.flags & TP3_STANDARD_RANDOM) && (c->params.flags & TP3_FASTER_THREADING)) {
        // ranf.rs:136-140 reseeds with seed + 123456*b in i32; beyond b = 6199 the seed leaves [0, 1e9)
        // and the reference itself produces out-of-range words. Refuse rather than imitate that.
        if (first + n > 6200) {
            c->err = "RANF jump() seeding is only defined for the first 6200 batches (seed < 1e9)";
            return TP3_E_INVALID;
        }
    }
    // Reach of the jump-ahead tables: 5 byte digits = 2^40 RANF rounds (12e4 / 55 rounds per batch; ~0.32 per event under
    // faster-evgen) resp. 2^40 xoshiro batch strides.
    if (first + n > ((c->params.flags & TP3_STANDARD_RANDOM) ? (1ull << 40) : 300000000ull)) {
        c->err = "batch index beyond the reach of the jump-ahead tables (3e8 batches with RANF, 2^40 with xoshiro)";
        return TP3_E_INVALID;
    }
    const bool faster = c->params.flags & TP3_FASTER_EVGEN;
    const bool seq_faster = faster && !(c->params.flags & TP3_FASTER_THREADING);
    std::vector<uint32_t> fe_ranf;
    std::vector<uint64_t> fe_xo;
    const bool device_scan = seq_faster && !(c->params.flags & TP3_STANDARD_RANDOM) && !c->opt_fe_host_scan;
    uint32_t fe_split = 1;
    if (device_scan && (c->opt_fe_legacy || c->opt_fe_split) && c->hist_bins) {
        c->err = "per-event observables under faster-evgen need the stream pipeline (fe_legacy / fe_split are set)";
        return TP3_E_INVALID;
    }
    if (seq_faster && !device_scan && c->hist_bins) {
        c->err = "per-event observables under faster-evgen need the stream pipeline (fe_host_scan is set)";
        return TP3_E_INVALID;
    }
    if (device_scan && !c->opt_fe_legacy && !c->opt_fe_split) {
        // walk -> event records -> physics (fe_stream.cuh): the shipped path for the sequential RANF stream
        s.fold_in_kernel = false;
        rc = fe_stream_simulate(c, s, first, n, last_len);
        if (rc) return rc;
        if (fold) {
            if (c->params.flags & TP3_F32) merge_kernel<float><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
            else merge_kernel<double><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
            ++c->launches;
            TP3_CUDA(c, cudaGetLastError());
        }
        s.last_first = first;
        s.last_n = n;
        return TP3_OK;
    }
    if (device_scan) {
        // One lane per 313 events: fills the device for any run size and keeps a warp's lanes on one batch (measured
        // faster than one thread per batch at every size).  Costs 7.3 KB of start states per batch, so very long
        // ranges fall back to one thread per batch.
        fe_split = n <= (1ull << 20) ? 32 : 1;
        if (c->opt_fe_split) fe_split = c->opt_fe_split == 32 ? 32 : 1;  // test hook (tp3_set_option)
        rc = fe_device_states_ranf(c, s, first, n, fe_split);
        if (rc) return rc;
    } else if (seq_faster && (c->params.flags & TP3_STANDARD_RANDOM) && !c->opt_fe_host_scan) {
        // one thread per batch fills the GPU from ~76 000 batches on; below that a warp shares a batch (32 lane starts each)
        fe_split = n <= 65536 ? 32 : 1;
        if (c->opt_fe_split) fe_split = c->opt_fe_split == 32 ? 32 : 1;  // test hook (tp3_set_option)
        rc = fe_device_states_xo(c, s, first, n, fe_split);
        if (rc) return rc;
    } else if (seq_faster) {
        fe_host_states(c, first, n, fe_ranf, fe_xo);
        if (!fe_ranf.empty()) {
            if (s.fe_states_cap < n) {
                if (s.d_fe_ranf_states) TP3_CUDA(c, cudaFree(s.d_fe_ranf_states));
                s.d_fe_ranf_states = nullptr;
                s.fe_states_cap = 0;
                TP3_CUDA(c, cudaMalloc(&s.d_fe_ranf_states, n * 57 * sizeof(uint32_t)));
                s.fe_states_cap = n;
            }
            TP3_CUDA(c, cudaMemcpyAsync(s.d_fe_ranf_states, fe_ranf.data(), n * 57 * 4, cudaMemcpyHostToDevice, s.stream));
        } else {
            TP3_CUDA(c, cudaMemcpyAsync(s.d_xo_states, fe_xo.data(), n * 32, cudaMemcpyHostToDevice, s.stream));
        }
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));  // the host vectors die at the end of this call
    }
    SimArgs a = make_args(c, s, first, n, last_len);
    uint64_t* xo_states = s.d_xo_states;
    if ((c->params.flags & TP3_STANDARD_RANDOM) && !seq_faster) {
        const unsigned blocks = (unsigned)((n + 127) / 128);
        if (c->params.flags & TP3_F32)
            xoshiro_seed_kernel<Xoshiro128Lane><<<blocks, 128, 0, s.stream>>>(first, n, s.d_xo_digit_polys, c->xo_digits,
                                                                              c->xo_base[0], c->xo_base[1], c->xo_base[2],
                                                                              c->xo_base[3], xo_states);
        else
            xoshiro_seed_kernel<Xoshiro256Lane><<<blocks, 128, 0, s.stream>>>(first, n, s.d_xo_digit_polys, c->xo_digits,
                                                                              c->xo_base[0], c->xo_base[1], c->xo_base[2],
                                                                              c->xo_base[3], xo_states);
        ++c->launches;
    }
    if (faster) {
        s.fold_in_kernel = false;
        FeArgs f;
        std::memset(&f, 0, sizeof f);
        f.first_batch = first;
        f.n_batches = n;
        f.last_batch_len = last_len;
        f.jump_seeding = a.jump_seeding;
        f.split = fe_split;
        f.ranf_states = s.d_fe_ranf_states;
        f.xo_states = s.d_xo_states;
        f.out = s.d_out;
        f.ranf_seed = RANF_DEFAULT_SEED;
        const bool f32 = c->params.flags & TP3_F32, xo = c->params.flags & TP3_STANDARD_RANDOM;
        if (f32) { if (xo) launch_fe<float, RNG_XOSHIRO>(f, c->params, s.stream); else launch_fe<float, RNG_RANF>(f, c->params, s.stream); }
        else { if (xo) launch_fe<double, RNG_XOSHIRO>(f, c->params, s.stream); else launch_fe<double, RNG_RANF>(f, c->params, s.stream); }
        ++c->launches;
        TP3_CUDA(c, cudaGetLastError());
        if (fold) {  // strict left fold of the per-batch accumulators by one CTA, after the batch kernel
            if (c->params.flags & TP3_F32) merge_kernel<float><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
            else merge_kernel<double><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
            ++c->launches;
            TP3_CUDA(c, cudaGetLastError());
        }
    } else {
        a.hist_bins = c->hist_bins;
        a.hist_counts = s.d_hist_counts;
        a.hist_weights = s.d_hist_weights;
        const Sched sc{s.sm_count, c->opt_unit_batches, c->opt_grid_warps,
                       !(c->params.flags & (TP3_STANDARD_RANDOM | TP3_FASTER_THREADING)), c->opt_sched_dynamic, &s.last_args, c->opt_ramp_units,
                       c->opt_taper_units, c->opt_tail_singles, c->opt_align_units != 0};
        // Small runs (the default 1e7 events are 1000 batches on 2368 warp slots): cut every batch into equal parts, one warp
        // each, and add the parts of a batch in part order afterwards.  Only where a part's start is a plain stream position
        // (sequential RANF stream), and only on request: the sums of a batch then depend on the number of parts in their last
        // bits, so a fixed setting is the caller's choice (tp3_run asks for 0 = auto, it has one launch).
        uint32_t parts = 1;
        if (c->opt_batch_parts != 1 && !(c->params.flags & (TP3_STANDARD_RANDOM | TP3_FASTER_THREADING))) {
            if (c->opt_batch_parts > 1) parts = (uint32_t)c->opt_batch_parts;
            else
                for (uint32_t p : {2u, 5u, 10u})  // (part sizes whose last warp iteration ends >= 55 draws in: RanfWarpStream::advance)
                    if (n * p <= (uint64_t)s.sm_count * 16) parts = p;
        }
        if (parts > 1) {
            if (n * parts > 0x7fffffffull) {
                c->err = "too many batch parts in one launch";
                return TP3_E_INVALID;
            }
            if (s.batch_parts_cap < n * parts) {
                if (s.d_batch_parts) TP3_CUDA(c, cudaFree(s.d_batch_parts));
                s.d_batch_parts = nullptr;
                s.batch_parts_cap = 0;
                TP3_CUDA(c, cudaMalloc(&s.d_batch_parts, n * parts * sizeof(tp3_acc)));
                s.batch_parts_cap = n * parts;
            }
            a.first_batch = first * parts;
            a.n_batches = n * parts;
            a.batch_events = TP3_EVENT_BATCH_SIZE / parts;
            a.batch_parts = parts;
            a.out = s.d_batch_parts;
            TP3_CUDA(c, pick_sim(c->params, c->hist_bins != 0, c->opt_f32_scalar != 0)(a, c->params, s.stream, sc));
            const unsigned cb = (unsigned)((n * 13 + 255) / 256);
            if (c->params.flags & TP3_F32) fe_combine_parts_kernel<float><<<cb, 256, 0, s.stream>>>(s.d_batch_parts, n, s.d_out, (int)parts);
            else fe_combine_parts_kernel<double><<<cb, 256, 0, s.stream>>>(s.d_batch_parts, n, s.d_out, (int)parts);
            c->launches += 2;
            TP3_CUDA(c, cudaGetLastError());
            if (fold) {
                if (c->params.flags & TP3_F32) merge_kernel<float><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
                else merge_kernel<double><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, n, &s.d_fold->running, true);
                ++c->launches;
                TP3_CUDA(c, cudaGetLastError());
            }
            s.fold_in_kernel = false;
            s.last_first = first;
            s.last_n = n;
            return TP3_OK;
        }
        s.fold_in_kernel = fold;
        if (fold) {
            // Completion marks: one word per unit, compared with this launch's epoch (no memset per launch).  The number
            // of units is at most n + the grid's warps; the grid never exceeds 64 warps per SM.
            const size_t need = (size_t)n + (size_t)s.sm_count * 64 + 64 + (c->opt_grid_warps > 0 ? (size_t)c->opt_grid_warps : 0);
            if (s.unit_done_cap < need) {
                if (s.d_unit_done) TP3_CUDA(c, cudaFree(s.d_unit_done));
                s.d_unit_done = nullptr;
                s.unit_done_cap = 0;
                TP3_CUDA(c, cudaMalloc(&s.d_unit_done, need * sizeof(uint32_t)));
                s.unit_done_cap = need;
                s.epoch = 0;
            }
            if (s.epoch == 0 || s.epoch == 0xffffffffu) {
                TP3_CUDA(c, cudaMemsetAsync(s.d_unit_done, 0, s.unit_done_cap * sizeof(uint32_t), s.stream));
                s.epoch = 0;
            }
            a.epoch = ++s.epoch;
            a.unit_done = s.d_unit_done;
            a.fold = s.d_fold;
            TP3_CUDA(c, cudaMemsetAsync(s.d_fold, 0, sizeof(FoldState), s.stream));
            if (s.copy_event2) TP3_CUDA(c, cudaEventRecord(s.copy_event2, s.stream));  // the fold state of THIS launch starts here
        }
        TP3_CUDA(c, pick_sim(c->params, c->hist_bins != 0, c->opt_f32_scalar != 0)(a, c->params, s.stream, sc));
        ++c->launches;
    }
    s.last_first = first;
    s.last_n = n;
    return TP3_OK;
}

// Contiguous split of [0, n) over the device slots: slot g gets [n*g/G, n*(g+1)/G).
void split(uint64_t n, size_t G, size_t g, uint64_t& off, uint64_t& cnt) {
    off = n * g / G;
    cnt = n * (g + 1) / G - off;
}

}  // namespace

extern "C" {
#ifdef TP3_TRACE_UNITS
void tp3_debug_set_trace(void* device_ptr) { g_trace = static_cast<unsigned long long*>(device_ptr); }
#endif

int tp3_abi_version(void) { return TP3_ABI_VERSION; }

const char* tp3_last_error(const tp3_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tp3_create(const tp3_params* params, int n_dev, const int* dev_ids, tp3_ctx** out) {
    if (!params || !out || n_dev < 1) {
        g_create_error = "tp3_create: bad arguments";
        return TP3_E_INVALID;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
        return TP3_E_NO_DEVICE;
    }
    tp3_ctx* c = new tp3_ctx;
    c->params = *params;
    ranf_seed_state(RANF_DEFAULT_SEED, c->ranf_base);
    const bool xo = params->flags & TP3_STANDARD_RANDOM;
    if (xo) build_xoshiro_tables(c);
    else {
        static const std::vector<uint32_t> table = ranf_round_jump_table();  // 1280 polynomial products: built once per process
        c->ranf_table = table;
    }
    auto fail = [&](int code, const std::string& msg) {
        g_create_error = msg;
        tp3_destroy(c);
        return code;
    };
    for (int i = 0; i < n_dev; ++i) {
        DeviceSlot s;
        s.dev = dev_ids ? dev_ids[i] : i;
        if (s.dev < 0 || s.dev >= count) return fail(TP3_E_NO_DEVICE, "device id out of range");
        // (two attributes, not cudaGetDeviceProperties: that call alone was most of the 2.7 ms a context took to create,
        // profiles/r02_default_run.txt)
        int major = 0, sms = 0;
        if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, s.dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.dev)) != cudaSuccess)
            return fail(TP3_E_CUDA, cudaGetErrorString(e));
        if (major != 10) return fail(TP3_E_NO_DEVICE, "device is not sm_100 (kernels are built for sm_100a only)");
        s.sm_count = sms;
        if ((e = cudaSetDevice(s.dev)) != cudaSuccess) return fail(TP3_E_CUDA, cudaGetErrorString(e));
        if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess)
            return fail(TP3_E_CUDA, cudaGetErrorString(e));
        s.own_stream = true;
        auto up = [&](const void* h, size_t bytes, void** d) {
            cudaError_t r = cudaMalloc(d, bytes);
            if (r == cudaSuccess) r = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
            return r;
        };
        if (xo) {
            e = up(c->xo_digit_polys.data(), c->xo_digit_polys.size() * 8, (void**)&s.d_xo_digit_polys);
            if (e == cudaSuccess) e = up(c->xo_lane_polys.data(), c->xo_lane_polys.size() * 8, (void**)&s.d_xo_lane_polys);
        } else {
            std::vector<uint32_t> t(c->ranf_table);  // the seeded round 0 rides behind the table (fe_scan.cuh)
            t.insert(t.end(), c->ranf_base, c->ranf_base + kRanfLag);
            e = up(t.data(), t.size() * 4, (void**)&s.d_ranf_table);
        }
        c->devs.push_back(s);
        if (e != cudaSuccess) return fail(TP3_E_CUDA, cudaGetErrorString(e));
    }
    *out = c;
    return TP3_OK;
}

void tp3_destroy(tp3_ctx* c) {
    if (!c) return;
    for (auto& s : c->devs) {
        cudaSetDevice(s.dev);
        if (s.stream) cudaStreamSynchronize(s.stream);  // nothing of this context is in flight when its buffers go
        if (s.own_stream && s.stream) cudaStreamDestroy(s.stream);
        cudaFree(s.d_ranf_table);
        cudaFree(s.d_batch_parts);
        cudaFree(s.d_xo_digit_polys);
        cudaFree(s.d_xo_lane_polys);
        cudaFree(s.d_xo_states);
        cudaFree(s.d_xo_scan);
        cudaFree(s.d_xo_bnd);
        cudaFree(s.d_out);
        cudaFree(s.d_fold);
        cudaFree(s.d_unit_done);
        for (auto& b : s.fs) {
            cudaFree(b.d_records);
            cudaFree(b.d_count);
            cudaFree(b.d_exit);
            cudaFree(b.d_fail);
            cudaFree(b.d_seg_events);
            cudaFree(b.d_redo_list);
            cudaFree(b.d_redo_entry);
            cudaFree(b.d_unit_seg);
            cudaFree(b.d_parts);
        }
        if (s.fs_stream2) {
            cudaStreamSynchronize(s.fs_stream2);
            cudaStreamDestroy(s.fs_stream2);
        }
        if (s.fs_event) cudaEventDestroy(s.fs_event);
        if (s.copy_stream) cudaStreamDestroy(s.copy_stream);
        if (s.copy_event) cudaEventDestroy(s.copy_event);
        if (s.copy_event2) cudaEventDestroy(s.copy_event2);
        if (s.h_progress) cudaFreeHost(s.h_progress);
        for (auto& b : s.fs)
            if (b.phys_done) cudaEventDestroy(b.phys_done);
        cudaFree(s.d_hist_counts);
        cudaFree(s.d_hist_weights);
    }
    delete c;
}

// ---- per-event observables (tp3.h) ---------------------------------------------------------------------
int tp3_histograms_reset(tp3_ctx* c) {
    if (!c) return TP3_E_INVALID;
    const size_t n = (size_t)TP3_HIST_OBSERVABLES * c->hist_bins;
    for (auto& s : c->devs) {
        if (!s.d_hist_counts) continue;
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaMemsetAsync(s.d_hist_counts, 0, n * sizeof(unsigned long long), s.stream));
        TP3_CUDA(c, cudaMemsetAsync(s.d_hist_weights, 0, n * kHistReplicas * sizeof(double), s.stream));
    }
    return TP3_OK;
}

int tp3_histograms_enable(tp3_ctx* c, uint32_t num_bins) {
    if (!c) return TP3_E_INVALID;
    if (num_bins > TP3_HIST_MAX_BINS) {
        c->err = "tp3_histograms_enable: at most " + std::to_string(TP3_HIST_MAX_BINS) + " bins";
        return TP3_E_INVALID;
    }
    // faster-evgen: the stream pipeline of the sequential RANF stream has the epilogue (fe_stream.cuh), the kernels that serve
    // xoshiro and jump() seeding (one thread per batch / one lane per 313 events) do not
    const bool fe = c->params.flags & TP3_FASTER_EVGEN;
    const bool fe_stream = fe && !(c->params.flags & (TP3_STANDARD_RANDOM | TP3_FASTER_THREADING));
    if (num_bins && ((fe && !fe_stream) || c->params.kernel != TP3_KERNEL_FAST)) {
        c->err = "per-event observables need the fast kernel and, under faster-evgen, the sequential RANF stream";
        return TP3_E_INVALID;
    }
    for (auto& s : c->devs) {
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));
        cudaFree(s.d_hist_counts);
        cudaFree(s.d_hist_weights);
        s.d_hist_counts = nullptr;
        s.d_hist_weights = nullptr;
        if (num_bins) {
            const size_t n = (size_t)TP3_HIST_OBSERVABLES * num_bins;
            TP3_CUDA(c, cudaMalloc(&s.d_hist_counts, n * sizeof(unsigned long long)));
            TP3_CUDA(c, cudaMalloc(&s.d_hist_weights, n * kHistReplicas * sizeof(double)));
        }
    }
    c->hist_bins = num_bins;
    return tp3_histograms_reset(c);
}

int tp3_histograms_fetch(tp3_ctx* c, uint64_t* counts, double* weights) {
    if (!c || !counts || !weights) return TP3_E_INVALID;
    if (!c->hist_bins) {
        c->err = "tp3_histograms_fetch: histograms are not enabled";
        return TP3_E_INVALID;
    }
    const size_t n = (size_t)TP3_HIST_OBSERVABLES * c->hist_bins;
    std::vector<unsigned long long> hc(n);
    std::vector<double> hw(n * kHistReplicas);
    std::fill(counts, counts + n, 0);
    std::fill(weights, weights + n, 0.0);
    for (auto& s : c->devs) {  // device order: the sum does not depend on how the work was timed
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaMemcpyAsync(hc.data(), s.d_hist_counts, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        TP3_CUDA(c, cudaMemcpyAsync(hw.data(), s.d_hist_weights, n * kHistReplicas * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));
        for (size_t i = 0; i < n; ++i) {
            counts[i] += hc[i];
            double w = 0;
            for (int r = 0; r < kHistReplicas; ++r) w += hw[(size_t)r * n + i];  // replica order
            weights[i] += w;
        }
    }
    return TP3_OK;
}

int tp3_set_stream(tp3_ctx* c, int slot, void* stream) {
    if (!c || slot < 0 || slot >= (int)c->devs.size()) return TP3_E_INVALID;
    DeviceSlot& s = c->devs[slot];
    TP3_CUDA(c, cudaSetDevice(s.dev));
    if (s.stream) TP3_CUDA(c, cudaStreamSynchronize(s.stream));  // work queued on the old stream finishes before the switch
    if (s.own_stream && s.stream) TP3_CUDA(c, cudaStreamDestroy(s.stream));
    s.stream = (cudaStream_t)stream;
    s.own_stream = false;
    return TP3_OK;
}

uint64_t tp3_launch_count(const tp3_ctx* c) { return c ? c->launches : 0; }

size_t tp3_kernel_arg_bytes(void) { return sizeof(SimArgs) + sizeof(PhysParams<double>); }

int tp3_synchronize(tp3_ctx* c) {
    if (!c) return TP3_E_INVALID;
    for (auto& s : c->devs) {
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));
    }
    return TP3_OK;
}

int tp3_simulate_batches_device(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len) {
    if (!c || n == 0 || last_len == 0 || last_len > TP3_EVENT_BATCH_SIZE) {
        if (c) c->err = "tp3_simulate_batches: bad range";
        return TP3_E_INVALID;
    }
    const size_t G = c->devs.size();
    for (size_t g = 0; g < G; ++g) {
        uint64_t off, cnt;
        split(n, G, g, off, cnt);
        c->devs[g].last_n = 0;
        if (!cnt) continue;
        const uint32_t ll = (off + cnt == n) ? last_len : TP3_EVENT_BATCH_SIZE;
        int rc = enqueue_range(c, c->devs[g], first + off, cnt, ll);
        if (rc) return rc;
    }
    return TP3_OK;
}

int tp3_fetch(tp3_ctx* c, tp3_acc* out, uint64_t n) {
    if (!c || !out) return TP3_E_INVALID;
    uint64_t done = 0;
    for (auto& s : c->devs) {
        if (!s.last_n) continue;
        if (done + s.last_n > n) {
            c->err = "tp3_fetch: output array too small";
            return TP3_E_INVALID;
        }
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaMemcpyAsync(out + done, s.d_out, s.last_n * sizeof(tp3_acc), cudaMemcpyDeviceToHost, s.stream));
        done += s.last_n;
    }
    return tp3_synchronize(c);
}

// Large single-device launches of the fused kernel: the per-batch accumulators are copied to the caller's array WHILE the
// kernel runs.  The in-kernel ordered fold doubles as the progress indicator: units [0, next_unit) are complete, in batch
// order, so their accumulators can go (104 MB per 1e6 batches would otherwise be copied after the kernel: 5 % of a run).
static int simulate_batches_streamed(tp3_ctx* c, DeviceSlot& s, uint64_t first, uint64_t n, uint32_t last_len, tp3_acc* out) {
    TP3_CUDA(c, cudaSetDevice(s.dev));
    if (!s.copy_stream) {
        TP3_CUDA(c, cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
        TP3_CUDA(c, cudaEventCreateWithFlags(&s.copy_event, cudaEventDisableTiming));
        TP3_CUDA(c, cudaEventCreateWithFlags(&s.copy_event2, cudaEventDisableTiming));
        TP3_CUDA(c, cudaHostAlloc(&s.h_progress, sizeof(unsigned long long), cudaHostAllocDefault));
    }
    int rc = enqueue_range(c, s, first, n, last_len, /*fold=*/true);
    if (rc) return rc;
    TP3_CUDA(c, cudaEventRecord(s.copy_event, s.stream));
    // the progress word is only meaningful once this launch's reset of the fold state has run (earlier work may still be queued)
    TP3_CUDA(c, cudaStreamWaitEvent(s.copy_stream, s.copy_event2, 0));
    const SimArgs& a = s.last_args;
    auto batches_of = [&](uint64_t units) -> uint64_t {  // batches covered by units [0, units) of the dynamic schedule (kernels.cuh)
        if (units == 0) return 0;
        uint64_t lo, hi;
        unit_range_dynamic(a.dynamic - 1, a.full_rounds, a.unit_batches, a.taper, units - 1, lo, hi);
        return std::min<uint64_t>(n, hi);
    };
    uint64_t copied = 0;
    const uint64_t chunk = std::max<uint64_t>(n / 64, 4096);
    for (;;) {
        const bool finished = cudaEventQuery(s.copy_event) == cudaSuccess;
        uint64_t ready = n;
        if (!finished) {
            TP3_CUDA(c, cudaMemcpyAsync(s.h_progress, &s.d_fold->next_unit, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.copy_stream));
            TP3_CUDA(c, cudaStreamSynchronize(s.copy_stream));
            ready = std::max(copied, batches_of(*s.h_progress));
        }
        if (ready - copied >= chunk || (finished && ready > copied)) {
            TP3_CUDA(c, cudaMemcpyAsync(out + copied, s.d_out + copied, (ready - copied) * sizeof(tp3_acc), cudaMemcpyDeviceToHost, s.copy_stream));
            TP3_CUDA(c, cudaStreamSynchronize(s.copy_stream));
            copied = ready;
        }
        if (finished && copied == n) break;
        if (!finished) std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) {
        c->err = std::string("tp3_simulate_batches: ") + cudaGetErrorString(e);
        return TP3_E_CUDA;
    }
    return TP3_OK;
}

// Long single-device launches of the fused kernel copy their accumulators to the host while the kernel runs.
static bool streamed_per_batch(const tp3_ctx* c, uint64_t n, uint32_t last_len) {
    return c && c->devs.size() == 1 && n >= 32768 && n <= 0x7fffffffull && last_len >= 1 && last_len <= TP3_EVENT_BATCH_SIZE &&
           !(c->params.flags & TP3_FASTER_EVGEN) && c->opt_sched_dynamic && !c->hist_bins && c->opt_batch_parts <= 1;
}

int tp3_simulate_batches(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len, tp3_acc* out) {
    if (!out) return TP3_E_INVALID;
    if (streamed_per_batch(c, n, last_len)) {
        c->devs[0].last_n = 0;
        return simulate_batches_streamed(c, c->devs[0], first, n, last_len, out);
    }
    int rc = tp3_simulate_batches_device(c, first, n, last_len);
    if (rc) return rc;
    return tp3_fetch(c, out, n);
}

// One launch per device with the in-kernel ordered fold; leaves every slot's result in s.d_fold (asynchronous).
static int enqueue_merged(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len) {
    const size_t G = c->devs.size();
    for (size_t g = 0; g < G; ++g) {
        DeviceSlot& s = c->devs[g];
        uint64_t off, cnt;
        split(n, G, g, off, cnt);
        s.last_n = 0;
        if (!cnt) continue;
        const uint32_t range_last = (off + cnt == n) ? last_len : TP3_EVENT_BATCH_SIZE;
        int rc = enqueue_range(c, s, first + off, cnt, range_last, /*fold=*/true);
        if (rc) return rc;
    }
    return TP3_OK;
}

// Waits for the launches of enqueue_merged / simulate_batches_streamed and collects the folded accumulator of every device.
static int collect_merged(tp3_ctx* c, tp3_acc* out) {
    std::vector<FoldState> parts(c->devs.size());
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceSlot& s = c->devs[g];
        if (!s.last_n) continue;
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaMemcpyAsync(&parts[g], s.d_fold, sizeof(FoldState), cudaMemcpyDeviceToHost, s.stream));
    }
    bool have = false;
    for (size_t g = 0; g < c->devs.size(); ++g) {
        DeviceSlot& s = c->devs[g];
        if (!s.last_n) continue;
        TP3_CUDA(c, cudaSetDevice(s.dev));
        TP3_CUDA(c, cudaStreamSynchronize(s.stream));
        if (s.fold_in_kernel && (parts[g].lock != 0 || parts[g].next_unit == 0)) {  // every unit was folded and the lock released
            c->err = "ordered fold did not complete";
            return TP3_E_CUDA;
        }
        // Device partials are merged in device (= batch) order: with several devices the result is a fold of per-device
        // folds, reproducible for a fixed device count (the reference's FastAccumulator, multi_threading.rs:130-190, is
        // order-insensitive in the same way); the per-batch path (tp3_simulate_batches + tp3_fold_batches) does not
        // depend on the device count.
        if (!have) *out = parts[g].running;
        else tp3_merge(out, &parts[g].running, c->params.flags);
        have = true;
    }
    return TP3_OK;
}

int tp3_simulate_merged(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len, tp3_acc* out) {
    if (!c || !out || n == 0 || last_len == 0 || last_len > TP3_EVENT_BATCH_SIZE) {
        if (c) c->err = "tp3_simulate_merged: bad range";
        return TP3_E_INVALID;
    }
    int rc = enqueue_merged(c, first, n, last_len);
    if (rc) return rc;
    return collect_merged(c, out);
}

int tp3_simulate_batches_merged(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len, tp3_acc* out, tp3_acc* merged) {
    if (!c || !out || !merged || n == 0 || last_len == 0 || last_len > TP3_EVENT_BATCH_SIZE) {
        if (c) c->err = "tp3_simulate_batches_merged: bad range";
        return TP3_E_INVALID;
    }
    if (streamed_per_batch(c, n, last_len)) {
        // the accumulators are copied while the kernel runs, and that launch already carries the ordered fold
        DeviceSlot& s = c->devs[0];
        s.last_n = 0;
        int rc = simulate_batches_streamed(c, s, first, n, last_len, out);
        if (rc) return rc;
        return collect_merged(c, merged);
    }
    int rc = enqueue_merged(c, first, n, last_len);
    if (rc) return rc;
    if ((rc = tp3_fetch(c, out, n))) return rc;
    return collect_merged(c, merged);
}

// 13 doubles for one ncclReduce(sum): the event count is exact as a double below 2^53.
__global__ void export_merged_kernel(const FoldState* fs, double* out13) {
    const int i = threadIdx.x;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&fs->running);
    if (i == 0) out13[0] = (double)src[0];
    else if (i < 13) out13[i] = __longlong_as_double((long long)src[i]);
}

int tp3_simulate_merged_device(tp3_ctx* c, uint64_t first, uint64_t n, uint32_t last_len, double* d_out13) {
    if (!c || !d_out13 || n == 0 || last_len == 0 || last_len > TP3_EVENT_BATCH_SIZE || c->devs.size() != 1) {
        if (c) c->err = "tp3_simulate_merged_device: bad range, or a context with several devices";
        return TP3_E_INVALID;
    }
    int rc = enqueue_merged(c, first, n, last_len);
    if (rc) return rc;
    DeviceSlot& s = c->devs[0];
    export_merged_kernel<<<1, 32, 0, s.stream>>>(s.d_fold, d_out13);
    ++c->launches;
    TP3_CUDA(c, cudaGetLastError());
    return TP3_OK;
}

int tp3_fe_tile_device(tp3_ctx* c, uint64_t first_round, uint64_t n_rounds, uint64_t max_events, double* d_out13, uint64_t* events_done) {
    const uint32_t need = TP3_FASTER_EVGEN, bad = TP3_STANDARD_RANDOM | TP3_FASTER_THREADING;
    if (!c || !d_out13 || !events_done || c->devs.size() != 1 || (c->params.flags & need) != need || (c->params.flags & bad) ||
        (n_rounds == 0 && max_events == 0) || first_round % 64 || n_rounds % 64) {
        if (c) c->err = "tp3_fe_tile_device: needs a single-device faster-evgen context on the sequential RANF stream, a round or an "
                        "event limit, and round numbers that are multiples of 64";
        return TP3_E_INVALID;
    }
    DeviceSlot& s = c->devs[0];
    TP3_CUDA(c, cudaSetDevice(s.dev));
    // accumulators per group of 10 000 consecutive events of the tile: at most 4 events start in a round
    const uint64_t cap = (max_events ? max_events : n_rounds * kFeSlotsPerRound) / TP3_EVENT_BATCH_SIZE + 2;
    int rc = ensure_out(c, s, cap);
    if (rc) return rc;
    FeTile t{first_round, n_rounds, max_events, 0, 0};
    rc = fe_stream_simulate(c, s, 0, 0, 0, &t);
    if (rc) return rc;
    *events_done = t.events_done;
    TP3_CUDA(c, cudaMemsetAsync(s.d_fold, 0, sizeof(FoldState), s.stream));
    if (t.batches_done) {
        if (c->params.flags & TP3_F32) merge_kernel<float><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, t.batches_done, &s.d_fold->running, true);
        else merge_kernel<double><<<1, kMergeThreads, 0, s.stream>>>(s.d_out, t.batches_done, &s.d_fold->running, true);
        ++c->launches;
        TP3_CUDA(c, cudaGetLastError());
    }
    export_merged_kernel<<<1, 32, 0, s.stream>>>(s.d_fold, d_out13);
    ++c->launches;
    TP3_CUDA(c, cudaGetLastError());
    return TP3_OK;
}

int tp3_set_option(tp3_ctx* c, const char* name, int64_t value) {
    if (!c || !name) return TP3_E_INVALID;
    const std::string k(name);
    if (k == "unit_batches" && value >= 0 && value <= 4096) c->opt_unit_batches = value;
    else if (k == "grid_warps" && value >= 0 && value <= (1 << 20)) c->opt_grid_warps = value;
    else if (k == "sched_dynamic") c->opt_sched_dynamic = value != 0;
    else if (k == "ramp_units" && value >= 0 && value <= (1 << 20)) c->opt_ramp_units = value;
    else if (k == "taper_units" && value >= -1 && value <= (1 << 20)) c->opt_taper_units = value;
    else if (k == "tail_singles" && value >= 0 && value <= (1 << 22)) c->opt_tail_singles = value;
    else if (k == "align_units") c->opt_align_units = value != 0;
    else if (k == "batch_parts" && (value == 0 || value == 1 || value == 2 || value == 5 || value == 10)) c->opt_batch_parts = value;
    else if (k == "f32_scalar") c->opt_f32_scalar = value != 0;
    else if (k == "fe_split" && (value == 0 || value == 1 || value == 32)) c->opt_fe_split = value;
    else if (k == "fe_host_scan") c->opt_fe_host_scan = value != 0;
    else if (k == "fe_xo_seg_units" && value >= 0 && value <= 64) c->opt_fe_xo_seg_units = value;
    else if (k == "fe_timing") c->opt_fe_timing = value != 0;
    else if (k == "fe_legacy") c->opt_fe_legacy = value != 0;
    else if (k == "fe_pass_segments" && value >= 0 && value <= (1 << 24)) c->opt_fe_pass_segments = value;
    else if (k == "fe_seg_rounds" && value >= 0 && value <= 4096) c->opt_fe_seg_rounds = value;
    else if (k == "fe_warm" && value >= 0 && value <= 1024) c->opt_fe_warm = value;
    else if (k == "fe_serial") c->opt_fe_serial = value != 0;
    else {
        c->err = "tp3_set_option: unknown option or value out of range: " + k;
        return TP3_E_INVALID;
    }
    return TP3_OK;
}

int tp3_get_stat(tp3_ctx* c, const char* name, int64_t* value) {
    if (!c || !name || !value) return TP3_E_INVALID;
    const std::string k(name);
    if (k == "fe_xo_pass_b") *value = c->stat_fe_xo_pass_b;
    else if (k == "fe_passes") *value = c->stat_fe_passes;
    else if (k == "fe_redone") *value = c->stat_fe_redone;
    else {
        c->err = "tp3_get_stat: unknown statistic: " + k;
        return TP3_E_INVALID;
    }
    return TP3_OK;
}

static int run_dump(tp3_ctx* c, uint64_t batch, uint32_t n_events, uint64_t* words, double* momenta, int32_t* kept,
                    double* m2) {
    if (c->params.flags & TP3_FASTER_EVGEN) {
        c->err = "per-event dumps are not available under faster-evgen (draw positions are data dependent)";
        return TP3_E_INVALID;
    }
    if (!(c->params.flags & TP3_STANDARD_RANDOM) && (c->params.flags & TP3_FASTER_THREADING) && batch >= 6200) {
        c->err = "RANF jump() seeding is only defined for the first 6200 batches (seed < 1e9)";
        return TP3_E_INVALID;
    }
    DeviceSlot& s = c->devs[0];
    TP3_CUDA(c, cudaSetDevice(s.dev));
    int rc = ensure_out(c, s, 1);
    if (rc) return rc;
    SimArgs a = make_args(c, s, batch, 1, TP3_EVENT_BATCH_SIZE);
    if (c->params.flags & TP3_STANDARD_RANDOM) {
        if (c->params.flags & TP3_F32)
            xoshiro_seed_kernel<Xoshiro128Lane><<<1, 32, 0, s.stream>>>(batch, 1, s.d_xo_digit_polys, c->xo_digits, c->xo_base[0],
                                                                        c->xo_base[1], c->xo_base[2], c->xo_base[3], s.d_xo_states);
        else
            xoshiro_seed_kernel<Xoshiro256Lane><<<1, 32, 0, s.stream>>>(batch, 1, s.d_xo_digit_polys, c->xo_digits, c->xo_base[0],
                                                                        c->xo_base[1], c->xo_base[2], c->xo_base[3], s.d_xo_states);
        ++c->launches;
    }
    DumpArgs d;
    std::memset(&d, 0, sizeof d);
    d.n_events = n_events;
    cudaError_t e = cudaSuccess;  // the buffers allocated so far are freed below whatever fails
    if (words) e = cudaMalloc(&d.words, (size_t)n_events * 12 * 8);
    if (momenta) {
        if (e == cudaSuccess) e = cudaMalloc(&d.momenta, (size_t)n_events * 12 * 8);
        if (e == cudaSuccess) e = cudaMalloc(&d.kept, (size_t)n_events * 4);
        if (e == cudaSuccess) e = cudaMalloc(&d.m2, (size_t)n_events * 5 * 8);
    }
    if (e == cudaSuccess) {
        pick_dump(c->params, c->opt_f32_scalar != 0)(a, c->params, d, s.stream);
        ++c->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e == cudaSuccess && words) e = cudaMemcpy(words, d.words, (size_t)n_events * 12 * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && momenta) {
        e = cudaMemcpy(momenta, d.momenta, (size_t)n_events * 12 * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(kept, d.kept, (size_t)n_events * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(m2, d.m2, (size_t)n_events * 5 * 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(d.words);
    cudaFree(d.momenta);
    cudaFree(d.kept);
    cudaFree(d.m2);
    if (e != cudaSuccess) {
        c->err = std::string("dump: ") + cudaGetErrorString(e);
        return TP3_E_CUDA;
    }
    return TP3_OK;
}

int tp3_rng_dump(tp3_ctx* c, uint64_t batch, uint32_t n_words, uint64_t* out) {
    if (!c || !out || n_words == 0 || n_words > 12u * TP3_EVENT_BATCH_SIZE) return TP3_E_INVALID;
    const uint32_t n_events = (n_words + 11) / 12;
    std::vector<uint64_t> tmp((size_t)n_events * 12);
    int rc = run_dump(c, batch, n_events, tmp.data(), nullptr, nullptr, nullptr);
    if (rc) return rc;
    std::memcpy(out, tmp.data(), (size_t)n_words * 8);
    return TP3_OK;
}

int tp3_events_dump(tp3_ctx* c, uint64_t batch, uint32_t n, double* momenta, int32_t* kept, double* m2) {
    if (!c || !momenta || !kept || !m2 || n == 0 || n > TP3_EVENT_BATCH_SIZE) return TP3_E_INVALID;
    return run_dump(c, batch, n, nullptr, momenta, kept, m2);
}

int tp3_peak_probe(tp3_ctx* c, int which, double* tflops) {
    if (!c || !tflops) return TP3_E_INVALID;
    DeviceSlot& s = c->devs[0];
    TP3_CUDA(c, cudaSetDevice(s.dev));
    cudaDeviceProp prop;
    TP3_CUDA(c, cudaGetDeviceProperties(&prop, s.dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    const int iters = which == 0 ? 4096 : 8192;
    void* buf = nullptr;
    TP3_CUDA(c, cudaMalloc(&buf, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    TP3_CUDA(c, cudaEventCreate(&e0));
    TP3_CUDA(c, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        TP3_CUDA(c, cudaEventRecord(e0, s.stream));
        if (which == 0) fma_probe_kernel<double><<<blocks, threads, 0, s.stream>>>((double*)buf, iters, 0.999999);
        else fma_probe_kernel<float><<<blocks, threads, 0, s.stream>>>((float*)buf, iters, 0.999999f);
        ++c->launches;
        TP3_CUDA(c, cudaEventRecord(e1, s.stream));
        TP3_CUDA(c, cudaEventSynchronize(e1));
        float ms;
        TP3_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    const double fma = (double)blocks * threads * (double)iters * 64.0;
    *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    return TP3_OK;
}

int tp3_fastmath_probe(tp3_ctx* c, int which, uint32_t n, const double* in, double* out) {
    if (!c || !in || !out || n == 0 || which < 0 || which > 10) return TP3_E_INVALID;
    DeviceSlot& s = c->devs[0];
    TP3_CUDA(c, cudaSetDevice(s.dev));
    double *d_in = nullptr, *d_out = nullptr;
    TP3_CUDA(c, cudaMalloc(&d_in, (size_t)n * 8));
    TP3_CUDA(c, cudaMalloc(&d_out, (size_t)n * 8));
    cudaError_t e = cudaMemcpyAsync(d_in, in, (size_t)n * 8, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) {
        fastmath_probe_kernel<<<148, 256, 0, s.stream>>>(which, n, d_in, d_out, FastCoef TP3_FAST_COEF_INIT);
        ++c->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        c->err = std::string("fastmath probe: ") + cudaGetErrorString(e);
        return TP3_E_CUDA;
    }
    return TP3_OK;
}

int tp3_host_ranf_round(int32_t seed, uint64_t round, uint32_t* out55) {
    if (!out55 || seed <= 0 || seed >= (int32_t)RANF_MOD) return TP3_E_INVALID;
    static const std::vector<uint32_t> table = ranf_round_jump_table();
    ranf_round_at(table, seed, round, out55);
    return TP3_OK;
}

int tp3_host_xoshiro_state(int f32, uint64_t n_steps, uint64_t n_jumps, uint64_t* out4) {
    if (!out4) return TP3_E_INVALID;
    if (f32) {
        Xoshiro128 g = xoshiro128_seed(12345);
        Gf2Mod mod = xoshiro_min_poly(g, 128);
        Gf2Poly p = gf2_mul_mod(gf2_x_pow(n_steps, 0, mod), gf2_pow(gf2_x_pow(1, 64, mod), n_jumps, mod), mod);
        xoshiro_apply(g, p, 128);
        for (int i = 0; i < 4; ++i) out4[i] = g.s[i];
    } else {
        Xoshiro256 g = xoshiro256_seed(12345);
        Gf2Mod mod = xoshiro_min_poly(g, 256);
        Gf2Poly p = gf2_mul_mod(gf2_x_pow(n_steps, 0, mod), gf2_pow(gf2_x_pow(1, 128, mod), n_jumps, mod), mod);
        xoshiro_apply(g, p, 256);
        for (int i = 0; i < 4; ++i) out4[i] = g.s[i];
    }
    return TP3_OK;
}

}  // extern "C"

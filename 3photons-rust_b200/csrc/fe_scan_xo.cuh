// `faster-evgen` + `standard-random`, sequential stream: where every batch starts, computed on the GPU.
//
// An event consumes 9 + 6 + 2 k xoshiro outputs, k = re-rolls of the three unit-disc points (src/evgen.rs:143-173,
// 221-249), so the stream position of batch b depends on every earlier event; the reference's scheduler thread walks
// the stream event by event (evgen.rs:257-267, multi_threading.rs:59-64).  xoshiro has no round structure to hang
// transition maps on (fe_scan.cuh does that for RANF), but walks that start at different positions COALESCE: two walks
// land on a common event start with probability ~1/40 per event (measured) and are one walk from there on.  So the stream is
// cut into segments of L positions and
//   pass A  one lane per segment walks from the segment start and records where its last event ends (the exit offset
//           into the next segment) — which, after the first few hundred positions, no longer depends on where it entered;
//   pass B  every lane walks again from the entry the previous segment's pass-A exit implies, counting its events, and
//           checks that its exit equals pass A's.  A mismatch (3.7 % of 2048-output segments — all event lengths are
//           odd, which slows the merging down —, ~1e-4 of 8192-output ones, none in practice from 32768 on) just repeats
//           pass B with the corrected exits; after pass t the first t segments are right whatever happened, so the loop
//           ends, and it ends with every entry, exit and count equal to the sequential walk's;
//   pass C  the host prefix-sums the counts, finds the segment of every wanted event index (10000 b) and one thread per
//           batch walks from that segment's entry to the event and writes the generator state the batch kernel starts from.
// Segment start states come from the same GF(2) jump-ahead as the batch seeding (digit tables of x^(2048 d 256^k)).
// The accept / re-roll test is evaluated exactly as in faster_evgen.cuh (no FMA contraction).
#pragma once

#include "faster_evgen.cuh"

namespace tp3 {

constexpr uint32_t kXoSegUnit = 2048;  // the jump tables step by this many outputs; a segment is a multiple of it

template <class F> struct XoScanGen;
template <> struct XoScanGen<double> {
    using Lane = Xoshiro256Lane;
    __device__ static __forceinline__ double uniform(Lane& g) { return to_uniform_xo(g.next()); }
};
template <> struct XoScanGen<float> {
    using Lane = Xoshiro128Lane;
    __device__ static __forceinline__ float uniform(Lane& g) { return to_uniform_xo(g.next()); }
};

template <class Lane> __device__ __forceinline__ void xo_load(Lane& g, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3) {
    g.s0 = (decltype(g.s0))b0; g.s1 = (decltype(g.s0))b1; g.s2 = (decltype(g.s0))b2; g.s3 = (decltype(g.s0))b3;
}

// base advanced by `units` x kXoSegUnit outputs (digit_polys[k][d] = x^(kXoSegUnit d 256^k))
template <class Lane>
__device__ __forceinline__ void xo_jump_units(Lane& g, uint64_t units, const uint64_t* __restrict__ digit_polys, int n_digits) {
    for (int k = 0; k < n_digits; ++k) {
        const unsigned dgt = (unsigned)((units >> (8 * k)) & 0xffu);
        if (dgt) g.apply(digit_polys + ((size_t)k * 256 + dgt) * 4);
    }
}

// One event of the scheduler's pre-advance (evgen.rs:257-267): returns the number of outputs it consumed.
template <class F> __device__ __forceinline__ uint32_t xo_skip_event(typename XoScanGen<F>::Lane& g) {
#pragma unroll
    for (int i = 0; i < 9; ++i) g.next();
    F v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = XoScanGen<F>::uniform(g);
    uint32_t used = 15;
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // the points are settled in order (evgen.rs:231-241)
        F x = (F)2 * v[k] - (F)1, y = (F)2 * v[3 + k] - (F)1;
        F r2 = add_rn(mul_rn(x, x), mul_rn(y, y));
        while (r2 > (F)1) {
            const F w0 = XoScanGen<F>::uniform(g), w1 = XoScanGen<F>::uniform(g);
            used += 2;
            x = (F)2 * w0 - (F)1;
            y = (F)2 * w1 - (F)1;
            r2 = add_rn(mul_rn(x, x), mul_rn(y, y));
        }
    }
    return used;
}

// Passes A and B.  Segment s covers outputs [s L, (s + 1) L) after the base state, L = seg_units x kXoSegUnit.
//   prev_exit == nullptr: enter at offset 0 (pass A);  else: enter segment s at prev_exit[s - 1] (segment 0 at 0).
// Writes exit_off[s] = how far the segment's last event reaches into the next one, count[s] = events started in it,
// and raises *mismatch when check_exit[s] differs from the new exit.
template <class F>
__global__ void __launch_bounds__(128) xo_walk_kernel(uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint64_t n_seg,
                                                      uint32_t seg_units, const uint64_t* __restrict__ digit_polys, int n_digits,
                                                      const uint32_t* __restrict__ prev_exit, const uint32_t* __restrict__ check_exit,
                                                      uint32_t* __restrict__ exit_off, uint32_t* __restrict__ count,
                                                      uint32_t* __restrict__ mismatch) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    typename XoScanGen<F>::Lane g;
    xo_load(g, b0, b1, b2, b3);
    xo_jump_units(g, s * seg_units, digit_polys, n_digits);
    const uint32_t len = seg_units * kXoSegUnit;
    uint32_t pos = (prev_exit && s) ? prev_exit[s - 1] : 0u;
    for (uint32_t i = 0; i < pos; ++i) g.next();
    uint32_t n = 0;
    while (pos < len) {
        pos += xo_skip_event<F>(g);
        ++n;
    }
    exit_off[s] = pos - len;
    count[s] = n;
    if (check_exit && check_exit[s] != pos - len) atomicOr(mismatch, 1u);
}

// Pass C.  Boundary j: the event `skip[j]` events after the entry of segment seg[j]; out[j] = generator state there.
struct XoBoundary {
    uint64_t seg;
    uint32_t skip;
    uint32_t pad;
};
template <class F>
__global__ void __launch_bounds__(128) xo_boundary_states_kernel(uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint32_t seg_units,
                                                                 const uint64_t* __restrict__ digit_polys, int n_digits,
                                                                 const uint32_t* __restrict__ exit_off,
                                                                 const XoBoundary* __restrict__ bnd, uint64_t n_bnd,
                                                                 uint64_t* __restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bnd) return;
    const XoBoundary b = bnd[j];
    typename XoScanGen<F>::Lane g;
    xo_load(g, b0, b1, b2, b3);
    xo_jump_units(g, b.seg * seg_units, digit_polys, n_digits);
    const uint32_t entry = b.seg ? exit_off[b.seg - 1] : 0u;
    for (uint32_t i = 0; i < entry; ++i) g.next();
    for (uint32_t e = 0; e < b.skip; ++e) xo_skip_event<F>(g);
    out[4 * j + 0] = g.s0; out[4 * j + 1] = g.s1; out[4 * j + 2] = g.s2; out[4 * j + 3] = g.s3;
}

}  // namespace tp3

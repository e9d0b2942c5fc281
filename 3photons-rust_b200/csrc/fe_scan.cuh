// `faster-evgen` with RANF: where does every batch start in the stream?  (GPU replacement for the
// reference scheduler's sequential pre-advance, src/evgen.rs:257-267.)
//
// Every request of the generator (9, 6 or 2 numbers) is served from ONE 55-number round: a request that does
// not fit discards the rest of the round (ranf.rs:87-92).  So at the start of every round the consumer is in
// one of nine states,
//     0            about to draw the 9 numbers of a new event
//     1            about to draw the 6 numbers of the three unit-disc points
//     2 + 2 f1 + f2  re-rolling point 0, with flags "point 1 / point 2 (as first drawn) is outside the disc"
//     6 + f2       re-rolling point 1
//     8            re-rolling point 2
// and what happens inside the round depends only on that state and on the round's 55 numbers.  Hence
//   1. fe_round_maps_kernel   for every round and every entry state that can occur: exit state + number of events
//                             started (one lane per segment of consecutive rounds, private generator per lane);
//   2. fe_segment_kernel      composes the maps of the rounds of a segment (one thread per segment);
//   3. (host)                 chains the segment maps: entry state and event count at every segment start;
//   4. fe_boundaries_kernel   re-walks the per-round maps of a segment with its true entry state and records
//                             for every boundary (event index 10000 b, or 10000 b + 313 l when the batch kernel
//                             runs one lane per 313 events) the round, entry state and rank of that event start
//                             inside the round;
//   5. fe_batch_states_kernel regenerates that round by jump-ahead, walks to the boundary and writes the
//                             generator state (numbers[0..55], index) the batch kernel starts from.
// The accept / re-roll test is evaluated exactly as the reference does (run's Float, no FMA contraction).
#pragma once

#include "faster_evgen.cuh"

namespace tp3 {

constexpr int kFeMaxSegRounds = 2048;  // rounds per segment (the host picks a power of two up to this)
constexpr int kFeTile = 32;             // segments walked in parallel by one warp

template <class F> __device__ __forceinline__ bool fe_outside(uint32_t a, uint32_t b) {
    const F ua = sizeof(F) == 8 ? (F)((double)(int)a * 1e-9) : (F)((float)(int)a * 1e-9f);
    const F ub = sizeof(F) == 8 ? (F)((double)(int)b * 1e-9) : (F)((float)(int)b * 1e-9f);
    const F x = (F)2 * ua - (F)1, y = (F)2 * ub - (F)1;
    return add_rn(mul_rn(x, x), mul_rn(y, y)) > (F)1;
}

// Walk one round (slots[i] = numbers[i + 1]) from entry state `s`. Returns the exit state; `count` = events
// started. If stop_k >= 0, stops right before serving the stop_k-th event start of the round and returns the
// generator's `index` at that moment in *stop_index.
template <class F>
__device__ __forceinline__ int fe_walk_round(const uint32_t* slots, int stride, int s, int& count, int stop_k, int* stop_index) {
    int idx = 55;
    count = 0;
    for (;;) {
        if (s == 0) {
            if (idx < 9) break;
            if (count == stop_k) {
                *stop_index = idx;
                return s;
            }
            idx -= 9;
            ++count;
            s = 1;
        } else if (s == 1) {
            if (idx < 6) break;
            idx -= 6;
            const bool n0 = fe_outside<F>(slots[(idx + 0) * stride], slots[(idx + 3) * stride]);
            const bool n1 = fe_outside<F>(slots[(idx + 1) * stride], slots[(idx + 4) * stride]);
            const bool n2 = fe_outside<F>(slots[(idx + 2) * stride], slots[(idx + 5) * stride]);
            s = n0 ? 2 + 2 * (int)n1 + (int)n2 : n1 ? 6 + (int)n2 : n2 ? 8 : 0;
        } else {
            if (idx < 2) break;
            idx -= 2;
            if (!fe_outside<F>(slots[idx * stride], slots[(idx + 1) * stride])) {
                if (s < 6) {
                    const int f1 = (s - 2) >> 1, f2 = (s - 2) & 1;
                    s = f1 ? 6 + f2 : f2 ? 8 : 0;
                } else if (s < 8) {
                    s = (s - 6) ? 8 : 0;
                } else {
                    s = 0;
                }
            }
        }
    }
    return s;
}

// The disc tests of a round, evaluated once: bit j of A = outside(slots[j], slots[j+1]) (a re-roll served at index j),
// bit j of B = outside(slots[j], slots[j+3]) (point p of a 6-number request served at index j - p).
template <class F> __device__ __forceinline__ void fe_round_masks(const uint32_t* slots, uint64_t& A, uint64_t& B) {
    A = 0;
    B = 0;
#pragma unroll 6
    for (int j = 0; j < kRanfLag - 1; ++j) A |= (uint64_t)fe_outside<F>(slots[j], slots[j + 1]) << j;
#pragma unroll 4
    for (int j = 0; j < kRanfLag - 3; ++j) B |= (uint64_t)fe_outside<F>(slots[j], slots[j + 3]) << j;
}

// fe_walk_round on the precomputed masks, driven by a 72-entry table in shared memory:
// entry [8 s + bits] = next state | (numbers the next state asks for) << 4, where bits = the three disc tests of a
// 6-number request (s = 1) or, in bit 0, the disc test of a re-roll (s >= 2).
__device__ __forceinline__ int fe_need(int s) { return s == 0 ? 9 : s == 1 ? 6 : 2; }
__device__ inline uint8_t fe_walk_entry(int s, int bits) {
    // next state after the 6-number request, indexed by (n0 | n1 << 1 | n2 << 2); after an accepted re-roll, indexed by s
    constexpr uint32_t kAfterSix = 0x0u | (2u << 4) | (6u << 8) | (4u << 12) | (8u << 16) | (3u << 20) | (7u << 24) | (5u << 28);
    constexpr uint64_t kAfterRoll = (0ull << 8) | (8ull << 12) | (6ull << 16) | (7ull << 20) | (0ull << 24) | (8ull << 28) | (0ull << 32);
    const int n = s == 0 ? 1 : s == 1 ? (int)((kAfterSix >> (4 * bits)) & 15u) : (bits & 1) ? s : (int)((kAfterRoll >> (4 * s)) & 15u);
    return (uint8_t)(n | fe_need(n) << 4);
}
__device__ __forceinline__ int fe_walk_masks(uint64_t A, uint64_t B, int s, int& count, const uint8_t* __restrict__ tab) {
    int idx = 55, need = fe_need(s);
    count = 0;
    while (idx >= need) {
        idx -= need;
        count += need >> 3;  // a 9-number request starts an event
        const uint64_t X = need == 6 ? B : A;
        const uint32_t e = tab[8 * s + ((uint32_t)(X >> idx) & 7u)];
        s = (int)(e & 15u);
        need = (int)(e >> 4);
    }
    return s;
}

// map word of a round: bits [4s, 4s+4) = exit state for entry state s, bits [36 + 3s, 36 + 3s + 3) = events started
__device__ __forceinline__ int fe_map_exit(uint64_t m, int s) { return (int)((m >> (4 * s)) & 15u); }
__device__ __forceinline__ int fe_map_count(uint64_t m, int s) { return (int)((m >> (36 + 3 * s)) & 7u); }

// One round for one lane: (GEN) advance the lane's private generator by one round in place (ranf.rs:106-119 on
// row[k] = numbers[k + 1]) and evaluate the disc tests of the new round, every number converted once.
template <class F> __device__ __forceinline__ F fe_disc_term(uint32_t w) {
    const F u = sizeof(F) == 8 ? (F)u32_times(w, 1e-9) : (F)((float)(int)w * 1e-9f);
    const F x = (F)2 * u - (F)1;
    return mul_rn(x, x);
}
template <class F, bool GEN> __device__ __forceinline__ void fe_round_fused(uint32_t* row, uint64_t& A, uint64_t& B) {
    uint32_t a_lo = 0, a_hi = 0, b_lo = 0, b_hi = 0;
    uint32_t nw[kRanfLag];
    F q[kRanfLag];
#pragma unroll
    for (int k = 0; k < kRanfLag; ++k) {
        uint32_t v = row[k];
        if (GEN) {
            v = ranf_sub(v, k < 24 ? row[k + 31] : nw[k - 24]);
            row[k] = v;
        }
        nw[k] = v;
        q[k] = fe_disc_term<F>(v);
        if (k >= 1 && add_rn(q[k - 1], q[k]) > (F)1) {
            if (k - 1 < 32) a_lo |= 1u << ((k - 1) & 31);
            else a_hi |= 1u << ((k - 1) & 31);
        }
        if (k >= 3 && add_rn(q[k - 3], q[k]) > (F)1) {
            if (k - 3 < 32) b_lo |= 1u << ((k - 3) & 31);
            else b_hi |= 1u << ((k - 3) & 31);
        }
    }
    A = ((uint64_t)a_hi << 32) | a_lo;
    B = ((uint64_t)b_hi << 32) | b_lo;
}

struct FeScanSmem {
    uint32_t win[2 * kRanfLag + 2];
    uint32_t tile[kFeTile][kRanfLag + 2];  // one private generator per lane; odd row stride (57): lane = row accesses are conflict free
};

constexpr uint64_t kFeNine4 = 0x111111111ull;  // replicates a 4-bit field nine times
constexpr uint64_t kFeNine3 = 0111111111ull;   // (octal) replicates a 3-bit field nine times

// 1. per-round maps for rounds [first_round, first_round + n_rounds), cut into segments of seg_rounds rounds.
// ONE LANE = ONE SEGMENT, walked round after round with a private generator (started by jump-ahead), so that the
// walk only follows the states that are still possible: all nine entry states of the segment at its first round,
// the image of those nine afterwards.  The images collapse to a single state within a few rounds (requests that do
// not fit leave every walk at a round start); from then on a round costs one walk and its map word carries that one
// (exit, count) pair in all nine fields.  Entries of states that no walk from the segment start can be in are 0;
// nothing downstream ever reads them.
template <class F>
__global__ void __launch_bounds__(128) fe_round_maps_kernel(const uint32_t* __restrict__ jump_table, uint64_t first_round,
                                                          uint64_t n_rounds, uint32_t seg_rounds, uint64_t* __restrict__ maps) {
    __shared__ FeScanSmem sm[4];
    __shared__ uint8_t walk_tab[72];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 72) walk_tab[threadIdx.x] = fe_walk_entry(threadIdx.x >> 3, threadIdx.x & 7);
    __syncthreads();
    FeScanSmem& w = sm[warp];
    const uint64_t seg0 = ((uint64_t)blockIdx.x * 4 + warp) * kFeTile;
    if (seg0 * seg_rounds >= n_rounds) return;
    // the seeded round 0 travels at the end of the jump table allocation (see api.cu)
    const uint32_t* base = jump_table + (size_t)kRanfDigits * 256 * kRanfLag;
    for (int k = 0; k < kFeTile; ++k) {
        const uint64_t r = (seg0 + k) * seg_rounds;
        if (r >= n_rounds) break;  // warp-uniform
        if (k == 0) ranf_jump_to_round(w.win, base, first_round + r, jump_table, lane);
        else ranf_jump_to_round(w.win, w.win, seg_rounds, jump_table, lane);  // the next segment starts seg_rounds later
        for (int i = lane; i < kRanfLag; i += 32) w.tile[k][i] = w.win[i];
        __syncwarp();
    }
    const uint64_t r0 = (seg0 + lane) * seg_rounds;
    const int n_mine = r0 < n_rounds ? (int)min((uint64_t)seg_rounds, n_rounds - r0) : 0;
    const int n_max = __shfl_sync(0xffffffffu, n_mine, 0);  // lane 0 owns the earliest segment, hence the longest
    uint32_t* row = w.tile[lane];
    uint64_t cur = 0x876543210ull;  // where each of the nine segment-entry states is now
    bool single = false;
    for (int t = 0; t < n_max; ++t) {
        if (t >= n_mine) continue;
        uint64_t A, B, m;
        if (t == 0) fe_round_fused<F, false>(row, A, B);
        else fe_round_fused<F, true>(row, A, B);
        if (single) {
            int cnt;
            const int e = fe_walk_masks(A, B, (int)(cur & 15u), cnt, walk_tab);
            m = (uint64_t)e * kFeNine4 | ((uint64_t)cnt * kFeNine3) << 36;
            cur = (uint64_t)e;
        } else {
            uint32_t img = 0;
#pragma unroll
            for (int s = 0; s < 9; ++s) img |= 1u << ((cur >> (4 * s)) & 15u);
            m = 0;
            while (img) {
                const int s = __ffs(img) - 1;
                img &= img - 1;
                int cnt;
                const int e = fe_walk_masks(A, B, s, cnt, walk_tab);
                m |= (uint64_t)e << (4 * s) | (uint64_t)cnt << (36 + 3 * s);
            }
            uint64_t nxt = 0;
#pragma unroll
            for (int s = 0; s < 9; ++s) nxt |= (uint64_t)fe_map_exit(m, (int)((cur >> (4 * s)) & 15u)) << (4 * s);
            cur = nxt;
            single = cur == (cur & 15u) * kFeNine4;
        }
        maps[r0 + t] = m;
    }
}

// 2. composed map of every segment: for each entry state, the exit state and the number of events started
__global__ void fe_segment_kernel(const uint64_t* __restrict__ maps, uint64_t n_rounds, uint32_t seg_rounds,
                                  uint8_t* __restrict__ seg_exit, uint32_t* __restrict__ seg_count) {
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = seg * seg_rounds;
    if (r0 >= n_rounds) return;
    const uint64_t r1 = min(n_rounds, r0 + seg_rounds);
    int cur[9];
    uint32_t cnt[9];
    for (int s = 0; s < 9; ++s) {
        cur[s] = s;
        cnt[s] = 0;
    }
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t m = maps[r];
#pragma unroll
        for (int s = 0; s < 9; ++s) {
            cnt[s] += fe_map_count(m, cur[s]);
            cur[s] = fe_map_exit(m, cur[s]);
        }
    }
    for (int s = 0; s < 9; ++s) {
        seg_exit[seg * 9 + s] = (uint8_t)cur[s];
        seg_count[seg * 9 + s] = cnt[s];
    }
}

struct FeBoundary {
    uint64_t round;   // global round index
    uint32_t state;   // entry state of that round
    uint32_t rank;    // the boundary is the rank-th event start of the round (0-based)
};

// 4. boundaries inside every segment. seg_state / seg_events: entry state and global event index at the segment start.
// Boundary j (global) is the event 10000 (j / split) + part_len (j % split): every batch start (split = 1), or every
// lane start inside every batch (split = 32, part_len = 313).  Boundaries are further apart than the events one round
// can start, so a round holds at most one.
__device__ __forceinline__ uint64_t fe_boundary_event(uint64_t j, uint32_t split, uint32_t part_len) {
    return (j / split) * (uint64_t)kBatch + (j % split) * (uint64_t)part_len;
}
__global__ void fe_boundaries_kernel(const uint64_t* __restrict__ maps, uint64_t first_round, uint64_t n_rounds, uint32_t seg_rounds,
                                     const uint8_t* __restrict__ seg_state, const uint64_t* __restrict__ seg_events,
                                     uint32_t split, uint32_t part_len, uint64_t first_boundary, uint64_t n_boundaries,
                                     FeBoundary* __restrict__ out) {
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = seg * seg_rounds;
    if (r0 >= n_rounds) return;
    const uint64_t r1 = min(n_rounds, r0 + seg_rounds);
    int cur = seg_state[seg];
    uint64_t ev = seg_events[seg];
    // next boundary at or after ev
    const uint64_t rem = ev % kBatch;
    uint64_t j = (ev / kBatch) * split + min((uint64_t)split, (rem + part_len - 1) / part_len);
    uint64_t target = fe_boundary_event(j, split, part_len);
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t m = maps[r];
        const int c = fe_map_count(m, cur);
        if (target < ev + c) {  // (target >= ev by construction)
            if (j >= first_boundary && j < first_boundary + n_boundaries) {
                FeBoundary o;
                o.round = first_round + r;
                o.state = (uint32_t)cur;
                o.rank = (uint32_t)(target - ev);
                out[j - first_boundary] = o;
            }
            ++j;
            target = fe_boundary_event(j, split, part_len);
        }
        ev += c;
        cur = fe_map_exit(m, cur);
    }
}

// 5. generator state at every boundary: numbers[0..55] + index, the layout faster_evgen_kernel reads.  One warp per
// run of `chain` consecutive boundaries: a full jump-ahead to the first one, short relative jumps to the others.
template <class F>
__global__ void __launch_bounds__(128) fe_batch_states_kernel(const uint32_t* __restrict__ jump_table, const FeBoundary* __restrict__ bnd,
                                                            uint64_t n_boundaries, uint32_t chain, uint32_t* __restrict__ states) {
    __shared__ uint32_t win[4][2 * kRanfLag + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t i0 = ((uint64_t)blockIdx.x * 4 + warp) * chain;
    if (i0 >= n_boundaries) return;
    const uint64_t i1 = min(n_boundaries, i0 + chain);
    uint64_t at = 0;
    for (uint64_t i = i0; i < i1; ++i) {
        const FeBoundary b = bnd[i];
        if (i == i0) ranf_jump_to_round(win[warp], jump_table + (size_t)kRanfDigits * 256 * kRanfLag, b.round, jump_table, lane);
        else ranf_jump_to_round(win[warp], win[warp], b.round - at, jump_table, lane);  // boundaries are in stream order
        at = b.round;
        uint32_t* o = states + i * 57;
        for (int k = lane; k < kRanfLag; k += 32) o[1 + k] = win[warp][k];
        if (lane == 0) {
            int cnt, idx = 55;
            fe_walk_round<F>(win[warp], 1, (int)b.state, cnt, (int)b.rank, &idx);
            o[0] = 0;
            o[56] = (uint32_t)idx;
        }
        __syncwarp();
    }
}

}  // namespace tp3

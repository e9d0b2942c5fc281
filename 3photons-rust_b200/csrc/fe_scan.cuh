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
//   1. fe_round_maps_kernel   for every round and every entry state: exit state + number of events started
//                             (one lane per round, so the 32 lanes of a warp run the same code on different data);
//   2. fe_segment_kernel      composes the maps of 1024 consecutive rounds (one thread per segment);
//   3. (host)                 chains the segment maps: entry state and event count at every segment start;
//   4. fe_boundaries_kernel   re-walks the per-round maps of a segment with its true entry state and records
//                             for every batch boundary (event index 10000 b) the round, entry state and rank of
//                             that event start inside the round;
//   5. fe_batch_states_kernel regenerates that round by jump-ahead, walks to the boundary and writes the
//                             generator state (numbers[0..55], index) the batch kernel starts from.
// The accept / re-roll test is evaluated exactly as the reference does (run's Float, no FMA contraction).
#pragma once

#include "faster_evgen.cuh"

namespace tp3 {

constexpr int kFeSegRounds = 1024;  // rounds per segment
constexpr int kFeTile = 32;         // rounds walked in parallel by one warp

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

// fe_walk_round on the precomputed masks: integer only, the same instruction stream for every entry state
__device__ __forceinline__ int fe_walk_masks(uint64_t A, uint64_t B, int s, int& count) {
    // next state after the 6-number request, indexed by (n0 | n1 << 1 | n2 << 2); after an accepted re-roll, indexed by s
    constexpr uint32_t kAfterSix = 0x0u | (2u << 4) | (6u << 8) | (4u << 12) | (8u << 16) | (3u << 20) | (7u << 24) | (5u << 28);
    constexpr uint64_t kAfterRoll = (0ull << 8) | (8ull << 12) | (6ull << 16) | (7ull << 20) | (0ull << 24) | (8ull << 28) | (0ull << 32);
    int idx = 55;
    count = 0;
    for (;;) {
        const int need = s == 0 ? 9 : s == 1 ? 6 : 2;
        if (idx < need) break;
        idx -= need;
        const int after_six = (int)((kAfterSix >> (4 * (int)((B >> idx) & 7u))) & 15u);
        const int after_roll = ((A >> idx) & 1u) ? s : (int)((kAfterRoll >> (4 * s)) & 15u);
        count += s == 0;
        s = s == 0 ? 1 : s == 1 ? after_six : after_roll;
    }
    return s;
}

// map word of a round: bits [4s, 4s+4) = exit state for entry state s, bits [36 + 3s, 36 + 3s + 3) = events started
__device__ __forceinline__ int fe_map_exit(uint64_t m, int s) { return (int)((m >> (4 * s)) & 15u); }
__device__ __forceinline__ int fe_map_count(uint64_t m, int s) { return (int)((m >> (36 + 3 * s)) & 7u); }

struct FeScanSmem {
    uint32_t win[2 * kRanfLag + 2];
    uint32_t tile[kFeTile][kRanfLag + 2];  // 32 consecutive rounds, slot order; odd row stride (57): lane = row reads are conflict free
};

// 1. per-round maps for rounds [first_round, first_round + n_rounds); one warp per segment of kFeSegRounds rounds
template <class F>
__global__ void __launch_bounds__(128) fe_round_maps_kernel(const uint32_t* __restrict__ jump_table, uint64_t first_round,
                                                          uint64_t n_rounds, uint64_t* __restrict__ maps) {
    __shared__ FeScanSmem sm[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FeScanSmem& w = sm[warp];
    const uint64_t seg = (uint64_t)blockIdx.x * 4 + warp;
    const uint64_t r0 = seg * kFeSegRounds;
    if (r0 >= n_rounds) return;
    const int n_seg = (int)min((uint64_t)kFeSegRounds, n_rounds - r0);
    // the seeded round 0 travels at the end of the jump table allocation (see api.cu)
    ranf_jump_to_round(w.win, jump_table + (size_t)kRanfDigits * 256 * kRanfLag, first_round + r0, jump_table, lane);
    for (int i = lane; i < kRanfLag; i += 32) w.tile[0][i] = w.win[i];
    __syncwarp();
    for (int t0 = 0; t0 < n_seg; t0 += kFeTile) {
        // rows 1..31 from row 0 (row 0 of the next tile from row 31 at the end)
        for (int r = 1; r < kFeTile; ++r) {
            for (int i = lane + 1; i <= kRanfLag; i += 32) w.tile[r][i - 1] = ranf_next_slot(w.tile[r - 1], i);
            __syncwarp();
        }
        if (t0 + lane < n_seg) {
            uint64_t A, B, m = 0;
            fe_round_masks<F>(w.tile[lane], A, B);
#pragma unroll 1
            for (int s = 0; s < 9; ++s) {
                int cnt;
                const int e = fe_walk_masks(A, B, s, cnt);
                m |= (uint64_t)e << (4 * s);
                m |= (uint64_t)cnt << (36 + 3 * s);
            }
            maps[r0 + t0 + lane] = m;
        }
        __syncwarp();
        if (t0 + kFeTile < n_seg) {
            uint32_t a = 0, b = 0;
            if (lane + 1 <= kRanfLag) a = ranf_next_slot(w.tile[kFeTile - 1], lane + 1);
            if (lane + 33 <= kRanfLag) b = ranf_next_slot(w.tile[kFeTile - 1], lane + 33);
            __syncwarp();
            w.tile[0][lane] = a;
            if (lane + 32 < kRanfLag) w.tile[0][lane + 32] = b;
            __syncwarp();
        }
    }
}

// 2. composed map of every segment: for each entry state, the exit state and the number of events started
__global__ void fe_segment_kernel(const uint64_t* __restrict__ maps, uint64_t n_rounds, uint8_t* __restrict__ seg_exit,
                                  uint32_t* __restrict__ seg_count) {
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = seg * kFeSegRounds;
    if (r0 >= n_rounds) return;
    const uint64_t r1 = min(n_rounds, r0 + kFeSegRounds);
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

// 4. batch boundaries inside every segment. seg_state / seg_events: entry state and global event index at the segment start.
__global__ void fe_boundaries_kernel(const uint64_t* __restrict__ maps, uint64_t first_round, uint64_t n_rounds,
                                     const uint8_t* __restrict__ seg_state, const uint64_t* __restrict__ seg_events,
                                     uint64_t first_batch, uint64_t n_batches, FeBoundary* __restrict__ out) {
    const uint64_t seg = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r0 = seg * kFeSegRounds;
    if (r0 >= n_rounds) return;
    const uint64_t r1 = min(n_rounds, r0 + kFeSegRounds);
    int cur = seg_state[seg];
    uint64_t ev = seg_events[seg];
    // next boundary at or after ev
    uint64_t b = (ev + kBatch - 1) / kBatch;
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t m = maps[r];
        const int c = fe_map_count(m, cur);
        const uint64_t target = b * kBatch;
        if (target < ev + c) {  // (target >= ev by construction)
            if (b >= first_batch && b < first_batch + n_batches) {
                FeBoundary o;
                o.round = first_round + r;
                o.state = (uint32_t)cur;
                o.rank = (uint32_t)(target - ev);
                out[b - first_batch] = o;
            }
            ++b;
        }
        ev += c;
        cur = fe_map_exit(m, cur);
    }
}

// 5. generator state at every batch boundary: numbers[0..55] + index, the layout faster_evgen_kernel reads
template <class F>
__global__ void __launch_bounds__(128) fe_batch_states_kernel(const uint32_t* __restrict__ jump_table, const FeBoundary* __restrict__ bnd,
                                                            uint64_t n_batches, uint32_t* __restrict__ states) {
    __shared__ uint32_t win[4][2 * kRanfLag + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t i = (uint64_t)blockIdx.x * 4 + warp;
    if (i >= n_batches) return;
    const FeBoundary b = bnd[i];
    ranf_jump_to_round(win[warp], jump_table + (size_t)kRanfDigits * 256 * kRanfLag, b.round, jump_table, lane);
    uint32_t* o = states + i * 57;
    for (int k = lane; k < kRanfLag; k += 32) o[1 + k] = win[warp][k];
    if (lane == 0) {
        int cnt, idx = 55;
        fe_walk_round<F>(win[warp], 1, (int)b.state, cnt, (int)b.rank, &idx);
        o[0] = 0;
        o[56] = (uint32_t)idx;
    }
}

}  // namespace tp3

// Device-side random streams, bit-exact with the reference's generators.
//
// RANF (src/random/ranf.rs): one warp owns a contiguous range of events and regenerates the
// reference's single global stream for exactly that range:
//   * start: the 55-word round containing the warp's first draw is reached by polynomial
//     jump-ahead over Z/1e9 (jump_tables.hpp), <= RANF_DIGITS cooperative 55x55 products;
//   * steady state: whole rounds are produced in parallel.  `reset` (ranf.rs:106-119) updates
//     slots 1..24 from the old round and slots 25..55 from freshly updated slots; substituting
//     gives every new slot as a +-combination of <= 4 OLD slots, so one round = 55 independent
//     lanes of work (2 passes of a 32-wide warp);
//   * layout: draws are stored in CONSUMPTION order (the reference hands out slots 55,54,..,1,
//     ranf.rs:95-101) in a 512-word shared-memory ring, so the 12 draws of one event are three
//     aligned 128-bit shared loads, bank-conflict free across the warp (stride 48 B).
//
// xoshiro256+/128+ (src/random/standard.rs): per-lane state in registers; each lane owns a
// contiguous run of events.  Batch start states come from a seeding kernel (GF(2) jump
// polynomials by byte digit of the batch index), lanes then apply one tabulated polynomial.
#pragma once

#include <cstdint>

namespace tp3 {

constexpr int kRanfLag = 55;
constexpr int kRanfDigits = 5;
constexpr uint32_t kRanfMod = 1000000000u;
constexpr int kRing = 512;       // words per warp ring
constexpr int kRingBias = 64;    // ring position of consumption coordinate 0 (keeps 16 B alignment)
constexpr int kDrawsPerEvent = 12;
constexpr int kWarpDraws = 32 * kDrawsPerEvent;  // draws consumed by one warp iteration

struct RanfWarpSmem {
    uint32_t ring[kRing];
    uint32_t win[2 * kRanfLag + 2];
};

__device__ __forceinline__ int ring_pos(int c) { return (c + kRingBias) & (kRing - 1); }

__device__ __forceinline__ uint32_t ranf_norm(int v) {
    // v in (-2e9, 2e9) -> [0, 1e9)
    v += (v < 0) ? (int)kRanfMod : 0;
    v += (v < 0) ? (int)kRanfMod : 0;
    v -= (v >= (int)kRanfMod) ? (int)kRanfMod : 0;
    return (uint32_t)v;
}

// Next round in slot order: y[0..54] = numbers[1..55] -> out (may not alias)
__device__ __forceinline__ uint32_t ranf_next_slot(const uint32_t* y, int i /*1..55*/) {
    int v;
    if (i <= 24) v = (int)y[i - 1] - (int)y[i + 31 - 1];
    else if (i <= 48) v = (int)y[i - 1] - (int)y[i - 24 - 1] + (int)y[i + 7 - 1];
    else v = (int)y[i - 1] - (int)y[i - 24 - 1] + (int)y[i - 48 - 1] - (int)y[i - 17 - 1];
    return ranf_norm(v);
}

struct RanfWarpStream {
    uint32_t* ring;
    int round_base;  // consumption coordinate of the first draw of the newest generated round
    int gen_end;     // one past the newest generated draw

    // base_y: round 0 of the generator (55 words, slot order) in shared or global memory;
    // d0: index of this warp's first draw in that generator's stream.
    __device__ void init(RanfWarpSmem* sm, const uint32_t* base_y, uint64_t d0,
                         const uint32_t* __restrict__ jump_table, int lane) {
        ring = sm->ring;
        uint32_t* win = sm->win;
        const uint64_t rho0 = d0 / kRanfLag;
        const int q0 = (int)(d0 - rho0 * kRanfLag);
        for (int i = lane; i < kRanfLag; i += 32) win[i] = base_y[i];
        __syncwarp();
        for (int k = 0; k < kRanfDigits; ++k) {
            const unsigned d = (unsigned)((rho0 >> (8 * k)) & 0xffu);
            if (d == 0) continue;  // warp-uniform
            // extend the window to y[0..109]
            for (int i = lane + 1; i <= kRanfLag; i += 32) win[kRanfLag + i - 1] = ranf_next_slot(win, i);
            __syncwarp();
            const uint32_t* __restrict__ c = jump_table + ((size_t)k * 256 + d) * kRanfLag;
            uint32_t o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                uint64_t acc = 0;
                if (i < kRanfLag) {
#pragma unroll 11
                    for (int j = 0; j < kRanfLag; ++j) {
                        acc += (uint64_t)__ldg(c + j) * win[i + j];
                        if ((j % 16) == 15) acc %= kRanfMod;  // 16 products < 1.6e19 < 2^64
                    }
                    acc %= kRanfMod;
                }
                o[h] = (uint32_t)acc;
            }
            __syncwarp();
            win[lane] = o[0];
            if (lane + 32 < kRanfLag) win[lane + 32] = o[1];
            __syncwarp();
        }
        // slot order -> consumption order: draw r of the round is slot 55 - r
        for (int r = lane; r < kRanfLag; r += 32) ring[ring_pos(-q0 + r)] = win[kRanfLag - 1 - r];
        round_base = -q0;
        gen_end = -q0 + kRanfLag;
        __syncwarp();
    }

    // One new round from the newest one, in consumption order (r = 55 - slot).
    __device__ __forceinline__ uint32_t next_draw(int r) const {
        const int b = round_base;
        const int k1 = (r >= 31) ? r - 31 : r + 24;
        const int k2 = (r >= 7) ? r - 7 : r + 48;
        int v = (int)ring[ring_pos(b + r)] - (int)ring[ring_pos(b + k1)];
        if (r < 31) v += (int)ring[ring_pos(b + k2)];
        if (r < 7) v -= (int)ring[ring_pos(b + r + 17)];
        return ranf_norm(v);
    }

    // Make draws [.., need_end) available.
    __device__ __forceinline__ void ensure(int need_end, int lane) {
        __syncwarp();
        while (gen_end < need_end) {
            const uint32_t v0 = next_draw(lane);
            const uint32_t v1 = (lane + 32 < kRanfLag) ? next_draw(lane + 32) : 0u;
            ring[ring_pos(round_base + kRanfLag + lane)] = v0;
            if (lane + 32 < kRanfLag) ring[ring_pos(round_base + kRanfLag + lane + 32)] = v1;
            round_base += kRanfLag;
            gen_end += kRanfLag;
            __syncwarp();
        }
    }

    // The 12 raw draws of event slot `lane` of warp iteration `it` (consumption coordinates).
    __device__ __forceinline__ void draws(int it, int lane, uint32_t out[kDrawsPerEvent]) const {
        const int c0 = it * kWarpDraws + lane * kDrawsPerEvent;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const uint4 v = *reinterpret_cast<const uint4*>(&ring[ring_pos(c0 + 4 * g)]);
            out[4 * g + 0] = v.x;
            out[4 * g + 1] = v.y;
            out[4 * g + 2] = v.z;
            out[4 * g + 3] = v.w;
        }
    }
};

// seeded_new (ranf.rs:36-66) by one warp: IN55 chain on lane 0, warm-up rounds in parallel.
__device__ inline void ranf_seed_warp(uint32_t* y /*smem 55*/, uint32_t* tmp /*smem 55*/, int32_t seed, int lane) {
    if (lane == 0) {
        y[kRanfLag - 1] = (uint32_t)seed;
        int j = seed, k = 1;
        for (int i = 1; i < kRanfLag; ++i) {
            const int ii = (21 * i) % kRanfLag;
            y[ii - 1] = (uint32_t)k;
            const int nk = j - k;
            j = k;
            k = nk < 0 ? nk + (int)kRanfMod : nk;
        }
    }
    __syncwarp();
    for (int r = 0; r < 10; ++r) {
        for (int i = lane + 1; i <= kRanfLag; i += 32) tmp[i - 1] = ranf_next_slot(y, i);
        __syncwarp();
        for (int i = lane; i < kRanfLag; i += 32) y[i] = tmp[i];
        __syncwarp();
    }
}

// --------------------------------------------------------------------------- xoshiro
struct Xoshiro256Lane {
    uint64_t s0, s1, s2, s3;
    __device__ __forceinline__ uint64_t next() {
        const uint64_t res = s0 + s3;
        const uint64_t t = s1 << 17;
        s2 ^= s0;
        s3 ^= s1;
        s1 ^= s2;
        s0 ^= s3;
        s2 ^= t;
        s3 = (s3 << 45) | (s3 >> 19);
        return res;
    }
    // s <- sum over set bits j of poly of (state advanced j steps); poly = 4 x u64
    __device__ void apply(const uint64_t* __restrict__ poly) {
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int w = 0; w < 4; ++w) {
            const uint64_t bits = poly[w];
            for (int b = 0; b < 64; ++b) {
                if ((bits >> b) & 1) {
                    a0 ^= s0;
                    a1 ^= s1;
                    a2 ^= s2;
                    a3 ^= s3;
                }
                next();
            }
        }
        s0 = a0;
        s1 = a1;
        s2 = a2;
        s3 = a3;
    }
};

struct Xoshiro128Lane {
    uint32_t s0, s1, s2, s3;
    __device__ __forceinline__ uint32_t next() {
        const uint32_t res = s0 + s3;
        const uint32_t t = s1 << 9;
        s2 ^= s0;
        s3 ^= s1;
        s1 ^= s2;
        s0 ^= s3;
        s2 ^= t;
        s3 = (s3 << 11) | (s3 >> 21);
        return res;
    }
    __device__ void apply(const uint64_t* __restrict__ poly /*2 x u64*/) {
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int w = 0; w < 2; ++w) {
            const uint64_t bits = poly[w];
            for (int b = 0; b < 64; ++b) {
                if ((bits >> b) & 1) {
                    a0 ^= s0;
                    a1 ^= s1;
                    a2 ^= s2;
                    a3 ^= s3;
                }
                next();
            }
        }
        s0 = a0;
        s1 = a1;
        s2 = a2;
        s3 = a3;
    }
};

}  // namespace tp3

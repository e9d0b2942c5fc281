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
//     ranf.rs:95-101) in a linear 8-round buffer per warp (440 draws >= 54 + the 384 one warp
//     iteration consumes); a refill keeps the unconsumed tail at the front and regenerates 7
//     rounds with compile-time shared-memory offsets (no ring arithmetic): ~17 warp instructions
//     per round of 55 numbers (ncu on the first version, which used a masked ring: 29 issue slots
//     per number, 20 % of the kernel).
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
constexpr int kDrawsPerEvent = 12;
constexpr int kWarpDraws = 32 * kDrawsPerEvent;   // draws consumed by one warp iteration (384)
constexpr int kBufRounds = 8;                     // rounds held per warp: 440 draws >= 54 + 384
constexpr int kBufWords = kBufRounds * kRanfLag;

struct alignas(16) RanfWarpSmem {
    // rounds rho .. rho+7 in CONSUMPTION order (draw r of a round = slot 55 - r), starting at word
    // `shift` (0..3) so that the first unconsumed draw is always 16-byte aligned
    uint32_t buf[kBufWords + 8];
    uint32_t win[2 * kRanfLag + 2];    // jump-ahead window (slot order), used by init only
};

// a - b for a, b in [0, 1e9): add 1e9 back if the difference is negative (unsigned min trick)
__device__ __forceinline__ uint32_t ranf_sub(uint32_t a, uint32_t b) {
    const uint32_t v = a - b;
    return min(v, v + kRanfMod);
}
// a + b for a, b in [0, 1e9): subtract 1e9 if the sum reaches it
__device__ __forceinline__ uint32_t ranf_add(uint32_t a, uint32_t b) {
    const uint32_t v = a + b;
    return min(v, v - kRanfMod);
}

// a mod 1e9 for any 64-bit a (nvcc does not strength-reduce 64-bit division by a constant): the quotient estimate
// floor(a * floor(2^64 / 1e9) / 2^64) is the true quotient or one less, so the remainder estimate is < 2e9.
__device__ __forceinline__ uint32_t ranf_mod64(uint64_t a) {
    const uint64_t q = __umul64hi(a, 18446744073ull);
    const uint32_t r = (uint32_t)(a - q * kRanfMod);
    return min(r, r - kRanfMod);
}

// Next round in slot order: y[0..54] = numbers[1..55] (ranf.rs:106-119 with the in-place updates
// substituted: every new slot is a +- combination of at most four OLD slots)
__device__ __forceinline__ uint32_t ranf_next_slot(const uint32_t* y, int i /*1..55*/) {
    if (i <= 24) return ranf_sub(y[i - 1], y[i + 31 - 1]);
    if (i <= 48) return ranf_add(ranf_sub(y[i - 1], y[i - 24 - 1]), y[i + 7 - 1]);
    return ranf_add(ranf_sub(y[i - 1], y[i - 24 - 1]), ranf_sub(y[i - 48 - 1], y[i - 17 - 1]));
}

// Warp-cooperative jump-ahead: win[0..54] <- round `rho0` of the generator whose round 0 is base_y (slot order),
// with <= kRanfDigits applications of the tabulated polynomials x^(55 d 256^k) mod P (jump_tables.hpp).
// win must hold 2 * 55 words.
__device__ inline void ranf_jump_to_round(uint32_t* win, const uint32_t* base_y, uint64_t rho0,
                                          const uint32_t* __restrict__ jump_table, int lane) {
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
#pragma unroll 5
                for (int j = 0; j < kRanfLag; ++j) {
                    acc += (uint64_t)__ldg(c + j) * win[i + j];
                    if ((j % 16) == 15) acc = ranf_mod64(acc);  // 16 products < 1.6e19 < 2^64
                }
                acc = ranf_mod64(acc);
            }
            o[h] = (uint32_t)acc;
        }
        __syncwarp();
        win[lane] = o[0];
        if (lane + 32 < kRanfLag) win[lane + 32] = o[1];
        __syncwarp();
    }
}

// One warp regenerates the reference's stream for a contiguous range of events.
// The buffer is linear (no ring arithmetic): all shared-memory offsets inside a refill are
// compile-time immediates relative to per-lane pointers, and every lane keeps its own two values
// of the newest round in registers (draws r = lane and r = lane + 32).
struct RanfWarpStream {
    uint32_t* buf;
    int shift;     // word offset of round 0 in buf (0..3)
    int p0;        // word offset in buf of the first draw of the current warp iteration; p0 % 4 == 0
    // This lane's three slots of the newest generated round: slot lane+1 (lanes < 24), slot lane+25
    // (lanes < 24) and slot lane+49 (lanes < 7). The reference's in-place update (ranf.rs:106-119) then
    // chains inside the lane: new[i] = old[i] - new[i-24] reads the value this lane has just produced.
    uint32_t va, vb, vc;

    // base_y: round 0 of the generator (55 words, slot order) in shared or global memory;
    // d0: index of this warp's first draw in that generator's stream.
    __device__ void init(RanfWarpSmem* sm, const uint32_t* base_y, uint64_t d0,
                         const uint32_t* __restrict__ jump_table, int lane) {
        buf = sm->buf;
        uint32_t* win = sm->win;
        const uint64_t rho0 = d0 / kRanfLag;
        const int q0 = (int)(d0 - rho0 * kRanfLag);
        ranf_jump_to_round(win, base_y, rho0, jump_table, lane);
        // slot order -> consumption order: draw r of the round is slot 55 - r
        shift = (-q0) & 3;
        p0 = shift + q0;
        const int l24 = lane < 24 ? lane : 23;  // lanes >= 24 shadow lane 23: same values, same addresses
        va = win[l24];
        vb = win[24 + l24];
        vc = win[lane < 7 ? 48 + lane : 48];
        store_round(buf + shift, lane);
        __syncwarp();
        refill(1, lane);
    }

    // st.shared predicated in one instruction (the compiler turns `if (p) *a = v` into a branch here)
    __device__ __forceinline__ static void store_if(bool p, uint32_t* addr, uint32_t val) {
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u32 [%0], %1; }" ::"r"((uint32_t)__cvta_generic_to_shared(addr)),
                     "r"(val), "r"((uint32_t)p)
                     : "memory");
    }

    // Write this lane's three slots of a round to its place in consumption order (r = 55 - slot).
    __device__ __forceinline__ void store_round(uint32_t* round0, int lane) const {
        uint32_t* pn = round0 - (lane < 24 ? lane : 23);
        pn[54] = va;                     // slot lane + 1
        pn[30] = vb;                     // slot lane + 25
        store_if(lane < 7, pn + 6, vc);  // slot lane + 49
    }

    // Round K of the buffer from round K-1 (ranf.rs:106-119), three dependent steps inside each lane:
    //   slots  1..24: new[i] = old[i] - old[i+31]      old[slot lane+32] lives in another lane's register
    //   slots 25..48: new[i] = old[i] - new[i-24]      = vb - (the va just computed)
    //   slots 49..55: new[i] = old[i] - new[i-24]      = vc - (the vb just computed)
    // Slot lane+32 is slot 25+L of lane L = lane+7 (L in 7..23) or slot 49+L of lane L = lane-17 (L in 0..6):
    // every lane publishes h = (L < 7 ? vc : vb) and reads it from lane (lane+7) mod 24 with ONE shuffle.
    // The recurrence therefore never reads shared memory: the stores below only feed the consumers, and no
    // warp fence is needed between rounds, so the compiler can interleave the chain with FP64 work.
    template <int K> __device__ __forceinline__ void gen_round(uint32_t* pn, int lane) {
        constexpr int B = K * kRanfLag;
        const int l24 = lane < 24 ? lane : 23;
        const uint32_t h = lane < 7 ? vc : vb;
        const uint32_t o = __shfl_sync(0xffffffffu, h, l24 < 17 ? l24 + 7 : l24 - 17);
        va = ranf_sub(va, o);
        vb = ranf_sub(vb, va);
        vc = ranf_sub(vc, vb);
        pn[B + 54] = va;
        pn[B + 30] = vb;
        store_if(lane < 7, pn + B + 6, vc);
    }

    __device__ __forceinline__ uint32_t* lane_base(int lane) const { return buf + shift - (lane < 24 ? lane : 23); }

    // Generate rounds first..7 of the buffer (warp-uniform first >= 1) from the registers va, vb, vc.
    __device__ __forceinline__ void refill(int first, int lane) {
        uint32_t* pn = lane_base(lane);
        if (first <= 1) gen_round<1>(pn, lane);
        if (first <= 2) gen_round<2>(pn, lane);
        if (first <= 3) gen_round<3>(pn, lane);
        if (first <= 4) gen_round<4>(pn, lane);
        if (first <= 5) gen_round<5>(pn, lane);
        if (first <= 6) gen_round<6>(pn, lane);
        gen_round<7>(pn, lane);
        __syncwarp();
    }

    // Pipelined form of advance(384): begin_next() right after the current iteration's draws have been
    // read, then tick<1..7>() spread over the event physics, so that the serial latency of a round
    // (shuffle -> 3 x (sub, min) -> stores) hides behind independent FP64 work of the same warp. The
    // consumer's __syncwarp() (draws()) publishes the stores.
    int pend_first;
    __device__ __forceinline__ void begin_next(int lane) {
        __syncwarp();
        const int rel = p0 - shift + kWarpDraws;        // in [384, 438]
        const int k = rel >= 7 * kRanfLag ? 7 : 6;
        const int new_rel = rel - k * kRanfLag;
        const int new_shift = (-new_rel) & 3;
        if (k == 7) {
            store_round(buf + new_shift, lane);
        } else {  // once every 55 iterations two rounds stay: move round 6 through registers
            const int l24 = lane < 24 ? lane : 23;
            const uint32_t* src = buf + shift + 6 * kRanfLag - l24;
            const uint32_t a = src[54], b = src[30], c = src[6];
            __syncwarp();
            uint32_t* dst = buf + new_shift - l24;
            dst[54] = a;
            dst[30] = b;
            store_if(lane < 7, dst + 6, c);
            store_round(buf + new_shift + kRanfLag, lane);
        }
        shift = new_shift;
        p0 = new_shift + new_rel;
        pend_first = kBufRounds - k;
        __syncwarp();
    }
    template <int K> __device__ __forceinline__ void tick(int lane) {
        if (K >= pend_first) gen_round<K>(lane_base(lane), lane);
    }

    // Step past `consumed` draws (384 after a full warp iteration, 12 * n after a partial one at the
    // end of a batch): the rounds that still hold unconsumed draws move to the front (re-aligned to
    // 16 bytes), the rest is regenerated.
    __device__ __forceinline__ void advance(int consumed, int lane) {
        __syncwarp();
        const int rel = p0 - shift + consumed;          // relative to round 0, < 440
        const int k = rel / kRanfLag;                   // round holding the next unconsumed draw
        const int new_rel = rel - k * kRanfLag;
        const int new_shift = (-new_rel) & 3;
        if (k == 7) {  // the usual case: only the newest round stays, and it is in registers
            store_round(buf + new_shift, lane);
        } else {
            // Batch boundary (consumed = 192 for full batches, so k is 3 or 4): forward copy in chunks of 32.
            // Needs k >= 1: the destination then starts below the source and a chunk never overwrites
            // words that are still to be read.
            const int n = (kBufRounds - k) * kRanfLag;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const uint32_t t = (i < n) ? buf[shift + k * kRanfLag + i] : 0u;
                __syncwarp();
                if (i < n) buf[new_shift + i] = t;
                __syncwarp();
            }
        }
        shift = new_shift;
        p0 = new_shift + new_rel;
        __syncwarp();
        refill(kBufRounds - k, lane);
    }

    // The 12 raw draws of event slot `lane` of the current warp iteration: three aligned 128-bit
    // loads, bank-conflict free across the warp (lane stride 48 bytes).
    __device__ __forceinline__ void draws(int lane, uint32_t out[kDrawsPerEvent]) const {
        __syncwarp();  // the rounds stored by other lanes since the last fence become visible
        const uint4* p = reinterpret_cast<const uint4*>(buf + p0 + lane * kDrawsPerEvent);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const uint4 q = p[g];
            out[4 * g + 0] = q.x;
            out[4 * g + 1] = q.y;
            out[4 * g + 2] = q.z;
            out[4 * g + 3] = q.w;
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
// One step of xoshiro (rand_xoshiro 0.6.0): s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3).  Written with every
// new word as a three-input XOR of OLD words (one LOP3 each instead of a chain of two-input ones: 4 instead of 5 logic
// instructions per 32-bit word and step, and no dependency between them).
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint64_t xor3(uint64_t a, uint64_t b, uint64_t c) {
    return (uint64_t)xor3((uint32_t)(a >> 32), (uint32_t)(b >> 32), (uint32_t)(c >> 32)) << 32 | xor3((uint32_t)a, (uint32_t)b, (uint32_t)c);
}

struct Xoshiro256Lane {
    uint64_t s0, s1, s2, s3;
    __device__ __forceinline__ uint64_t next() {
        const uint64_t res = s0 + s3;
        const uint64_t t = s1 << 17;
        const uint64_t n1 = xor3(s1, s2, s0), n0 = xor3(s0, s3, s1), n2 = xor3(s2, s0, t), x3 = s3 ^ s1;
        s0 = n0;
        s1 = n1;
        s2 = n2;
        s3 = (x3 << 45) | (x3 >> 19);
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
        const uint32_t n1 = xor3(s1, s2, s0), n0 = xor3(s0, s3, s1), n2 = xor3(s2, s0, t), x3 = s3 ^ s1;
        s0 = n0;
        s1 = n1;
        s2 = n2;
        s3 = __funnelshift_l(x3, x3, 11);
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

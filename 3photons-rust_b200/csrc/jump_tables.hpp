// Host-side jump-ahead algebra for the two random generators of the reference.
//
// RANF (src/random/ranf.rs): `reset` (ranf.rs:106-119) is the linear recurrence
//     y[m] = (y[m-55] - y[m-24]) mod 1e9,      y[55k+i] = numbers_k[i], i = 1..55,
// with characteristic polynomial P(x) = x^55 + x^31 - 1 over Z/1e9 (monic, so polynomial
// reduction works in the ring).  Jumping R rounds ahead is y[m+55R] = sum_j c_j y[m+j]
// with sum_j c_j x^j = x^(55R) mod P.  We tabulate x^(55 * d * 256^k) for every byte digit
// so that any round index is reached with at most RANF_DIGITS polynomial applications.
//
// xoshiro256+/xoshiro128+ (src/random/standard.rs over rand_xoshiro 0.6.0): the state
// transition is linear over GF(2); the minimal polynomial is recovered with
// Berlekamp-Massey, and x^n mod P gives the jump polynomial for any distance n (the
// published jump() constants are the special case n = 2^128 / 2^64, which we check).
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace tp3 {

// ----------------------------------------------------------------------------- RANF
constexpr uint32_t RANF_MOD = 1000000000u;
constexpr int RANF_LAG = 55;
constexpr int RANF_DIGITS = 5;  // byte digits of the round index: 2^40 rounds = 5e12 events
constexpr int32_t RANF_DEFAULT_SEED = 234612947;  // ranf.rs:31
constexpr uint32_t RANF_JUMP_SEED_STEP = 123456;  // ranf.rs:138

using RanfPoly = std::array<uint32_t, RANF_LAG>;

inline RanfPoly ranf_poly_mul(const RanfPoly& a, const RanfPoly& b) {
    uint64_t c[2 * RANF_LAG - 1] = {0};
    for (int i = 0; i < RANF_LAG; ++i) {
        if (!a[i]) continue;
        for (int j = 0; j < RANF_LAG; ++j) c[i + j] = (c[i + j] + (uint64_t)a[i] * b[j]) % RANF_MOD;
    }
    // x^55 = 1 - x^31  =>  x^k = x^(k-55) - x^(k-24)
    for (int k = 2 * RANF_LAG - 2; k >= RANF_LAG; --k) {
        uint64_t t = c[k];
        if (!t) continue;
        c[k] = 0;
        c[k - 55] = (c[k - 55] + t) % RANF_MOD;
        c[k - 24] = (c[k - 24] + RANF_MOD - t) % RANF_MOD;
    }
    RanfPoly r;
    for (int i = 0; i < RANF_LAG; ++i) r[i] = (uint32_t)c[i];
    return r;
}

// table[(k*256 + d)*55 + j] = coefficient j of x^(55 * d * 256^k) mod P   (d = 0 -> identity)
inline std::vector<uint32_t> ranf_round_jump_table() {
    std::vector<uint32_t> t((size_t)RANF_DIGITS * 256 * RANF_LAG, 0u);
    RanfPoly unit{};  // x^55 = 1 - x^31
    unit[0] = 1;
    unit[31] = RANF_MOD - 1;
    RanfPoly one{};
    one[0] = 1;
    for (int k = 0; k < RANF_DIGITS; ++k) {
        RanfPoly cur = one;
        for (int d = 0; d < 256; ++d) {
            std::memcpy(&t[((size_t)k * 256 + d) * RANF_LAG], cur.data(), sizeof(uint32_t) * RANF_LAG);
            cur = ranf_poly_mul(cur, unit);
        }
        unit = cur;  // x^(55 * 256^(k+1))
    }
    return t;
}

// One `reset` (ranf.rs:106-119) on y[0..54] = numbers[1..55], values in [0, 1e9)
inline void ranf_reset(uint32_t* y) {
    for (int i = 1; i < 25; ++i) {
        int32_t v = (int32_t)y[i - 1] - (int32_t)y[i + 31 - 1];
        y[i - 1] = (uint32_t)(v < 0 ? v + (int32_t)RANF_MOD : v);
    }
    for (int i = 25; i < 56; ++i) {
        int32_t v = (int32_t)y[i - 1] - (int32_t)y[i - 24 - 1];
        y[i - 1] = (uint32_t)(v < 0 ? v + (int32_t)RANF_MOD : v);
    }
}

// seeded_new (ranf.rs:36-66): IN55 initialisation + 10 warm-up rounds. Valid for 0 < seed < 1e9.
inline void ranf_seed_state(int32_t seed, uint32_t* y /*55*/) {
    int32_t n[56] = {0};
    n[55] = seed;
    int32_t j = seed, k = 1;
    for (int i = 1; i < 55; ++i) {
        int ii = (21 * i) % 55;
        n[ii] = k;
        k = j - k;
        if (k < 0) k += (int32_t)RANF_MOD;
        j = n[ii];
    }
    for (int i = 1; i <= 55; ++i) y[i - 1] = (uint32_t)n[i];
    for (int r = 0; r < 10; ++r) ranf_reset(y);
}

// Host mirror of the device jump: round `round` from the seeded round 0.
inline void ranf_round_at(const std::vector<uint32_t>& table, int32_t seed, uint64_t round, uint32_t* out) {
    uint32_t w[2 * RANF_LAG];
    ranf_seed_state(seed, w);
    for (int k = 0; k < RANF_DIGITS; ++k) {
        unsigned d = (unsigned)((round >> (8 * k)) & 0xff);
        if (!d) continue;
        std::memcpy(w + RANF_LAG, w, sizeof(uint32_t) * RANF_LAG);
        ranf_reset(w + RANF_LAG);  // window y[0..109]
        const uint32_t* c = &table[((size_t)k * 256 + d) * RANF_LAG];
        uint32_t nw[RANF_LAG];
        for (int i = 0; i < RANF_LAG; ++i) {
            unsigned __int128 acc = 0;
            for (int jx = 0; jx < RANF_LAG; ++jx) acc += (uint64_t)c[jx] * w[i + jx];
            nw[i] = (uint32_t)(acc % RANF_MOD);
        }
        std::memcpy(w, nw, sizeof nw);
    }
    std::memcpy(out, w, sizeof(uint32_t) * RANF_LAG);
}

// -------------------------------------------------------------------------- xoshiro
// Polynomials over GF(2) of degree < 256 as 4 x u64 (bit j of word j/64 = coefficient of x^j).
struct Gf2Poly {
    uint64_t w[4] = {0, 0, 0, 0};
};

struct Xoshiro256 {
    uint64_t s[4];
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    void step() {
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
    }
};
struct Xoshiro128 {
    uint32_t s[4];
    static uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    void step() {
        uint32_t t = s[1] << 9;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 11);
    }
};

inline uint64_t splitmix64_next(uint64_t& x) {
    x += 0x9e3779b97f4a7c15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// rand_core SeedableRng::seed_from_u64 as used by standard.rs:23
inline Xoshiro256 xoshiro256_seed(uint64_t seed) {
    Xoshiro256 g;
    for (auto& w : g.s) w = splitmix64_next(seed);
    return g;
}
inline Xoshiro128 xoshiro128_seed(uint64_t seed) {
    Xoshiro128 g;
    uint64_t a = splitmix64_next(seed), b = splitmix64_next(seed);
    g.s[0] = (uint32_t)a;
    g.s[1] = (uint32_t)(a >> 32);
    g.s[2] = (uint32_t)b;
    g.s[3] = (uint32_t)(b >> 32);
    return g;
}

// GF(2) modulus: monic polynomial of degree DEG (256 or 128), low DEG coefficients in `low`.
struct Gf2Mod {
    int deg = 0;
    Gf2Poly low;
};

// Berlekamp-Massey on bit 0 of s[0]; returns the minimal polynomial (degree must equal the state size).
template <class Gen> inline Gf2Mod xoshiro_min_poly(Gen g, int deg) {
    const int n = 2 * deg + 64;
    std::vector<uint8_t> bits(n);
    for (int i = 0; i < n; ++i) {
        bits[i] = (uint8_t)(g.s[0] & 1);
        g.step();
    }
    std::vector<uint8_t> C(n + 1, 0), B(n + 1, 0), T;
    C[0] = B[0] = 1;
    int L = 0, m = 1;
    for (int i = 0; i < n; ++i) {
        uint8_t d = bits[i];
        for (int j = 1; j <= L; ++j) d ^= (uint8_t)(C[j] & bits[i - j]);
        if (d) {
            T = C;
            for (int j = 0; j + m <= n; ++j) C[j + m] ^= B[j];
            if (2 * L <= i) {
                L = i + 1 - L;
                B = T;
                m = 1;
            } else {
                ++m;
            }
        } else {
            ++m;
        }
    }
    // b[t] = sum_{j=1..L} C[j] b[t-j]  =>  P(x) = x^L + sum_j C[j] x^(L-j)
    Gf2Mod mod;
    mod.deg = L;
    for (int j = 1; j <= L; ++j)
        if (C[j]) {
            int e = L - j;
            mod.low.w[e / 64] |= 1ull << (e % 64);
        }
    return mod;
}

inline Gf2Poly gf2_mul_mod(const Gf2Poly& a, const Gf2Poly& b, const Gf2Mod& mod) {
    // shift-and-add from the top bit of b, reducing after each shift
    Gf2Poly r;
    for (int bit = mod.deg - 1; bit >= 0; --bit) {
        // r = r * x mod P
        bool carry = (r.w[(mod.deg - 1) / 64] >> ((mod.deg - 1) % 64)) & 1;
        for (int i = 3; i > 0; --i) r.w[i] = (r.w[i] << 1) | (r.w[i - 1] >> 63);
        r.w[0] <<= 1;
        if (mod.deg < 256) {  // clear bits >= deg
            int wi = mod.deg / 64;
            for (int i = wi; i < 4; ++i) r.w[i] = 0;
        }
        if (carry)
            for (int i = 0; i < 4; ++i) r.w[i] ^= mod.low.w[i];
        if ((b.w[bit / 64] >> (bit % 64)) & 1)
            for (int i = 0; i < 4; ++i) r.w[i] ^= a.w[i];
    }
    return r;
}

// x^(n * 2^shift) mod P
inline Gf2Poly gf2_x_pow(uint64_t n, int shift, const Gf2Mod& mod) {
    Gf2Poly result;
    result.w[0] = 1;
    Gf2Poly base;
    base.w[0] = 2;  // x
    for (int i = 0; i < shift; ++i) base = gf2_mul_mod(base, base, mod);
    while (n) {
        if (n & 1) result = gf2_mul_mod(result, base, mod);
        base = gf2_mul_mod(base, base, mod);
        n >>= 1;
    }
    return result;
}

inline Gf2Poly gf2_pow(Gf2Poly base, uint64_t n, const Gf2Mod& mod) {
    Gf2Poly result;
    result.w[0] = 1;
    while (n) {
        if (n & 1) result = gf2_mul_mod(result, base, mod);
        base = gf2_mul_mod(base, base, mod);
        n >>= 1;
    }
    return result;
}

// Apply a jump polynomial to a generator state (host mirror of the device routine).
template <class Gen> inline void xoshiro_apply(Gen& g, const Gf2Poly& p, int deg) {
    decltype(g.s[0] + 0) acc[4] = {0, 0, 0, 0};
    Gen n = g;
    for (int j = 0; j < deg; ++j) {
        if ((p.w[j / 64] >> (j % 64)) & 1)
            for (int i = 0; i < 4; ++i) acc[i] ^= n.s[i];
        n.step();
    }
    for (int i = 0; i < 4; ++i) g.s[i] = (decltype(g.s[0] + 0))acc[i];
}

}  // namespace tp3

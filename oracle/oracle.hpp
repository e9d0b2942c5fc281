// oracle/oracle.hpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A literal, scalar C++ restatement of the reference's per-event hot path and
// of the host code that turns its sums into `res.data` / stdout.  It exists to
// CHECK the CUDA product (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs).  Nothing under 3photons-rust_b200/
// may include, link or call it.
//
// Parity status: PINNED.  The Rust reference cannot be built here (no
// rustc/cargo, crates not vendored), so the oracle is pinned against the
// reference's own golden outputs instead: tests/test_oracle_golden.py checks
// the `res.data` and stdout produced by this code against all 7 distinct
// golden file pairs of /root/reference/reference/ (copied to tests/golden/)
// with the tolerances of the reference's CI (.github/workflows/ci.yml:179-203).
//
// Arithmetic rules (why this file must be compiled with -ffp-contract=off and
// without -ffast-math): Rust never contracts a*b+c into an FMA and never
// reassociates, so every expression below is written in the reference's own
// evaluation order and relies on plain IEEE operations.
//
// Third-party semantics restated here (crates pinned in Cargo.lock, sources
// not under /root/reference): nalgebra 0.32.3 (column-major from_fn, small
// fixed-size dot products), num-complex 0.4.4 (mul, div, powi), rand 0.8.5 /
// rand_core 0.6.4 / rand_xoshiro 0.6.0 (SplitMix64 seeding, xoshiro256+ /
// xoshiro128+, float conversion, jump()).  See SURVEY.md Appendix B.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------------------
// Feature flags (Cargo.toml:9-21)
// ---------------------------------------------------------------------------
struct Features {
    bool f32 = false;                // numeric.rs:6-13
    bool faster_evgen = false;       // evgen.rs:143
    bool faster_threading = false;   // scheduling/multi_threading.rs:30-39,66-69
    bool multi_threading = false;    // scheduling/mod.rs:4-7
    bool no_photon_sorting = false;  // evgen.rs:109, event.rs:96
    bool standard_random = false;    // random/mod.rs:5-16
};

constexpr int EVENT_BATCH_SIZE = 10000;  // scheduling/mod.rs:21

// ---------------------------------------------------------------------------
// Numeric prelude (numeric.rs:6-37)
// ---------------------------------------------------------------------------
template <class F> struct K;
template <> struct K<double> {
    static constexpr double PI = 3.14159265358979323846264338327950288;
    static constexpr double FRAC_PI_2 = 1.57079632679489661923132169163975144;
    static constexpr double SQRT_2 = 1.41421356237309504880168872420969808;
    static constexpr double MIN_POSITIVE = 2.2250738585072014e-308;
    static constexpr int DIGITS = 15;
};
template <> struct K<float> {
    static constexpr float PI = 3.14159265358979323846264338327950288f;
    static constexpr float FRAC_PI_2 = 1.57079632679489661923132169163975144f;
    static constexpr float SQRT_2 = 1.41421356237309504880168872420969808f;
    static constexpr float MIN_POSITIVE = 1.17549435e-38f;
    static constexpr int DIGITS = 6;
};

// num-complex 0.4.4 arithmetic, restated
template <class F> struct Cx {
    F re, im;
};
template <class F> inline Cx<F> operator+(Cx<F> a, Cx<F> b) { return {a.re + b.re, a.im + b.im}; }
template <class F> inline Cx<F> operator-(Cx<F> a, Cx<F> b) { return {a.re - b.re, a.im - b.im}; }
template <class F> inline Cx<F> operator-(Cx<F> a) { return {-a.re, -a.im}; }
template <class F> inline Cx<F> operator*(Cx<F> a, Cx<F> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class F> inline Cx<F> operator*(Cx<F> a, F b) { return {a.re * b, a.im * b}; }
template <class F> inline Cx<F> operator*(F a, Cx<F> b) { return {a * b.re, a * b.im}; }
template <class F> inline Cx<F> operator/(Cx<F> a, F b) { return {a.re / b, a.im / b}; }
template <class F> inline Cx<F> operator/(Cx<F> a, Cx<F> b) {
    F n = b.re * b.re + b.im * b.im;
    F re = a.re * b.re + a.im * b.im;
    F im = a.im * b.re - a.re * b.im;
    return {re / n, im / n};
}
template <class F> inline Cx<F> conj(Cx<F> a) { return {a.re, -a.im}; }
template <class F> inline F norm_sqr(Cx<F> a) { return a.re * a.re + a.im * a.im; }
template <class F> inline Cx<F> sqr(Cx<F> a) { return a * a; }  // powi(2) = square-and-multiply

// ---------------------------------------------------------------------------
// Configuration (config.rs:8-128)
// ---------------------------------------------------------------------------
template <class F> struct Config {
    uint64_t num_events;
    F e_total, beam_photons_cut, photon_photon_cut, e_min, beam_photon_plane_cut;
    F alpha, alpha_z, gev2_to_picobarn, m_z0, g_z0, sin2_weinberg, branching_ep_em, beta_plus,
        beta_minus;
    int32_t num_bins;
    bool impr, plot;
};

inline float parse_float(const char* s, float*) { return strtof(s, nullptr); }
inline double parse_float(const char* s, double*) { return strtod(s, nullptr); }

// config.rs:57-128. Returns an empty string on success, else the error text.
template <class F> std::string load_config(const std::string& text, Config<F>& cfg) {
    std::vector<std::string> items;
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        size_t b = line.find_first_not_of(" \t\r\f\v");
        if (b != std::string::npos) {
            size_t e = line.find_first_of(" \t\r\f\v", b);
            items.push_back(line.substr(b, e == std::string::npos ? std::string::npos : e - b));
        }
        pos = eol + 1;
    }
    static const char* names[18] = {"num_events", "e_total", "beam_photons_cut", "photon_photon_cut",
                                    "e_min", "beam_photon_plane_cut", "alpha", "alpha_z",
                                    "gev2_to_picobarn", "m_z0", "g_z0", "sin2_weinberg",
                                    "branching_ep_em", "beta_plus", "beta_moins", "num_bins", "impr",
                                    "plot"};
    if (items.size() < 18) return std::string("missing configuration of ") + names[items.size()];
    auto fl = [&](int i) { return parse_float(items[i].c_str(), (F*)nullptr); };
    auto bo = [&](int i, bool& ok) {
        std::string s = items[i];
        for (auto& c : s) c = (char)tolower(c);
        ok = true;
        if (s == ".true." || s == "true") return true;
        if (s == ".false." || s == "false") return false;
        ok = false;
        return false;
    };
    cfg.num_events = strtoull(items[0].c_str(), nullptr, 10);
    cfg.e_total = fl(1);
    cfg.beam_photons_cut = fl(2);
    cfg.photon_photon_cut = fl(3);
    cfg.e_min = fl(4);
    cfg.beam_photon_plane_cut = fl(5);
    cfg.alpha = fl(6);
    cfg.alpha_z = fl(7);
    cfg.gev2_to_picobarn = fl(8);
    cfg.m_z0 = fl(9);
    cfg.g_z0 = fl(10);
    cfg.sin2_weinberg = fl(11);
    cfg.branching_ep_em = fl(12);
    cfg.beta_plus = fl(13);
    cfg.beta_minus = fl(14);
    cfg.num_bins = (int32_t)strtol(items[15].c_str(), nullptr, 10);
    bool ok1, ok2;
    cfg.impr = bo(16, ok1);
    cfg.plot = bo(17, ok2);
    if (!ok1) return "could not parse configuration of impr";
    if (!ok2) return "could not parse configuration of plot";
    if (cfg.num_events == 0) return "Please simulate at least one event";
    if (cfg.plot) return "Plotting is not supported by this version";
    if (cfg.impr) return "Individual result printing is not supported.";
    return "";
}

// ---------------------------------------------------------------------------
// Random number generators
// ---------------------------------------------------------------------------
// random/ranf.rs (whole file). i32 arithmetic wraps like release-mode Rust.
template <class F> struct Ranf {
    static constexpr int32_t MODULO = 1000000000;
    int32_t seed;
    int32_t numbers[56];
    int index;

    Ranf() { seeded_new(234612947); }  // ranf.rs:28-32
    // ranf.rs:36-66
    void seeded_new(int32_t s) {
        seed = s;
        std::memset(numbers, 0, sizeof numbers);
        index = 55;
        numbers[55] = s;
        int32_t j = s, k = 1;
        for (int i = 1; i < 55; ++i) {
            int ii = (21 * i) % 55;
            numbers[ii] = k;
            k = (int32_t)((uint32_t)j - (uint32_t)k);
            if (k < 0) k = (int32_t)((uint32_t)k + (uint32_t)MODULO);
            j = numbers[ii];
        }
        for (int r = 0; r < 10; ++r) reset();
    }
    // ranf.rs:106-119
    void reset() {
        for (int i = 1; i < 25; ++i) {
            numbers[i] = (int32_t)((uint32_t)numbers[i] - (uint32_t)numbers[i + 31]);
            if (numbers[i] < 0) numbers[i] = (int32_t)((uint32_t)numbers[i] + (uint32_t)MODULO);
        }
        for (int i = 25; i < 56; ++i) {
            numbers[i] = (int32_t)((uint32_t)numbers[i] - (uint32_t)numbers[i - 24]);
            if (numbers[i] < 0) numbers[i] = (int32_t)((uint32_t)numbers[i] + (uint32_t)MODULO);
        }
    }
    // ranf.rs:78-102. Also returns the raw integers (for the bit-exactness tests).
    template <int N> void random_array(F* out, int32_t* raw = nullptr) {
        if (index < N) {
            reset();
            index = 55;
        }
        index -= N;
        for (int i = 0; i < N; ++i) {
            int32_t n = numbers[index + 1 + i];
            if (raw) raw[i] = n;
            out[i] = (F)n * (F)1e-9;
        }
    }
    F random() {  // ranf.rs:72-74
        F r;
        random_array<1>(&r);
        return r;
    }
    uint64_t next_raw() {  // integer view of random(), for stream dumps
        F r;
        int32_t raw;
        random_array<1>(&r, &raw);
        return (uint64_t)(uint32_t)raw;
    }
    void jump() { seeded_new((int32_t)((uint32_t)seed + 123456u)); }  // ranf.rs:136-140
};

inline uint64_t splitmix64(uint64_t& x) {  // rand_core 0.6.4 SeedableRng::seed_from_u64 / rand_xoshiro SplitMix64
    x += 0x9e3779b97f4a7c15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// random/standard.rs over rand_xoshiro 0.6.0
template <class F> struct Xoshiro;
template <> struct Xoshiro<double> {  // Xoshiro256Plus
    uint64_t s[4];
    Xoshiro() {
        uint64_t x = 12345;  // standard.rs:23
        for (auto& w : s) w = splitmix64(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_raw() {
        uint64_t res = s[0] + s[3];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return res;
    }
    double random() { return (double)(next_raw() >> 11) * (1.0 / 9007199254740992.0); }
    template <int N> void random_array(double* out, int32_t* = nullptr) {
        for (int i = 0; i < N; ++i) out[i] = random();
    }
    void jump() {
        static const uint64_t J[4] = {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull,
                                      0xa9582618e03fc9aaull, 0x39abdc4529b1661cull};
        uint64_t n[4] = {0, 0, 0, 0};
        for (uint64_t jw : J)
            for (int b = 0; b < 64; ++b) {
                if (jw & (1ull << b))
                    for (int i = 0; i < 4; ++i) n[i] ^= s[i];
                next_raw();
            }
        for (int i = 0; i < 4; ++i) s[i] = n[i];
    }
};
template <> struct Xoshiro<float> {  // Xoshiro128Plus
    uint32_t s[4];
    Xoshiro() {
        uint64_t x = 12345;
        uint64_t a = splitmix64(x), b = splitmix64(x);
        s[0] = (uint32_t)a;
        s[1] = (uint32_t)(a >> 32);
        s[2] = (uint32_t)b;
        s[3] = (uint32_t)(b >> 32);
    }
    static uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
    uint64_t next_raw() {
        uint32_t res = s[0] + s[3];
        uint32_t t = s[1] << 9;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 11);
        return res;
    }
    float random() { return (float)((uint32_t)next_raw() >> 8) * (1.0f / 16777216.0f); }
    template <int N> void random_array(float* out, int32_t* = nullptr) {
        for (int i = 0; i < N; ++i) out[i] = random();
    }
    void jump() {  // published xoshiro128+ jump polynomial (2^64 steps); not exercised by any golden
        static const uint32_t J[4] = {0x8764000bu, 0xf542d2d3u, 0x6fa035c3u, 0x77f2db5bu};
        uint32_t n[4] = {0, 0, 0, 0};
        for (uint32_t jw : J)
            for (int b = 0; b < 32; ++b) {
                if (jw & (1u << b))
                    for (int i = 0; i < 4; ++i) n[i] ^= s[i];
                next_raw();
            }
        for (int i = 0; i < 4; ++i) s[i] = n[i];
    }
};

// ---------------------------------------------------------------------------
// Event generation (evgen.rs) — event layout event.rs:51: 5 particles x (X,Y,Z,E)
// ---------------------------------------------------------------------------
template <class F> struct Event {
    F p[5][4];
};

// evgen.rs:40-77: event weight
template <class F> F event_weight(F e_total) {
    F z_n = (F)2 * std::log(K<F>::FRAC_PI_2);
    for (int k = 2; k < 3; ++k) z_n -= (F)2 * std::log((F)(k - 1));
    z_n = z_n - std::log((F)2);
    F ln_weight = ((F)2 * (F)3 - (F)4) * std::log(e_total) + z_n;
    return std::exp(ln_weight);
}

// evgen.rs:221-249
template <class F, class Rng> void random_unit_2d_outgoing(Rng& rng, F pts[3][2]) {
    F v[6];
    rng.template random_array<6>(v);
    for (int c = 0; c < 2; ++c)  // from_iterator fills column-major: 3 rows x 2 columns
        for (int p = 0; p < 3; ++p) pts[p][c] = (F)2 * v[c * 3 + p] - (F)1;
    F r2[3];
    for (int p = 0; p < 3; ++p) r2[p] = pts[p][0] * pts[p][0] + pts[p][1] * pts[p][1];
    const F MIN_POSITIVE_2 = K<F>::MIN_POSITIVE * K<F>::MIN_POSITIVE;
    for (int p = 0; p < 3; ++p) {
        while (r2[p] > (F)1 || r2[p] < MIN_POSITIVE_2) {
            F w[2];
            rng.template random_array<2>(w);
            pts[p][0] = (F)2 * w[0] - (F)1;
            pts[p][1] = (F)2 * w[1] - (F)1;
            r2[p] = pts[p][0] * pts[p][0] + pts[p][1] * pts[p][1];
        }
    }
    for (int p = 0; p < 3; ++p) {
        F n = (F)1 / std::sqrt(r2[p]);
        pts[p][0] *= n;
        pts[p][1] *= n;
    }
}

// evgen.rs:139-208: q[coord][particle]
template <class F, class Rng> void generate_raw(Rng& rng, const Features& ft, F q[4][3]) {
    F cos_theta[3], exp_min_e[3], sx[3], sy[3];
    if (ft.faster_evgen) {
        F u[9];
        rng.template random_array<9>(u);
        for (int p = 0; p < 3; ++p) {
            cos_theta[p] = (F)2 * u[p] - (F)1;
            exp_min_e[p] = u[3 + p] * u[6 + p];
        }
        F pts[3][2];
        random_unit_2d_outgoing<F>(rng, pts);
        for (int p = 0; p < 3; ++p) {
            sx[p] = pts[p][0];
            sy[p] = pts[p][1];
        }
    } else {
        F phi[3];
        for (int p = 0; p < 3; ++p) {  // column-major from_fn: per photon cos_theta, phi, r*r'
            cos_theta[p] = (F)2 * rng.random() - (F)1;
            phi[p] = (F)2 * K<F>::PI * rng.random();
            F a = rng.random();
            F b = rng.random();
            exp_min_e[p] = a * b;
        }
        for (int p = 0; p < 3; ++p) {
            sy[p] = std::cos(phi[p]);  // Y <- cos, X <- sin (evgen.rs:200-201)
            sx[p] = std::sin(phi[p]);
        }
    }
    for (int p = 0; p < 3; ++p) {
        F sin_theta = std::sqrt((F)1 - cos_theta[p] * cos_theta[p]);
        F energy = -std::log(exp_min_e[p] + K<F>::MIN_POSITIVE);
        q[0][p] = energy * (sin_theta * sx[p]);
        q[1][p] = energy * (sin_theta * sy[p]);
        q[2][p] = energy * cos_theta[p];
        q[3][p] = energy * (F)1;
    }
}

// evgen.rs:89-132
template <class F, class Rng> Event<F> generate(Rng& rng, const Features& ft, F e_total) {
    F q[4][3];
    generate_raw<F>(rng, ft, q);
    F r[4];
    for (int c = 0; c < 4; ++c) r[c] = (q[c][0] + q[c][1]) + q[c][2];  // column_sum
    F r_norm_2 = r[3] * r[3] - ((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
    F alpha = e_total / r_norm_2;
    F r_norm = std::sqrt(r_norm_2);
    F beta = (F)1 / (r_norm + r[3]);
    F p_e[3], p_xyz[3][3];
    for (int i = 0; i < 3; ++i) {
        F rq = (q[0][i] * r[0] + q[1][i] * r[1]) + q[2][i] * r[2];  // gemv, column by column
        p_e[i] = alpha * (r[3] * q[3][i] - rq);
        F b_rq_e = beta * rq - q[3][i];
        for (int c = 0; c < 3; ++c) p_xyz[i][c] = alpha * (r_norm * q[c][i] + b_rq_e * r[c]);
    }
    if (!ft.no_photon_sorting) {
        for (int a = 0; a < 2; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (p_e[b] > p_e[a]) {
                    std::swap(p_e[a], p_e[b]);
                    for (int c = 0; c < 3; ++c) std::swap(p_xyz[a][c], p_xyz[b][c]);
                }
    }
    Event<F> ev;
    F half = e_total / (F)2;
    ev.p[0][0] = -half; ev.p[0][1] = 0; ev.p[0][2] = 0; ev.p[0][3] = half;
    ev.p[1][0] = half;  ev.p[1][1] = 0; ev.p[1][2] = 0; ev.p[1][3] = half;
    for (int i = 0; i < 3; ++i) {
        for (int c = 0; c < 3; ++c) ev.p[2 + i][c] = p_xyz[i][c];
        ev.p[2 + i][3] = p_e[i];
    }
    return ev;
}

// ---------------------------------------------------------------------------
// Cuts (evcut.rs:42-96)
// ---------------------------------------------------------------------------
template <class F> bool keep(const Config<F>& cfg, const Features& ft, const Event<F>& ev) {
    F e_min_ph;
    if (ft.no_photon_sorting) {  // event.rs:96-105
        e_min_ph = ev.p[2][3];
        for (int i = 1; i < 3; ++i) e_min_ph = (e_min_ph < ev.p[2 + i][3]) ? e_min_ph : ev.p[2 + i][3];
    } else {
        e_min_ph = ev.p[4][3];
    }
    if (e_min_ph < cfg.e_min) return false;
    const F* pel = ev.p[0];
    for (int i = 0; i < 3; ++i) {
        const F* ph = ev.p[2 + i];
        F num = (ph[0] * pel[0] + ph[1] * pel[1]) + ph[2] * pel[2];
        F den = ph[3] * pel[3];
        if (std::fabs(num) > cfg.beam_photons_cut * den) return false;
    }
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b) {
            const F* p1 = ev.p[2 + a];
            const F* p2 = ev.p[2 + b];
            F num = (p1[0] * p2[0] + p1[1] * p2[1]) + p1[2] * p2[2];
            F den = p1[3] * p2[3];
            if (num > cfg.photon_photon_cut * den) return false;
        }
    const F* a = ev.p[2];
    const F* b = ev.p[3];
    F n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    F cos_num = (pel[0] * n[0] + pel[1] * n[1]) + pel[2] * n[2];
    F nn = std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    F cos_den = pel[3] * nn;
    if (std::fabs(cos_num) < cfg.beam_photon_plane_cut * cos_den) return false;
    return true;
}

// ---------------------------------------------------------------------------
// Spinor products and helicity amplitudes (spinor.rs)
// ---------------------------------------------------------------------------
template <class F> struct Spinor {
    Cx<F> sx[5][5];
    static constexpr F RAC8() { return (F)2 * K<F>::SQRT_2; }
    explicit Spinor(const Event<F>& ev) {  // spinor.rs:31-54
        F xx[5];
        Cx<F> fx[5];
        for (int i = 0; i < 5; ++i) xx[i] = std::sqrt(ev.p[i][3] + ev.p[i][2]);
        for (int i = 0; i < 5; ++i) {
            if (xx[i] > K<F>::MIN_POSITIVE)
                fx[i] = Cx<F>{ev.p[i][0], ev.p[i][1]} / xx[i];
            else
                fx[i] = Cx<F>{std::sqrt((F)2 * ev.p[i][3]), (F)0};
        }
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) sx[i][j] = fx[i] * xx[j] - fx[j] * xx[i];
    }
    Cx<F> s(int i, int j) const { return sx[i][j]; }
    Cx<F> t(int i, int j) const { return -conj(sx[i][j]); }
    // spinor.rs:120-162 (E_M = 0, E_P = 1)
    Cx<F> a_ppm(int k1, int k2, int k3) const {
        return ((-RAC8()) * s(0, 1)) * sqr(s(0, k3)) / (((s(0, k1) * s(0, k2)) * s(1, k1)) * s(1, k2));
    }
    Cx<F> a_pmm(int k1, int k2, int k3) const {
        return ((-RAC8()) * t(0, 1)) * sqr(t(1, k1)) / (((t(1, k2) * t(1, k3)) * t(0, k2)) * t(0, k3));
    }
    Cx<F> bp_ppm(int k1, int k2, int k3) const {
        return ((-RAC8()) * t(0, 1)) * sqr(t(k1, k2) * s(k3, 0));
    }
    Cx<F> bp_pmm(int k1, int k2, int k3) const {
        return ((-RAC8()) * s(0, 1)) * sqr(t(k1, 1) * s(k2, k3));
    }
    Cx<F> bm_ppp(int k1, int k2, int k3) const {
        return ((-RAC8()) * s(0, 1)) *
               ((sqr(t(k1, k2) * t(k3, 1)) + sqr(t(k1, k3) * t(k2, 1))) + sqr(t(k2, k3) * t(k1, 1)));
    }
    Cx<F> bm_mmm(int k1, int k2, int k3) const {
        return ((-RAC8()) * t(0, 1)) *
               ((sqr(s(k1, 0) * s(k2, k3)) + sqr(s(k2, 0) * s(k1, k3))) + sqr(s(k3, 0) * s(k1, k2)));
    }
    // spinor.rs:71-116; helicity index 0..7 = MMM,MMP,MPM,MPP,PMM,PMP,PPM,PPP
    Cx<F> a(int h) const {
        switch (h) {
            case 1: return a_pmm(4, 2, 3);
            case 2: return a_pmm(3, 2, 4);
            case 3: return a_ppm(3, 4, 2);
            case 4: return a_pmm(2, 3, 4);
            case 5: return a_ppm(2, 4, 3);
            case 6: return a_ppm(2, 3, 4);
            default: return {0, 0};
        }
    }
    Cx<F> b_p(int h) const {
        switch (h) {
            case 1: return bp_pmm(4, 2, 3);
            case 2: return bp_pmm(3, 2, 4);
            case 3: return bp_ppm(3, 4, 2);
            case 4: return bp_pmm(2, 3, 4);
            case 5: return bp_ppm(2, 4, 3);
            case 6: return bp_ppm(2, 3, 4);
            default: return {0, 0};
        }
    }
    Cx<F> b_m(int h) const {
        switch (h) {
            case 0: return bm_mmm(2, 3, 4);
            case 7: return bm_ppp(2, 3, 4);
            default: return {0, 0};
        }
    }
};

// coupling.rs:23-33
template <class F> struct Couplings {
    F g_a, g_beta_p, g_beta_m;
    explicit Couplings(const Config<F>& cfg) {
        F e2 = (F)4 * K<F>::PI * cfg.alpha;
        F e2_z = (F)4 * K<F>::PI * cfg.alpha_z;
        F cos2w = (F)1 - cfg.sin2_weinberg;
        F mz2 = cfg.m_z0 * cfg.m_z0;
        F g_beta = -std::sqrt(e2_z / ((F)4 * cos2w * cfg.sin2_weinberg)) / (mz2 * mz2);
        F se = std::sqrt(e2);
        g_a = -(se * (se * se));  // powi(3): r = a; a = a*a; r = r*a
        g_beta_p = g_beta;
        g_beta_m = g_beta;
    }
};

// matelems.rs:53-86: the five helicity-summed squared matrix elements of one event
template <class F> void m2_sums(const Couplings<F>& cp, const Event<F>& ev, F out[5]) {
    Spinor<F> sp(ev);
    F m2[5][8];
    for (int h = 0; h < 8; ++h) {
        Cx<F> a = sp.a(h) * cp.g_a;
        Cx<F> bp = sp.b_p(h) * cp.g_beta_p;
        Cx<F> bm = sp.b_m(h) * cp.g_beta_m;
        Cx<F> mixed = ((F)2 * a) * conj(bp);
        m2[0][h] = norm_sqr(a);
        m2[1][h] = norm_sqr(bp);
        m2[2][h] = norm_sqr(bm);
        m2[3][h] = mixed.re;
        m2[4][h] = mixed.im;
    }
    for (int k = 0; k < 5; ++k) {
        F s = m2[k][0];
        for (int h = 1; h < 8; ++h) s += m2[k][h];
        out[k] = s;
    }
}

// ---------------------------------------------------------------------------
// Results accumulation (resacc.rs)
// ---------------------------------------------------------------------------
template <class F> struct Accumulator {
    uint64_t selected_events = 0;
    F spm2[5] = {0, 0, 0, 0, 0};
    F vars[5] = {0, 0, 0, 0, 0};
    F sigma_contribs[5];
    F sigma = 0, variance = 0;
    F fact_com, norm_weight, propagator, delta_with_z0_peak;

    Accumulator() = default;
    // resacc.rs:59-117
    Accumulator(const Config<F>& cfg, F ev_weight) {
        fact_com = (F)1 / (F)6 * cfg.gev2_to_picobarn;
        F relat_width = cfg.g_z0 / cfg.m_z0;
        F p_aa = 2;
        F p_ab = (F)1 - (F)4 * cfg.sin2_weinberg;
        F p_bb = p_ab + (F)8 * (cfg.sin2_weinberg * cfg.sin2_weinberg);
        F mz2 = cfg.m_z0 * cfg.m_z0;
        F c_aa = fact_com * p_aa;
        F c_ab = fact_com * p_ab / mz2;
        F c_bb = fact_com * p_bb / (mz2 * mz2);
        F ez = cfg.e_total / cfg.m_z0;
        F dzeta = ez * ez;
        delta_with_z0_peak = (dzeta - (F)1) / relat_width;
        propagator = (F)1 / ((F)1 + delta_with_z0_peak * delta_with_z0_peak);
        F n_ev = (F)cfg.num_events;
        // (2*PI).powi(-5) has constant operands: LLVM folds it through the host pow()
        F two_pi = (F)2 * K<F>::PI;
        F norm = (F)std::pow((double)two_pi, -5.0) / n_ev;
        norm_weight = ev_weight * norm;
        F com = norm_weight / (F)4;
        F aa = com * c_aa;
        F bb = com * c_bb * propagator / (relat_width * relat_width);
        F ab = com * c_ab * (F)2 * cfg.beta_plus * propagator / relat_width;
        sigma_contribs[0] = aa;
        sigma_contribs[1] = bb * (cfg.beta_plus * cfg.beta_plus);
        sigma_contribs[2] = bb * (cfg.beta_minus * cfg.beta_minus);
        sigma_contribs[3] = ab * delta_with_z0_peak;
        sigma_contribs[4] = -ab;
    }
    // resacc.rs:121-129
    void integrate(const F m[5]) {
        selected_events += 1;
        for (int k = 0; k < 5; ++k) spm2[k] += m[k];
        for (int k = 0; k < 5; ++k) vars[k] += m[k] * m[k];
        // nalgebra dot, 5-vector special case
        F a = m[0] * sigma_contribs[0], b = m[1] * sigma_contribs[1], c = m[2] * sigma_contribs[2],
          d = m[3] * sigma_contribs[3], e = m[4] * sigma_contribs[4];
        a += c;
        a += e;
        b += d;
        F weight = a + b;
        sigma += weight;
        variance += weight * weight;
    }
    // resacc.rs:133-139
    void merge(const Accumulator& o) {
        selected_events += o.selected_events;
        for (int k = 0; k < 5; ++k) spm2[k] += o.spm2[k];
        for (int k = 0; k < 5; ++k) vars[k] += o.vars[k];
        sigma += o.sigma;
        variance += o.variance;
    }
};

// resfin.rs:26-62
template <class F> struct FinalResults {
    uint64_t selected_events;
    F spm2[2][5], vars[2][5];
    F sigma, prec, variance, beta_min, ss_p, inc_ss_p, ss_m, inc_ss_m;
};

// resacc.rs:142-223
template <class F> FinalResults<F> finalize(const Config<F>& cfg, Accumulator<F> acc) {
    FinalResults<F> r;
    F n_ev = (F)cfg.num_events;
    for (int k = 0; k < 5; ++k) {
        F v = (acc.vars[k] - acc.spm2[k] * acc.spm2[k] / n_ev) / (n_ev - (F)1);
        acc.vars[k] = std::sqrt(v / n_ev) / std::fabs(acc.spm2[k] / n_ev);
    }
    for (int sp = 0; sp < 2; ++sp)
        for (int k = 0; k < 5; ++k) {
            r.spm2[sp][k] = acc.spm2[k];
            r.vars[sp][k] = acc.vars[k];
        }
    F polar_p = (F)-2 * cfg.sin2_weinberg;
    F polar_m = (F)1 + polar_p;
    F polars[2] = {polar_m, polar_p};
    for (int k = 1; k < 5; ++k)
        for (int sp = 0; sp < 2; ++sp) r.spm2[sp][k] *= polars[sp];
    for (int k = 1; k < 3; ++k)
        for (int sp = 0; sp < 2; ++sp) r.spm2[sp][k] *= polars[sp];
    F incident_flux = (F)1 / ((F)2 * (cfg.e_total * cfg.e_total));
    F scale = acc.fact_com * incident_flux * acc.norm_weight;
    for (int sp = 0; sp < 2; ++sp)
        for (int k = 0; k < 5; ++k) r.spm2[sp][k] *= scale;
    F gm_z0 = cfg.g_z0 * cfg.m_z0;
    for (int k = 1; k < 5; ++k)
        for (int sp = 0; sp < 2; ++sp) r.spm2[sp][k] *= acc.propagator / gm_z0;
    for (int k = 1; k < 3; ++k)
        for (int sp = 0; sp < 2; ++sp) r.spm2[sp][k] /= gm_z0;
    for (int sp = 0; sp < 2; ++sp) r.spm2[sp][3] *= acc.delta_with_z0_peak;

    auto colsum = [&](int k) { return ((F)0 + r.spm2[0][k]) + r.spm2[1][k]; };
    r.beta_min = std::sqrt(colsum(0) / colsum(1));
    F ss_denom = colsum(0);
    F ss_norm = (F)1 / ((F)2 * std::sqrt(ss_denom));
    r.ss_p = colsum(1) * ss_norm;
    r.ss_m = colsum(2) * ss_norm;
    auto inc_num = [&](int k) {
        F a = r.spm2[0][k] * r.vars[0][k], b = r.spm2[1][k] * r.vars[1][k];
        return std::sqrt(a * a + b * b);
    };
    F inc_ss_common = inc_num(0) / ((F)2 * std::fabs(ss_denom));
    r.inc_ss_p = inc_num(1) / std::fabs(colsum(1)) + inc_ss_common;
    r.inc_ss_m = inc_num(2) / std::fabs(colsum(2)) + inc_ss_common;
    r.variance = (acc.variance - acc.sigma * acc.sigma / n_ev) / (n_ev - (F)1);
    r.prec = std::sqrt(r.variance / n_ev) / std::fabs(acc.sigma / n_ev);
    r.sigma = acc.sigma * incident_flux;
    r.selected_events = acc.selected_events;
    return r;
}

// ---------------------------------------------------------------------------
// The per-batch kernel (main.rs:103-128) and the schedulers (scheduling/*.rs)
// ---------------------------------------------------------------------------
template <class F, class Rng>
Accumulator<F> simulate_events(uint64_t n, Rng& rng, const Config<F>& cfg, const Features& ft,
                               const Couplings<F>& cp, F ev_weight) {
    Accumulator<F> acc(cfg, ev_weight);
    for (uint64_t i = 0; i < n; ++i) {
        Event<F> ev = generate<F>(rng, ft, cfg.e_total);
        if (keep(cfg, ft, ev)) {
            F m[5];
            m2_sums(cp, ev, m);
            acc.integrate(m);
        }
    }
    return acc;
}

// evgen.rs:257-267: advance the master RNG past one batch (reproducible multi-threading)
template <class F, class Rng> void simulate_event_batch(Rng& rng, const Features& ft, uint64_t n) {
    if (ft.faster_evgen) {
        for (uint64_t i = 0; i < n; ++i) {
            F u[9];
            rng.template random_array<9>(u);
            F pts[3][2];
            random_unit_2d_outgoing<F>(rng, pts);
        }
    } else {
        for (uint64_t i = 0; i < n * 12; ++i) rng.random();
    }
}

}  // namespace oracle

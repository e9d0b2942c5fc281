// oracle/oracle_main.cpp — CPU ORACLE driver. TEST INFRASTRUCTURE ONLY.
//
// Schedulers (scheduling/mod.rs:31-59, sequential.rs:15-40,
// multi_threading.rs:16-75), a C API for the Python tests (ctypes) and a CLI
// that behaves like the reference binary (reads `valeurs`, prints the stdout
// report, writes `res.data`).  Build: see oracle/Makefile.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>

#include "oracle_text.hpp"

using namespace oracle;

extern "C" {
struct oracle_acc {
    uint64_t selected_events;
    double spm2[5], vars[5], sigma, variance;
};
enum {
    ORACLE_F32 = 1,
    ORACLE_FASTER_EVGEN = 2,
    ORACLE_FASTER_THREADING = 4,
    ORACLE_MULTI_THREADING = 8,
    ORACLE_NO_PHOTON_SORTING = 16,
    ORACLE_STANDARD_RANDOM = 32
};
}

static Features features_from_mask(uint32_t m) {
    Features f;
    f.f32 = m & ORACLE_F32;
    f.faster_evgen = m & ORACLE_FASTER_EVGEN;
    f.faster_threading = m & ORACLE_FASTER_THREADING;
    f.multi_threading = m & ORACLE_MULTI_THREADING;
    f.no_photon_sorting = m & ORACLE_NO_PHOTON_SORTING;
    f.standard_random = m & ORACLE_STANDARD_RANDOM;
    return f;
}

template <class F> static oracle_acc widen(const Accumulator<F>& a) {
    oracle_acc o;
    o.selected_events = a.selected_events;
    for (int k = 0; k < 5; ++k) {
        o.spm2[k] = (double)a.spm2[k];
        o.vars[k] = (double)a.vars[k];
    }
    o.sigma = (double)a.sigma;
    o.variance = (double)a.variance;
    return o;
}

struct RunOutput {
    std::vector<oracle_acc> per_batch;
    oracle_acc merged;
    std::string stdout_text, res_data_text;
    double seconds = 0;  // the reference's timed region (main.rs:85-138)
    uint64_t num_events = 0;
};

// Minimal task pool standing in for rayon::scope (multi_threading.rs:44-71)
class Pool {
  public:
    explicit Pool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(m_);
            done_ = true;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void spawn(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back(std::move(f));
            ++pending_;
        }
        cv_.notify_one();
    }
    void wait() {
        std::unique_lock<std::mutex> l(m_);
        idle_.wait(l, [this] { return pending_ == 0; });
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return done_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
            {
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) idle_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_, idle_;
    size_t pending_ = 0;
    bool done_ = false;
};

template <class F, class Rng>
static void run_typed(const Config<F>& cfg, const Features& ft, int n_threads, bool want_text,
                      bool keep_batches, RunOutput& out) {
    auto t0 = std::chrono::steady_clock::now();
    Couplings<F> cp(cfg);
    F ev_weight = event_weight<F>(cfg.e_total);
    Rng rng;
    uint64_t num_events = cfg.num_events;
    std::vector<Accumulator<F>> batches;

    if (!ft.multi_threading) {
        // sequential.rs:15-40
        uint64_t first = std::min<uint64_t>(EVENT_BATCH_SIZE, num_events);
        num_events -= first;
        batches.push_back(simulate_events<F>(first, rng, cfg, ft, cp, ev_weight));
        uint64_t full = num_events / EVENT_BATCH_SIZE;
        for (uint64_t b = 0; b < full; ++b)
            batches.push_back(simulate_events<F>(EVENT_BATCH_SIZE, rng, cfg, ft, cp, ev_weight));
        num_events %= EVENT_BATCH_SIZE;
        batches.push_back(simulate_events<F>(num_events, rng, cfg, ft, cp, ev_weight));
    } else {
        // multi_threading.rs:16-75
        uint64_t num_batches = num_events / EVENT_BATCH_SIZE + (num_events % EVENT_BATCH_SIZE != 0);
        batches.resize(num_batches);
        Pool pool(n_threads > 0 ? n_threads : 1);
        for (uint64_t b = 0; b < num_batches; ++b) {
            uint64_t sz = std::min<uint64_t>(num_events, EVENT_BATCH_SIZE);
            num_events -= sz;
            Rng task_rng = rng;
            pool.spawn([&, b, sz, task_rng]() mutable {
                batches[b] = simulate_events<F>(sz, task_rng, cfg, ft, cp, ev_weight);
            });
            if (!ft.faster_threading)
                simulate_event_batch<F>(rng, ft, sz);
            else
                rng.jump();
        }
        pool.wait();
    }
    // Left fold in batch order (sequential.rs:24-36, multi_threading.rs:107-126)
    Accumulator<F> acc = batches[0];
    for (size_t b = 1; b < batches.size(); ++b) acc.merge(batches[b]);
    FinalResults<F> fin = finalize(cfg, acc);
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    out.merged = widen(acc);
    out.num_events = cfg.num_events;
    if (keep_batches) {
        out.per_batch.reserve(batches.size());
        for (auto& b : batches) out.per_batch.push_back(widen(b));
    }
    if (want_text) {
        out.stdout_text = config_echo(cfg) + "IBegin\n" + eric(cfg, fin) + fawzi(cfg, fin);
        out.res_data_text = res_data(cfg, fin);
    }
}

template <class F>
static std::string run_float(const std::string& valeurs, const Features& ft, int n_threads,
                             uint64_t override_events, bool want_text, bool keep_batches, RunOutput& out) {
    Config<F> cfg;
    std::string err = load_config<F>(valeurs, cfg);
    if (!err.empty()) return err;
    if (override_events) cfg.num_events = override_events;
    if (ft.standard_random)
        run_typed<F, Xoshiro<F>>(cfg, ft, n_threads, want_text, keep_batches, out);
    else
        run_typed<F, Ranf<F>>(cfg, ft, n_threads, want_text, keep_batches, out);
    return "";
}

static std::string run_any(const std::string& valeurs, const Features& ft, int n_threads,
                           uint64_t override_events, bool want_text, bool keep_batches, RunOutput& out) {
    return ft.f32 ? run_float<float>(valeurs, ft, n_threads, override_events, want_text, keep_batches, out)
                  : run_float<double>(valeurs, ft, n_threads, override_events, want_text, keep_batches, out);
}

static void copy_text(const std::string& s, char* buf, size_t cap) {
    if (!buf || cap == 0) return;
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
}

// The oracle's own Ranf, wrapped only to count its refills (rounds): see oracle_fe_tile.
template <class F> struct CountingRanf {
    Ranf<F> g;
    uint64_t round = 0;
    template <int N> void random_array(F* out, int32_t* raw = nullptr) {
        if (g.index < N) ++round;  // ranf.rs:87-92: this request does not fit, the rest of the round is discarded
        g.template random_array<N>(out, raw);
    }
    F random() {
        F r;
        random_array<1>(&r);
        return r;
    }
};

extern "C" {

// Whole run. Returns 0 on success; per_batch may be NULL. *n_batches in: capacity, out: count.
int oracle_run(const char* valeurs_text, uint32_t feature_mask, int n_threads, uint64_t override_events,
               oracle_acc* per_batch, uint64_t* n_batches, oracle_acc* merged, char* res_data_buf,
               size_t res_data_cap, char* stdout_buf, size_t stdout_cap, double* seconds) {
    RunOutput out;
    Features ft = features_from_mask(feature_mask);
    std::string err = run_any(valeurs_text, ft, n_threads, override_events, res_data_buf || stdout_buf,
                              per_batch != nullptr, out);
    if (!err.empty()) {
        copy_text(err, stdout_buf, stdout_cap);
        return 1;
    }
    if (per_batch && n_batches) {
        uint64_t n = std::min<uint64_t>(*n_batches, out.per_batch.size());
        for (uint64_t i = 0; i < n; ++i) per_batch[i] = out.per_batch[i];
        *n_batches = out.per_batch.size();
    }
    if (merged) *merged = out.merged;
    copy_text(out.res_data_text, res_data_buf, res_data_cap);
    copy_text(out.stdout_text, stdout_buf, stdout_cap);
    if (seconds) *seconds = out.seconds;
    return 0;
}

// Raw integer stream at the start of batch `batch` (default evgen: 12 draws per event, one
// number per request), as the scheduler positions it: sequential-stream position, or after
// `batch` jump()s under faster-threading.  RANF words are the i32 values, xoshiro words are
// the full 64-bit (f64) or 32-bit (f32) outputs.
int oracle_rng_words(uint32_t feature_mask, uint64_t batch, uint32_t n_words, uint64_t* out) {
    Features ft = features_from_mask(feature_mask);
    auto go = [&](auto rng) {
        if (ft.faster_threading)
            for (uint64_t b = 0; b < batch; ++b) rng.jump();
        else
            for (uint64_t i = 0; i < batch * EVENT_BATCH_SIZE * 12; ++i) rng.next_raw();
        for (uint32_t i = 0; i < n_words; ++i) out[i] = rng.next_raw();
    };
    if (ft.standard_random) {
        if (ft.f32) go(Xoshiro<float>()); else go(Xoshiro<double>());
    } else {
        if (ft.f32) go(Ranf<float>()); else go(Ranf<double>());
    }
    return 0;
}

// The first `n` events of the sequential stream: momenta of the 3 photons [n][3][4] (X,Y,Z,E),
// cut decision [n], helicity-summed matrix elements [n][5] (zero if rejected).
int oracle_events(const char* valeurs_text, uint32_t feature_mask, uint32_t n, double* momenta, int32_t* kept,
                  double* m2) {
    Features ft = features_from_mask(feature_mask);
    auto go = [&](auto fzero, auto rng) -> int {
        using F = decltype(fzero);
        Config<F> cfg;
        if (!load_config<F>(valeurs_text, cfg).empty()) return 1;
        Couplings<F> cp(cfg);
        for (uint32_t i = 0; i < n; ++i) {
            Event<F> ev = generate<F>(rng, ft, cfg.e_total);
            for (int p = 0; p < 3; ++p)
                for (int c = 0; c < 4; ++c) momenta[(i * 3 + p) * 4 + c] = (double)ev.p[2 + p][c];
            bool k = keep(cfg, ft, ev);
            kept[i] = k;
            F m[5] = {0, 0, 0, 0, 0};
            if (k) m2_sums(cp, ev, m);
            for (int c = 0; c < 5; ++c) m2[i * 5 + c] = (double)m[c];
        }
        return 0;
    };
    if (ft.standard_random)
        return ft.f32 ? go(0.0f, Xoshiro<float>()) : go(0.0, Xoshiro<double>());
    return ft.f32 ? go(0.0f, Ranf<float>()) : go(0.0, Ranf<double>());
}

// The events of the sequential faster-evgen RANF stream that START in rounds [first_round, first_round + n_rounds)
// (n_rounds = 0: no round limit), at most max_events of them (0: no limit), accumulated in groups of 10 000 consecutive
// events counted from the first one and merged in order: the checker of tp3_fe_tile_device.  A "round" is one refill of the
// generator (ranf.rs:87-92,106-119), round 0 the state after seeding; an event starts in the round that serves its
// random_array::<9>() (evgen.rs:149).  The generator is the oracle's own Ranf, wrapped only to count its refills.
int oracle_fe_tile(const char* valeurs_text, uint32_t feature_mask, uint64_t first_round, uint64_t n_rounds, uint64_t max_events,
                   oracle_acc* merged, uint64_t* events_done) {
    Features ft = features_from_mask(feature_mask);
    if (!ft.faster_evgen || ft.standard_random) return 2;
    auto go = [&](auto fzero) -> int {
        using F = decltype(fzero);
        Config<F> cfg;
        if (!load_config<F>(valeurs_text, cfg).empty()) return 1;
        Couplings<F> cp(cfg);
        const F w = event_weight<F>(cfg.e_total);
        CountingRanf<F> rng;
        Accumulator<F> total(cfg, w), batch(cfg, w);
        uint64_t done = 0, in_batch = 0;
        bool first_batch = true;
        auto flush = [&]() {
            if (first_batch) total = batch;
            else total.merge(batch);
            first_batch = false;
            batch = Accumulator<F>(cfg, w);
            in_batch = 0;
        };
        for (;;) {
            if (max_events && done == max_events) break;
            // where would the next event start?  (the 9-number request moves to the next round if fewer than 9 are left)
            const uint64_t start_round = rng.round + (rng.g.index < 9 ? 1 : 0);
            if (n_rounds && start_round >= first_round + n_rounds) break;
            Event<F> ev = generate<F>(rng, ft, cfg.e_total);
            if (start_round < first_round) continue;  // an event of an earlier tile
            if (keep(cfg, ft, ev)) {
                F m[5];
                m2_sums(cp, ev, m);
                batch.integrate(m);
            }
            ++done;
            if (++in_batch == EVENT_BATCH_SIZE) flush();
        }
        if (in_batch || first_batch) flush();
        *merged = widen(total);
        *events_done = done;
        return 0;
    };
    return ft.f32 ? go(0.0f) : go(0.0);
}

// finalize (resacc.rs:142-223) + the text surfaces (resfin.rs:66-194, output.rs:30-177) applied to GIVEN sums: lets the
// tests cross-check the product's host formatting against this restatement on values the golden runs never produce.
int oracle_finalize_text(const char* valeurs_text, uint32_t feature_mask, const oracle_acc* sums, char* res_data_buf,
                         size_t res_data_cap, char* stdout_buf, size_t stdout_cap) {
    Features ft = features_from_mask(feature_mask);
    auto go = [&](auto fzero) -> int {
        using F = decltype(fzero);
        Config<F> cfg;
        if (!load_config<F>(valeurs_text, cfg).empty()) return 1;
        Accumulator<F> acc(cfg, event_weight<F>(cfg.e_total));
        acc.selected_events = sums->selected_events;
        for (int k = 0; k < 5; ++k) {
            acc.spm2[k] = (F)sums->spm2[k];
            acc.vars[k] = (F)sums->vars[k];
        }
        acc.sigma = (F)sums->sigma;
        acc.variance = (F)sums->variance;
        FinalResults<F> fin = finalize(cfg, acc);
        copy_text(config_echo(cfg) + "IBegin\n" + eric(cfg, fin) + fawzi(cfg, fin), stdout_buf, stdout_cap);
        copy_text(res_data(cfg, fin), res_data_buf, res_data_cap);
        return 0;
    };
    return ft.f32 ? go(0.0f) : go(0.0);
}

// Host-side constants for a configuration (coupling.rs:23-33, evgen.rs:60-62, resacc.rs:59-117),
// out = {g_a, g_beta_p, g_beta_m, ev_weight, norm_weight, sigma_contribs[5]}
int oracle_constants(const char* valeurs_text, uint32_t feature_mask, double* out) {
    Features ft = features_from_mask(feature_mask);
    auto go = [&](auto fzero) -> int {
        using F = decltype(fzero);
        Config<F> cfg;
        if (!load_config<F>(valeurs_text, cfg).empty()) return 1;
        Couplings<F> cp(cfg);
        F w = event_weight<F>(cfg.e_total);
        Accumulator<F> acc(cfg, w);
        out[0] = cp.g_a; out[1] = cp.g_beta_p; out[2] = cp.g_beta_m; out[3] = w; out[4] = acc.norm_weight;
        for (int k = 0; k < 5; ++k) out[5 + k] = acc.sigma_contribs[k];
        return 0;
    };
    return ft.f32 ? go(0.0f) : go(0.0);
}

}  // extern "C"

#ifdef ORACLE_CLI
// Usage: oracle_cli [--features a,b,c] [--threads N] [--events N] [--valeurs PATH] [--out res.data]
int main(int argc, char** argv) {
    std::string feats, valeurs_path = "valeurs", out_path = "res.data";
    int threads = (int)std::thread::hardware_concurrency();
    uint64_t events = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--features") feats = next();
        else if (a == "--threads") threads = atoi(next().c_str());
        else if (a == "--events") events = strtoull(next().c_str(), nullptr, 10);
        else if (a == "--valeurs") valeurs_path = next();
        else if (a == "--out") out_path = next();
    }
    uint32_t mask = 0;
    std::stringstream ss(feats);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok == "f32") mask |= ORACLE_F32;
        else if (tok == "faster-evgen") mask |= ORACLE_FASTER_EVGEN;
        else if (tok == "faster-threading") mask |= ORACLE_FASTER_THREADING;
        else if (tok == "multi-threading") mask |= ORACLE_MULTI_THREADING;
        else if (tok == "no-photon-sorting") mask |= ORACLE_NO_PHOTON_SORTING;
        else if (tok == "standard-random") mask |= ORACLE_STANDARD_RANDOM;
        else if (!tok.empty()) { fprintf(stderr, "unknown feature %s\n", tok.c_str()); return 2; }
    }
    std::ifstream in(valeurs_path);
    if (!in) { fprintf(stderr, "failed to load the configuration\n"); return 1; }
    std::stringstream buf;
    buf << in.rdbuf();
    RunOutput out;
    std::string err = run_any(buf.str(), features_from_mask(mask), threads, events, true, false, out);
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    fputs(out.stdout_text.c_str(), stdout);
    std::ofstream(out_path) << out.res_data_text;
    fprintf(stderr, "elapsed %.6f s, %.4g events/s\n", out.seconds,
            (double)out.num_events / out.seconds);
    return 0;
}
#endif

// oracle/gen_far_kat.cpp — TEST INFRASTRUCTURE ONLY (generator of tests/golden/far_stream_kat.json).
//
// Known answers for the random streams at the batch indices the benchmark actually reaches:
// the ORACLE's generators (oracle.hpp: Ranf = random/ranf.rs, Xoshiro = random/standard.rs) are stepped
// SEQUENTIALLY, one random() at a time, exactly as the reference's scheduler pre-advances its master generator
// (evgen.rs:257-267, ranf.rs:122-126 `skip`), from the seed to the start of each wanted batch.  Nothing here uses
// jump-ahead algebra: that is what these vectors are there to check.
//   RANF              : batches 999 999, 1 968 526, 1 968 527 (first round index >= 2^32: the 5th byte digit of the device
//                       jump-ahead), 7 999 999 (last batch of the 8-GPU weak run)   -> 9.6e11 sequential draws
//   xoshiro256+/128+  : batches 999 999 and 7 999 999 of the sequential stream, and batch 7 999 999 after as many jump()s
// For each: the first 240 raw words of the batch, the generator state, and (f64, batches 999 999 / 7 999 999) the
// batch's ResultsAccumulator from the oracle's simulate_events with cfg.num_events = 8e10.
//
//   g++ -O3 -std=c++17 -ffp-contract=off -pthread -o _build/gen_far_kat gen_far_kat.cpp && _build/gen_far_kat VALEURS OUT.json
// Cost: ~15 CPU-minutes per stream (three streams run in parallel threads).
#include <cinttypes>
#include <fstream>
#include <sstream>
#include <thread>

#include "oracle.hpp"

using namespace oracle;

static const uint64_t kDrawsPerBatch = 12ull * EVENT_BATCH_SIZE;
static const uint64_t kKatEvents = 80000000000ull;  // num_events of the run the accumulators belong to (8e6 batches)

struct Entry {
    std::string rng, seeding, dtype;
    uint64_t batch;
    std::vector<uint64_t> words, state;
    bool has_acc = false;
    Accumulator<double> acc;
};

template <class Rng> static std::vector<uint64_t> peek_words(Rng rng /*copy*/, int n) {
    std::vector<uint64_t> w(n);
    for (int i = 0; i < n; ++i) w[i] = rng.next_raw();
    return w;
}
static std::vector<uint64_t> state_of(const Ranf<double>& g) {
    std::vector<uint64_t> s;
    for (int i = 1; i <= 55; ++i) s.push_back((uint32_t)g.numbers[i]);
    s.push_back((uint64_t)g.index);
    return s;
}
static std::vector<uint64_t> state_of(const Xoshiro<double>& g) { return {g.s[0], g.s[1], g.s[2], g.s[3]}; }
static std::vector<uint64_t> state_of(const Xoshiro<float>& g) { return {g.s[0], g.s[1], g.s[2], g.s[3]}; }

template <class Rng> static Accumulator<double> batch_acc(Rng rng /*copy*/, const std::string& valeurs, bool standard_random) {
    Config<double> cfg;
    std::string err = load_config<double>(valeurs, cfg);
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); exit(1); }
    cfg.num_events = kKatEvents;
    Features ft;
    ft.standard_random = standard_random;
    Couplings<double> cp(cfg);
    return simulate_events<double>(EVENT_BATCH_SIZE, rng, cfg, ft, cp, event_weight<double>(cfg.e_total));
}

// Sequential walk: `targets` ascending batch indices.
template <class Rng>
static void walk(const char* name, const char* dtype, const std::vector<uint64_t>& targets, const std::vector<bool>& want_acc,
                 const std::string& valeurs, bool standard_random, std::vector<Entry>& out) {
    Rng rng;
    uint64_t pos = 0;  // draws consumed
    for (size_t t = 0; t < targets.size(); ++t) {
        const uint64_t goal = targets[t] * kDrawsPerBatch;
        for (; pos < goal; ++pos) rng.random();  // ranf.rs:122-126 / standard.rs:38-42: one random() per skipped draw
        Entry e;
        e.rng = name;
        e.seeding = "sequential";
        e.dtype = dtype;
        e.batch = targets[t];
        e.words = peek_words(rng, 240);
        e.state = state_of(rng);
        if (want_acc[t]) {
            if constexpr (!std::is_same<Rng, Xoshiro<float>>::value) {
                e.has_acc = true;
                e.acc = batch_acc(rng, valeurs, standard_random);
            }
        }
        out.push_back(e);
        fprintf(stderr, "[gen_far_kat] %s batch %" PRIu64 " done\n", name, targets[t]);
    }
}

template <class Rng> static void jumps(const char* name, const char* dtype, uint64_t batch, std::vector<Entry>& out) {
    Rng rng;
    for (uint64_t b = 0; b < batch; ++b) rng.jump();  // multi_threading.rs:66-69
    Entry e;
    e.rng = name;
    e.seeding = "jump";
    e.dtype = dtype;
    e.batch = batch;
    e.words = peek_words(rng, 240);
    e.state = state_of(rng);
    out.push_back(e);
}

static void emit(std::ostream& os, const Entry& e, bool last) {
    os << "  {\"rng\": \"" << e.rng << "\", \"seeding\": \"" << e.seeding << "\", \"dtype\": \"" << e.dtype << "\", \"batch\": " << e.batch
       << ",\n   \"words\": [";
    for (size_t i = 0; i < e.words.size(); ++i) os << (i ? ", " : "") << e.words[i];
    os << "],\n   \"state\": [";
    for (size_t i = 0; i < e.state.size(); ++i) os << (i ? ", " : "") << e.state[i];
    os << "]";
    if (e.has_acc) {
        char buf[64];
        auto num = [&](double v) { snprintf(buf, sizeof buf, "%.17g", v); return std::string(buf); };
        os << ",\n   \"acc\": {\"num_events_total\": " << kKatEvents << ", \"selected_events\": " << e.acc.selected_events << ", \"spm2\": [";
        for (int k = 0; k < 5; ++k) os << (k ? ", " : "") << num(e.acc.spm2[k]);
        os << "], \"vars\": [";
        for (int k = 0; k < 5; ++k) os << (k ? ", " : "") << num(e.acc.vars[k]);
        os << "], \"sigma\": " << num(e.acc.sigma) << ", \"variance\": " << num(e.acc.variance) << "}";
    }
    os << "}" << (last ? "\n" : ",\n");
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: gen_far_kat VALEURS OUT.json [--quick]\n"); return 2; }
    std::ifstream in(argv[1]);
    std::stringstream buf;
    buf << in.rdbuf();
    const std::string valeurs = buf.str();
    const bool quick = argc > 3 && std::string(argv[3]) == "--quick";  // small indices: checks the generator itself in seconds
    std::vector<uint64_t> ranf_t = quick ? std::vector<uint64_t>{999, 2001} : std::vector<uint64_t>{999999, 1968526, 1968527, 7999999};
    std::vector<bool> ranf_acc = quick ? std::vector<bool>{true, true} : std::vector<bool>{true, false, false, true};
    std::vector<uint64_t> xo_t = quick ? std::vector<uint64_t>{999, 2001} : std::vector<uint64_t>{999999, 7999999};
    std::vector<bool> xo_acc = {true, true};
    std::vector<Entry> a, b, c, d;
    std::thread t1([&] { walk<Ranf<double>>("ranf", "f64", ranf_t, ranf_acc, valeurs, false, a); });
    std::thread t2([&] { walk<Xoshiro<double>>("xoshiro256+", "f64", xo_t, xo_acc, valeurs, true, b); });
    std::thread t3([&] { walk<Xoshiro<float>>("xoshiro128+", "f32", xo_t, xo_acc, valeurs, true, c); });
    jumps<Xoshiro<double>>("xoshiro256+", "f64", xo_t.back(), d);
    jumps<Xoshiro<float>>("xoshiro128+", "f32", xo_t.back(), d);
    t1.join(); t2.join(); t3.join();
    std::vector<Entry> all;
    for (auto* v : {&a, &b, &c, &d}) all.insert(all.end(), v->begin(), v->end());
    std::ofstream os(argv[2]);
    os << "{\"generator\": \"oracle/gen_far_kat.cpp (sequential walk of the oracle's generators; no jump-ahead)\",\n \"entries\": [\n";
    for (size_t i = 0; i < all.size(); ++i) emit(os, all[i], i + 1 == all.size());
    os << "]}\n";
    return 0;
}

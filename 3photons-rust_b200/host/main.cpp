// trois_photons_b200 — command-line twin of the reference binary (main.rs:75-145): reads
// `valeurs` from the working directory, prints the reference's stdout report, writes res.data,
// res.times and appends to pil.mc.  The cargo features become --features a,b,c.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/tp3.h"

int main(int argc, char** argv) {
    std::string feats, valeurs = "valeurs", out_dir = "";
    int gpus = 1;
    uint32_t kernel = TP3_KERNEL_FAST;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--features") feats = val();
        else if (a == "--gpus") gpus = std::atoi(val().c_str());
        else if (a == "--valeurs") valeurs = val();
        else if (a == "--out-dir") out_dir = val();
        else if (a == "--kernel") kernel = val() == "literal" ? TP3_KERNEL_LITERAL : TP3_KERNEL_FAST;
        else {
            std::fprintf(stderr, "usage: %s [--features f32,faster-evgen,no-photon-sorting,standard-random,"
                                 "multi-threading,faster-threading] [--gpus N] [--valeurs PATH] [--out-dir DIR] "
                                 "[--kernel fast|literal]\n", argv[0]);
            return 2;
        }
    }
    uint32_t flags = 0;
    std::stringstream ss(feats);
    std::string t;
    while (std::getline(ss, t, ',')) {
        if (t == "f32") flags |= TP3_F32;
        else if (t == "faster-evgen") flags |= TP3_FASTER_EVGEN;
        else if (t == "faster-threading") flags |= TP3_FASTER_THREADING;
        else if (t == "multi-threading") flags |= TP3_MULTI_THREADING;
        else if (t == "no-photon-sorting") flags |= TP3_NO_PHOTON_SORTING;
        else if (t == "standard-random") flags |= TP3_STANDARD_RANDOM;
        else if (!t.empty()) {
            std::fprintf(stderr, "unknown feature '%s'\n", t.c_str());
            return 2;
        }
    }
    // faster-threading only changes seeding when multi-threading is on (scheduling/mod.rs:45-54)
    if (!(flags & TP3_MULTI_THREADING)) flags &= ~TP3_FASTER_THREADING;
    std::vector<char> out(1 << 16);
    double secs = 0;
    int rc = tp3_run(valeurs.c_str(), out_dir.c_str(), flags, kernel, gpus, out.data(), out.size(), &secs);
    if (rc != TP3_OK) {
        std::fprintf(stderr, "Error: %s (code %d)\n", out.data(), rc);
        return 1;
    }
    std::fputs(out.data(), stdout);
    return 0;
}

// Host side of the reference surface, kept next to the GPU hot path so that a caller can go
// `valeurs` -> kernel parameters -> merged accumulator -> `res.data` / stdout exactly like the
// reference binary does (main.rs:75-145).  Everything here is O(1) scalar work:
//   tp3_config_parse        config.rs:57-128   (first token of each non-blank line, 18 items)
//   tp3_params_from_config  coupling.rs:23-33, evgen.rs:40-77, resacc.rs:59-117
//   tp3_merge               resacc.rs:133-139
//   tp3_finalize            resacc.rs:142-223
//   tp3_format_stdout       config.rs:131-155, evgen.rs:47, resfin.rs:66-194
//   tp3_format_res_data     output.rs:64-142,180-269
//   tp3_run                 main.rs:75-145 + scheduling/mod.rs:31-59 + output.rs:30-177
// All arithmetic is done in the run's Float (f32 under TP3_F32) and carried across the C ABI
// as exactly-widened doubles.
#include <charconv>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/tp3.h"

namespace {

template <class F> struct Lim;
template <> struct Lim<double> { static constexpr int digits10 = 15; };
template <> struct Lim<float> { static constexpr int digits10 = 6; };

template <class F> F pi() { return (F)3.14159265358979323846264338327950288L; }

// ---- Rust-compatible number formatting -------------------------------------------------
std::string special(double x) { return std::isnan(x) ? "NaN" : (x < 0 ? "-inf" : "inf"); }

// `{}`: shortest digits that round-trip, positional notation
template <class F> std::string fmt_display(F x) {
    if (!std::isfinite(x)) return special(x);
    char raw[64];
    auto r = std::to_chars(raw, raw + sizeof raw, x, std::chars_format::scientific);
    std::string s(raw, r.ptr), sign;
    if (s[0] == '-') {
        sign = "-";
        s.erase(0, 1);
    }
    const size_t epos = s.find('e');
    const int ex = std::atoi(s.c_str() + epos + 1);
    std::string dig;
    for (size_t i = 0; i < epos; ++i)
        if (s[i] != '.') dig.push_back(s[i]);
    const int nd = (int)dig.size();
    std::string body;
    if (dig == "0") body = "0";
    else if (ex >= nd - 1) body = dig + std::string((size_t)(ex - nd + 1), '0');
    else if (ex >= 0) body = dig.substr(0, (size_t)ex + 1) + "." + dig.substr((size_t)ex + 1);
    else body = "0." + std::string((size_t)(-ex - 1), '0') + dig;
    return sign + body;
}
// `{:.p}`
template <class F> std::string fmt_fixed(F x, int p) {
    if (!std::isfinite(x)) return special(x);
    char b[400];
    std::snprintf(b, sizeof b, "%.*f", p, (double)x);
    return b;
}
// `{:.pe}`: exponent printed bare (e-4, e0, e12)
template <class F> std::string fmt_sci(F x, int p) {
    if (!std::isfinite(x)) return special(x);
    char b[96];
    std::snprintf(b, sizeof b, "%.*e", p, (double)x);
    char* e = std::strchr(b, 'e');
    const int ex = std::atoi(e + 1);
    *e = 0;
    return std::string(b) + "e" + std::to_string(ex);
}
std::string rjust(const std::string& s, size_t w) { return s.size() < w ? std::string(w - s.size(), ' ') + s : s; }
std::string ljust(const std::string& s, size_t w) { return s.size() < w ? s + std::string(w - s.size(), ' ') : s; }

// output.rs:230-269 — %g-like: positional for 1e-3 <= |x| <= 10^sig, else scientific
template <class F> std::string fmt_engineering(F x, int sig) {
    if (x == (F)0) return "0";
    const F lg = std::log10(std::fabs(x));
    if (lg >= (F)-3 && lg <= (F)sig) {
        int prec = sig - 1 - (int)std::trunc(lg);
        if (lg < (F)0) prec += 1;
        if (prec < 0) prec = 0;
        std::string s = fmt_fixed(x, prec);
        if (s.find('.') != std::string::npos) {
            s.erase(s.find_last_not_of('0') + 1);
            if (s.back() == '.') s.pop_back();
        }
        return s;
    }
    return fmt_sci(x, sig - 1);
}

// ---- typed views ------------------------------------------------------------------------
template <class F> struct Cfg {
    uint64_t n;
    F e_total, acut, bcut, e_min, sincut, alpha, alpha_z, conv, m_z0, g_z0, s2w, br, beta_p, beta_m;
    int32_t nbins;
    bool impr, plot;
    explicit Cfg(const tp3_config& c)
        : n(c.num_events), e_total((F)c.e_total), acut((F)c.beam_photons_cut), bcut((F)c.photon_photon_cut),
          e_min((F)c.e_min), sincut((F)c.beam_photon_plane_cut), alpha((F)c.alpha), alpha_z((F)c.alpha_z),
          conv((F)c.gev2_to_picobarn), m_z0((F)c.m_z0), g_z0((F)c.g_z0), s2w((F)c.sin2_weinberg),
          br((F)c.branching_ep_em), beta_p((F)c.beta_plus), beta_m((F)c.beta_minus), nbins(c.num_bins),
          impr(c.impr != 0), plot(c.plot != 0) {}
};

// Quantities ResultsAccumulator::new caches for finalize (resacc.rs:36-55)
template <class F> struct AccConsts {
    F fact_com, norm_weight, propagator, delta, sigma_contribs[5];
};

template <class F> F event_weight(F e_total) {  // evgen.rs:49-62 for 3 photons
    F z = (F)2 * std::log(pi<F>() / (F)2);      // (INP-1) ln(pi/2)
    z -= (F)2 * std::log((F)1);                 // k = 2
    z -= std::log((F)2);                        // ln(INP-1)
    const F lnw = ((F)2 * (F)3 - (F)4) * std::log(e_total) + z;
    return std::exp(lnw);
}

template <class F> AccConsts<F> acc_consts(const Cfg<F>& c) {  // resacc.rs:59-100
    AccConsts<F> k;
    k.fact_com = (F)1 / (F)6 * c.conv;
    const F rw = c.g_z0 / c.m_z0;
    const F p_aa = 2, p_ab = (F)1 - (F)4 * c.s2w, p_bb = p_ab + (F)8 * (c.s2w * c.s2w);
    const F mz2 = c.m_z0 * c.m_z0;
    const F c_aa = k.fact_com * p_aa, c_ab = k.fact_com * p_ab / mz2, c_bb = k.fact_com * p_bb / (mz2 * mz2);
    const F er = c.e_total / c.m_z0;
    k.delta = (er * er - (F)1) / rw;
    k.propagator = (F)1 / ((F)1 + k.delta * k.delta);
    // (2 pi)^(4 - 3*3): constant operands, folded by the Rust compiler through pow()
    const F norm = (F)std::pow((double)((F)2 * pi<F>()), -5.0) / (F)c.n;
    k.norm_weight = event_weight<F>(c.e_total) * norm;
    const F com = k.norm_weight / (F)4;
    const F aa = com * c_aa;
    const F bb = com * c_bb * k.propagator / (rw * rw);
    const F ab = com * c_ab * (F)2 * c.beta_p * k.propagator / rw;
    k.sigma_contribs[0] = aa;
    k.sigma_contribs[1] = bb * (c.beta_p * c.beta_p);
    k.sigma_contribs[2] = bb * (c.beta_m * c.beta_m);
    k.sigma_contribs[3] = ab * k.delta;
    k.sigma_contribs[4] = -ab;
    return k;
}

template <class F> void make_params(const tp3_config& raw, uint32_t flags, uint32_t kernel, tp3_params& out) {
    const Cfg<F> c(raw);
    std::memset(&out, 0, sizeof out);
    out.num_events_total = c.n;
    out.e_total = c.e_total;
    out.acut = c.acut;
    out.bcut = c.bcut;
    out.e_min = c.e_min;
    out.sincut = c.sincut;
    // coupling.rs:23-33
    const F e2 = (F)4 * pi<F>() * c.alpha, e2z = (F)4 * pi<F>() * c.alpha_z;
    const F cos2 = (F)1 - c.s2w;
    const F mz2 = c.m_z0 * c.m_z0;
    const F gb = -std::sqrt(e2z / ((F)4 * cos2 * c.s2w)) / (mz2 * mz2);
    const F se = std::sqrt(e2);
    out.g_a = -(se * (se * se));
    out.g_beta_p = gb;
    out.g_beta_m = gb;
    const AccConsts<F> k = acc_consts(c);
    for (int i = 0; i < 5; ++i) out.sigma_contribs[i] = k.sigma_contribs[i];
    out.flags = flags;
    out.kernel = kernel;
}

template <class F> void finalize_t(const tp3_config& raw, const tp3_acc& acc, tp3_final& out) {  // resacc.rs:142-223
    const Cfg<F> c(raw);
    const AccConsts<F> k = acc_consts(c);
    const F n_ev = (F)c.n;
    F tot[5], rel[5];
    for (int i = 0; i < 5; ++i) {
        tot[i] = (F)acc.spm2[i];
        F v = ((F)acc.vars[i] - tot[i] * tot[i] / n_ev) / (n_ev - (F)1);
        rel[i] = std::sqrt(v / n_ev) / std::fabs(tot[i] / n_ev);
    }
    F sp[2][5];
    const F pol_p = (F)-2 * c.s2w, pol_m = (F)1 + pol_p;
    const F pol[2] = {pol_m, pol_p};
    const F flux = (F)1 / ((F)2 * (c.e_total * c.e_total));
    const F scale = k.fact_com * flux * k.norm_weight;
    const F gm = c.g_z0 * c.m_z0;
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < 5; ++i) {
            F x = tot[i];
            if (i >= 1) x *= pol[s];
            if (i == 1 || i == 2) x *= pol[s];
            x *= scale;
            if (i >= 1) x *= k.propagator / gm;
            if (i == 1 || i == 2) x /= gm;
            if (i == 3) x *= k.delta;
            sp[s][i] = x;
        }
    auto sum2 = [&](int i) { return ((F)0 + sp[0][i]) + sp[1][i]; };
    auto quad = [&](int i) {
        const F a = sp[0][i] * rel[i], b = sp[1][i] * rel[i];
        return std::sqrt(a * a + b * b);
    };
    const F sa = sum2(0);
    out.beta_min = std::sqrt(sa / sum2(1));
    const F ss_norm = (F)1 / ((F)2 * std::sqrt(sa));
    out.ss_p = sum2(1) * ss_norm;
    out.ss_m = sum2(2) * ss_norm;
    const F common = quad(0) / ((F)2 * std::fabs(sa));
    out.inc_ss_p = quad(1) / std::fabs(sum2(1)) + common;
    out.inc_ss_m = quad(2) / std::fabs(sum2(2)) + common;
    const F sig = (F)acc.sigma;
    const F var = ((F)acc.variance - sig * sig / n_ev) / (n_ev - (F)1);
    out.variance = var;
    out.prec = std::sqrt(var / n_ev) / std::fabs(sig / n_ev);
    out.sigma = sig * flux;
    out.selected_events = acc.selected_events;
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < 5; ++i) {
            out.spm2[s][i] = sp[s][i];
            out.vars[s][i] = rel[i];
        }
}

template <class F> std::string stdout_t(const tp3_config& raw, const tp3_final& fin) {
    const Cfg<F> c(raw);
    std::ostringstream o;
    // config.rs:131-155
    o << "ITOT           : " << c.n << "\n";
    o << "ETOT           : " << fmt_display(c.e_total) << "\n";
    o << "oCutpar.ACUT   : " << fmt_display(c.acut) << "\n";
    o << "oCutpar.BCUT   : " << fmt_display(c.bcut) << "\n";
    o << "oCutpar.EMIN   : " << fmt_display(c.e_min) << "\n";
    o << "oCutpar.SINCUT : " << fmt_display(c.sincut) << "\n";
    o << "ALPHA          : " << fmt_display(c.alpha) << "\n";
    o << "ALPHAZ         : " << fmt_display(c.alpha_z) << "\n";
    o << "CONVERS        : " << fmt_display(c.conv) << "\n";
    o << "oParam.MZ0     : " << fmt_display(c.m_z0) << "\n";
    o << "oParam.GZ0     : " << fmt_display(c.g_z0) << "\n";
    o << "SIN2W          : " << fmt_display(c.s2w) << "\n";
    o << "BREPEM         : " << fmt_display(c.br) << "\n";
    o << "BETAPLUS       : " << fmt_display(c.beta_p) << "\n";
    o << "BETAMOINS      : " << fmt_display(c.beta_m) << "\n";
    o << "NBIN           : " << c.nbins << "\n";
    o << "oParam.IMPR    : " << (c.impr ? "true" : "false") << "\n";
    o << "PLOT           : " << (c.plot ? "true" : "false") << "\n";
    o << "IBegin\n";  // evgen.rs:47

    F sp[2][5], rel[5];
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < 5; ++i) sp[s][i] = (F)fin.spm2[s][i];
    for (int i = 0; i < 5; ++i) rel[i] = (F)fin.vars[0][i];
    const F PI = pi<F>();
    // resfin.rs:66-97
    {
        const F mu_th = c.br * c.conv / ((F)8 * (F)9 * (F)5 * (F)std::pow((double)PI, 2.0) * c.m_z0 * c.g_z0);
        F sigma0[2], alpha0[2], beta0[2], lambda0[2], mu0[2];
        for (int s = 0; s < 2; ++s) {
            sigma0[s] = sp[s][0] / (F)2;
            alpha0[s] = sp[s][4] / (F)2;
            beta0[s] = -sp[s][3] / (F)2;
            lambda0[s] = (sp[s][2] - sp[s][1]) / (F)2;
            mu0[s] = (sp[s][2] + sp[s][1]) / (F)2;
        }
        const F mu_num = (((((F)0 + sp[0][1]) + sp[1][1]) + sp[0][2]) + sp[1][2]) / (F)4;
        o << "\n";
        o << "       :        -          +\n";
        o << "sigma0  : " << fmt_fixed(sigma0[0], 6) << " | " << fmt_fixed(sigma0[1], 6) << "\n";
        o << "alpha0  : " << fmt_sci(alpha0[0], 5) << " | " << fmt_sci(alpha0[1], 4) << "\n";
        o << "beta0   : " << fmt_display(beta0[0]) << " | " << fmt_display(beta0[1]) << "\n";
        o << "lambda0 : " << fmt_fixed(lambda0[0], 4) << " | " << fmt_fixed(lambda0[1], 4) << "\n";
        o << "mu0     : " << fmt_fixed(mu0[0], 4) << " | " << fmt_fixed(mu0[1], 5) << "\n";
        o << "mu/lamb : " << fmt_fixed(mu0[0] / lambda0[0], 5) << " | " << fmt_fixed(mu0[1] / lambda0[1], 5) << "\n";
        o << "mu (num): " << fmt_fixed(mu_num, 4) << "\n";
        o << "rapport : " << fmt_fixed(mu_num / mu_th, 6) << "\n";
        o << "mu (th) : " << fmt_fixed(mu_th, 4) << "\n";
    }
    // resfin.rs:101-194
    {
        auto ipow = [](F a, int n) {
            F r = 1;
            for (;;) {
                if (n & 1) r *= a;
                n /= 2;
                if (!n) break;
                a *= a;
            }
            return r;
        };
        const F mre = c.m_z0 / c.e_total;
        const F gre = c.g_z0 * c.m_z0 / (c.e_total * c.e_total);
        const F x = (F)1 - mre * mre;
        const F den = x * x + gre * gre;
        const F sdz_re = x / den, sdz_im = -gre / den;
        const F del = ((F)1 - c.bcut) / (F)2;
        const F eps = (F)2 * c.e_min / c.e_total;
        const F bra = c.m_z0 / ((F)3 * (F)6 * (F)std::pow((double)PI, 3.0) * (F)16 * (F)120);
        const F sig = (F)12 * PI / (c.m_z0 * c.m_z0) * c.br * c.g_z0 * bra / (c.e_total * c.e_total) *
                      ipow(c.e_total / c.m_z0, 8) * (sdz_re * sdz_re + sdz_im * sdz_im) * c.conv;
        const F e4 = ipow(eps, 4), d2 = del * del, d3 = ipow(del, 3);
        const F f1 = (F)1 - (F)15 * e4 - (F)9 / (F)7 * ((F)1 - (F)70 * e4) * d2 + (F)6 / (F)7 * ((F)1 + (F)70 * e4) * d3;
        const F g1 = (F)1 - (F)30 * e4 - (F)9 / (F)7 * ((F)1 - (F)70 * e4) * del - (F)90 * e4 * d2 -
                     (F)1 / (F)7 * ((F)1 - (F)420 * e4) * d3;
        const F g2 = (F)1 - (F)25 * e4 - (F)6 / (F)7 * ((F)1 - (F)70 * e4) * del - (F)3 / (F)7 * ((F)1 + (F)210 * e4) * d2 -
                     (F)8 / (F)21 * ((F)1 - (F)52.5 * e4) * d3;
        const F g3 = (F)1 - (F)195 / (F)11 * e4 - (F)18 / (F)77 * ((F)1 - (F)7 * e4) * del -
                     (F)9 / (F)11 * ((F)9 / (F)7 - (F)70 * e4) * d2 - (F)8 / (F)11 * ((F)1 - (F)105 / (F)11 * e4) * d3;
        const F sc3 = ipow(c.sincut, 3);
        const F ff = f1 * ((F)1 - sc3);
        const F gg = g1 - (F)27 / (F)16 * g2 * c.sincut + (F)11 / (F)16 * g3 * sc3;
        const F sig_p = sig * (ff + (F)2 * gg);
        const F sig_m = sig_p + (F)2 * sig * gg;
        auto sum2 = [&](int i) { return ((F)0 + sp[0][i]) + sp[1][i]; };
        auto incr = [&](int i) {
            const F a = sp[0][i] * rel[i], b = sp[1][i] * rel[i];
            return std::sqrt(a * a + b * b) / std::fabs(sum2(i));
        };
        const F mc_p = sum2(1) / (F)4, mc_m = sum2(2) / (F)4;
        const F inc_p = incr(1), inc_m = incr(2);
        o << "\n";
        o << "s (pb) :   Sig_cut_Th    Sig_Th      Rapport\n";
        o << "       :   Sig_Num\n";
        o << "       :   Ecart_relatif  Incertitude\n";
        o << "\n";
        o << "s+(pb) : " << fmt_fixed(sig_p, 5) << " | " << fmt_fixed(sig * (F)3, 5) << " | "
          << fmt_fixed(sig_p / ((F)3 * sig), 6) << "\n";
        o << "       : " << fmt_fixed(mc_p, 5) << "\n";
        o << "       : " << fmt_fixed(mc_p / sig_p - (F)1, 6) << " | " << fmt_fixed(inc_p, 8) << " | "
          << fmt_fixed((mc_p / sig_p - (F)1) / inc_p, 2) << "\n";
        o << "\n";
        o << "s-(pb) : " << fmt_fixed(sig_m, 5) << " | " << fmt_fixed(sig * (F)5, 4) << " | "
          << fmt_fixed(sig_m / ((F)5 * sig), 6) << "\n";
        o << "       : " << fmt_fixed(mc_m, 5) << "\n";
        o << "       : " << fmt_fixed(mc_m / sig_m - (F)1, 6) << " | " << fmt_fixed(inc_m, 9) << " | "
          << fmt_fixed((mc_m / sig_m - (F)1) / inc_m, 2) << "\n";
        o << "\n";
    }
    return o.str();
}

template <class F> std::string res_data_t(const tp3_config& raw, const tp3_final& fin) {  // output.rs:64-142
    const Cfg<F> c(raw);
    const int SIG = Lim<F>::digits10 - 1;  // output.rs:26
    std::ostringstream o;
    auto key = [&](const char* k) -> std::ostringstream& {
        o << " " << ljust(k, 31) << ": ";
        return o;
    };
    auto num = [&](const char* k, F v) { key(k) << fmt_engineering(v, SIG) << "\n"; };
    key("Nombre d'evenements") << c.n << "\n";
    key("... apres coupure") << fin.selected_events << "\n";
    num("energie dans le CdM      (GeV)", c.e_total);
    num("coupure / cos(photon,faisceau)", c.acut);
    num("coupure / cos(photon,photon)", c.bcut);
    num("coupure / sin(normale,faisceau)", c.sincut);
    num("coupure sur l'energie    (GeV)", c.e_min);
    num("1/(constante de structure fine)", (F)1 / c.alpha);
    num("1/(structure fine au pic)", (F)1 / c.alpha_z);
    num("facteur de conversion GeV-2/pb", c.conv);
    num("Masse du Z0              (GeV)", c.m_z0);
    num("Largeur du Z0            (GeV)", c.g_z0);
    num("Sinus^2 Theta Weinberg", c.s2w);
    num("Taux de branchement Z--->e+e-", c.br);
    num("Beta plus", c.beta_p);
    num("Beta moins", c.beta_m);
    o << " ---------------------------------------------\n";
    const F sigma = (F)fin.sigma, prec = (F)fin.prec;
    num("Section Efficace          (pb)", sigma);
    num("Ecart-Type                (pb)", sigma * prec);
    num("Precision Relative", prec);
    o << " ---------------------------------------------\n";
    num("Beta minimum", (F)fin.beta_min);
    num("Stat. Significance  B+(pb-1/2)", (F)fin.ss_p);
    num("Incert. Stat. Sign. B+(pb-1/2)", (F)fin.ss_p * (F)fin.inc_ss_p);
    num("Stat. Significance  B-(pb-1/2)", (F)fin.ss_m);
    num("Incert. Stat. Sign. B-(pb-1/2)", (F)fin.ss_m * (F)fin.inc_ss_m);
    o << "\n";
    const int dec = SIG - 1 < 7 ? SIG - 1 : 7;
    const size_t w = (size_t)dec + 8;
    for (int s = 0; s < 2; ++s) {
        for (int i = 0; i < 5; ++i) {
            const F v = (F)fin.spm2[s][i], r = (F)fin.vars[s][i];
            o << rjust(std::to_string(s + 1), 3) << rjust(std::to_string(i + 1), 3) << rjust(fmt_sci(v, dec), w)
              << rjust(fmt_sci((F)(std::fabs(v) * r), dec), w) << rjust(fmt_sci(r, dec), w) << "\n";
        }
        o << "\n";
    }
    for (int i = 0; i < 5; ++i) {
        const F a = (F)fin.spm2[0][i], b = (F)fin.spm2[1][i], ra = (F)fin.vars[0][i], rb = (F)fin.vars[1][i];
        const F t1 = ((F)0 + a) + b;
        const F x = a * ra, y = b * rb;
        const F t2 = std::sqrt(x * x + y * y);
        o << "   " << rjust(std::to_string(i + 1), 3) << rjust(fmt_sci(t1 / (F)4, dec), w) << rjust(fmt_sci(t2 / (F)4, dec), w)
          << rjust(fmt_sci(t2 / std::fabs(t1), dec), w) << "\n";
    }
    return o.str();
}

size_t emit(const std::string& s, char* buf, size_t cap) {
    if (buf && cap) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

bool parse_bool_item(const std::string& s, bool& ok) {  // config.rs:187-195
    // lower-cased only for the FORTRAN forms; anything else goes to Rust's bool parser, which takes exactly "true" / "false"
    std::string low = s;
    for (auto& ch : low) ch = (char)std::tolower((unsigned char)ch);
    ok = true;
    if (low == ".true.") return true;
    if (low == ".false.") return false;
    if (s == "true") return true;
    if (s == "false") return false;
    ok = false;
    return false;
}

template <class T> bool parse_num(const std::string& s, T& out) {
    // Rust's FromStr rejects trailing garbage and leading whitespace; accepts "0.e0", "91.187e0", "+1"
    if (s.empty() || std::isspace((unsigned char)s[0])) return false;
    char* end = nullptr;
    errno = 0;
    if constexpr (std::is_floating_point<T>::value) {
        // what strtod takes and Rust's f64::from_str does not: hexadecimal floats and nan(...) payloads
        for (char ch : s)
            if (ch == 'x' || ch == 'X' || ch == '(') return false;
        if constexpr (std::is_same<T, float>::value) out = std::strtof(s.c_str(), &end);
        else out = std::strtod(s.c_str(), &end);
        // (out-of-range decimals parse to inf / 0 in Rust as well: ERANGE is not an error here)
    } else if constexpr (std::is_same<T, uint64_t>::value) {
        if (s[0] == '-') return false;
        out = std::strtoull(s.c_str(), &end, 10);
        if (errno == ERANGE) return false;  // Rust: "number too large to fit in target type"
    } else {
        const long v = std::strtol(s.c_str(), &end, 10);
        if (errno == ERANGE || v < INT32_MIN || v > INT32_MAX) return false;
        out = (T)v;
    }
    return end && *end == 0;
}

}  // namespace

extern "C" {

int tp3_config_parse(const char* text, uint32_t flags, tp3_config* out, char* err, size_t err_cap) {
    if (!text || !out) return TP3_E_INVALID;
    auto fail = [&](const std::string& m) {
        emit(m, err, err_cap);
        return TP3_E_CONFIG;
    };
    std::vector<std::string> items;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        if (ls >> tok) items.push_back(tok);
    }
    static const char* names[18] = {"num_events", "e_total", "beam_photons_cut", "photon_photon_cut", "e_min",
                                    "beam_photon_plane_cut", "alpha", "alpha_z", "gev2_to_picobarn", "m_z0", "g_z0",
                                    "sin2_weinberg", "branching_ep_em", "beta_plus", "beta_moins", "num_bins", "impr", "plot"};
    if (items.size() < 18) return fail(std::string("missing configuration of ") + names[items.size()]);
    std::memset(out, 0, sizeof *out);
    if (!parse_num<uint64_t>(items[0], out->num_events)) return fail("could not parse configuration of num_events");
    double* fields[14] = {&out->e_total, &out->beam_photons_cut, &out->photon_photon_cut, &out->e_min,
                          &out->beam_photon_plane_cut, &out->alpha, &out->alpha_z, &out->gev2_to_picobarn, &out->m_z0,
                          &out->g_z0, &out->sin2_weinberg, &out->branching_ep_em, &out->beta_plus, &out->beta_minus};
    for (int i = 0; i < 14; ++i) {
        bool ok;
        if (flags & TP3_F32) {  // decimal text -> f32 directly (no double rounding)
            float v;
            ok = parse_num<float>(items[1 + i], v);
            *fields[i] = v;
        } else {
            ok = parse_num<double>(items[1 + i], *fields[i]);
        }
        if (!ok) return fail(std::string("could not parse configuration of ") + names[1 + i]);
    }
    if (!parse_num<int32_t>(items[15], out->num_bins)) return fail("could not parse configuration of num_bins");
    bool ok;
    out->impr = parse_bool_item(items[16], ok);
    if (!ok) return fail("could not parse configuration of impr");
    out->plot = parse_bool_item(items[17], ok);
    if (!ok) return fail("could not parse configuration of plot");
    // config.rs:111-124
    if (out->num_events == 0) return fail("Please simulate at least one event");
    if (out->plot) return fail("Plotting is not supported by this version");
    if (out->impr)
        return fail("Individual result printing is not supported. This debugging feature has a run-time performance cost "
                    "even when unused. It should be implemented at compile-time instead.");
    return TP3_OK;
}

int tp3_params_from_config(const tp3_config* cfg, uint32_t flags, uint32_t kernel, tp3_params* out) {
    if (!cfg || !out) return TP3_E_INVALID;
    if (flags & TP3_F32) make_params<float>(*cfg, flags, kernel, *out);
    else make_params<double>(*cfg, flags, kernel, *out);
    return TP3_OK;
}

int tp3_merge(tp3_acc* into, const tp3_acc* other, uint32_t flags) {
    if (!into || !other) return TP3_E_INVALID;
    into->selected_events += other->selected_events;
    if (flags & TP3_F32) {
        for (int k = 0; k < 5; ++k) into->spm2[k] = (float)into->spm2[k] + (float)other->spm2[k];
        for (int k = 0; k < 5; ++k) into->vars[k] = (float)into->vars[k] + (float)other->vars[k];
        into->sigma = (float)into->sigma + (float)other->sigma;
        into->variance = (float)into->variance + (float)other->variance;
    } else {
        for (int k = 0; k < 5; ++k) into->spm2[k] += other->spm2[k];
        for (int k = 0; k < 5; ++k) into->vars[k] += other->vars[k];
        into->sigma += other->sigma;
        into->variance += other->variance;
    }
    return TP3_OK;
}

int tp3_fold_batches(const tp3_acc* per_batch, uint64_t n, uint32_t flags, tp3_acc* out) {
    if (!per_batch || !out || n == 0) return TP3_E_INVALID;
    *out = per_batch[0];  // the fold starts FROM the first batch (sequential.rs:24-26)
    for (uint64_t b = 1; b < n; ++b) tp3_merge(out, &per_batch[b], flags);
    return TP3_OK;
}

int tp3_finalize(const tp3_config* cfg, uint32_t flags, const tp3_acc* merged, tp3_final* out) {
    if (!cfg || !merged || !out) return TP3_E_INVALID;
    if (flags & TP3_F32) finalize_t<float>(*cfg, *merged, *out);
    else finalize_t<double>(*cfg, *merged, *out);
    return TP3_OK;
}

size_t tp3_format_res_data(const tp3_config* cfg, uint32_t flags, const tp3_final* fin, char* buf, size_t cap) {
    if (!cfg || !fin) return 0;
    return emit((flags & TP3_F32) ? res_data_t<float>(*cfg, *fin) : res_data_t<double>(*cfg, *fin), buf, cap);
}

size_t tp3_format_stdout(const tp3_config* cfg, uint32_t flags, const tp3_final* fin, char* buf, size_t cap) {
    if (!cfg || !fin) return 0;
    return emit((flags & TP3_F32) ? stdout_t<float>(*cfg, *fin) : stdout_t<double>(*cfg, *fin), buf, cap);
}

int tp3_run(const char* valeurs_path, const char* out_dir, uint32_t flags, uint32_t kernel, int n_dev, char* stdout_buf,
            size_t stdout_cap, double* elapsed) {
    return tp3_run_stages(valeurs_path, out_dir, flags, kernel, n_dev, stdout_buf, stdout_cap, elapsed, nullptr);
}

int tp3_run_stages(const char* valeurs_path, const char* out_dir, uint32_t flags, uint32_t kernel, int n_dev, char* stdout_buf,
                   size_t stdout_cap, double* elapsed, double* stages) {
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    };
    if (stages) std::fill(stages, stages + TP3_RUN_STAGES, 0.0);
    std::ifstream in(valeurs_path ? valeurs_path : "valeurs");
    if (!in) {
        emit("failed to load the configuration", stdout_buf, stdout_cap);
        return TP3_E_IO;
    }
    std::stringstream ss;
    ss << in.rdbuf();
    tp3_config cfg;
    char err[512] = {0};
    int rc = tp3_config_parse(ss.str().c_str(), flags, &cfg, err, sizeof err);
    if (rc) {
        emit(std::string("failed to load the configuration: ") + err, stdout_buf, stdout_cap);
        return rc;
    }
    if (stages) stages[0] = since(t_begin);
    // The reference starts its clock after configuration I/O (main.rs:83-85) ...
    const auto t0 = std::chrono::steady_clock::now();
    tp3_params params;
    tp3_params_from_config(&cfg, flags, kernel, &params);
    tp3_ctx* ctx = nullptr;
    rc = tp3_create(&params, n_dev < 1 ? 1 : n_dev, nullptr, &ctx);
    if (rc) {
        emit(tp3_last_error(nullptr), stdout_buf, stdout_cap);
        return rc;
    }
    if (stages) stages[1] = since(t0);
    const auto t_sim = std::chrono::steady_clock::now();
    // scheduling: ceil(N / 10000) batches, the last one possibly short (multi_threading.rs:25,47);
    // left fold in batch order (sequential.rs:24-36)
    const uint64_t nb = (cfg.num_events + TP3_EVENT_BATCH_SIZE - 1) / TP3_EVENT_BATCH_SIZE;
    const uint32_t last = (uint32_t)(cfg.num_events - (nb - 1) * TP3_EVENT_BATCH_SIZE);
    // (tp3_simulate_merged: the same left fold, done on the device while the batches are simulated)
    // A run too small to give every resident warp a batch (the default 1e7 events are 1000 batches for 2368 warp slots) is cut
    // into parts of batches, added in part order: this program issues one launch, so the choice is reproducible.
    tp3_set_option(ctx, "batch_parts", 0);
    tp3_acc total;
    rc = tp3_simulate_merged(ctx, 0, nb, last, &total);
    if (rc) {
        emit(tp3_last_error(ctx), stdout_buf, stdout_cap);
        tp3_destroy(ctx);
        return rc;
    }
    if (stages) stages[2] = since(t_sim);
    const auto t_fin = std::chrono::steady_clock::now();
    tp3_final fin;
    tp3_finalize(&cfg, flags, &total, &fin);
    // ... and stops it before output (main.rs:138)
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stages) stages[3] = since(t_fin);
    const auto t_out = std::chrono::steady_clock::now();
    tp3_destroy(ctx);
    if (stages) stages[5] = since(t_out);
    if (elapsed) *elapsed = secs;

    std::string so(tp3_format_stdout(&cfg, flags, &fin, nullptr, 0) + 1, '\0');
    tp3_format_stdout(&cfg, flags, &fin, so.data(), so.size());
    so.pop_back();
    emit(so, stdout_buf, stdout_cap);

    const std::string dir = (out_dir && *out_dir) ? std::string(out_dir) + "/" : std::string();
    std::string rd(tp3_format_res_data(&cfg, flags, &fin, nullptr, 0) + 1, '\0');
    tp3_format_res_data(&cfg, flags, &fin, rd.data(), rd.size());
    rd.pop_back();
    // timestamp "[day]-[month repr:short]-[year repr:last_two] [hour]:[minute]:[second]" UTC (output.rs:36-40)
    char stamp[64];
    {
        std::time_t now = std::time(nullptr);
        std::tm tm;
        gmtime_r(&now, &tm);
        std::strftime(stamp, sizeof stamp, "%d-%b-%y %H:%M:%S", &tm);
    }
    const bool f32 = flags & TP3_F32;
    auto eng = [&](double v) { return f32 ? fmt_engineering((float)v, 5) : fmt_engineering(v, 14); };
    auto disp = [&](double v) { return f32 ? fmt_display((float)v) : fmt_display(v); };
    {
        std::ofstream f(dir + "res.data");
        if (!f) return TP3_E_IO;
        f << rd;
    }
    {  // output.rs:43-60
        std::ofstream f(dir + "res.times");
        if (!f) return TP3_E_IO;
        f << " " << stamp << "\n";
        f << " ---------------------------------------------\n";
        f << " " << ljust("Temps ecoule", 31) << ": ???\n";
        f << " " << ljust("Temps ecoule utilisateur", 31) << ": " << eng(secs) << "\n";
        f << " " << ljust("Temps ecoule systeme", 31) << ": ???\n";
        f << " " << ljust("Temps ecoule par evenement", 31) << ": " << eng(secs / (double)cfg.num_events) << "\n";
    }
    {  // output.rs:148-173
        std::ofstream f(dir + "pil.mc", std::ios::app);
        if (!f) return TP3_E_IO;
        f << stamp << "\n";
        auto line = [&](auto zero) {  // in the run's Float, like every other number the reference prints
            using F = decltype(zero);
            auto col = [&](int i) { return (F)((F)fin.spm2[0][i] + (F)fin.spm2[1][i]); };
            const F bp = (F)cfg.beta_plus, bm = (F)cfg.beta_minus;
            const F r1 = col(0), r2 = col(1) * (bp * bp), r3 = col(2) * (bm * bm), r4 = col(3) * bp;
            f << disp(cfg.e_total) << " " << disp(r1 / (F)4) << " " << disp(r2 / (F)4) << " " << disp(r3 / (F)4) << " " << disp(r4 / (F)4)
              << " " << disp((r1 + r2 + r3 + r4) / (F)4) << " " << disp(fin.sigma) << "\n";
        };
        if (f32) line(0.0f);
        else line(0.0);
    }
    return TP3_OK;
}

}  // extern "C"

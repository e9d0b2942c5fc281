// oracle/oracle_text.hpp — CPU ORACLE, text surfaces. TEST INFRASTRUCTURE ONLY.
//
// Restates the reference's stdout reports and `res.data` writer so that the
// oracle can be pinned against the golden files of /root/reference/reference/.
// Follows config.rs:131-155 (echo), resfin.rs:66-194 (eric, fawzi),
// output.rs:64-142,180-269 (res.data, write_engineering) and Rust's float
// formatting rules.
#pragma once

#include <charconv>

#include "oracle.hpp"

namespace oracle {

// Rust `{}` on a float: shortest round-trip digits, never scientific.
template <class F> std::string rust_display(F x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    // shortest round-trip digits d.ddd e±x, then laid out positionally, zero-padded like Rust
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, res.ptr);
    bool neg = s[0] == '-';
    if (neg) s.erase(0, 1);
    size_t e = s.find('e');
    int ex = atoi(s.c_str() + e + 1);
    std::string digits;
    for (char c : s.substr(0, e))
        if (c != '.') digits += c;
    std::string o;
    int nd = (int)digits.size();
    if (ex >= nd - 1) {
        o = digits + std::string(ex - (nd - 1), '0');
    } else if (ex >= 0) {
        o = digits.substr(0, ex + 1) + "." + digits.substr(ex + 1);
    } else {
        o = "0." + std::string(-ex - 1, '0') + digits;
    }
    if (digits == "0") o = "0";
    return (neg ? "-" : "") + o;
}
// Rust `{:.p}`
template <class F> std::string rust_fixed(F x, int p) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    char buf[512];
    snprintf(buf, sizeof buf, "%.*f", p, (double)x);
    return buf;
}
// Rust `{:.pe}`: mantissa as printf, exponent without sign padding or leading zeros
template <class F> std::string rust_exp(F x, int p) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    char buf[128];
    snprintf(buf, sizeof buf, "%.*e", p, (double)x);
    std::string s(buf);
    size_t e = s.find('e');
    int ex = atoi(s.c_str() + e + 1);
    return s.substr(0, e) + "e" + std::to_string(ex);
}
inline std::string pad_left(const std::string& s, size_t w) {
    return s.size() >= w ? s : std::string(w - s.size(), ' ') + s;
}
inline std::string pad_right(const std::string& s, size_t w) {
    return s.size() >= w ? s : s + std::string(w - s.size(), ' ');
}

// output.rs:230-269
template <class F> std::string write_engineering(F x, int sig_digits) {
    int precision = sig_digits - 1;
    if (x == (F)0) return "0";
    F log_x = std::log10(std::fabs(x));
    if (log_x >= (F)-3 && log_x <= (F)sig_digits) {
        precision = precision - (int)std::trunc(log_x);
        if (log_x < (F)0) precision += 1;
        if (precision < 0) precision = 0;
        std::string s = rust_fixed(x, precision);
        if (s.find('.') != std::string::npos) {
            while (!s.empty() && s.back() == '0') s.pop_back();
            if (!s.empty() && s.back() == '.') s.pop_back();
        }
        return s;
    }
    return rust_exp(x, precision);
}

// config.rs:131-155
template <class F> std::string config_echo(const Config<F>& c) {
    std::string o;
    auto line = [&](const char* k, const std::string& v) { o += std::string(k) + v + "\n"; };
    line("ITOT           : ", std::to_string(c.num_events));
    line("ETOT           : ", rust_display(c.e_total));
    line("oCutpar.ACUT   : ", rust_display(c.beam_photons_cut));
    line("oCutpar.BCUT   : ", rust_display(c.photon_photon_cut));
    line("oCutpar.EMIN   : ", rust_display(c.e_min));
    line("oCutpar.SINCUT : ", rust_display(c.beam_photon_plane_cut));
    line("ALPHA          : ", rust_display(c.alpha));
    line("ALPHAZ         : ", rust_display(c.alpha_z));
    line("CONVERS        : ", rust_display(c.gev2_to_picobarn));
    line("oParam.MZ0     : ", rust_display(c.m_z0));
    line("oParam.GZ0     : ", rust_display(c.g_z0));
    line("SIN2W          : ", rust_display(c.sin2_weinberg));
    line("BREPEM         : ", rust_display(c.branching_ep_em));
    line("BETAPLUS       : ", rust_display(c.beta_plus));
    line("BETAMOINS      : ", rust_display(c.beta_minus));
    line("NBIN           : ", std::to_string(c.num_bins));
    line("oParam.IMPR    : ", c.impr ? "true" : "false");
    line("PLOT           : ", c.plot ? "true" : "false");
    return o;
}

// resfin.rs:66-97
template <class F> std::string eric(const Config<F>& cfg, const FinalResults<F>& r) {
    const F PI = K<F>::PI;
    F mu_th = cfg.branching_ep_em * cfg.gev2_to_picobarn /
              ((F)8 * (F)9 * (F)5 * (F)std::pow((double)PI, 2.0) * cfg.m_z0 * cfg.g_z0);
    F sigma0[2], alpha0[2], beta0[2], lambda0[2], mu0[2];
    for (int sp = 0; sp < 2; ++sp) {
        sigma0[sp] = r.spm2[sp][0] / (F)2;
        alpha0[sp] = r.spm2[sp][4] / (F)2;
        beta0[sp] = -r.spm2[sp][3] / (F)2;
        lambda0[sp] = (r.spm2[sp][2] - r.spm2[sp][1]) / (F)2;
        mu0[sp] = (r.spm2[sp][2] + r.spm2[sp][1]) / (F)2;
    }
    F mu_num = (((((F)0 + r.spm2[0][1]) + r.spm2[1][1]) + r.spm2[0][2]) + r.spm2[1][2]) / (F)4;
    std::string o;
    o += "\n";
    o += "       :        -          +\n";
    o += "sigma0  : " + rust_fixed(sigma0[0], 6) + " | " + rust_fixed(sigma0[1], 6) + "\n";
    o += "alpha0  : " + rust_exp(alpha0[0], 5) + " | " + rust_exp(alpha0[1], 4) + "\n";
    o += "beta0   : " + rust_display(beta0[0]) + " | " + rust_display(beta0[1]) + "\n";
    o += "lambda0 : " + rust_fixed(lambda0[0], 4) + " | " + rust_fixed(lambda0[1], 4) + "\n";
    o += "mu0     : " + rust_fixed(mu0[0], 4) + " | " + rust_fixed(mu0[1], 5) + "\n";
    o += "mu/lamb : " + rust_fixed(mu0[0] / lambda0[0], 5) + " | " + rust_fixed(mu0[1] / lambda0[1], 5) + "\n";
    o += "mu (num): " + rust_fixed(mu_num, 4) + "\n";
    o += "rapport : " + rust_fixed(mu_num / mu_th, 6) + "\n";
    o += "mu (th) : " + rust_fixed(mu_th, 4) + "\n";
    return o;
}

template <class F> inline F powi_f(F a, int n) {  // compiler-rt __powi?f2 order of operations
    F r = 1;
    while (true) {
        if (n & 1) r *= a;
        n /= 2;
        if (n == 0) break;
        a *= a;
    }
    return r;
}

// resfin.rs:101-194
template <class F> std::string fawzi(const Config<F>& cfg, const FinalResults<F>& r) {
    const F PI = K<F>::PI;
    F mre = cfg.m_z0 / cfg.e_total;
    F gre = cfg.g_z0 * cfg.m_z0 / (cfg.e_total * cfg.e_total);
    F x = (F)1 - mre * mre;
    F sdz_den = x * x + gre * gre;
    Cx<F> sdz = Cx<F>{x, -gre} / sdz_den;
    F del = ((F)1 - cfg.photon_photon_cut) / (F)2;
    F eps = (F)2 * cfg.e_min / cfg.e_total;
    F bra = cfg.m_z0 / ((F)3 * (F)6 * (F)std::pow((double)PI, 3.0) * (F)16 * (F)120);
    F sig = (F)12 * PI / (cfg.m_z0 * cfg.m_z0) * cfg.branching_ep_em * cfg.g_z0 * bra /
            (cfg.e_total * cfg.e_total) * powi_f(cfg.e_total / cfg.m_z0, 8) * norm_sqr(sdz) *
            cfg.gev2_to_picobarn;
    F eps_4 = powi_f(eps, 4);
    F del_2 = del * del;
    F del_3 = powi_f(del, 3);
    F f1 = (F)1 - (F)15 * eps_4 - (F)9 / (F)7 * ((F)1 - (F)70 * eps_4) * del_2 +
           (F)6 / (F)7 * ((F)1 + (F)70 * eps_4) * del_3;
    F g1 = (F)1 - (F)30 * eps_4 - (F)9 / (F)7 * ((F)1 - (F)70 * eps_4) * del - (F)90 * eps_4 * del_2 -
           (F)1 / (F)7 * ((F)1 - (F)420 * eps_4) * del_3;
    F g2 = (F)1 - (F)25 * eps_4 - (F)6 / (F)7 * ((F)1 - (F)70 * eps_4) * del -
           (F)3 / (F)7 * ((F)1 + (F)210 * eps_4) * del_2 - (F)8 / (F)21 * ((F)1 - (F)52.5 * eps_4) * del_3;
    F g3 = (F)1 - (F)195 / (F)11 * eps_4 - (F)18 / (F)77 * ((F)1 - (F)7 * eps_4) * del -
           (F)9 / (F)11 * ((F)9 / (F)7 - (F)70 * eps_4) * del_2 -
           (F)8 / (F)11 * ((F)1 - (F)105 / (F)11 * eps_4) * del_3;
    F cut3 = powi_f(cfg.beam_photon_plane_cut, 3);
    F ff = f1 * ((F)1 - cut3);
    F gg = g1 - (F)27 / (F)16 * g2 * cfg.beam_photon_plane_cut + (F)11 / (F)16 * g3 * cut3;
    F sig_p = sig * (ff + (F)2 * gg);
    F sig_m = sig_p + (F)2 * sig * gg;
    auto colsum = [&](int k) { return ((F)0 + r.spm2[0][k]) + r.spm2[1][k]; };
    F mc_p = colsum(1) / (F)4;
    F mc_m = colsum(2) / (F)4;
    auto incr = [&](int k) {
        F a = r.spm2[0][k] * r.vars[0][k], b = r.spm2[1][k] * r.vars[1][k];
        return std::sqrt(a * a + b * b) / std::fabs(colsum(k));
    };
    F incr_p = incr(1), incr_m = incr(2);
    std::string o;
    o += "\n";
    o += "s (pb) :   Sig_cut_Th    Sig_Th      Rapport\n";
    o += "       :   Sig_Num\n";
    o += "       :   Ecart_relatif  Incertitude\n";
    o += "\n";
    o += "s+(pb) : " + rust_fixed(sig_p, 5) + " | " + rust_fixed(sig * (F)3, 5) + " | " +
         rust_fixed(sig_p / ((F)3 * sig), 6) + "\n";
    o += "       : " + rust_fixed(mc_p, 5) + "\n";
    o += "       : " + rust_fixed(mc_p / sig_p - (F)1, 6) + " | " + rust_fixed(incr_p, 8) + " | " +
         rust_fixed((mc_p / sig_p - (F)1) / incr_p, 2) + "\n";
    o += "\n";
    o += "s-(pb) : " + rust_fixed(sig_m, 5) + " | " + rust_fixed(sig * (F)5, 4) + " | " +
         rust_fixed(sig_m / ((F)5 * sig), 6) + "\n";
    o += "       : " + rust_fixed(mc_m, 5) + "\n";
    o += "       : " + rust_fixed(mc_m / sig_m - (F)1, 6) + " | " + rust_fixed(incr_m, 9) + " | " +
         rust_fixed((mc_m / sig_m - (F)1) / incr_m, 2) + "\n";
    o += "\n";
    return o;
}

// output.rs:64-142
template <class F> std::string res_data(const Config<F>& cfg, const FinalResults<F>& r) {
    const int SIG = K<F>::DIGITS - 1;
    std::string o;
    auto kv_f = [&](const char* k, F v) { o += " " + pad_right(k, 31) + ": " + write_engineering(v, SIG) + "\n"; };
    auto kv_u = [&](const char* k, uint64_t v) { o += " " + pad_right(k, 31) + ": " + std::to_string(v) + "\n"; };
    auto sep = [&]() { o += " ---------------------------------------------\n"; };
    kv_u("Nombre d'evenements", cfg.num_events);
    kv_u("... apres coupure", r.selected_events);
    kv_f("energie dans le CdM      (GeV)", cfg.e_total);
    kv_f("coupure / cos(photon,faisceau)", cfg.beam_photons_cut);
    kv_f("coupure / cos(photon,photon)", cfg.photon_photon_cut);
    kv_f("coupure / sin(normale,faisceau)", cfg.beam_photon_plane_cut);
    kv_f("coupure sur l'energie    (GeV)", cfg.e_min);
    kv_f("1/(constante de structure fine)", (F)1 / cfg.alpha);
    kv_f("1/(structure fine au pic)", (F)1 / cfg.alpha_z);
    kv_f("facteur de conversion GeV-2/pb", cfg.gev2_to_picobarn);
    kv_f("Masse du Z0              (GeV)", cfg.m_z0);
    kv_f("Largeur du Z0            (GeV)", cfg.g_z0);
    kv_f("Sinus^2 Theta Weinberg", cfg.sin2_weinberg);
    kv_f("Taux de branchement Z--->e+e-", cfg.branching_ep_em);
    kv_f("Beta plus", cfg.beta_plus);
    kv_f("Beta moins", cfg.beta_minus);
    sep();
    kv_f("Section Efficace          (pb)", r.sigma);
    kv_f("Ecart-Type                (pb)", r.sigma * r.prec);
    kv_f("Precision Relative", r.prec);
    sep();
    kv_f("Beta minimum", r.beta_min);
    kv_f("Stat. Significance  B+(pb-1/2)", r.ss_p);
    kv_f("Incert. Stat. Sign. B+(pb-1/2)", r.ss_p * r.inc_ss_p);
    kv_f("Stat. Significance  B-(pb-1/2)", r.ss_m);
    kv_f("Incert. Stat. Sign. B-(pb-1/2)", r.ss_m * r.inc_ss_m);
    o += "\n";
    int decimals = SIG - 1 < 7 ? SIG - 1 : 7;
    size_t width = decimals + 8;
    for (int sp = 0; sp < 2; ++sp) {
        for (int k = 0; k < 5; ++k) {
            o += pad_left(std::to_string(sp + 1), 3) + pad_left(std::to_string(k + 1), 3) +
                 pad_left(rust_exp(r.spm2[sp][k], decimals), width) +
                 pad_left(rust_exp((F)(std::fabs(r.spm2[sp][k]) * r.vars[sp][k]), decimals), width) +
                 pad_left(rust_exp(r.vars[sp][k], decimals), width) + "\n";
        }
        o += "\n";
    }
    for (int k = 0; k < 5; ++k) {
        F tmp1 = ((F)0 + r.spm2[0][k]) + r.spm2[1][k];
        F a = r.spm2[0][k] * r.vars[0][k], b = r.spm2[1][k] * r.vars[1][k];
        F tmp2 = std::sqrt(a * a + b * b);
        o += "   " + pad_left(std::to_string(k + 1), 3) + pad_left(rust_exp(tmp1 / (F)4, decimals), width) +
             pad_left(rust_exp(tmp2 / (F)4, decimals), width) +
             pad_left(rust_exp(tmp2 / std::fabs(tmp1), decimals), width) + "\n";
    }
    return o;
}

}  // namespace oracle

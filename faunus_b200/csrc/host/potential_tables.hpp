// Configuration-time lowering of the reference's pair-potential JSON schema to flat tables for the
// device (fb_config). Nothing here runs per move: mixing matrices, the S(q) spline of the Coulomb
// scheme, and (for `nonbonded_splined`) the per-pair r² splines are built once on the host and
// uploaded. Mirrors: src/potentials.cpp:18-98 (PairMixer), :672-703 (LJ matrices), :956-977 (HS),
// :452-460 (plain Coulomb), :1628-1699 (CoulombGalore dispatch), :1202-1326 (functor assembly),
// :1453-1595 (splined assembly), src/tabulate.h:95-308 (Andrea generator),
// src/energy.cpp:1868-1933 (GroupCutoff), docs/_docs/energy.md:189-210 (S(q) definitions; the
// generator itself lives in the un-vendored coulombgalore dependency → "parity unpinned").
#pragma once
#include "space.hpp"
#include <cfloat>
#include <cstdint>
#include <functional>

namespace fb {

namespace term {
constexpr uint32_t COULOMB_SPLINED = 1, COULOMB_PLAIN = 2, LJ = 4, WCA = 8, HARDSPHERE = 16;
}
namespace potkind {
constexpr int COULOMB_LJ = 0, COULOMB_WCA = 1, PM = 2, PMWCA = 3, FUNCTOR = 4, SPLINED = 5;
}

/** Andrea piecewise-quintic table: knots ascending, 6 coefficients per interval */
struct SplineTable
{
    std::vector<double> knots;
    std::vector<double> coeffs;
    double xmin = 0, xmax = 0;
};

struct SplineOptions
{
    double utol = 1e-5;
    double ftol = -1;
    double umaxtol = -1;
    double fmaxtol = -1;
    double numdr = 0.0001; //!< step of the numerical derivatives
    int max_knots = 1200;
    int max_shrink = 100;
    double shrink = 0.9;
};

/**
 * Adaptive Andrea tabulation of f on ]xmin, xmax], walking from xmax down in steps that are uniform
 * in sqrt(x) and shrunk by 0.9 until an 11-point check meets the tolerance (src/tabulate.h:223-307).
 */
inline SplineTable tabulateAndrea(const std::function<double(double)>& f, double xmin, double xmax,
                                  const SplineOptions& opt)
{
    auto d1 = [&](double x) { return (f(x + opt.numdr * 0.5) - f(x - opt.numdr * 0.5)) / opt.numdr; };
    auto d2 = [&](double x) { return (d1(x + opt.numdr * 0.5) - d1(x - opt.numdr * 0.5)) / opt.numdr; };
    const double smin = std::sqrt(xmin);
    const double smax = std::sqrt(xmax);
    SplineTable table;
    table.xmin = smin * smin;
    table.xmax = smax * smax;
    std::vector<double> knots_desc = {smax * smax};
    std::vector<std::array<double, 6>> blocks_desc;
    double lower_limit = smin;
    double upper = smax;
    bool repulsive = false;
    int knot = 0;
    for (; knot < opt.max_knots; ++knot) {
        double step = upper - smin;
        double lower = upper;
        double zlow = 0;
        std::array<double, 6> block{};
        int trial = 0;
        for (; trial < opt.max_shrink; ++trial) {
            const double zupp = upper * upper;
            lower = std::max(upper - step, lower_limit);
            zlow = lower * lower;
            const double u0l = f(zlow), u1l = d1(zlow), u2l = d2(zlow);
            const double u0u = f(zupp), u1u = d1(zupp), u2u = d2(zupp);
            if (std::fabs(u0l) < 1e-9 && std::fabs(u1l) < 1e-9) {
                block = {0, 0, 0, 0, 0, 0};
            }
            else { // quintic Hermite through value/slope/curvature at both ends (tabulate.h:112-131)
                const double h = zupp - zlow, h2 = h * h, h3 = h2 * h;
                const double c0 = u0l, c1 = u1l, c2 = 0.5 * u2l;
                const double a = 6 * (u0u - c0 - c1 * h - c2 * h2) / h3;
                const double b = 2 * (u1u - c1 - 2 * c2 * h) / h2;
                const double c = (u2u - 2 * c2) / h;
                block = {c0, c1, c2, (10 * a - 12 * b + 3 * c) / 6, (-15 * a + 21 * b - 6 * c) / (6 * h),
                         (2 * a - 3 * b + c) / (2 * h2)};
            }
            bool ok = true;
            repulsive = false;
            const double dr = (upper - lower) / 10.0;
            for (int i = 0; i < 11 && ok; ++i) {
                const double s = lower + dr * static_cast<double>(i);
                const double x = s * s;
                const double dz = x - lower * lower;
                const double us =
                    block[0] + dz * (block[1] + dz * (block[2] + dz * (block[3] + dz * (block[4] + dz * block[5]))));
                const double fs = block[1] + dz * (2 * block[2] +
                                                   dz * (3 * block[3] + dz * (4 * block[4] + dz * (5 * block[5]))));
                if (std::fabs(us - f(x)) > opt.utol) {
                    ok = false;
                }
                else if (opt.ftol != -1 && std::fabs(fs - d1(x)) > opt.ftol) {
                    ok = false;
                }
                else {
                    if (opt.umaxtol != -1 && std::fabs(us) > opt.umaxtol) {
                        repulsive = true;
                    }
                    if (opt.fmaxtol != -1 && std::fabs(us) > opt.fmaxtol) {
                        repulsive = true;
                    }
                }
            }
            if (ok) {
                upper = lower;
                break;
            }
            repulsive = false;
            step *= opt.shrink;
        }
        if (trial >= opt.max_shrink) {
            throw std::runtime_error("Andrea spline: try to increase utol/ftol");
        }
        knots_desc.push_back(zlow);
        blocks_desc.push_back(block);
        if (repulsive) {
            lower_limit = lower;
            table.xmin = lower * lower;
        }
        if (lower <= lower_limit || repulsive) {
            break;
        }
    }
    if (knot >= opt.max_knots) {
        throw std::runtime_error("Andrea spline: try to increase utol/ftol");
    }
    table.knots.assign(knots_desc.rbegin(), knots_desc.rend());
    for (auto it = blocks_desc.rbegin(); it != blocks_desc.rend(); ++it) {
        table.coeffs.insert(table.coeffs.end(), it->begin(), it->end());
    }
    return table;
}

/** pos = (#knots < x) − 1 (first interval for x ≤ first knot), Horner from c5 down (tabulate.h:184-196) */
inline double evalAndrea(const SplineTable& t, double x)
{
    size_t idx = static_cast<size_t>(std::lower_bound(t.knots.begin(), t.knots.end(), x) - t.knots.begin());
    const size_t pos = idx == 0 ? 0 : idx - 1;
    const double dz = x - t.knots[pos];
    const double* c = &t.coeffs[6 * pos];
    double sum = 0;
    for (int i = 5; i > 0; --i) {
        sum = dz * (sum + c[i]);
    }
    return sum + c[0];
}

/** Splined Coulomb scheme: u = lB zz / r · S(r/Rc) · exp(−κ r), r < Rc */
struct CoulombTable
{
    std::string type;
    double bjerrum_length = 0;
    double cutoff = std::sqrt(DBL_MAX); //!< plain / unshifted yukawa (cf. examples/minimal/minimal.out.json)
    double kappa = 0;
    double self_prefactor = 0; //!< self energy per particle = lB · prefactor · q² / cutoff
    SplineTable S;
    SplineTable dS; //!< S'(q), tabulated like S: the force is lB zz/r³ [S(1 + κr) − q S'] e^{−κr} (fb_force.cuh)
};

inline double binomialCoefficient(int n, int k)
{
    if (k < 0 || k > n) {
        return 0;
    }
    double r = 1;
    for (int i = 1; i <= k; ++i) {
        r = r * (n - k + i) / i;
    }
    return r;
}

/** S(q) per `type` (docs/_docs/energy.md:189-210), self prefactor S'(0)/2 (docs :272-282) */
inline CoulombTable makeCoulombTable(const Json& j)
{
    CoulombTable t;
    t.type = j.at("type").string();
    t.bjerrum_length = pc::bjerrumLength(j.at("epsr").number());
    const auto salt = makeElectrolyte(j);
    const double inv_debye = salt ? 1.0 / salt->debyeLength(t.bjerrum_length) : 0.0;
    const double sqrt_pi = std::sqrt(pc::pi);
    std::function<double(double)> S;
    std::function<double(double)> dS = [](double) { return 0.0; }; // S ≡ 1 unless a type says otherwise
    auto cutoff = [&] { return j.at("cutoff").number(); };
    auto poisson = [&](int C, int D, double kappa) {
        t.cutoff = cutoff();
        t.kappa = kappa;
        const double kRc = kappa * t.cutoff;
        const bool screened = kRc > 1e-10;
        dS = [=](double q) { // chain rule through q' (q), product rule on (1 − q')^{D+1} · Σ
            double qp = q, dqp = 1.0;
            if (screened) {
                const double denominator = 1.0 - std::exp(2.0 * kRc);
                qp = (1.0 - std::exp(2.0 * kRc * q)) / denominator;
                dqp = -2.0 * kRc * std::exp(2.0 * kRc * q) / denominator;
            }
            double sum = 0, dsum = 0;
            for (int c = 0; c < C; ++c) {
                const double a = static_cast<double>(C - c) / C * binomialCoefficient(D - 1 + c, c);
                sum += a * std::pow(qp, c);
                if (c > 0) {
                    dsum += a * c * std::pow(qp, c - 1);
                }
            }
            return (-(D + 1) * std::pow(1.0 - qp, D) * sum + std::pow(1.0 - qp, D + 1) * dsum) * dqp;
        };
        S = [=](double q) {
            double qp = q;
            if (screened) {
                qp = (1.0 - std::exp(2.0 * kRc * q)) / (1.0 - std::exp(2.0 * kRc));
            }
            double sum = 0;
            for (int c = 0; c < C; ++c) {
                sum += static_cast<double>(C - c) / C * binomialCoefficient(D - 1 + c, c) * std::pow(qp, c);
            }
            return std::pow(1.0 - qp, D + 1) * sum;
        };
        double slope = 1.0;
        if (screened) {
            slope = 2.0 * kRc / (std::exp(2.0 * kRc) - 1.0);
        }
        t.self_prefactor = -0.5 * static_cast<double>(C + D) / C * slope;
    };
    const std::string& type = t.type;
    if (type == "plain") {
        if (j.contains("cutoff")) {
            throw std::runtime_error("unexpected cutoff for plain: it's *always* infinity");
        }
        t.kappa = j.contains("debyelength") ? 1.0 / j.at("debyelength").number() : 0.0;
        S = [](double) { return 1.0; };
    }
    else if (type == "yukawa") {
        if (!salt) {
            throw std::runtime_error("yukawa requires debyelength or molarity");
        }
        if (j.value("shift", false)) {
            poisson(1, 1, inv_debye);
        }
        else {
            if (j.contains("cutoff")) {
                throw std::runtime_error("unexpected 'cutoff' for non-shifted yukawa which is always infinity");
            }
            t.kappa = inv_debye;
            S = [](double) { return 1.0; };
        }
    }
    else if (type == "poisson") {
        poisson(j.value("C", 3), j.value("D", 3), inv_debye);
    }
    else if (type == "fanourgakis") {
        t.cutoff = cutoff();
        S = [](double q) {
            const double q2 = q * q;
            const double q5 = q2 * q2 * q;
            return 1.0 - 1.75 * q + 5.25 * q5 - 7.0 * q5 * q + 2.5 * q5 * q2;
        };
        dS = [](double q) {
            const double q2 = q * q;
            const double q4 = q2 * q2;
            return -1.75 + 26.25 * q4 - 42.0 * q4 * q + 17.5 * q4 * q2;
        };
        t.self_prefactor = -0.875;
    }
    else if (type == "qpotential") {
        t.cutoff = cutoff();
        const int order = j.at("order").integer();
        S = [order](double q) {
            double product = 1, power = 1;
            for (int n = 1; n <= order; ++n) {
                power *= q;
                product *= (1.0 - power);
            }
            return product;
        };
        dS = [order](double q) { // Σ_n −n q^{n−1} Π_{m≠n} (1 − q^m)
            double sum = 0;
            for (int n = 1; n <= order; ++n) {
                double rest = 1;
                for (int m = 1; m <= order; ++m) {
                    if (m != n) {
                        rest *= 1.0 - std::pow(q, m);
                    }
                }
                sum -= n * std::pow(q, n - 1) * rest;
            }
            return sum;
        };
        t.self_prefactor = -0.5;
    }
    else if (type == "ewald") {
        t.cutoff = cutoff();
        t.kappa = inv_debye;
        const double eta = j.at("alpha").number() * t.cutoff;
        const double zeta = t.kappa * t.cutoff;
        if (zeta < 1e-12) {
            S = [eta](double q) { return std::erfc(eta * q); };
            dS = [eta, sqrt_pi](double q) { return -2 * eta / sqrt_pi * std::exp(-eta * eta * q * q); };
            t.self_prefactor = -eta / sqrt_pi;
        }
        else {
            S = [eta, zeta](double q) {
                return 0.5 * std::erfc(eta * q + zeta / (2 * eta)) * std::exp(2 * zeta * q) +
                       0.5 * std::erfc(eta * q - zeta / (2 * eta));
            };
            dS = [eta, zeta, sqrt_pi](double q) {
                const double up = eta * q + zeta / (2 * eta);
                const double down = eta * q - zeta / (2 * eta);
                return zeta * std::erfc(up) * std::exp(2 * zeta * q) -
                       eta / sqrt_pi * (std::exp(-up * up + 2 * zeta * q) + std::exp(-down * down));
            };
            t.self_prefactor = -eta / sqrt_pi * (std::exp(-zeta * zeta / (4 * eta * eta)) -
                                                 sqrt_pi * zeta / (2 * eta) * std::erfc(zeta / (2 * eta)));
        }
    }
    else if (type == "wolf" || type == "zahn" || type == "fennell" || type == "zerodipole") {
        t.cutoff = cutoff();
        const double eta = j.at("alpha").number() * t.cutoff;
        const double e1 = std::erfc(eta);
        const double e2 = e1 + 2 * eta / sqrt_pi * std::exp(-eta * eta);
        auto d_erfc = [eta, sqrt_pi](double q) { return -2 * eta / sqrt_pi * std::exp(-eta * eta * q * q); };
        if (type == "wolf") {
            S = [=](double q) { return std::erfc(eta * q) - e1 * q; };
            dS = [=](double q) { return d_erfc(q) - e1; };
            t.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - e1);
        }
        else if (type == "zahn") {
            S = [=](double q) { return std::erfc(eta * q) - (q - 1) * q * e2; };
            dS = [=](double q) { return d_erfc(q) - (2 * q - 1) * e2; };
            t.self_prefactor = 0.5 * (-2 * eta / sqrt_pi + e2);
        }
        else if (type == "fennell") {
            S = [=](double q) { return std::erfc(eta * q) - q * e1 + (q - 1) * q * e2; };
            dS = [=](double q) { return d_erfc(q) - e1 + (2 * q - 1) * e2; };
            t.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - e1 - e2);
        }
        else {
            S = [=](double q) { return std::erfc(eta * q) - q * e1 + 0.5 * (q * q - 1) * q * e2; };
            dS = [=](double q) { return d_erfc(q) - e1 + 0.5 * (3 * q * q - 1) * e2; };
            t.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - e1 - 0.5 * e2);
        }
    }
    else if (type == "reactionfield") {
        t.cutoff = cutoff();
        const double epsr = j.at("epsr").number();
        const double epsrf = j.at("epsrf").number();
        const double a = (epsrf - epsr) / (2 * epsrf + epsr);
        const double b = 3 * epsrf / (2 * epsrf + epsr);
        S = [=](double q) { return 1 + a * q * q * q - b * q; };
        dS = [=](double q) { return 3 * a * q * q - b; };
        t.self_prefactor = -0.5 * b;
    }
    else {
        throw std::runtime_error("unknown type '" + type + "'");
    }
    SplineOptions opt;
    opt.utol = j.value("utol", 0.005 / t.bjerrum_length); // src/potentials.cpp:1634
    t.S = tabulateAndrea(S, 0.0, 1.0, opt);
    t.dS = tabulateAndrea(dS, 0.0, 1.0, opt);
    return t;
}

/** All tables the device needs for one `nonbonded*` energy entry */
struct PairTables
{
    int kind = potkind::FUNCTOR;
    int n_types = 0;
    std::vector<uint32_t> flags;
    std::vector<std::vector<uint32_t>> term_order; //!< per pair: term bits in input order (spline building)
    std::vector<double> lj_s2, lj_e4, wca_s2, wca_e4, hs_s2;
    bool has_coulomb = false;
    CoulombTable coulomb;
    std::string coulomb_json; //!< detects conflicting per-pair Coulomb definitions
    double plain_bjerrum_length = 0;
    // nonbonded_splined
    std::vector<int> sp_offset;
    std::vector<double> sp_knots, sp_coeffs, sp_rmin2, sp_rmax2;
    std::vector<unsigned char> sp_hs;
    // group cutoffs
    std::vector<double> g2g_cutoff_squared;
    const Json* ewald_json = nullptr; //!< coulomb block if type == ewald (Hamiltonian::addEwald)
};

namespace detail {

enum class Rule
{
    UNDEFINED,
    ARITHMETIC,
    GEOMETRIC,
    LB
};

inline Rule ruleFromJson(const Json& j, Rule fallback)
{
    const auto* m = j.find("mixing");
    if (!m) {
        return fallback;
    }
    const std::string& s = m->string();
    if (s == "LB" || s == "lorentz_berthelot") {
        return Rule::LB;
    }
    if (s == "arithmetic") {
        return Rule::ARITHMETIC;
    }
    if (s == "geometric") {
        return Rule::GEOMETRIC;
    }
    if (s == "undefined") {
        return Rule::UNDEFINED;
    }
    throw std::runtime_error("unknown combination rule " + s);
}

/** One mixed coefficient matrix: f(mix(x_i, x_j)), diagonal unmixed, `custom` overrides */
inline std::vector<double> mixedMatrix(const Topology& topo, const Json& cfg, const std::string& key,
                                       const std::string& custom_key, double unit, Rule rule, bool is_sigma,
                                       const std::function<double(double)>& f)
{
    const size_t n = topo.atoms.size();
    std::vector<double> m(n * n);
    auto mix = [&](double a, double b) {
        switch (rule) {
        case Rule::ARITHMETIC:
            return 0.5 * (a + b);
        case Rule::GEOMETRIC:
            return std::sqrt(a * b);
        case Rule::LB:
            return is_sigma ? 0.5 * (a + b) : std::sqrt(a * b);
        default:
            return std::nan("");
        }
    };
    for (size_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < n; ++k) {
            const auto& a = topo.atoms[i];
            const auto& b = topo.atoms[k];
            if (a.implicit || b.implicit) {
                m[i * n + k] = std::nan("");
            }
            else if (i == k) {
                m[i * n + k] = f(a.parameter(key) * unit);
            }
            else {
                m[i * n + k] = f(mix(a.parameter(key) * unit, b.parameter(key) * unit));
            }
        }
    }
    if (const auto* custom = cfg.find("custom")) {
        auto apply = [&](const std::string& names, const Json& values) {
            const auto w = splitWords(names);
            if (w.size() != 2) {
                throw std::runtime_error("custom interaction parameters require exactly 2 space-separated atoms");
            }
            const size_t a = static_cast<size_t>(topo.atomId(w[0]));
            const size_t b = static_cast<size_t>(topo.atomId(w[1]));
            const double v = f(values.at(custom_key).number() * unit);
            m[a * n + b] = m[b * n + a] = v;
        };
        if (custom->is_array()) {
            for (const auto& item : custom->items()) {
                apply(item.single().first, item.single().second);
            }
        }
        else {
            for (const auto& [names, values] : custom->members()) {
                apply(names, values);
            }
        }
    }
    return m;
}

inline const Json& own(const Json& j, const char* key)
{
    const auto* p = j.find(key);
    return p ? *p : j;
}

inline void addLennardJones(PairTables& t, const Topology& topo, const Json& j, bool wca)
{
    const Json& cfg = own(j, wca ? "wca" : "lennardjones");
    const Rule rule = ruleFromJson(cfg, Rule::LB);
    const std::string sigma_key = cfg.is_object() ? cfg.value("sigma", "sigma") : "sigma";
    const std::string eps_key = cfg.is_object() ? cfg.value("eps", "eps") : "eps";
    auto s2 = mixedMatrix(topo, cfg, sigma_key, "sigma", 1.0, rule, true, [](double x) { return x * x; });
    auto e4 = mixedMatrix(topo, cfg, eps_key, "eps", units::kJmol(), rule, false, [](double x) { return 4 * x; });
    (wca ? t.wca_s2 : t.lj_s2) = s2;
    (wca ? t.wca_e4 : t.lj_e4) = e4;
}

inline void addHardSphere(PairTables& t, const Topology& topo, const Json& j)
{
    const Json& cfg = own(j, "hardsphere");
    const Rule rule = ruleFromJson(cfg, Rule::ARITHMETIC);
    const std::string sigma_key = cfg.is_object() ? cfg.value("sigma", "sigma") : "sigma";
    t.hs_s2 = mixedMatrix(topo, cfg, sigma_key, "sigma", 1.0, rule, true, [](double x) { return x * x; });
}

inline void addSplinedCoulomb(PairTables& t, const Json& j)
{
    const Json& cfg = own(j, "coulomb");
    const std::string text = cfg.dump();
    if (t.has_coulomb) {
        if (text != t.coulomb_json) {
            throw std::runtime_error("different `coulomb` definitions per atom pair are not supported on the device");
        }
        return;
    }
    t.coulomb = makeCoulombTable(cfg);
    t.coulomb_json = text;
    t.has_coulomb = true;
}

inline void addPlainCoulomb(PairTables& t, const Json& j)
{
    const Json& cfg = own(j, "coulomb");
    if (!(cfg.is_object() && cfg.size() == 1)) {
        throw std::runtime_error("Plain Coulomb potential expects 'epsr' key (only)");
    }
    const double lB = pc::bjerrumLength(cfg.at("epsr").number());
    if (t.plain_bjerrum_length != 0 && t.plain_bjerrum_length != lB) {
        throw std::runtime_error("different plain Coulomb definitions per atom pair are not supported on the device");
    }
    t.plain_bjerrum_length = lB;
}

/** One potential array of the functor schema → term flags; tables are merged into `t` */
inline uint32_t lowerPotentialArray(PairTables& t, const Topology& topo, const Json& array,
                                    std::vector<uint32_t>& order)
{
    if (!array.is_array()) {
        throw std::runtime_error("potential array required");
    }
    uint32_t flags = 0;
    order.clear();
    auto once = [&](uint32_t bit, const std::string& name) {
        if (flags & bit) {
            throw std::runtime_error("potential '" + name + "' listed twice for one pair is not supported");
        }
        flags |= bit;
        order.push_back(bit);
    };
    for (const auto& record : array.items()) {
        if (!record.is_object() || record.size() != 1) {
            continue;
        }
        const auto& [name, cfg] = record.single();
        if (name == "coulomb") {
            once(term::COULOMB_SPLINED, name);
            addSplinedCoulomb(t, cfg);
        }
        else if (name == "lennardjones") {
            once(term::LJ, name);
            if (t.lj_s2.empty()) {
                addLennardJones(t, topo, record, false);
            }
        }
        else if (name == "wca") {
            once(term::WCA, name);
            if (t.wca_s2.empty()) {
                addLennardJones(t, topo, record, true);
            }
        }
        else if (name == "hardsphere") {
            once(term::HARDSPHERE, name);
            if (t.hs_s2.empty()) {
                addHardSphere(t, topo, record);
            }
        }
        else if (name == "pm") {
            once(term::COULOMB_PLAIN, name);
            once(term::HARDSPHERE, name);
            addPlainCoulomb(t, cfg);
            if (t.hs_s2.empty()) {
                addHardSphere(t, topo, cfg);
            }
        }
        else if (name == "pmwca") {
            once(term::COULOMB_PLAIN, name);
            once(term::WCA, name);
            addPlainCoulomb(t, cfg);
            if (t.wca_s2.empty()) {
                addLennardJones(t, topo, cfg, true);
            }
        }
        else {
            throw std::runtime_error("potential '" + name + "' is outside the B200 hot-path scope");
        }
    }
    return flags;
}

} // namespace detail

/** Host evaluation of the exact per-pair sum — used ONLY to build spline tables at configuration time */
inline double evalExactHost(const PairTables& t, int a, int b, double qa, double qb, double r2)
{
    const size_t idx = static_cast<size_t>(a) * t.n_types + b;
    double u = 0; // terms are added in input order, as the functor composition does (potentials.cpp:1297-1299)
    for (const uint32_t f : t.term_order[idx]) {
        double v = 0;
        if (f == term::COULOMB_SPLINED) {
            const double r = std::sqrt(r2) + std::numeric_limits<double>::epsilon();
            if (r < t.coulomb.cutoff) {
                v = qa * qb / r * evalAndrea(t.coulomb.S, r * (1.0 / t.coulomb.cutoff));
                if (t.coulomb.kappa > 0) {
                    v *= std::exp(-t.coulomb.kappa * r);
                }
                v = t.coulomb.bjerrum_length * v;
            }
        }
        else if (f == term::COULOMB_PLAIN) {
            v = t.plain_bjerrum_length * qa * qb / std::sqrt(r2);
        }
        else if (f == term::LJ) {
            double x = t.lj_s2[idx] / r2;
            x = x * x * x;
            v = t.lj_e4[idx] * (x * x - x);
        }
        else if (f == term::WCA) {
            double x = t.wca_s2[idx];
            if (!(r2 > x * 1.2599210498948732)) {
                x = x / r2;
                x = x * x * x;
                v = t.wca_e4[idx] * (x * x - x + 0.25);
            }
        }
        else if (f == term::HARDSPHERE) {
            v = r2 < t.hs_s2[idx] ? pc::infty : 0.0;
        }
        u = u + v;
    }
    return u;
}

inline std::vector<double> groupCutoffMatrix(const Json& j, const Topology& topo)
{
    const size_t n = topo.molecules.size();
    std::vector<double> c2(n * n, pc::max_value);
    auto single = [&](double cutoff) {
        const double v = (cutoff < std::sqrt(pc::max_value)) ? cutoff * cutoff : pc::max_value;
        std::fill(c2.begin(), c2.end(), v);
    };
    if (const auto* it = j.find("cutoff_g2g")) {
        if (it->is_number()) {
            single(it->number());
        }
        else if (it->is_object()) {
            single(it->value("default", pc::max_value));
            for (const auto& [pair, value] : it->members()) {
                if (pair == "default") {
                    continue;
                }
                const auto names = splitWords(pair);
                if (names.size() != 2) {
                    throw std::runtime_error("invalid molecules names");
                }
                const auto a = static_cast<size_t>(topo.moleculeId(names[0]));
                const auto b = static_cast<size_t>(topo.moleculeId(names[1]));
                c2[a * n + b] = c2[b * n + a] = std::pow(value.number(), 2);
            }
        }
    }
    return c2;
}

/** name → flavour map of Hamiltonian::createEnergy, src/energy.cpp:1293-1327 */
inline bool isNonbondedName(const std::string& name)
{
    return name == "nonbonded_coulomblj" || name == "nonbonded_newcoulomblj" || name == "nonbonded_coulombwca" ||
           name == "nonbonded_pm" || name == "nonbonded_coulombhs" || name == "nonbonded_pmwca" ||
           name == "nonbonded" || name == "nonbonded_exact" || name == "nonbonded_splined" || name == "nonbonded_cached";
}

inline PairTables buildPairTables(const std::string& name, const Json& cfg, const Topology& topo)
{
    PairTables t;
    t.n_types = static_cast<int>(topo.atoms.size());
    const size_t n2 = static_cast<size_t>(t.n_types) * t.n_types;
    if (name == "nonbonded_coulomblj" || name == "nonbonded_newcoulomblj") {
        t.kind = potkind::COULOMB_LJ;
        detail::addSplinedCoulomb(t, cfg);
        detail::addLennardJones(t, topo, cfg, false);
        t.flags.assign(n2, term::COULOMB_SPLINED | term::LJ);
    }
    else if (name == "nonbonded_coulombwca") {
        t.kind = potkind::COULOMB_WCA;
        detail::addSplinedCoulomb(t, cfg);
        detail::addLennardJones(t, topo, cfg, true);
        t.flags.assign(n2, term::COULOMB_SPLINED | term::WCA);
    }
    else if (name == "nonbonded_pm" || name == "nonbonded_coulombhs") {
        t.kind = potkind::PM;
        detail::addPlainCoulomb(t, cfg);
        detail::addHardSphere(t, topo, cfg);
        t.flags.assign(n2, term::COULOMB_PLAIN | term::HARDSPHERE);
    }
    else if (name == "nonbonded_pmwca") {
        t.kind = potkind::PMWCA;
        detail::addPlainCoulomb(t, cfg);
        detail::addLennardJones(t, topo, cfg, true);
        t.flags.assign(n2, term::COULOMB_PLAIN | term::WCA);
    }
    else if (name == "nonbonded" || name == "nonbonded_exact" || name == "nonbonded_splined" || name == "nonbonded_cached") {
        // `nonbonded_cached` = NonbondedCached<PairEnergy<SplinedPotential>> (src/energy.cpp:1311-1315): the splined
        // potential behind a cache of group-group energies (src/energy.h:1614-1758) — the same energies as
        // `nonbonded_splined`; the cache is a CPU device to avoid pair loops and has no counterpart here
        t.kind = (name == "nonbonded_splined" || name == "nonbonded_cached") ? potkind::SPLINED : potkind::FUNCTOR;
        std::vector<uint32_t> order;
        t.flags.assign(n2, detail::lowerPotentialArray(t, topo, cfg.at("default"), order));
        t.term_order.assign(n2, order);
        for (const auto& [key, value] : cfg.members()) {
            const auto pair = splitWords(key);
            if (pair.size() == 2 && value.is_array()) {
                const auto a = static_cast<size_t>(topo.atomId(pair[0]));
                const auto b = static_cast<size_t>(topo.atomId(pair[1]));
                t.flags[a * t.n_types + b] = t.flags[b * t.n_types + a] =
                    detail::lowerPotentialArray(t, topo, value, order);
                t.term_order[a * t.n_types + b] = t.term_order[b * t.n_types + a] = order;
            }
        }
    }
    else {
        throw std::runtime_error("'" + name + "' is not a non-bonded energy");
    }
    if (t.kind == potkind::SPLINED) { // src/potentials.cpp:1513-1595
        SplineOptions opt;
        opt.utol = cfg.value("utol", 1e-3);
        opt.ftol = cfg.value("ftol", 1e-2);
        const bool hardsphere = cfg.value("hardsphere", false);
        const double u_at_rmin = cfg.value("u_at_rmin", 20.0);
        const double u_at_rmax = cfg.value("u_at_rmax", 1e-6);
        const double dr = 1e-2;
        const int n = t.n_types;
        std::vector<SplineTable> tables(n2);
        t.sp_rmin2.assign(n2, 0.0);
        t.sp_rmax2.assign(n2, 0.0);
        t.sp_hs.assign(n2, 0);
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k <= i; ++k) {
                if (topo.atoms[i].implicit || topo.atoms[k].implicit) {
                    continue;
                }
                const double qa = topo.atoms[i].charge;
                const double qb = topo.atoms[k].charge;
                auto exact = [&](double r2) { return evalExactHost(t, i, k, qa, qb, r2); };
                double rmin = 0.5 * (topo.atoms[i].sigma + topo.atoms[k].sigma);
                double rmax = cfg.value("rmax", rmin * 10);
                if (const auto* it = cfg.find("cutoff_g2g")) {
                    rmax = it->is_number() ? it->number() : (it->is_object() ? it->at("default").number() : rmax);
                }
                for (int it = 0; rmin >= dr; ++it) { // findLowerDistance
                    if (it == 1000000) {
                        throw std::runtime_error("Pair potential spline error: cannot determine minimum distance");
                    }
                    const double u = std::fabs(exact(rmin * rmin));
                    if (u > u_at_rmin * 1.1) {
                        rmin += dr;
                    }
                    else if (u < u_at_rmin / 1.1) {
                        rmin -= dr;
                    }
                    else {
                        break;
                    }
                }
                for (int it = 0; rmax >= dr; ++it) { // findUpperDistance
                    if (it == 1000000) {
                        throw std::runtime_error("Pair potential spline error: cannot determine maximum distance");
                    }
                    if (std::fabs(exact(rmax * rmax)) > u_at_rmax) {
                        rmax += dr;
                    }
                    else {
                        break;
                    }
                }
                SplineTable tab = tabulateAndrea(exact, rmin * rmin, rmax * rmax, opt);
                bool hs = hardsphere;
                if (evalAndrea(tab, tab.xmin + dr) < 0) {
                    hs = false;
                }
                const size_t a = static_cast<size_t>(i) * n + k;
                const size_t b = static_cast<size_t>(k) * n + i;
                t.sp_rmin2[a] = t.sp_rmin2[b] = tab.xmin;
                t.sp_rmax2[a] = t.sp_rmax2[b] = tab.xmax;
                t.sp_hs[a] = t.sp_hs[b] = hs ? 1 : 0;
                tables[a] = tables[b] = tab;
            }
        }
        t.sp_offset.assign(n2 + 1, 0);
        for (size_t p = 0; p < n2; ++p) {
            t.sp_offset[p] = static_cast<int>(t.sp_knots.size());
            t.sp_knots.insert(t.sp_knots.end(), tables[p].knots.begin(), tables[p].knots.end());
            t.sp_coeffs.insert(t.sp_coeffs.end(), tables[p].coeffs.begin(), tables[p].coeffs.end());
        }
        t.sp_offset[n2] = static_cast<int>(t.sp_knots.size());
    }
    t.g2g_cutoff_squared = groupCutoffMatrix(cfg, topo);
    // Hamiltonian::addEwald: `default[i].coulomb` or `coulomb` with type == ewald (energy.cpp:1134-1160)
    const Json* coulomb = nullptr;
    if (const auto* def = cfg.find("default")) {
        for (const auto& i : def->items()) {
            if (const auto* c = i.find("coulomb")) {
                coulomb = c;
                break;
            }
        }
    }
    else if (const auto* c = cfg.find("coulomb")) {
        coulomb = c;
    }
    if (coulomb && coulomb->value("type", "") == "ewald") {
        t.ewald_json = coulomb;
    }
    return t;
}

} // namespace fb

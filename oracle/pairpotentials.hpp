// ORACLE — test infrastructure, not product code. CPU restatement of the reference's pair
// potentials for the ΔU hot path. Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline
// may use anything under oracle/.
//
// Restates (reference file:line):
//   PairMixer / combination rules ........ src/potentials.cpp:18-98
//   LennardJones ........................... src/potentials.h:15-50, src/potentials.cpp:672-703
//   WeeksChandlerAndersen .................. src/potentials.h:147-185
//   HardSphere ............................. src/potentials.h:193-208, src/potentials.cpp:956-977
//   Coulomb (plain, no cutoff) ............. src/potentials.h:463-479, src/potentials.cpp:452-460
//   NewCoulombGalore ....................... src/potentials.h:581-598, src/potentials.cpp:1599-1699
//   Tabulate::Andrea (generate + eval) ..... src/tabulate.h:28-308
//   CombinedPairPotential .................. src/potentials_base.h:226-286
//   FunctorPotential / SplinedPotential .... src/potentials.h:728-824, src/potentials.cpp:1202-1326,
//                                            1453-1595
//
// PARITY UNPINNED for every non-`plain` Coulomb type: the short-range functions S(q), their
// splining (`CoulombGalore::Splined`) and the self-energy prefactors live in the un-vendored
// dependency mlund/coulombgalore @ 4055f58538d781acccb2937ab4580855fcba31f8
// (cmake/ExternalTools.cmake:202-215), absent from /root/reference. What is restated here is the
// published definition: S(q) from docs/_docs/energy.md:189-210, splined on q∈[0,1] with the
// reference's own Andrea tabulator (src/tabulate.h) at `utol = 0.005/lB` (src/potentials.cpp:1634),
// u = lB·zA·zB/r·S(r/Rc)·exp(−κr) for r < Rc, self energy −½·lim(u−ũ) = lB·q²·S'(0)/(2Rc)
// (docs/_docs/energy.md:272-282).
#pragma once
#include "../faunus_b200/csrc/host/space.hpp"
#include <functional>

namespace oracle {
using namespace fb;

// ---------------------------------------------------------------------------------------------
// Andrea spline, src/tabulate.h
// ---------------------------------------------------------------------------------------------
struct SplineData
{
    std::vector<double> r2; //!< knots (abscissa; r² for pair splines, q for the Coulomb S(q))
    std::vector<double> c;  //!< 6 coefficients per interval
    double rmin2 = 0, rmax2 = 0;
    size_t numKnots() const { return r2.size(); }
};

class Andrea
{
    double utol = 1e-5, ftol = -1, umaxtol = -1, fmaxtol = -1;
    double numdr = 0.0001;
    int mngrid = 1200;
    int ndr = 100;
    double drfrac = 0.9;
    using Func = std::function<double(double)>;

    double f1(const Func& f, double x) const { return (f(x + numdr * 0.5) - f(x - numdr * 0.5)) / numdr; }
    double f2(const Func& f, double x) const
    {
        return (f1(f, x + numdr * 0.5) - f1(f, x - numdr * 0.5)) / numdr;
    }

    /** src/tabulate.h:103-132 */
    static std::array<double, 7> setUBuffer(double zlow, double zupp, double u0low, double u1low,
                                            double u2low, double u0upp, double u1upp, double u2upp)
    {
        if (std::fabs(u0low) < 1e-9 && std::fabs(u1low) < 1e-9) {
            return {0, 0, 0, 0, 0, 0, 0};
        }
        const double dz1 = zupp - zlow;
        const double dz2 = dz1 * dz1;
        const double dz3 = dz2 * dz1;
        const double c0 = u0low;
        const double c1 = u1low;
        const double c2 = u2low * 0.5;
        const double a = 6 * (u0upp - c0 - c1 * dz1 - c2 * dz2) / dz3;
        const double b = 2 * (u1upp - c1 - 2 * c2 * dz1) / dz2;
        const double c = (u2upp - 2 * c2) / dz1;
        const double c3 = (10 * a - 12 * b + 3 * c) / 6;
        const double c4 = (-15 * a + 21 * b - 6 * c) / (6 * dz1);
        const double c5 = (2 * a - 3 * b + c) / (2 * dz2);
        return {zlow, c0, c1, c2, c3, c4, c5};
    }

    /** @return {tolerance approved, repulsive part found}; src/tabulate.h:139-175 */
    std::pair<bool, bool> checkUBuffer(const std::array<double, 7>& ub, double rlow, double rupp,
                                       const Func& f) const
    {
        const int ncheck = 11;
        const double dr = (rupp - rlow) / (ncheck - 1);
        bool repulsive = false;
        for (int i = 0; i < ncheck; i++) {
            const double r1 = rlow + dr * static_cast<double>(i);
            const double r2 = r1 * r1;
            const double u0 = f(r2);
            const double u1 = f1(f, r2);
            const double dz = r2 - rlow * rlow;
            const double usum = ub[1] + dz * (ub[2] + dz * (ub[3] + dz * (ub[4] + dz * (ub[5] + dz * ub[6]))));
            const double fsum =
                ub[2] + dz * (2 * ub[3] + dz * (3 * ub[4] + dz * (4 * ub[5] + dz * (5 * ub[6]))));
            if (std::fabs(usum - u0) > utol) {
                return {false, false};
            }
            if (ftol != -1 && std::fabs(fsum - u1) > ftol) {
                return {false, false};
            }
            if (umaxtol != -1 && std::fabs(usum) > umaxtol) {
                repulsive = true;
            }
            if (fmaxtol != -1 && std::fabs(usum) > fmaxtol) {
                repulsive = true;
            }
        }
        return {true, repulsive};
    }

  public:
    void setTolerance(double u, double f = -1, double umax = -1, double fmax = -1)
    {
        utol = u;
        ftol = f;
        umaxtol = umax;
        fmaxtol = fmax;
    }

    /**
     * src/tabulate.h:184-196: pos = lower_bound(knots, x) − 1; Horner as the loop
     * `sum = dz·(sum + c[i])` for i = 5..1, then + c0.
     * Deviation: for x ≤ first knot the reference indexes out of range; here the first interval
     * is used.
     */
    static double eval(const SplineData& d, double x)
    {
        size_t idx = static_cast<size_t>(std::lower_bound(d.r2.begin(), d.r2.end(), x) - d.r2.begin());
        const size_t pos = (idx == 0) ? 0 : idx - 1;
        const size_t pos6 = 6 * pos;
        const double dz = x - d.r2[pos];
        double sum = 0;
        for (size_t i = 5; i > 0; i--) {
            sum = dz * (sum + d.c[pos6 + i]);
        }
        return sum + d.c[pos6];
    }

    /** Tabulate f(x) in ]min,max]; src/tabulate.h:223-307 */
    SplineData generate(const Func& f, double rmin, double rmax) const
    {
        rmin = std::sqrt(rmin);
        rmax = std::sqrt(rmax);
        SplineData td;
        td.rmin2 = rmin * rmin;
        td.rmax2 = rmax * rmax;
        double rumin = rmin;
        const double rmax2 = rmax * rmax;
        double dr = rmax - rmin;
        double rupp = rmax;
        double zupp = rmax2;
        bool repul = false;
        td.r2.push_back(zupp);
        int i;
        for (i = 0; i < mngrid; i++) {
            double rlow = rupp;
            double zlow = 0;
            std::array<double, 7> ubuft{};
            int j;
            dr = (rupp - rmin);
            for (j = 0; j < ndr; j++) {
                zupp = rupp * rupp;
                rlow = rupp - dr;
                if (rumin > rlow) {
                    rlow = rumin;
                }
                zlow = rlow * rlow;
                const double u0low = f(zlow);
                const double u1low = f1(f, zlow);
                const double u2low = f2(f, zlow);
                const double u0upp = f(zupp);
                const double u1upp = f1(f, zupp);
                const double u2upp = f2(f, zupp);
                ubuft = setUBuffer(zlow, zupp, u0low, u1low, u2low, u0upp, u1upp, u2upp);
                const auto [ok, rep] = checkUBuffer(ubuft, rlow, rupp, f);
                repul = rep;
                if (ok) {
                    rupp = rlow;
                    break;
                }
                dr *= drfrac;
            }
            if (j >= ndr) {
                throw std::runtime_error("Andrea spline: try to increase utol/ftol");
            }
            td.r2.push_back(zlow);
            for (size_t k = 1; k < ubuft.size(); k++) {
                td.c.push_back(ubuft[k]);
            }
            if (repul) {
                rumin = rlow;
                td.rmin2 = rlow * rlow;
            }
            if (rlow <= rumin || repul) {
                break;
            }
        }
        if (i >= mngrid) {
            throw std::runtime_error("Andrea spline: try to increase utol/ftol");
        }
        std::reverse(td.r2.begin(), td.r2.end());
        for (size_t k = 0; k < td.c.size() / 2; k += 6) { // reverse knot order in packets of six
            std::swap_ranges(td.c.begin() + k, td.c.begin() + k + 6, td.c.end() - k - 6);
        }
        return td;
    }
};

// ---------------------------------------------------------------------------------------------
// Mixing, src/potentials.cpp:18-98
// ---------------------------------------------------------------------------------------------
enum class Mixing
{
    UNDEFINED,
    ARITHMETIC,
    GEOMETRIC,
    LORENTZ_BERTHELOT
};

inline Mixing mixingFromString(const std::string& s)
{
    if (s == "undefined") {
        return Mixing::UNDEFINED;
    }
    if (s == "arithmetic") {
        return Mixing::ARITHMETIC;
    }
    if (s == "geometric") {
        return Mixing::GEOMETRIC;
    }
    if (s == "lorentz_berthelot" || s == "LB") {
        return Mixing::LORENTZ_BERTHELOT;
    }
    throw std::runtime_error("unknown combination rule " + s);
}

struct PairMatrix
{
    size_t n = 0;
    std::vector<double> v;
    PairMatrix() = default;
    explicit PairMatrix(size_t n, double init = 0)
        : n(n)
        , v(n * n, init)
    {
    }
    double& operator()(size_t i, size_t j) { return v[i * n + j]; }
    double operator()(size_t i, size_t j) const { return v[i * n + j]; }
};

enum class Coefficient
{
    SIGMA,
    EPSILON,
    ANY
};

inline double combine(Mixing rule, Coefficient coeff, double a, double b)
{
    switch (rule) {
    case Mixing::UNDEFINED:
        return std::nan("");
    case Mixing::ARITHMETIC:
        return 0.5 * (a + b);
    case Mixing::GEOMETRIC:
        return std::sqrt(a * b);
    case Mixing::LORENTZ_BERTHELOT:
        if (coeff == Coefficient::SIGMA) {
            return 0.5 * (a + b);
        }
        if (coeff == Coefficient::EPSILON) {
            return std::sqrt(a * b);
        }
        throw std::logic_error("unsupported mixer initialization");
    }
    return std::nan("");
}

struct CustomPair
{
    int id1, id2;
    std::map<std::string, double> values;
};

inline std::vector<CustomPair> customPairsFromJson(const Json& j, const Topology& topo)
{
    std::vector<CustomPair> out;
    auto append = [&](const std::string& key, const Json& value) {
        const auto names = splitWords(key);
        if (names.size() != 2) {
            throw std::runtime_error("custom interaction parameters require exactly 2 space-separated atoms");
        }
        CustomPair cp{topo.atomId(names[0]), topo.atomId(names[1]), {}};
        for (const auto& [k, v] : value.members()) {
            if (v.is_number()) {
                cp.values[k] = v.number();
            }
        }
        out.push_back(cp);
    };
    if (j.is_array()) {
        for (const auto& item : j.items()) {
            const auto& [key, value] = item.single();
            append(key, value);
        }
    }
    else if (j.is_object()) {
        for (const auto& [key, value] : j.members()) {
            append(key, value);
        }
    }
    else {
        throw std::runtime_error("invalid JSON for custom interaction parameters");
    }
    return out;
}

/**
 * n_types × n_types matrix of modifier(combinator(extract(i), extract(j))), homogeneous pairs
 * without combination, implicit atoms → NaN, custom pairs override; src/potentials.cpp:50-91.
 * `unit` converts the raw parameter (kJ/mol → kT for eps; 1 for sigma).
 */
inline PairMatrix makePairMatrix(const Topology& topo, const std::string& param, double unit, Mixing rule,
                                 Coefficient coeff, const std::function<double(double)>& modifier,
                                 const std::vector<CustomPair>& custom)
{
    const auto& atoms = topo.atoms;
    PairMatrix m(atoms.size());
    for (const auto& i : atoms) {
        for (const auto& j : atoms) {
            if (i.implicit || j.implicit) {
                m(i.id, j.id) = std::nan("");
            }
            else if (i.id == j.id) {
                m(i.id, j.id) = modifier(i.parameter(param) * unit);
            }
            else {
                m(i.id, j.id) = modifier(combine(rule, coeff, i.parameter(param) * unit, j.parameter(param) * unit));
            }
        }
    }
    for (const auto& cp : custom) {
        auto it = cp.values.find(param);
        if (it == cp.values.end()) {
            throw std::runtime_error("custom pair misses '" + param + "'");
        }
        m(cp.id1, cp.id2) = m(cp.id2, cp.id1) = modifier(it->second * unit);
    }
    return m;
}

/** Looks for the potential's own key first, as pairpotential::from_json does (potentials.cpp:258-273) */
inline const Json& subConfig(const Json& j, const char* name)
{
    if (const auto* p = j.find(name)) {
        return *p;
    }
    return j;
}

// ---------------------------------------------------------------------------------------------
// Short-ranged potentials
// ---------------------------------------------------------------------------------------------
class LennardJones
{
  protected:
    PairMatrix sigma_squared;     //!< σ_ij²
    PairMatrix epsilon_quadruple; //!< 4 ε_ij (kT)

  public:
    static constexpr const char* key = "lennardjones";
    void from_json(const Json& jin, const Topology& topo, const char* name = key)
    {
        const Json& j = subConfig(jin, name);
        Mixing rule = Mixing::LORENTZ_BERTHELOT;
        if (const auto* m = j.find("mixing")) {
            rule = mixingFromString(m->string());
        }
        std::vector<CustomPair> custom;
        if (const auto* c = j.find("custom")) {
            custom = customPairsFromJson(*c, topo);
        }
        const std::string sigma_name = j.is_object() ? j.value("sigma", "sigma") : "sigma";
        const std::string eps_name = j.is_object() ? j.value("eps", "eps") : "eps";
        // custom pairs always use the literal keys "sigma"/"eps" of the InteractionData
        auto rename = [&](std::vector<CustomPair> v, const std::string& from, const std::string& to) {
            for (auto& cp : v) {
                if (from != to && cp.values.count(from)) {
                    cp.values[to] = cp.values[from];
                }
            }
            return v;
        };
        sigma_squared = makePairMatrix(topo, sigma_name, 1.0, rule, Coefficient::SIGMA,
                                       [](double x) { return x * x; }, rename(custom, "sigma", sigma_name));
        epsilon_quadruple = makePairMatrix(topo, eps_name, units::kJmol(), rule, Coefficient::EPSILON,
                                           [](double x) { return 4 * x; }, rename(custom, "eps", eps_name));
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        double x = sigma_squared(a.id, b.id) / r2;
        x = x * x * x;
        return epsilon_quadruple(a.id, b.id) * (x * x - x);
    }
    /** force on a: 6·4ε σ⁶ (2σ⁶ − r⁶) / r¹⁴ · (b → a); src/potentials.h:32-40 */
    inline Point force(const Particle& a, const Particle& b, double r2, const Point& b_towards_a) const
    {
        const double s2 = sigma_squared(a.id, b.id);
        const double s6 = s2 * s2 * s2;
        const double r6 = r2 * r2 * r2;
        const double r14 = r6 * r6 * r2;
        return b_towards_a * (6.0 * epsilon_quadruple(a.id, b.id) * s6 * (2.0 * s6 - r6) / r14);
    }
    const PairMatrix& sigma2() const { return sigma_squared; }
    const PairMatrix& eps4() const { return epsilon_quadruple; }
};

class WeeksChandlerAndersen : public LennardJones
{
    static constexpr double onefourth = 0.25, twototwosixth = 1.2599210498948732;

  public:
    static constexpr const char* key = "wca";
    void from_json(const Json& j, const Topology& topo) { LennardJones::from_json(j, topo, key); }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        double x = sigma_squared(a.id, b.id);
        if (r2 > x * twototwosixth) {
            return 0;
        }
        x = x / r2;
        x = x * x * x;
        return epsilon_quadruple(a.id, b.id) * (x * x - x + onefourth);
    }
    /** src/potentials.h:173-184 */
    inline Point force(const Particle& a, const Particle& b, double r2, const Point& b_towards_a) const
    {
        double x = sigma_squared(a.id, b.id);
        if (r2 > x * twototwosixth) {
            return {0.0, 0.0, 0.0};
        }
        x = x / r2;
        x = x * x * x;
        return b_towards_a * (epsilon_quadruple(a.id, b.id) * 6.0 * (2.0 * x * x - x) / r2);
    }
};

class HardSphere
{
    PairMatrix sigma_squared;

  public:
    static constexpr const char* key = "hardsphere";
    void from_json(const Json& jin, const Topology& topo)
    {
        const Json& j = subConfig(jin, key);
        Mixing rule = Mixing::ARITHMETIC;
        if (const auto* m = j.find("mixing")) {
            rule = mixingFromString(m->string());
        }
        std::vector<CustomPair> custom;
        if (const auto* c = j.find("custom")) {
            custom = customPairsFromJson(*c, topo);
        }
        const std::string sigma_name = j.is_object() ? j.value("sigma", "sigma") : "sigma";
        sigma_squared = makePairMatrix(topo, sigma_name, 1.0, rule, Coefficient::ANY,
                                       [](double x) { return x * x; }, custom);
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        return r2 < sigma_squared(a.id, b.id) ? pc::infty : 0.0;
    }
    /** no force of its own: PairPotential::force throws, src/potentials.cpp:246-251 */
    inline Point force(const Particle&, const Particle&, double, const Point&) const
    {
        throw std::logic_error("Force computation not implemented for this setup!");
    }
    const PairMatrix& sigma2() const { return sigma_squared; }
};

/** Plain Coulomb without cutoff: lB qa qb / sqrt(r²); src/potentials.h:463-479 */
class Coulomb
{
  public:
    static constexpr const char* key = "coulomb";
    double bjerrum_length = 0;
    void from_json(const Json& jin, const Topology&)
    {
        const Json& j = subConfig(jin, key);
        if (j.size() == 1 && j.is_object()) {
            bjerrum_length = pc::bjerrumLength(j.at("epsr").number());
        }
        else {
            throw std::runtime_error("Plain Coulomb potential expects 'epsr' key (only)");
        }
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        return bjerrum_length * a.charge * b.charge / std::sqrt(r2);
    }
    /** no force of its own: PairPotential::force throws, src/potentials.cpp:246-251 */
    inline Point force(const Particle&, const Particle&, double, const Point&) const
    {
        throw std::logic_error("Force computation not implemented for this setup!");
    }
    std::function<double(const Particle&)> selfEnergy() const { return nullptr; }
};

// ---------------------------------------------------------------------------------------------
// CoulombGalore restatement (see PARITY UNPINNED note in the file header)
// ---------------------------------------------------------------------------------------------
struct CoulombScheme
{
    std::string type;
    double cutoff = std::sqrt(pc::max_value); //!< plain/yukawa: sqrt(DBL_MAX), cf. minimal.out.json
    double kappa = 0;                         //!< inverse Debye length
    double self_prefactor = 0;                //!< self energy = lB · prefactor · q² / cutoff
    std::function<double(double)> S;          //!< short-range function S(q)
    std::function<double(double)> dS = [](double) { return 0.0; }; //!< S'(q) (CoulombGalore short_range_function_derivative)
};

inline double binomial(int n, int k)
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

/** S(q) table: docs/_docs/energy.md:189-210; dispatch: src/potentials.cpp:1628-1699 */
inline CoulombScheme makeCoulombScheme(const Json& j, double bjerrum_length)
{
    CoulombScheme s;
    s.type = j.at("type").string();
    const auto electrolyte = makeElectrolyte(j);
    const double debye = electrolyte ? electrolyte->debyeLength(bjerrum_length) : pc::infty;
    const double sqrt_pi = std::sqrt(pc::pi);
    auto need_cutoff = [&]() { return j.at("cutoff").number(); };

    auto poisson = [&](int C, int D, double kappa) {
        s.cutoff = need_cutoff();
        s.kappa = kappa;
        const double Rc = s.cutoff;
        const double kRc = kappa * Rc;
        const bool screened = kRc > 1e-10;
        s.S = [=](double q) {
            double qp = q;
            if (screened) {
                qp = (1.0 - std::exp(2.0 * kRc * q)) / (1.0 - std::exp(2.0 * kRc));
            }
            double sum = 0;
            for (int c = 0; c < C; ++c) {
                sum += static_cast<double>(C - c) / C * binomial(D - 1 + c, c) * std::pow(qp, c);
            }
            return std::pow(1.0 - qp, D + 1) * sum;
        };
        s.dS = [=](double q) {
            double qp = q, slope = 1.0;
            if (screened) {
                qp = (1.0 - std::exp(2.0 * kRc * q)) / (1.0 - std::exp(2.0 * kRc));
                slope = -2.0 * kRc * std::exp(2.0 * kRc * q) / (1.0 - std::exp(2.0 * kRc));
            }
            double sum = 0, dsum = 0;
            for (int c = 0; c < C; ++c) {
                const double a = static_cast<double>(C - c) / C * binomial(D - 1 + c, c);
                sum += a * std::pow(qp, c);
                dsum += c > 0 ? a * c * std::pow(qp, c - 1) : 0.0;
            }
            return slope * (std::pow(1.0 - qp, D + 1) * dsum - (D + 1) * std::pow(1.0 - qp, D) * sum);
        };
        double dqp = 1.0;
        if (screened) {
            dqp = 2.0 * kRc / (std::exp(2.0 * kRc) - 1.0);
        }
        s.self_prefactor = -0.5 * static_cast<double>(C + D) / C * dqp;
    };

    if (s.type == "yukawa") {
        if (!electrolyte) {
            throw std::runtime_error("yukawa requires debyelength or molarity");
        }
        if (j.value("shift", false)) { // poisson with C=1, D=1
            poisson(1, 1, 1.0 / debye);
        }
        else {
            if (j.contains("cutoff")) {
                throw std::runtime_error("unexpected 'cutoff' for non-shifted yukawa which is always infinity");
            }
            s.kappa = 1.0 / debye;
            s.S = [](double) { return 1.0; };
            s.self_prefactor = 0;
        }
    }
    else if (s.type == "plain") {
        if (j.contains("cutoff")) {
            throw std::runtime_error("unexpected cutoff for plain: it's *always* infinity");
        }
        s.kappa = j.contains("debyelength") ? 1.0 / j.at("debyelength").number() : 0.0;
        s.S = [](double) { return 1.0; };
    }
    else if (s.type == "fanourgakis") {
        s.cutoff = need_cutoff();
        s.S = [](double q) {
            const double q2 = q * q;
            const double q5 = q2 * q2 * q;
            return 1.0 - 1.75 * q + 5.25 * q5 - 7.0 * q5 * q + 2.5 * q5 * q2;
        };
        s.dS = [](double q) {
            const double q2 = q * q;
            const double q4 = q2 * q2;
            return -1.75 + 26.25 * q4 - 42.0 * q4 * q + 17.5 * q4 * q2;
        };
        s.self_prefactor = -0.875;
    }
    else if (s.type == "qpotential") {
        s.cutoff = need_cutoff();
        const int order = j.at("order").integer();
        s.S = [order](double q) {
            double prod = 1, qn = 1;
            for (int n = 1; n <= order; ++n) {
                qn *= q;
                prod *= (1.0 - qn);
            }
            return prod;
        };
        s.dS = [order](double q) { // product rule, one factor at a time
            double total = 0;
            for (int n = 1; n <= order; ++n) {
                double others = 1;
                for (int m = 1; m <= order; ++m) {
                    if (m != n) {
                        others *= 1.0 - std::pow(q, m);
                    }
                }
                total -= n * std::pow(q, n - 1) * others;
            }
            return total;
        };
        s.self_prefactor = -0.5;
    }
    else if (s.type == "poisson") {
        poisson(j.value("C", 3), j.value("D", 3), electrolyte ? 1.0 / debye : 0.0);
    }
    else if (s.type == "ewald") {
        s.cutoff = need_cutoff();
        const double eta = j.at("alpha").number() * s.cutoff;
        s.kappa = electrolyte ? 1.0 / debye : 0.0;
        const double zeta = s.kappa * s.cutoff;
        if (zeta < 1e-12) {
            s.S = [eta](double q) { return std::erfc(eta * q); };
            s.dS = [eta, sqrt_pi](double q) { return -2 * eta / sqrt_pi * std::exp(-eta * eta * q * q); };
            s.self_prefactor = -eta / sqrt_pi;
        }
        else {
            s.S = [eta, zeta](double q) {
                return 0.5 * std::erfc(eta * q + zeta / (2 * eta)) * std::exp(2 * zeta * q) +
                       0.5 * std::erfc(eta * q - zeta / (2 * eta));
            };
            s.dS = [eta, zeta, sqrt_pi](double q) {
                const double a = eta * q + zeta / (2 * eta);
                const double b = eta * q - zeta / (2 * eta);
                return zeta * std::erfc(a) * std::exp(2 * zeta * q) -
                       eta / sqrt_pi * (std::exp(-a * a + 2 * zeta * q) + std::exp(-b * b));
            };
            s.self_prefactor = -eta / sqrt_pi * (std::exp(-zeta * zeta / (4 * eta * eta)) -
                                                 sqrt_pi * zeta / (2 * eta) * std::erfc(zeta / (2 * eta)));
        }
    }
    else if (s.type == "wolf" || s.type == "zahn" || s.type == "fennell" || s.type == "zerodipole") {
        s.cutoff = need_cutoff();
        const double eta = j.at("alpha").number() * s.cutoff;
        const double erfc_eta = std::erfc(eta);
        const double gauss = erfc_eta + 2 * eta / sqrt_pi * std::exp(-eta * eta);
        const auto gaussian_slope = [eta, sqrt_pi](double q) { return -2 * eta / sqrt_pi * std::exp(-eta * eta * q * q); };
        if (s.type == "wolf") {
            s.S = [=](double q) { return std::erfc(eta * q) - erfc_eta * q; };
            s.dS = [=](double q) { return gaussian_slope(q) - erfc_eta; };
            s.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - erfc_eta);
        }
        else if (s.type == "zahn") {
            s.S = [=](double q) { return std::erfc(eta * q) - (q - 1) * q * gauss; };
            s.dS = [=](double q) { return gaussian_slope(q) - (2 * q - 1) * gauss; };
            s.self_prefactor = 0.5 * (-2 * eta / sqrt_pi + gauss);
        }
        else if (s.type == "fennell") {
            s.S = [=](double q) { return std::erfc(eta * q) - q * erfc_eta + (q - 1) * q * gauss; };
            s.dS = [=](double q) { return gaussian_slope(q) - erfc_eta + (2 * q - 1) * gauss; };
            s.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - erfc_eta - gauss);
        }
        else {
            s.S = [=](double q) { return std::erfc(eta * q) - q * erfc_eta + 0.5 * (q * q - 1) * q * gauss; };
            s.dS = [=](double q) { return gaussian_slope(q) - erfc_eta + 0.5 * (3 * q * q - 1) * gauss; };
            s.self_prefactor = 0.5 * (-2 * eta / sqrt_pi - erfc_eta - 0.5 * gauss);
        }
    }
    else if (s.type == "reactionfield") {
        s.cutoff = need_cutoff();
        const double epsr = j.at("epsr").number();
        const double epsrf = j.at("epsrf").number();
        const double a = (epsrf - epsr) / (2 * epsrf + epsr);
        const double b = 3 * epsrf / (2 * epsrf + epsr);
        s.S = [=](double q) { return 1 + a * q * q * q - b * q; };
        s.dS = [=](double q) { return 3 * a * q * q - b; };
        s.self_prefactor = -0.5 * b;
    }
    else {
        throw std::runtime_error("unknown type '" + s.type + "'");
    }
    return s;
}

/** `coulomb: {type, epsr, cutoff, …}` = lB·zz/r·S̃(r/Rc)·exp(−κr); src/potentials.h:581-598 */
class NewCoulombGalore
{
  public:
    static constexpr const char* key = "coulomb";
    double bjerrum_length = 0;
    CoulombScheme scheme;
    SplineData table; //!< Andrea spline of S(q) on [0,1]
    SplineData slope_table; //!< … and of S'(q) (CoulombGalore::Splined tabulates the derivatives too)
    double inverse_cutoff = 0;

    void from_json(const Json& jin, const Topology&)
    {
        const Json& j = subConfig(jin, key);
        bjerrum_length = pc::bjerrumLength(j.at("epsr").number());
        scheme = makeCoulombScheme(j, bjerrum_length);
        Andrea spline;
        spline.setTolerance(j.value("utol", 0.005 / bjerrum_length));
        table = spline.generate(scheme.S, 0.0, 1.0);
        slope_table = spline.generate(scheme.dS, 0.0, 1.0);
        inverse_cutoff = 1.0 / scheme.cutoff;
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        const double r = std::sqrt(r2) + std::numeric_limits<double>::epsilon();
        if (r < scheme.cutoff) {
            double u = a.charge * b.charge / r * Andrea::eval(table, r * inverse_cutoff);
            if (scheme.kappa > 0) {
                u *= std::exp(-scheme.kappa * r);
            }
            return bjerrum_length * u;
        }
        return 0.0;
    }
    /**
     * src/potentials.h:600-606: lB · ion_ion_force(qa, qb, b → a). CoulombGalore (absent, PARITY UNPINNED beyond
     * `plain`): the field of an ion, z r / r³ · [S(q)(1 + κr) − q S'(q)] e^{−κr} for r² < Rc², is minus the gradient
     * of the energy above. Pinned for `plain` by src/potentials.cpp:1612-1626 (0.1429734149 at r = 7, ε_r = 80).
     */
    inline Point force(const Particle& a, const Particle& b, double r2, const Point& b_towards_a) const
    {
        if (r2 < scheme.cutoff * scheme.cutoff) {
            const double r = std::sqrt(r2);
            const double q = r * inverse_cutoff;
            const double kr = scheme.kappa * r;
            double f = a.charge * b.charge / (r2 * r) *
                       (Andrea::eval(table, q) * (1.0 + kr) - q * Andrea::eval(slope_table, q));
            if (scheme.kappa > 0) {
                f *= std::exp(-kr);
            }
            return b_towards_a * (bjerrum_length * f);
        }
        return {0.0, 0.0, 0.0};
    }
    /** src/potentials.cpp:1599-1604 */
    std::function<double(const Particle&)> selfEnergy() const
    {
        const double pref = bjerrum_length * scheme.self_prefactor / scheme.cutoff;
        return [pref](const Particle& p) { return pref * p.charge * p.charge; };
    }
};

template <class T1, class T2> class CombinedPairPotential
{
  public:
    T1 first;
    T2 second;
    void from_json(const Json& j, const Topology& topo)
    {
        first.from_json(j, topo);
        second.from_json(j, topo);
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        return first(a, b, r2) + second(a, b, r2);
    }
    /** src/potentials_base.h:250-256 */
    inline Point force(const Particle& a, const Particle& b, double r2, const Point& b_towards_a) const
    {
        return first.force(a, b, r2, b_towards_a) + second.force(a, b, r2, b_towards_a);
    }
    std::function<double(const Particle&)> selfEnergy() const { return first.selfEnergy(); }
};

using CoulombLJ = CombinedPairPotential<NewCoulombGalore, LennardJones>;
using CoulombWCA = CombinedPairPotential<NewCoulombGalore, WeeksChandlerAndersen>;
using PrimitiveModel = CombinedPairPotential<Coulomb, HardSphere>;
using PrimitiveModelWCA = CombinedPairPotential<Coulomb, WeeksChandlerAndersen>;

/** Per-(type,type) functor composition from JSON; src/potentials.cpp:1202-1326 */
class FunctorPotential
{
  public:
    using EnergyFunctor = std::function<double(const Particle&, const Particle&, double)>;

  protected:
    size_t n_types = 0;
    std::vector<EnergyFunctor> umatrix;
    std::function<double(const Particle&)> self_energy;
    bool have_monopole_self_energy = false;

    EnergyFunctor combine(const Json& potential_array, const Topology& topo)
    {
        if (!potential_array.is_array()) {
            throw std::runtime_error("potential array required");
        }
        EnergyFunctor func = [](const Particle&, const Particle&, double) { return 0.0; };
        for (const auto& record : potential_array.items()) {
            if (!record.is_object() || record.size() != 1) {
                continue;
            }
            const auto& [name, cfg] = record.single();
            EnergyFunctor new_func;
            if (name == "coulomb") {
                NewCoulombGalore pot;
                pot.from_json(cfg, topo);
                if (!have_monopole_self_energy) {
                    self_energy = pot.selfEnergy();
                    have_monopole_self_energy = true;
                }
                new_func = pot;
            }
            else if (name == "lennardjones") {
                LennardJones pot;
                pot.from_json(record, topo);
                new_func = pot;
            }
            else if (name == "wca") {
                WeeksChandlerAndersen pot;
                pot.from_json(record, topo);
                new_func = pot;
            }
            else if (name == "hardsphere") {
                HardSphere pot;
                pot.from_json(record, topo);
                new_func = pot;
            }
            else if (name == "pm") {
                PrimitiveModel pot;
                pot.from_json(cfg, topo);
                new_func = pot;
            }
            else if (name == "pmwca") {
                PrimitiveModelWCA pot;
                pot.from_json(cfg, topo);
                new_func = pot;
            }
            else {
                throw std::runtime_error("potential '" + name + "' is outside the hot-path scope");
            }
            func = [func, new_func](const Particle& a, const Particle& b, double r2) {
                return func(a, b, r2) + new_func(a, b, r2);
            };
        }
        return func;
    }

  public:
    void from_json(const Json& j, const Topology& topo)
    {
        have_monopole_self_energy = false;
        self_energy = nullptr;
        n_types = topo.atoms.size();
        umatrix.assign(n_types * n_types, combine(j.at("default"), topo));
        for (const auto& [key, value] : j.members()) {
            const auto pair = splitWords(key);
            if (pair.size() == 2 && value.is_array()) {
                const int a = topo.atomId(pair[0]);
                const int b = topo.atomId(pair[1]);
                umatrix[a * n_types + b] = umatrix[b * n_types + a] = combine(value, topo);
            }
        }
    }
    inline double exact(const Particle& a, const Particle& b, double r2) const
    {
        return umatrix[a.id * n_types + b.id](a, b, r2);
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const { return exact(a, b, r2); }
    std::function<double(const Particle&)> selfEnergy() const { return self_energy; }
};

/** Per-pair Andrea tables in r² with auto-detected [rmin, rmax]; src/potentials.cpp:1453-1595 */
class SplinedPotential : public FunctorPotential
{
  public:
    struct KnotData : SplineData
    {
        bool hardsphere_repulsion = false;
    };

  private:
    std::vector<KnotData> matrix_of_knots;
    bool hardsphere_repulsion = false;
    static constexpr int max_iterations = 1000000;
    static constexpr double dr = 1e-2;

    double findLowerDistance(const Particle& p1, const Particle& p2, double threshold, double rmin) const
    {
        int it = 0;
        while (rmin >= dr) {
            if (it++ == max_iterations) {
                throw std::runtime_error("Pair potential spline error: cannot determine minimum distance");
            }
            const double u = std::fabs(exact(p1, p2, rmin * rmin));
            if (u > threshold * 1.1) {
                rmin += dr;
            }
            else if (u < threshold / 1.1) {
                rmin -= dr;
            }
            else {
                break;
            }
        }
        return rmin;
    }
    double findUpperDistance(const Particle& p1, const Particle& p2, double threshold, double rmax) const
    {
        int it = 0;
        while (rmax >= dr) {
            if (it++ == max_iterations) {
                throw std::runtime_error("Pair potential spline error: cannot determine maximum distance");
            }
            const double u = exact(p1, p2, rmax * rmax);
            if (std::fabs(u) > threshold) {
                rmax += dr;
            }
            else {
                break;
            }
        }
        return rmax;
    }

  public:
    void from_json(const Json& js, const Topology& topo)
    {
        FunctorPotential::from_json(js, topo);
        Andrea spline;
        spline.setTolerance(js.value("utol", 1e-3), js.value("ftol", 1e-2));
        hardsphere_repulsion = js.value("hardsphere", false);
        const double energy_at_rmin = js.value("u_at_rmin", 20.0);
        const double energy_at_rmax = js.value("u_at_rmax", 1e-6);
        matrix_of_knots.assign(n_types * n_types, KnotData());
        for (size_t i = 0; i < n_types; ++i) {
            for (size_t k = 0; k <= i; ++k) {
                if (topo.atoms[i].implicit || topo.atoms[k].implicit) {
                    continue;
                }
                const Particle p1 = topo.makeParticle(static_cast<int>(i));
                const Particle p2 = topo.makeParticle(static_cast<int>(k));
                double rmin = 0.5 * (topo.atoms[i].sigma + topo.atoms[k].sigma);
                double rmax = js.value("rmax", rmin * 10);
                if (const auto* it = js.find("cutoff_g2g")) {
                    if (it->is_number()) {
                        rmax = it->number();
                    }
                    else if (it->is_object()) {
                        rmax = it->at("default").number();
                    }
                }
                rmin = findLowerDistance(p1, p2, energy_at_rmin, rmin);
                rmax = findUpperDistance(p1, p2, energy_at_rmax, rmax);
                KnotData kd;
                static_cast<SplineData&>(kd) =
                    spline.generate([&](double r2) { return exact(p1, p2, r2); }, rmin * rmin, rmax * rmax);
                kd.hardsphere_repulsion = hardsphere_repulsion;
                if (Andrea::eval(kd, kd.rmin2 + dr) < 0) {
                    kd.hardsphere_repulsion = false;
                }
                matrix_of_knots[i * n_types + k] = matrix_of_knots[k * n_types + i] = kd;
            }
        }
    }
    inline double operator()(const Particle& a, const Particle& b, double r2) const
    {
        const auto& knots = matrix_of_knots[a.id * n_types + b.id];
        if (r2 >= knots.rmax2) {
            return 0.0;
        }
        if (r2 > knots.rmin2) {
            return Andrea::eval(knots, r2);
        }
        if (knots.hardsphere_repulsion) {
            return pc::infty;
        }
        return exact(a, b, r2);
    }
    const KnotData& knots(int a, int b) const { return matrix_of_knots[a * n_types + b]; }
};

} // namespace oracle

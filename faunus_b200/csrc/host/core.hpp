// Host-side basics shared by the B200 adaptor layer and (as caller-side scaffolding) by the CPU
// oracle: 3-vector, physical constants and units, and the reference's random number wrapper.
//
// Mirrors: src/core.h:23 (Point), src/units.h:10-47,91-260 (constants, unit literals),
// src/random.h:17-66 (Random = std::mt19937 + libstdc++ distributions), src/core.cpp:249-261
// (randomUnitVector), src/core.cpp:430-506 (Electrolyte).
#pragma once
#include "minijson.hpp"
#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <numeric>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <vector>

namespace fb {

struct Point
{
    double x = 0, y = 0, z = 0;
    Point() = default;
    Point(double x, double y, double z)
        : x(x)
        , y(y)
        , z(z)
    {
    }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Point operator+(const Point& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Point operator-(const Point& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Point operator-() const { return {-x, -y, -z}; }
    Point operator*(double s) const { return {x * s, y * s, z * s}; }
    Point operator/(double s) const { return {x / s, y / s, z / s}; }
    Point& operator+=(const Point& o)
    {
        x += o.x;
        y += o.y;
        z += o.z;
        return *this;
    }
    Point& operator-=(const Point& o)
    {
        x -= o.x;
        y -= o.y;
        z -= o.z;
        return *this;
    }
    Point cwiseProduct(const Point& o) const { return {x * o.x, y * o.y, z * o.z}; }
    double dot(const Point& o) const { return x * o.x + y * o.y + z * o.z; }
    Point cross(const Point& o) const
    {
        return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x};
    }
    double squaredNorm() const { return x * x + y * y + z * z; }
    double norm() const { return std::sqrt(squaredNorm()); }
    double maxCoeff() const { return std::max(x, std::max(y, z)); }
    double prod() const { return x * y * z; }
    bool operator==(const Point& o) const { return x == o.x && y == o.y && z == o.z; }
};
inline Point operator*(double s, const Point& p)
{
    return p * s;
}

inline Point pointFromJson(const Json& j)
{
    if (j.is_number()) {
        const double v = j.number();
        return {v, v, v};
    }
    const auto v = j.numbers();
    if (v.size() != 3) {
        throw std::runtime_error("3-vector expected");
    }
    return {v[0], v[1], v[2]};
}
inline Json pointToJson(const Point& p)
{
    Json j = Json::array();
    j.push_back(p.x);
    j.push_back(p.y);
    j.push_back(p.z);
    return j;
}

/** Physical constants; values are the reference's (non-CODATA) ones, src/units.h:10-24 */
namespace pc {
constexpr double infty = std::numeric_limits<double>::infinity();
constexpr double neg_infty = -std::numeric_limits<double>::infinity();
constexpr double epsilon_dbl = std::numeric_limits<double>::epsilon();
constexpr double max_value = std::numeric_limits<double>::max();
constexpr double max_exp_argument = 709.782712893384;
constexpr double pi = 3.141592653589793238462643383279502884;
constexpr double vacuum_permittivity = 8.85419e-12;
constexpr double elementary_charge = 1.602177e-19;
constexpr double boltzmann_constant = 1.380658e-23;
constexpr double avogadro = 6.022137e23;
inline double temperature = 298.15; //!< global, set from the `temperature` input key (faunus.cpp:106)

inline double kT()
{
    return temperature * boltzmann_constant;
}
/** Bjerrum length in angstrom, src/units.h:37-41 */
inline double bjerrumLength(double relative_dielectric_constant)
{
    return elementary_charge * elementary_charge /
           (4 * pi * vacuum_permittivity * relative_dielectric_constant * 1e-10 * kT());
}
} // namespace pc

/** Unit conversion factors to internal units (angstrom, kT, particles/angstrom^3); src/units.h:91-260 */
namespace units {
inline double kJmol()
{
    return 1.0 / pc::kT() / pc::avogadro * 1e3;
}
constexpr double liter = 1e27;
constexpr double molar = pc::avogadro / liter; //!< 1 mol/l in particles per cubic angstrom
constexpr double m3 = 1e30;
inline double Pa()
{
    return 1.0 / pc::kT() / m3;
}
inline double atm()
{
    return 101325.0 * Pa();
}
inline double bar()
{
    return 100000.0 * Pa();
}
} // namespace units

/**
 * Random number generator with the reference's exact engine and distributions so that a fixed
 * seed gives the reference's proposal stream (libstdc++): src/random.h:34-66.
 */
class Random
{
    std::uniform_real_distribution<double> dist01{0.0, 1.0};

  public:
    std::mt19937 engine; //!< default seed 5489 (`random: {seed: fixed}`)
    double operator()() { return dist01(engine); }
    int range(int min, int max) { return std::uniform_int_distribution<>(min, max)(engine); }
    /** index version of Random::sample(begin,end): draws even when n == 1 */
    int sampleIndex(int n) { return range(0, n - 1); }
    std::string state() const
    {
        std::ostringstream o;
        o << engine;
        return o.str();
    }
    void setState(const std::string& s)
    {
        std::istringstream i(s);
        i >> engine;
    }
};

inline Random random_global; //!< Faunus::random (src/random.cpp:62)

/** Rejection sampling in the unit cube, then normalised; src/core.cpp:249-261 */
inline Point randomUnitVector(Random& rand, const Point& directions = {1, 1, 1})
{
    constexpr double squared_radius = 0.25;
    double squared_norm;
    Point position;
    do {
        for (int i = 0; i < 3; i++) {
            position[i] = (rand() - 0.5) * directions[i];
        }
        squared_norm = position.squaredNorm();
    } while (squared_norm > squared_radius);
    return position / std::sqrt(squared_norm);
}

/**
 * Unit quaternion rotation, restating Eigen 3.4's `Quaterniond(AngleAxisd(angle, axis)) * v`
 * (Quaternion::_transformVector): uv = 2 (q.vec x v); v + w uv + q.vec x uv.
 * Used by src/geometry.h:613-626 and src/move.cpp:1652-1668.
 */
struct Quaternion
{
    double w = 1;
    Point vec;
    Quaternion() = default;
    Quaternion(double angle, const Point& unit_axis)
        : w(std::cos(0.5 * angle))
        , vec(unit_axis * std::sin(0.5 * angle))
    {
    }
    Point operator*(const Point& v) const
    {
        Point uv = vec.cross(v);
        uv += uv;
        return v + uv * w + vec.cross(uv);
    }
};

/** Salt description → Debye length; src/core.cpp:411-506 */
class Electrolyte
{
    double ionic_strength = 0;
    double molarity = 0;

  public:
    Electrolyte(double molarity, const std::vector<int>& valencies)
        : molarity(molarity)
    {
        int sum_pos = 0;
        int sum_neg = 0;
        for (int z : valencies) {
            if (z > 0) {
                sum_pos += z;
            }
            else {
                sum_neg -= z;
            }
        }
        if (sum_pos == 0 || sum_neg == 0) {
            throw std::runtime_error("cannot resolve stoichiometry; did you provide both + and - ions?");
        }
        const int g = std::gcd(sum_pos, sum_neg);
        double nu_z2 = 0;
        for (int z : valencies) {
            const int nu = (z > 0) ? sum_neg / g : sum_pos / g;
            nu_z2 += nu * z * z;
        }
        ionic_strength = 0.5 * molarity * nu_z2;
    }
    Electrolyte(double debye_length, double bjerrum_length)
    {
        ionic_strength = molarity =
            std::pow(1.0 / debye_length, 2) / (8.0 * pc::pi * bjerrum_length * units::molar);
    }
    double ionicStrength() const { return ionic_strength; }
    double getMolarity() const { return molarity; }
    double debyeLength(double bjerrum_length) const
    {
        return 1.0 / std::sqrt(8.0 * pc::pi * bjerrum_length * ionic_strength * units::molar);
    }
};

inline std::optional<Electrolyte> makeElectrolyte(const Json& j)
{
    if (const auto* it = j.find("debyelength")) {
        const double debye_length = it->number();
        const double lB = pc::bjerrumLength(j.at("epsr").number());
        return Electrolyte(debye_length, lB);
    }
    const double molarity = j.value("molarity", j.value("salt", 0.0));
    if (molarity > 0.0) {
        std::vector<int> valencies = {1, -1};
        if (const auto* v = j.find("valencies")) {
            valencies.clear();
            for (const auto& x : v->items()) {
                valencies.push_back(x.integer());
            }
        }
        return Electrolyte(molarity, valencies);
    }
    return std::nullopt;
}

/** split on whitespace (src/auxiliary.h splitConvert) */
inline std::vector<std::string> splitWords(const std::string& s)
{
    std::vector<std::string> out;
    std::istringstream in(s);
    std::string w;
    while (in >> w) {
        out.push_back(w);
    }
    return out;
}

} // namespace fb

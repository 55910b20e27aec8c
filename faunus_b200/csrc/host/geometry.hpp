// Simulation cell with the reference's `Geometry::Chameleon` distance and boundary rules for the
// in-scope shapes: cuboid (PBC xyz), slit (PBC xy) and sphere (no PBC).
// Mirrors src/geometry.h:407-485 (boundary, vdist, sqdist), src/geometry.cpp:42-140 (Cuboid
// setVolume/boundary/randompos/collision), :229-266 (Sphere), :965-980 (len_or_zero).
#pragma once
#include "core.hpp"

namespace fb {

enum class VolumeMethod
{
    ISOTROPIC,
    ISOCHORIC,
    XY,
    Z
};

inline VolumeMethod volumeMethodFromString(const std::string& s)
{
    if (s == "isotropic") {
        return VolumeMethod::ISOTROPIC;
    }
    if (s == "isochoric") {
        return VolumeMethod::ISOCHORIC;
    }
    if (s == "xy") {
        return VolumeMethod::XY;
    }
    if (s == "z") {
        return VolumeMethod::Z;
    }
    throw std::runtime_error("invalid volume scaling method '" + s + "'");
}

class Geometry
{
  public:
    enum class Type
    {
        CUBOID,
        SLIT,
        SPHERE
    };
    Type type = Type::CUBOID;

  private:
    Point len, len_half, len_inv, len_or_zero;
    std::array<bool, 3> periodic{true, true, true};
    double radius = 0; //!< sphere only

    static int anint(double x) { return static_cast<int>(x > 0.0 ? x + 0.5 : x - 0.5); }

    void updateCache()
    {
        len_half = len * 0.5;
        len_inv = {1.0 / len.x, 1.0 / len.y, 1.0 / len.z};
        for (int i = 0; i < 3; ++i) {
            len_or_zero[i] = periodic[i] ? len[i] : 0.0;
        }
    }

  public:
    Geometry() { setLength({0, 0, 0}); }

    void setLength(const Point& l)
    {
        len = l;
        updateCache();
    }
    const Point& getLength() const { return len; }
    const Point& lengthOrZero() const { return len_or_zero; } //!< L if periodic else 0, per axis
    const Point& halfLength() const { return len_half; }
    bool isPeriodic(int axis) const { return periodic[axis]; }
    double getRadius() const { return radius; }

    double getVolume() const
    {
        if (type == Type::SPHERE) {
            return 4.0 / 3.0 * pc::pi * radius * radius * radius;
        }
        return len.x * len.y * len.z;
    }

    /** @return scaling factors for positions; src/geometry.cpp:52-88, :229-243 */
    Point setVolume(double volume, VolumeMethod method = VolumeMethod::ISOTROPIC)
    {
        if (type == Type::SPHERE) {
            if (method != VolumeMethod::ISOTROPIC) {
                throw std::invalid_argument("unsupported volume scaling method for the spherical geometry");
            }
            const double old_radius = radius;
            radius = std::cbrt(volume / (4.0 / 3.0 * pc::pi));
            setLength({2 * radius, 2 * radius, 2 * radius});
            const double s = radius / old_radius;
            return {s, s, s};
        }
        const double old_volume = getVolume();
        double alpha;
        Point scaling;
        switch (method) {
        case VolumeMethod::ISOTROPIC:
            alpha = std::cbrt(volume / old_volume);
            scaling = {alpha, alpha, alpha};
            break;
        case VolumeMethod::XY:
            alpha = std::sqrt(volume / old_volume);
            scaling = {alpha, alpha, 1.0};
            break;
        case VolumeMethod::Z:
            alpha = volume / old_volume;
            scaling = {1.0, 1.0, alpha};
            break;
        case VolumeMethod::ISOCHORIC:
            alpha = std::cbrt(volume / old_volume);
            scaling = {alpha, alpha, 1 / (alpha * alpha)};
            break;
        default:
            throw std::invalid_argument("unsupported volume scaling method");
        }
        setLength(len.cwiseProduct(scaling));
        return scaling;
    }

    /** Wrap into the cell; src/geometry.h:407-427 */
    void boundary(Point& a) const
    {
        for (int i = 0; i < 3; ++i) {
            if (periodic[i] && std::fabs(a[i]) > len_half[i]) {
                a[i] -= len[i] * anint(a[i] * len_inv[i]);
            }
        }
    }

    /** Minimum image distance vector a-b; src/geometry.h:429-458 */
    Point vdist(const Point& a, const Point& b) const
    {
        Point d = a - b;
        for (int i = 0; i < 3; ++i) {
            if (periodic[i]) {
                if (d[i] > len_half[i]) {
                    d[i] -= len[i];
                }
                else if (d[i] < -len_half[i]) {
                    d[i] += len[i];
                }
            }
        }
        return d;
    }

    /** Squared minimum image distance, single fold; src/geometry.h:460-470 */
    double sqdist(const Point& a, const Point& b) const
    {
        Point d = a - b;
        for (int i = 0; i < 3; ++i) {
            d[i] = std::fabs(d[i]);
            d[i] = d[i] - len_or_zero[i] * static_cast<double>(d[i] > len_half[i]);
        }
        return d.x * d.x + d.y * d.y + d.z * d.z;
    }

    void randompos(Point& m, Random& rand) const
    {
        if (type == Type::SPHERE) { // src/geometry.cpp:258-266
            const double r2 = radius * radius;
            const double d = 2 * radius;
            do {
                m.x = (rand() - 0.5) * d;
                m.y = (rand() - 0.5) * d;
                m.z = (rand() - 0.5) * d;
            } while (m.squaredNorm() > r2);
            return;
        }
        m.x = (rand() - 0.5) * len.x; // src/geometry.cpp:125-130
        m.y = (rand() - 0.5) * len.y;
        m.z = (rand() - 0.5) * len.z;
    }

    bool collision(const Point& a) const
    {
        if (type == Type::SPHERE) {
            return a.squaredNorm() > radius * radius;
        }
        return std::fabs(a.x) > len_half.x || std::fabs(a.y) > len_half.y ||
               std::fabs(a.z) > len_half.z;
    }

    static Geometry fromJson(const Json& j)
    {
        Geometry g;
        const std::string type = j.at("type").string();
        if (type == "cuboid" || type == "slit") {
            g.type = (type == "slit") ? Type::SLIT : Type::CUBOID;
            g.periodic = {true, true, type == "cuboid"};
            g.setLength(pointFromJson(j.at("length")));
        }
        else if (type == "sphere") {
            g.type = Type::SPHERE;
            g.periodic = {false, false, false};
            g.radius = j.at("radius").number();
            g.setLength({2 * g.radius, 2 * g.radius, 2 * g.radius});
        }
        else {
            throw std::runtime_error("geometry '" + type + "' is outside the B200 hot-path scope "
                                     "(cuboid, slit, sphere)");
        }
        return g;
    }

    Json toJson() const
    {
        Json j = Json::object();
        if (type == Type::SPHERE) {
            j["type"] = "sphere";
            j["radius"] = radius;
        }
        else {
            j["type"] = (type == Type::SLIT) ? "slit" : "cuboid";
            j["length"] = pointToJson(len);
        }
        return j;
    }
};

} // namespace fb

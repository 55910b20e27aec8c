// Space / Group / Change: the state and change bookkeeping that every energy term reads.
// Mirrors src/space.h:29-75 (Change), :92-373 (Space), src/space.cpp:149-181 (addGroup), :199-240
// (sync), :249-302 (scaleVolume), :449-523 (JSON), :757-946 (InsertMoleculesInSpace),
// src/group.h:60-177 + src/group.cpp:41-140 (Group), src/molecule.cpp:842-907 (RandomInserter),
// src/geometry.h:504-626 (massCenter, rotate).
//
// Groups are index ranges [begin, begin+size) with capacity into Space::particles (the reference
// uses iterator ranges; the contract — inactive particles sit between size and capacity and never
// interact — is the same).
#pragma once
#include "geometry.hpp"
#include "topology.hpp"
#include <functional>

namespace fb {

struct Change
{
    using index_type = std::size_t;
    bool everything = false;
    bool volume_change = false;
    bool matter_change = false;
    bool moved_to_moved_interactions = true;
    bool disable_translational_entropy = false;

    struct GroupChange
    {
        index_type group_index = 0;
        bool dNatomic = false;
        bool dNswap = false;
        bool internal = false;
        bool all = false;
        std::vector<index_type> relative_atom_indices;
        bool operator<(const GroupChange& other) const { return group_index < other.group_index; }
    };
    std::vector<GroupChange> groups;

    void clear()
    {
        everything = volume_change = matter_change = disable_translational_entropy = false;
        moved_to_moved_interactions = true;
        groups.clear();
    }
    bool empty() const
    {
        return !(everything || volume_change || matter_change) && groups.empty();
    }
    explicit operator bool() const { return !empty(); }
};

struct Group
{
    int id = -1;            //!< molecule type id
    size_t begin = 0;       //!< index of first particle in Space::particles
    size_t size_ = 0;       //!< active particles
    size_t capacity_ = 0;   //!< active + inactive particles
    Point mass_center;      //!< only meaningful for molecular groups
    bool atomic = false;    //!< cached MoleculeData::atomic
    bool rigid = false;     //!< cached MoleculeData::rigid
    bool compressible = false;

    size_t size() const { return size_; }
    size_t capacity() const { return capacity_; }
    bool empty() const { return size_ == 0; }
    bool isAtomic() const { return atomic; }
    bool isMolecular() const { return !atomic; }
    bool isFull() const { return size_ == capacity_; }
    void resize(size_t n)
    {
        if (n > capacity_) {
            throw std::runtime_error("group resize beyond capacity");
        }
        size_ = n;
    }
};

class Space
{
  public:
    enum class Selection
    {
        ALL,
        ACTIVE,
        INACTIVE
    };
    std::shared_ptr<const Topology> topology;
    Geometry geometry;
    ParticleVector particles;
    std::vector<Group> groups;

    const MoleculeData& traits(const Group& g) const { return topology->molecules.at(g.id); }
    const AtomData& traits(const Particle& p) const { return topology->atoms.at(p.id); }

    Particle& at(const Group& g, size_t i) { return particles[g.begin + i]; }
    const Particle& at(const Group& g, size_t i) const { return particles[g.begin + i]; }

    size_t numParticles(Selection sel = Selection::ACTIVE) const
    {
        size_t n = 0;
        for (const auto& g : groups) {
            n += (sel == Selection::ALL) ? g.capacity() : g.size();
        }
        return n;
    }

    /** Indices of groups of type `molid` (src/space.h findMolecules) */
    std::vector<size_t> findMolecules(int molid, Selection sel) const
    {
        std::vector<size_t> out;
        for (size_t i = 0; i < groups.size(); ++i) {
            const auto& g = groups[i];
            if (g.id != molid) {
                continue;
            }
            bool ok = true;
            switch (sel) {
            case Selection::ALL:
                break;
            case Selection::ACTIVE: // src/space.h:243-265: active ⇔ size == capacity
                ok = g.size() == g.capacity();
                break;
            case Selection::INACTIVE:
                ok = g.size() != g.capacity();
                break;
            }
            if (ok) {
                out.push_back(i);
            }
        }
        return out;
    }

    /** Σ m_i wrap(r_i + shift) / Σ m_i − shift, wrapped; src/geometry.h:504-526 */
    Point massCenter(size_t first, size_t last, const Point& shift, bool use_boundary = true) const
    {
        double weight_sum = 0.0;
        Point center;
        for (size_t i = first; i < last; ++i) {
            const double w = topology->atoms[particles[i].id].mw;
            Point shifted = particles[i].pos + shift;
            if (use_boundary) {
                geometry.boundary(shifted);
            }
            center += shifted * w;
            weight_sum += w;
        }
        if (std::fabs(weight_sum) > pc::epsilon_dbl) {
            center = center / weight_sum - shift;
            if (use_boundary) {
                geometry.boundary(center);
            }
            return center;
        }
        return {};
    }
    Point massCenter(const Group& g, const Point& shift) const
    {
        return massCenter(g.begin, g.begin + g.size(), shift);
    }

    /** Group::translate, src/group.cpp:115-123 */
    void translate(Group& g, const Point& displacement)
    {
        g.mass_center += displacement;
        geometry.boundary(g.mass_center);
        for (size_t i = 0; i < g.size(); ++i) {
            auto& pos = at(g, i).pos;
            pos += displacement;
            geometry.boundary(pos);
        }
    }

    /** Group::rotate → Geometry::rotate with shift = −mass_center; src/geometry.h:613-626 */
    void rotate(Group& g, const Quaternion& q)
    {
        const Point shift = -g.mass_center;
        for (size_t i = 0; i < g.size(); ++i) {
            auto& pos = at(g, i).pos;
            pos += shift;
            geometry.boundary(pos);
            pos = (q * pos) - shift;
            geometry.boundary(pos);
        }
    }

    /** Group::unwrap, src/group.h:349-356 */
    void unwrap(Group& g)
    {
        if (g.isMolecular()) {
            for (size_t i = 0; i < g.size(); ++i) {
                auto& pos = at(g, i).pos;
                pos = g.mass_center + geometry.vdist(pos, g.mass_center);
            }
        }
    }

    /** src/space.cpp:149-181 */
    Group& addGroup(int molid, const ParticleVector& new_particles)
    {
        if (new_particles.empty()) {
            throw std::runtime_error("cannot add empty molecule");
        }
        const auto& mol = topology->molecules.at(molid);
        Group g;
        g.id = molid;
        g.begin = particles.size();
        g.size_ = g.capacity_ = new_particles.size();
        g.atomic = mol.atomic;
        g.rigid = mol.rigid;
        g.compressible = mol.compressible;
        particles.insert(particles.end(), new_particles.begin(), new_particles.end());
        if (g.isMolecular()) {
            if (new_particles.size() != mol.atoms.size()) {
                throw std::runtime_error("particle size mismatch");
            }
            g.mass_center = massCenter(g, -new_particles.front().pos);
        }
        else if (new_particles.size() % mol.atoms.size() != 0) {
            throw std::runtime_error("indivisible by atomic group size: " + mol.name);
        }
        groups.push_back(g);
        return groups.back();
    }

    /**
     * Copy from `other` what `change` says was modified; src/space.cpp:199-240.
     */
    void sync(const Space& other, const Change& change)
    {
        if (&other == this || change.empty()) {
            return;
        }
        if (particles.size() != other.particles.size() || groups.size() != other.groups.size()) {
            throw std::runtime_error("space sync error");
        }
        if (change.volume_change || change.everything) {
            geometry = other.geometry;
        }
        if (change.everything) {
            particles = other.particles;
            groups = other.groups;
            return;
        }
        for (const auto& changed : change.groups) {
            auto& group = groups.at(changed.group_index);
            const auto& other_group = other.groups.at(changed.group_index);
            if (group.capacity() != other_group.capacity()) {
                throw std::runtime_error("Group::shallowCopy: capacity mismatch");
            }
            group.size_ = other_group.size_; // shallow copy: size, id, mass center
            group.id = other_group.id;
            group.mass_center = other_group.mass_center;
            if (changed.all) { // deep copy incl. inactive particles
                std::copy(other.particles.begin() + other_group.begin,
                          other.particles.begin() + other_group.begin + other_group.capacity(),
                          particles.begin() + group.begin);
            }
            else {
                for (auto i : changed.relative_atom_indices) {
                    if (i >= group.capacity()) {
                        throw std::out_of_range("atom index out of range in sync");
                    }
                    particles[group.begin + i] = other.particles[other_group.begin + i];
                }
            }
        }
    }

    /**
     * Replace all particles and recompute mass centers of molecular groups;
     * Space::updateParticles, src/space.h:193-217 (used by tempering and tests).
     */
    void updateMassCenters()
    {
        for (auto& g : groups) {
            if (g.isMolecular() && !g.empty()) {
                g.mass_center = massCenter(g, -at(g, 0).pos);
            }
        }
    }

    /** src/space.cpp:249-302 */
    Point scaleVolume(double new_volume, VolumeMethod method = VolumeMethod::ISOTROPIC)
    {
        for (auto& g : groups) {
            unwrap(g);
        }
        const Point scale = geometry.setVolume(new_volume, method);
        auto scale_position = [&](Particle& p) {
            p.pos = p.pos.cwiseProduct(scale);
            geometry.boundary(p.pos);
        };
        for (auto& g : groups) {
            if (g.empty()) {
                continue;
            }
            if (g.isAtomic()) {
                for (size_t i = 0; i < g.size(); ++i) {
                    scale_position(at(g, i));
                }
            }
            else {
                const Point original_mass_center = g.mass_center;
                if (g.compressible) {
                    for (size_t i = 0; i < g.size(); ++i) {
                        scale_position(at(g, i));
                    }
                    g.mass_center = massCenter(g, -original_mass_center);
                }
                else {
                    g.mass_center = g.mass_center.cwiseProduct(scale);
                    geometry.boundary(g.mass_center);
                    const Point displacement = g.mass_center - original_mass_center;
                    for (size_t i = 0; i < g.size(); ++i) {
                        auto& pos = at(g, i).pos;
                        pos += displacement;
                        geometry.boundary(pos);
                    }
                }
            }
        }
        return scale;
    }

    // ----- JSON state (src/space.cpp:449-523, src/group.cpp:230-250, src/particle.cpp) -----

    Json toJson() const
    {
        Json j = Json::object();
        j["geometry"] = geometry.toJson();
        Json jg = Json::array();
        for (const auto& g : groups) {
            Json x = Json::object();
            x["id"] = g.id;
            x["cm"] = pointToJson(g.mass_center);
            x["atomic"] = g.atomic;
            x["compressible"] = g.compressible;
            x["size"] = g.size();
            if (g.capacity() > g.size()) {
                x["capacity"] = g.capacity();
            }
            jg.push_back(x);
        }
        j["groups"] = jg;
        Json jp = Json::array();
        for (const auto& p : particles) {
            Json x = Json::object();
            x["id"] = p.id;
            x["pos"] = pointToJson(p.pos);
            x["q"] = p.charge;
            jp.push_back(x);
        }
        j["particles"] = jp;
        return j;
    }

    /** Load `groups` + `particles` (+ `geometry`) as written by the reference's savestate */
    void loadState(const Json& j)
    {
        particles.clear();
        groups.clear();
        geometry = Geometry::fromJson(j.at("geometry"));
        for (const auto& jp : j.at("particles").items()) {
            Particle p;
            p.id = jp.at("id").integer();
            p.charge = jp.value("q", 0.0);
            p.pos = pointFromJson(jp.at("pos"));
            particles.push_back(p);
        }
        size_t begin = 0;
        for (const auto& jg : j.at("groups").items()) {
            Group g;
            g.id = jg.at("id").integer();
            const auto& mol = topology->molecules.at(g.id);
            g.begin = begin;
            g.size_ = static_cast<size_t>(jg.at("size").integer());
            g.capacity_ = static_cast<size_t>(jg.value("capacity", static_cast<int>(g.size_)));
            g.mass_center = pointFromJson(jg.at("cm"));
            g.atomic = mol.atomic;
            g.rigid = mol.rigid;
            g.compressible = jg.value("compressible", mol.compressible);
            groups.push_back(g);
            begin += g.capacity_;
        }
        if (begin != particles.size()) {
            throw std::runtime_error("load error");
        }
        for (const auto& g : groups) { // src/space.cpp:499-516
            if (!g.empty() && g.isMolecular()) {
                const double d2 = geometry.sqdist(g.mass_center, massCenter(g, -g.mass_center));
                if (d2 > 1e-4) {
                    throw std::runtime_error("couldn't calculate mass center for " + traits(g).name);
                }
            }
        }
    }
};

/**
 * Random insertion of a molecule with the reference's RNG consumption (global `Faunus::random`);
 * src/molecule.cpp:842-907. `dir` is the inserter direction (molecule `insdir`, overridden by
 * Widom's `dir`).
 */
struct RandomInserter
{
    Point dir{1, 1, 1};
    Point offset{0, 0, 0};
    bool rotate = true;
    bool keep_positions = false;
    bool allow_overlap = false;
    int max_trials = 20000;

    static RandomInserter fromMolecule(const MoleculeData& mol)
    {
        RandomInserter ins;
        ins.dir = mol.insdir;
        ins.offset = mol.insoffset;
        ins.rotate = mol.rotate;
        ins.keep_positions = mol.keeppos;
        return ins;
    }

    ParticleVector operator()(const Space& spc, const MoleculeData& mol, Random& random) const
    {
        const auto& geo = spc.geometry;
        ParticleVector particles = mol.structure; // single conformation: no RNG draw (libstdc++
                                                  // discrete_distribution with one weight)
        if (particles.empty()) {
            throw std::runtime_error("nothing to insert for molecule '" + mol.name + "'");
        }
        auto overlap = [&](const ParticleVector& v) {
            return std::any_of(v.begin(), v.end(), [&](const Particle& p) { return geo.collision(p.pos); });
        };
        if (keep_positions) {
            if (!overlap(particles)) {
                return particles;
            }
            throw std::runtime_error("inserted molecule does not fit in container");
        }
        for (int attempts = 0; attempts < max_trials; attempts++) {
            if (mol.atomic) {
                for (auto& particle : particles) {
                    if (rotate) { // rotation of isotropic particles: RNG consumed, no effect
                        (void)randomUnitVector(random); // g++ evaluates call arguments right-to-left
                        (void)(2.0 * pc::pi * random());
                    }
                    geo.randompos(particle.pos, random);
                    particle.pos = particle.pos.cwiseProduct(dir) + offset;
                    geo.boundary(particle.pos);
                }
            }
            else {
                // translate mass center to origin without PBC (Geometry::translateToOrigin)
                double wsum = 0;
                Point cm;
                for (const auto& p : particles) {
                    const double w = spc.topology->atoms[p.id].mw;
                    cm += p.pos * w;
                    wsum += w;
                }
                cm = cm / wsum;
                for (auto& p : particles) {
                    p.pos -= cm;
                }
                if (rotate) {
                    // `rotator.set(2π·random(), randomUnitVector(random))`: a GCC build evaluates
                    // the arguments right-to-left, i.e. the axis is drawn before the angle
                    const Point axis = randomUnitVector(random);
                    const double angle = 2.0 * pc::pi * random();
                    const Quaternion q(angle, axis);
                    for (auto& p : particles) {
                        p.pos = q * p.pos;
                    }
                }
                Point new_mass_center;
                geo.randompos(new_mass_center, random);
                new_mass_center = new_mass_center.cwiseProduct(dir) + offset;
                for (auto& p : particles) {
                    p.pos += new_mass_center;
                    geo.boundary(p.pos);
                }
            }
            if (allow_overlap || !overlap(particles)) {
                return particles;
            }
        }
        throw std::runtime_error("Max. # of overlap checks reached upon insertion.");
    }
};

/** `insertmolecules` section; src/space.cpp:767-946 (N and inactive only) */
inline void insertMolecules(const Json& j, Space& spc, Random& random)
{
    spc.particles.clear();
    spc.groups.clear();
    for (const auto& item : j.items()) {
        const auto& [molname, props] = item.single();
        const auto& mol = spc.topology->molecules.at(spc.topology->moleculeId(molname));
        if (!props.contains("N")) {
            throw std::runtime_error("insertmolecules: only `N` is supported for " + molname);
        }
        const auto num = static_cast<size_t>(props.at("N").integer());
        size_t num_inactive = 0;
        if (const auto* in = props.find("inactive")) {
            num_inactive = in->is_bool() ? (in->boolean() ? num : 0) : static_cast<size_t>(in->integer());
        }
        if (num_inactive > num) {
            throw std::runtime_error("too many inactive molecules requested");
        }
        const auto inserter = RandomInserter::fromMolecule(mol);
        if (mol.atomic) {
            ParticleVector repeated;
            for (size_t i = 0; i < num; ++i) {
                const auto p = inserter(spc, mol, random);
                repeated.insert(repeated.end(), p.begin(), p.end());
            }
            auto& g = spc.addGroup(mol.id, repeated);
            if (num_inactive > 0) {
                g.resize((num - num_inactive) * mol.atoms.size());
            }
        }
        else {
            for (size_t i = 0; i < num; ++i) {
                spc.addGroup(mol.id, inserter(spc, mol, random));
            }
            for (size_t i = 0; i < num_inactive; ++i) {
                auto& g = spc.groups[spc.groups.size() - 1 - i];
                spc.unwrap(g);
                g.resize(0);
            }
        }
    }
}

/** Space(json): topology must already be set; src/space.cpp:457-523 */
inline void spaceFromJson(const Json& j, Space& spc, Random& random)
{
    spc.geometry = Geometry::fromJson(j.at("geometry"));
    if (!j.contains("groups")) {
        insertMolecules(j.at("insertmolecules"), spc, random);
    }
    else {
        spc.loadState(j);
    }
}

} // namespace fb

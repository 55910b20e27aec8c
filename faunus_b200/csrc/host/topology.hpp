// Atom and molecule type tables (the reference's global `Faunus::atoms` / `Faunus::molecules`),
// here owned per simulation so several simulations (replicas, oracle + device) can coexist in a
// process. Mirrors src/atomdata.{h,cpp} (AtomData, from_json :149-196), src/molecule.{h,cpp}
// (MoleculeData :211-260, MoleculeBuilder :394-520, ExclusionsVicinity :140-203) and
// src/particle.h:220-250 (Particle) for the subset the ΔU path reads.
#pragma once
#include "core.hpp"
#include <map>
#include <memory>

namespace fb {

/** Particle = {atom type id, charge, position}; extensions (dipoles, …) are out of scope */
struct Particle
{
    int id = -1;
    double charge = 0.0;
    Point pos;
};
using ParticleVector = std::vector<Particle>;

struct AtomData
{
    std::string name;
    int id = -1;
    double charge = 0;
    double mw = 1;
    double sigma = 0;
    double activity = 0;
    std::optional<double> dp;    //!< translational displacement parameter (angstrom)
    std::optional<double> dprot; //!< rotational displacement parameter (rad)
    bool implicit = false;
    std::map<std::string, double> interaction; //!< named numeric pair parameters: sigma, eps, …

    /** named interaction parameter; NaN if absent (mirrors the NaN-on-missing behaviour tested at
     * src/potentials.cpp:1796-1851) */
    double parameter(const std::string& key) const
    {
        auto it = interaction.find(key);
        return it == interaction.end() ? std::nan("") : it->second;
    }
};

inline AtomData atomFromJson(const Json& j)
{
    const auto& [name, val] = j.single();
    AtomData a;
    a.name = name;
    static const char* known[] = {"alphax", "q",     "id",          "mu",        "mulen",
                                  "psc",    "mw",    "tension",     "tfe",       "hydrophobic",
                                  "implicit", "scattering_f0", "dp", "dprot",    "activity",
                                  "pactivity", "r",  "sigma"};
    a.charge = val.value("q", 0.0);
    a.mw = val.value("mw", 1.0);
    a.implicit = val.value("implicit", false);
    if (val.contains("dp")) {
        a.dp = val.at("dp").number();
    }
    if (val.contains("dprot")) {
        a.dprot = val.at("dprot").number();
    }
    if (val.contains("activity")) {
        a.activity = val.at("activity").number() * units::molar;
    }
    double sigma = val.value("sigma", 0.0);
    if (std::fabs(sigma) < 1e-20) {
        sigma = 2.0 * val.value("r", 0.0);
    }
    a.sigma = sigma;
    a.interaction["sigma"] = sigma;
    if (val.is_object()) {
        for (const auto& [key, v] : val.members()) {
            const bool is_known =
                std::any_of(std::begin(known), std::end(known), [&](const char* k) { return key == k; });
            if (!is_known && v.is_number()) {
                a.interaction[key] = v.number();
            }
        }
    }
    return a;
}

struct MoleculeData
{
    std::string name;
    int id = -1;
    bool atomic = false;
    bool rigid = false;
    bool compressible = false;
    bool implicit = false;
    std::vector<int> atoms;   //!< atom type ids in the molecule
    ParticleVector structure; //!< (single) conformation
    // RandomInserter settings (src/molecule.cpp:909-915)
    Point insdir{1, 1, 1};
    Point insoffset{0, 0, 0};
    bool rotate = true;
    bool keeppos = false;
    std::vector<unsigned char> excluded; //!< n x n symmetric exclusion matrix (may be empty)

    bool isAtomic() const { return atomic; }
    bool isMolecular() const { return !atomic; }
    bool isPairExcluded(int i, int j) const
    {
        if (excluded.empty()) {
            return false;
        }
        const auto n = atoms.size();
        return excluded[static_cast<size_t>(i) * n + j] != 0;
    }
};

struct Topology
{
    std::vector<AtomData> atoms;
    std::vector<MoleculeData> molecules;

    int atomId(const std::string& name) const
    {
        for (const auto& a : atoms) {
            if (a.name == name) {
                return a.id;
            }
        }
        throw std::runtime_error("unknown atom '" + name + "'");
    }
    int moleculeId(const std::string& name) const
    {
        for (const auto& m : molecules) {
            if (m.name == name) {
                return m.id;
            }
        }
        throw std::runtime_error("unknown molecule '" + name + "'");
    }
    Particle makeParticle(int atom_id, const Point& pos = {}) const
    {
        Particle p;
        p.id = atom_id;
        p.charge = atoms.at(atom_id).charge;
        p.pos = pos;
        return p;
    }
};

inline MoleculeData moleculeFromJson(const Json& j, const Topology& topo)
{
    const auto& [name, val] = j.single();
    MoleculeData m;
    m.name = name;
    m.atomic = val.value("atomic", false);
    m.rigid = val.value("rigid", false);
    m.compressible = val.value("compressible", false);
    m.implicit = val.value("implicit", false);
    if (m.implicit) {
        throw std::runtime_error("implicit molecules are outside the B200 hot-path scope");
    }
    if (const auto* p = val.find("insdir")) {
        m.insdir = pointFromJson(*p);
    }
    if (const auto* p = val.find("insoffset")) {
        m.insoffset = pointFromJson(*p);
    }
    m.rotate = val.value("rotate", true);
    m.keeppos = val.value("keeppos", false);
    if (m.atomic) { // src/molecule.cpp:453-486: one particle per listed atom name at the origin
        for (const auto& atomname : val.at("atoms").items()) {
            const int id = topo.atomId(atomname.string());
            m.atoms.push_back(id);
            m.structure.push_back(topo.makeParticle(id));
        }
    }
    else {
        const auto& structure = val.at("structure");
        if (!structure.is_array()) {
            throw std::runtime_error("only inline `structure` lists are supported for molecule " + name);
        }
        for (const auto& item : structure.items()) {
            const auto& [atomname, pos] = item.single();
            const int id = topo.atomId(atomname);
            m.atoms.push_back(id);
            m.structure.push_back(topo.makeParticle(id, pointFromJson(pos)));
        }
    }
    if (val.value("excluded_neighbours", 0) > 0) {
        throw std::runtime_error("excluded_neighbours needs the bonded topology (out of scope); "
                                 "use exclusionlist");
    }
    if (const auto* list = val.find("exclusionlist")) { // src/molecule.cpp:540-546
        const auto n = m.atoms.size();
        m.excluded.assign(n * n, 0);
        for (const auto& pair : list->items()) {
            const auto ij = pair.numbers();
            const auto i = static_cast<size_t>(ij.at(0));
            const auto k = static_cast<size_t>(ij.at(1));
            if (i >= n || k >= n) {
                throw std::runtime_error("exclusionlist index out of range");
            }
            m.excluded[i * n + k] = m.excluded[k * n + i] = 1;
        }
    }
    return m;
}

inline std::shared_ptr<Topology> topologyFromJson(const Json& j)
{
    auto topo = std::make_shared<Topology>();
    for (const auto& item : j.at("atomlist").items()) {
        auto a = atomFromJson(item);
        a.id = static_cast<int>(topo->atoms.size());
        topo->atoms.push_back(a);
    }
    for (const auto& item : j.at("moleculelist").items()) {
        auto m = moleculeFromJson(item, *topo);
        m.id = static_cast<int>(topo->molecules.size());
        topo->molecules.push_back(m);
    }
    return topo;
}

} // namespace fb

// The drop-in boundary: `Energy::EnergyTerm` and `Energy::Hamiltonian` with the reference's
// semantics, plus the cheap scalar host terms that compose with the non-bonded path.
// Mirrors src/externalpotential.h:21-43 (EnergyTerm), src/energy.cpp:1171-1268 (Hamiltonian:
// construction, energy with early break at maxenergy/NaN, init, updateState, sync),
// :800-860 (ContainerOverlap), :866-908 (Isobaric), :764-795 (Example2D).
//
// The non-bonded / Ewald / self-energy terms are supplied by a factory: the product build plugs in
// the B200 adaptor terms (faunus_b200/csrc/b200_terms.hpp), the oracle build its CPU restatement
// (oracle/nonbonded.hpp, oracle/ewald.hpp; wired up in oracle/oracle.cpp). Everything else in this file is caller-side scaffolding.
#pragma once
#include "space.hpp"
#include <chrono>

namespace fb {

class EnergyTerm
{
  public:
    enum class MonteCarloState
    {
        ACCEPTED,
        TRIAL,
        NONE
    };
    MonteCarloState state = MonteCarloState::NONE;
    std::string name;
    std::string citation_information;
    double seconds = 0; //!< accumulated wall time in energy() (the reference's `timer`)

    virtual double energy(const Change& change) = 0;
    virtual void to_json(Json&) const {}
    virtual void sync(EnergyTerm*, const Change&) {}
    virtual void init() {}
    virtual void updateState(const Change&) {}
    virtual void force(std::vector<Point>&) {} //!< no-op unless a term has forces (src/externalpotential.cpp:30)
    virtual ~EnergyTerm() = default;
};

/** Infinite energy if a changed particle left a non-cuboid container; src/energy.cpp:800-860 */
class ContainerOverlap : public EnergyTerm
{
    const Space& spc;

  public:
    explicit ContainerOverlap(const Space& spc)
        : spc(spc)
    {
        name = "ContainerOverlap";
    }
    double energy(const Change& change) override
    {
        if (!change || spc.geometry.type == Geometry::Type::CUBOID) {
            return 0.0;
        }
        if (change.volume_change || change.everything) {
            for (const auto& g : spc.groups) {
                for (size_t i = 0; i < g.size(); ++i) {
                    if (spc.geometry.collision(spc.at(g, i).pos)) {
                        return pc::infty;
                    }
                }
            }
            return 0.0;
        }
        for (const auto& gc : change.groups) {
            const auto& g = spc.groups.at(gc.group_index);
            if (gc.all) {
                for (size_t i = 0; i < g.size(); ++i) {
                    if (spc.geometry.collision(spc.at(g, i).pos)) {
                        return pc::infty;
                    }
                }
            }
            else {
                for (auto i : gc.relative_atom_indices) {
                    if (i < g.size() && spc.geometry.collision(spc.at(g, i).pos)) {
                        return pc::infty;
                    }
                }
            }
        }
        return 0.0;
    }
};

/** p V / kT − (N + 1) ln V; src/energy.cpp:866-908 */
class Isobaric : public EnergyTerm
{
    const Space& spc;
    double pressure = 0;

  public:
    Isobaric(const Json& j, const Space& spc)
        : spc(spc)
    {
        name = "isobaric";
        const std::pair<const char*, double> pressure_units[] = {
            {"P/atm", units::atm()},
            {"P/bar", units::bar()},
            {"P/kT", 1.0},
            {"P/mM", 1e-3 * units::molar},
            {"P/Pa", units::Pa()}};
        for (const auto& [key, factor] : pressure_units) {
            if (const auto* p = j.find(key)) {
                pressure = p->number() * factor;
                return;
            }
        }
        throw std::runtime_error("specify pressure");
    }
    double energy(const Change& change) override
    {
        if (change.volume_change || change.everything || change.matter_change) {
            int n = 0;
            for (const auto& g : spc.groups) {
                if (!g.empty()) {
                    n += g.isAtomic() ? static_cast<int>(g.size()) : 1;
                }
            }
            const double volume = spc.geometry.getVolume();
            return pressure * volume - static_cast<double>(n + 1) * std::log(volume);
        }
        return 0.0;
    }
    void to_json(Json& j) const override { j["P/kT"] = pressure; }
};

/** Hard-coded 1D/2D test potential on the first particle; src/energy.cpp:764-795 */
class Example2D : public EnergyTerm
{
    const Space& spc;
    double scale_energy = 1.0;
    bool use_2d = true;

  public:
    Example2D(const Json& j, const Space& spc)
        : spc(spc)
    {
        scale_energy = j.value("scale", 1.0);
        use_2d = j.value("2D", true);
        name = "Example2D";
    }
    double energy(const Change&) override
    {
        const auto& p = spc.particles.at(0).pos;
        double s = 1 + std::sin(2.0 * pc::pi * p.x) +
                   std::cos(2.0 * pc::pi * p.y) * static_cast<double>(use_2d);
        s *= scale_energy;
        if (p.x >= -2.00 && p.x <= -1.25) {
            return 1 * s;
        }
        if (p.x >= -1.25 && p.x <= -0.25) {
            return 2 * s;
        }
        if (p.x >= -0.25 && p.x <= 0.75) {
            return 3 * s;
        }
        if (p.x >= 0.75 && p.x <= 1.75) {
            return 4 * s;
        }
        if (p.x >= 1.75 && p.x <= 2.00) {
            return 5 * s;
        }
        return 1e10;
    }
    void to_json(Json& j) const override
    {
        j["scale"] = scale_energy;
        j["2D"] = use_2d;
    }
};

class Hamiltonian;

/**
 * Creates the term(s) for one `{name: config}` entry of the `energy` array that the base
 * Hamiltonian does not know (all `nonbonded*` flavours and their self-energy / Ewald siblings)
 * and appends them to the Hamiltonian in the reference's order (energy.h:462-477,
 * energy.cpp:1134-1160). Returns false if the name is unknown to the factory.
 */
using TermFactory =
    std::function<bool(Hamiltonian&, Space&, const std::string& name, const Json& config)>;

class Hamiltonian : public EnergyTerm
{
    std::vector<std::shared_ptr<EnergyTerm>> energy_terms;
    std::vector<double> latest_energies;
    double maximum_allowed_energy = pc::infty;

  public:
    Hamiltonian(Space& spc, const Json& j, const TermFactory& factory)
    {
        name = "hamiltonian";
        if (!j.is_array()) {
            throw std::runtime_error("energy: json array expected");
        }
        if (spc.geometry.type != Geometry::Type::CUBOID) {
            energy_terms.push_back(std::make_shared<ContainerOverlap>(spc));
        }
        for (const auto& j_energy : j.items()) {
            const auto& [key, value] = j_energy.single();
            try {
                if (key == "maxenergy") {
                    maximum_allowed_energy = value.number();
                }
                else if (key == "isobaric") {
                    energy_terms.push_back(std::make_shared<Isobaric>(value, spc));
                }
                else if (key == "example2d") {
                    energy_terms.push_back(std::make_shared<Example2D>(value, spc));
                }
                else if (!factory || !factory(*this, spc, key, value)) {
                    throw std::runtime_error("'" + key + "' unknown or outside the hot-path scope");
                }
            }
            catch (std::exception& e) {
                throw std::runtime_error("energy -> " + key + " -> " + e.what());
            }
        }
    }

    void push_back(std::shared_ptr<EnergyTerm> term) { energy_terms.push_back(std::move(term)); }
    size_t size() const { return energy_terms.size(); }
    const std::vector<std::shared_ptr<EnergyTerm>>& terms() const { return energy_terms; }
    const std::vector<double>& latestEnergies() const { return latest_energies; }
    double maximumAllowedEnergy() const { return maximum_allowed_energy; }

    template <class T> std::vector<std::shared_ptr<T>> find() const
    {
        std::vector<std::shared_ptr<T>> out;
        for (const auto& t : energy_terms) {
            if (auto p = std::dynamic_pointer_cast<T>(t)) {
                out.push_back(p);
            }
        }
        return out;
    }

    /** Σ terms with early stop at maxenergy / NaN; src/energy.cpp:1227-1241 */
    double energy(const Change& change) override
    {
        latest_energies.clear();
        for (auto& term : energy_terms) {
            term->state = state;
            const auto t0 = std::chrono::steady_clock::now();
            const double u = term->energy(change);
            term->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            latest_energies.push_back(u);
            if (u >= maximum_allowed_energy || std::isnan(u)) {
                break;
            }
        }
        return std::accumulate(latest_energies.begin(), latest_energies.end(), 0.0);
    }
    void init() override
    {
        for (auto& t : energy_terms) {
            t->init();
        }
    }
    /** every term in turn on the same vector; src/energy.cpp:1162-1166 */
    void force(std::vector<Point>& forces) override
    {
        for (auto& t : energy_terms) {
            t->force(forces);
        }
    }
    void updateState(const Change& change) override
    {
        for (auto& t : energy_terms) {
            t->state = state;
            t->updateState(change);
        }
    }
    void sync(EnergyTerm* other_hamiltonian, const Change& change) override
    {
        if (auto* other = dynamic_cast<Hamiltonian*>(other_hamiltonian)) {
            if (other->size() == size()) {
                latest_energies = other->latestEnergies();
                for (size_t i = 0; i < energy_terms.size(); ++i) {
                    energy_terms[i]->sync(other->energy_terms[i].get(), change);
                }
                return;
            }
        }
        throw std::runtime_error("hamiltonian mismatch");
    }
    void to_json(Json& j) const override
    {
        j = Json::array();
        for (const auto& t : energy_terms) {
            Json inner = Json::object();
            t->to_json(inner);
            Json wrapped = Json::object();
            wrapped[t->name] = inner;
            j.push_back(wrapped);
        }
    }
};

} // namespace fb

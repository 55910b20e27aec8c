// ORACLE — test infrastructure, not product code. CPU restatement of the reference's non-bonded
// energy term: which particle pairs a `Change` touches and how they are summed.
//
// Restates (reference file:line):
//   PairEnergy::potential + Chameleon::sqdist ... src/energy.h:428-438, src/geometry.h:460-470
//   Instant/Delayed accumulators ................ src/energy.h:556-716 (serial | openmp)
//   GroupCutoff ................................. src/energy.h:746-786, src/energy.cpp:1868-1933
//   GroupPairingPolicy loops .................... src/energy.h:838-1326
//   GroupPairing::accumulate dispatch ........... src/energy.h:1347-1475 (speciation: out of scope)
//   Nonbonded::energy ........................... src/energy.h:1512-1598
//   ParticleSelfEnergy (ExternalPotential) ...... src/externalpotential.cpp:60-126, 514-537
#pragma once
#include "../faunus_b200/csrc/host/energyterm.hpp"
#include "pairpotentials.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

/** r² by single-fold minimum image, then the pair functor; src/energy.h:428-438 */
template <class TPairPotential> class PairEnergy
{
  public:
    TPairPotential pair_potential;
    const Geometry& geometry;
    explicit PairEnergy(const Space& spc)
        : geometry(spc.geometry)
    {
    }
    inline double potential(const Particle& a, const Particle& b) const
    {
        return pair_potential(a, b, geometry.sqdist(a.pos, b.pos));
    }
    /** force on a due to b from the minimum-image vector b → a; src/energy.h:441-447 */
    inline Point force(const Particle& a, const Particle& b) const
    {
        const Point b_towards_a = geometry.vdist(a.pos, b.pos);
        return pair_potential.force(a, b, b_towards_a.squaredNorm(), b_towards_a);
    }
};

/**
 * Large synthetic systems only (tests/test_gpu_fullsize.py, bench.py): the Σ_{i<j} over one big ATOMIC group is
 * spread over the host threads (one row i per task, row sums added in row order: the result does not depend on the
 * thread count) and a repeated evaluation of the same configuration returns the stored sum (the MC driver evaluates
 * the start configuration once per state, src/montecarlo.cpp:43-71). Off by default: the reference's serial
 * pair order (src/energy.h:856-871) is what every other test runs.
 */
inline bool parallel_group_internal = false;
/** N = 1e6: the start-up sum over all 5e11 pairs is not evaluated at all (returns 0): only per-move energies count */
inline bool skip_group_internal = false;
inline size_t parallel_group_internal_min = 20000;
struct GroupInternalMemo
{
    std::vector<double> key;
    double value = 0.0;
};
inline GroupInternalMemo group_internal_memo;

/** Sums instantly (summation_policy: serial) or buffers pairs and reduces with OpenMP */
template <class TPairEnergy> class EnergyAccumulator
{
    const TPairEnergy& pair_energy;
    double value = 0.0;
    std::vector<std::pair<const Particle*, const Particle*>> pairs;
    size_t buffer_capacity = 0;

  public:
    bool openmp = false;
    explicit EnergyAccumulator(const TPairEnergy& pe)
        : pair_energy(pe)
    {
    }
    void reserve(size_t number_of_particles) // src/energy.h:622-637
    {
        number_of_particles = std::min<size_t>(number_of_particles, 10000);
        buffer_capacity = std::max<size_t>(1, (number_of_particles - 1U) * number_of_particles / 2U);
        if (openmp) {
            pairs.reserve(buffer_capacity);
        }
    }
    void clear()
    {
        value = 0.0;
        pairs.clear();
    }
    inline void add(const Particle& a, const Particle& b)
    {
        if (!openmp) {
            value += pair_energy.potential(a, b);
            return;
        }
        if (pairs.size() == buffer_capacity) {
            flush();
        }
        pairs.emplace_back(&a, &b);
    }
    void flush()
    {
        double sum = 0.0;
        const long n = static_cast<long>(pairs.size());
#pragma omp parallel for reduction(+ : sum)
        for (long i = 0; i < n; ++i) {
            sum += pair_energy.potential(*pairs[i].first, *pairs[i].second);
        }
        value += sum;
        pairs.clear();
    }
    double result()
    {
        if (openmp) {
            flush();
        }
        return value;
    }
    /** a sum evaluated elsewhere (parallel_group_internal) */
    void addValue(double sum) { value += sum; }
    const TPairEnergy& pairEnergy() const { return pair_energy; }
};

/** Mass-centre cutoff between molecular groups; src/energy.h:746-786, energy.cpp:1868-1933 */
class GroupCutoff
{
    size_t n = 0;
    std::vector<double> cutoff_squared;
    const Geometry& geometry;

    void setSingleCutoff(double cutoff)
    {
        const double c2 = (cutoff < std::sqrt(pc::max_value)) ? cutoff * cutoff : pc::max_value;
        std::fill(cutoff_squared.begin(), cutoff_squared.end(), c2);
    }

  public:
    explicit GroupCutoff(const Geometry& geometry)
        : geometry(geometry)
    {
    }
    void from_json(const Json& j, const Topology& topo)
    {
        n = topo.molecules.size();
        cutoff_squared.assign(n * n, pc::max_value);
        if (const auto* it = j.find("cutoff_g2g")) {
            if (it->is_number()) {
                setSingleCutoff(it->number());
            }
            else if (it->is_object()) {
                setSingleCutoff(it->value("default", pc::max_value));
                for (const auto& [named_pair, pair_cutoff] : it->members()) {
                    if (named_pair == "default") {
                        continue;
                    }
                    const auto names = splitWords(named_pair);
                    if (names.size() != 2) {
                        throw std::runtime_error("invalid molecules names");
                    }
                    const auto a = static_cast<size_t>(topo.moleculeId(names[0]));
                    const auto b = static_cast<size_t>(topo.moleculeId(names[1]));
                    cutoff_squared[a * n + b] = cutoff_squared[b * n + a] = std::pow(pair_cutoff.number(), 2);
                }
            }
        }
    }
    inline bool cut(const Group& g1, const Group& g2) const
    {
        if (g1.isAtomic() || g2.isAtomic()) {
            return false;
        }
        return geometry.sqdist(g1.mass_center, g2.mass_center) >= cutoff_squared[g1.id * n + g2.id];
    }
    double cutoffSquared(int a, int b) const { return cutoff_squared[a * n + b]; }
};

/** src/energy.h:804-1326, restated over index-range groups */
template <class TAccumulator> class GroupPairingPolicy
{
    const Space& spc;

  public:
    GroupCutoff cut;
    explicit GroupPairingPolicy(const Space& spc)
        : spc(spc)
        , cut(spc.geometry)
    {
    }

    static std::vector<size_t> indexComplement(size_t size, const std::vector<size_t>& index)
    {
        std::vector<size_t> out;
        for (size_t i = 0; i < size; ++i) {
            if (std::find(index.begin(), index.end(), i) == index.end()) {
                out.push_back(i);
            }
        }
        return out;
    }

    void groupInternal(TAccumulator& acc, const Group& group) // :856-871
    {
        const auto& moldata = spc.traits(group);
        if (parallel_group_internal && group.isAtomic() && group.size() >= parallel_group_internal_min) {
            if (skip_group_internal) {
                return;
            }
            const long n = static_cast<long>(group.size());
            std::vector<double> key;
            key.reserve(4 * group.size() + 3);
            const Point box = spc.geometry.getLength();
            key.insert(key.end(), {box.x, box.y, box.z});
            for (long i = 0; i < n; ++i) {
                const auto& p = spc.at(group, i);
                key.insert(key.end(), {p.pos.x, p.pos.y, p.pos.z, p.charge + 1000.0 * p.id});
            }
            if (key != group_internal_memo.key) {
                std::vector<double> row(static_cast<size_t>(n), 0.0);
                const auto& pe = acc.pairEnergy();
#pragma omp parallel for schedule(dynamic, 16)
                for (long i = 0; i < n - 1; ++i) {
                    double sum = 0.0;
                    for (long j = i + 1; j < n; ++j) {
                        sum += pe.potential(spc.at(group, i), spc.at(group, j));
                    }
                    row[i] = sum;
                }
                double total = 0.0;
                for (long i = 0; i < n - 1; ++i) {
                    total += row[i];
                }
                group_internal_memo.key = std::move(key);
                group_internal_memo.value = total;
            }
            acc.addValue(group_internal_memo.value);
            return;
        }
        if (!moldata.rigid) {
            const int group_size = static_cast<int>(group.size());
            for (int i = 0; i < group_size - 1; ++i) {
                for (int j = i + 1; j < group_size; ++j) {
                    if (group.isAtomic() || !moldata.isPairExcluded(i, j)) {
                        acc.add(spc.at(group, i), spc.at(group, j));
                    }
                }
            }
        }
    }

    void groupInternal(TAccumulator& acc, const Group& group, size_t index) // :885-914
    {
        const auto& moldata = spc.traits(group);
        if (!moldata.rigid) {
            const bool atomic = group.isAtomic();
            for (size_t i = 0; i < index; ++i) {
                if (atomic || !moldata.isPairExcluded(static_cast<int>(index), static_cast<int>(i))) {
                    acc.add(spc.at(group, index), spc.at(group, i));
                }
            }
            for (size_t i = index + 1; i < group.size(); ++i) {
                if (atomic || !moldata.isPairExcluded(static_cast<int>(index), static_cast<int>(i))) {
                    acc.add(spc.at(group, index), spc.at(group, i));
                }
            }
        }
    }

    void groupInternal(TAccumulator& acc, const Group& group, const std::vector<size_t>& index) // :931-961
    {
        const auto& moldata = spc.traits(group);
        if (!moldata.rigid) {
            if (index.size() == 1) {
                groupInternal(acc, group, index[0]);
            }
            else {
                const auto index_complement = indexComplement(group.size(), index);
                for (auto i : index) {
                    for (auto j : index_complement) {
                        if (!moldata.isPairExcluded(static_cast<int>(i), static_cast<int>(j))) {
                            acc.add(spc.at(group, i), spc.at(group, j));
                        }
                    }
                }
                for (auto i_it = index.begin(); i_it < index.end(); ++i_it) {
                    for (auto j_it = std::next(i_it); j_it < index.end(); ++j_it) {
                        if (!moldata.isPairExcluded(static_cast<int>(*i_it), static_cast<int>(*j_it))) {
                            acc.add(spc.at(group, *i_it), spc.at(group, *j_it));
                        }
                    }
                }
            }
        }
    }

    void group2group(TAccumulator& acc, const Group& g1, const Group& g2) // :979-989
    {
        if (!cut.cut(g1, g2)) {
            for (size_t i = 0; i < g1.size(); ++i) {
                for (size_t j = 0; j < g2.size(); ++j) {
                    acc.add(spc.at(g1, i), spc.at(g2, j));
                }
            }
        }
    }

    void group2group(TAccumulator& acc, const Group& g1, const Group& g2,
                     const std::vector<size_t>& index1) // :1015-1027
    {
        if (!cut.cut(g1, g2)) {
            for (auto i : index1) {
                for (size_t j = 0; j < g2.size(); ++j) {
                    acc.add(spc.at(g1, i), spc.at(g2, j));
                }
            }
        }
    }

    /** (group1 × ⊕group2) + (⊕group1 × ∁⊕group2): pairs with at least one indexed particle; :1055-1082 */
    void group2group(TAccumulator& acc, const Group& g1, const Group& g2, const std::vector<size_t>& index1,
                     const std::vector<size_t>& index2)
    {
        if (!cut.cut(g1, g2)) {
            if (!index2.empty()) {
                group2group(acc, g2, g1, index2);
                const auto index2_complement = indexComplement(g2.size(), index2);
                for (auto i : index1) {
                    for (auto j : index2_complement) {
                        acc.add(spc.at(g2, j), spc.at(g1, i));
                    }
                }
            }
            else if (!index1.empty()) {
                group2group(acc, g1, g2, index1);
            }
        }
    }

    /** ⊕group × (∪ groups given by index); :1131-1141 */
    void group2groups(TAccumulator& acc, const Group& group, const std::vector<size_t>& group_index,
                      const std::vector<size_t>& index)
    {
        for (auto other_index : group_index) {
            const auto& other = spc.groups[other_index];
            if (&other != &group) {
                group2group(acc, group, other, index);
            }
        }
    }

    /**
     * The number of particles has changed: pairs that involve a listed, ACTIVE particle of a changed group —
     * with the static groups, with the other changed groups, inside its own group; src/energy.h:1390-1435.
     * Removed particles need no care: they are present in the other Space.
     */
    void accumulateSpeciation(TAccumulator& acc, const Change& change)
    {
        std::vector<size_t> moved;
        for (const auto& g : change.groups) {
            moved.push_back(g.group_index);
        }
        const auto fixed = indexComplement(spc.groups.size(), moved);
        auto active = [](const std::vector<size_t>& listed, size_t size) {
            std::vector<size_t> out;
            for (auto i : listed) {
                if (i < size) {
                    out.push_back(i);
                }
            }
            return out;
        };
        for (auto it1 = change.groups.begin(); it1 < change.groups.end(); ++it1) {
            const auto& group1 = spc.groups.at(it1->group_index);
            const auto index1 = active(it1->relative_atom_indices, group1.size());
            if (!index1.empty()) {
                group2groups(acc, group1, fixed, index1);
            }
            for (auto it2 = std::next(it1); it2 < change.groups.end(); ++it2) {
                const auto& group2 = spc.groups.at(it2->group_index);
                const auto index2 = active(it2->relative_atom_indices, group2.size());
                if (!index1.empty() || !index2.empty()) {
                    group2group(acc, group1, group2, index1, index2);
                }
            }
            if (!index1.empty() && !spc.traits(group1).rigid) {
                if (it1->all) {
                    groupInternal(acc, group1);
                }
                else {
                    groupInternal(acc, group1, index1);
                }
            }
        }
    }

    void group2all(TAccumulator& acc, const Group& group) // :1155-1163
    {
        for (const auto& other : spc.groups) {
            if (&other != &group) {
                group2group(acc, group, other);
            }
        }
    }

    void group2all(TAccumulator& acc, const Group& group, size_t index) // :1182-1195
    {
        const auto& particle = spc.at(group, index);
        for (const auto& other : spc.groups) {
            if (&other != &group) {
                if (!cut.cut(other, group)) {
                    for (size_t j = 0; j < other.size(); ++j) {
                        acc.add(particle, spc.at(other, j));
                    }
                }
            }
        }
    }

    void group2all(TAccumulator& acc, const Group& group, const std::vector<size_t>& index) // :1212-1226
    {
        if (index.size() == 1) {
            group2all(acc, group, index[0]);
        }
        else {
            for (const auto& other : spc.groups) {
                if (&other != &group) {
                    group2group(acc, group, other, index);
                }
            }
        }
    }

    void groups2self(TAccumulator& acc, const std::vector<size_t>& group_index) // :1241-1254
    {
        for (auto it1 = group_index.begin(); it1 < group_index.end(); ++it1) {
            for (auto it2 = std::next(it1); it2 < group_index.end(); ++it2) {
                group2group(acc, spc.groups[*it1], spc.groups[*it2]);
            }
        }
    }

    void groups2all(TAccumulator& acc, const std::vector<size_t>& group_index) // :1268-1278
    {
        groups2self(acc, group_index);
        const auto index_complement = indexComplement(spc.groups.size(), group_index);
        for (auto g1 : group_index) {
            for (auto g2 : index_complement) {
                group2group(acc, spc.groups[g1], spc.groups[g2]);
            }
        }
    }

    template <class Condition> void all(TAccumulator& acc, Condition condition) // :1290-1325
    {
        for (auto it = spc.groups.begin(); it < spc.groups.end(); ++it) {
            if (condition(*it)) {
                groupInternal(acc, *it);
            }
            for (auto other = std::next(it); other < spc.groups.end(); ++other) {
                group2group(acc, *it, *other);
            }
        }
    }
};

/** Non-bonded energy term; src/energy.h:1447-1475 (dispatch), :1512-1598 */
template <class TPairPotential> class Nonbonded : public EnergyTerm
{
    const Space& spc;
    using TPairEnergy = PairEnergy<TPairPotential>;
    using TAccumulator = EnergyAccumulator<TPairEnergy>;
    TPairEnergy pair_energy;
    TAccumulator accumulator;
    GroupPairingPolicy<TAccumulator> pairing;

  public:
    Nonbonded(const Json& j, Space& spc)
        : spc(spc)
        , pair_energy(spc)
        , accumulator(pair_energy)
        , pairing(spc)
    {
        name = "nonbonded";
        pair_energy.pair_potential.from_json(j, *spc.topology);
        pairing.cut.from_json(j, *spc.topology);
        const std::string policy = j.value("summation_policy", "serial");
        accumulator.openmp = (policy == "openmp");
        accumulator.reserve(spc.numParticles(Space::Selection::ALL));
    }
    const TPairPotential& pairPotential() const { return pair_energy.pair_potential; }
    double particleParticleEnergy(const Particle& a, const Particle& b) const
    {
        return pair_energy.potential(a, b);
    }

    /** every pair of the particle vector, active or not (`@todo A stub`); src/energy.h:1584-1597 */
    void force(std::vector<Point>& forces) override
    {
        if (forces.size() != spc.particles.size()) {
            throw std::runtime_error("the forces size must match the particle size");
        }
        if constexpr (requires(const TPairPotential& pot, const Particle& p, const Point& r) { pot.force(p, p, 1.0, r); }) {
            for (size_t i = 0; i + 1 < spc.particles.size(); ++i) {
                for (size_t j = i + 1; j < spc.particles.size(); ++j) {
                    const Point f = pair_energy.force(spc.particles[i], spc.particles[j]);
                    forces[i] = forces[i] + f;
                    forces[j] = forces[j] - f;
                }
            }
        }
        else { // FunctorPotential, SplinedPotential: PairPotential::force, src/potentials.cpp:246-251
            throw std::logic_error("Force computation not implemented for this setup!");
        }
    }

    double energy(const Change& change) override
    {
        accumulator.clear();
        if (change.everything) {
            pairing.all(accumulator, [](const Group&) { return true; });
        }
        else if (change.volume_change) {
            pairing.all(accumulator, [](const Group& g) { return g.isAtomic() || g.compressible; });
        }
        else if (!change.matter_change) {
            if (change.groups.size() == 1) {
                const auto& cd = change.groups.at(0);
                const auto& group = spc.groups.at(cd.group_index);
                if (cd.relative_atom_indices.size() == 1) {
                    pairing.group2all(accumulator, group, cd.relative_atom_indices[0]);
                    if (cd.internal) {
                        pairing.groupInternal(accumulator, group, cd.relative_atom_indices[0]);
                    }
                }
                else if (cd.relative_atom_indices.empty()) {
                    pairing.group2all(accumulator, group);
                    if (cd.internal) {
                        pairing.groupInternal(accumulator, group);
                    }
                }
                else {
                    pairing.group2all(accumulator, group, cd.relative_atom_indices);
                    if (cd.internal) {
                        pairing.groupInternal(accumulator, group, cd.relative_atom_indices);
                    }
                }
            }
            else {
                std::vector<size_t> moved;
                for (const auto& g : change.groups) {
                    moved.push_back(g.group_index);
                }
                pairing.groups2all(accumulator, moved);
            }
        }
        else {
            pairing.accumulateSpeciation(accumulator, change);
        }
        return accumulator.result();
    }
};

/** Σ self_energy(particle) over changed (or all) active particles; src/externalpotential.cpp:60-126 */
class ParticleSelfEnergy : public EnergyTerm
{
    const Space& spc;
    std::function<double(const Particle&)> func;

    double groupEnergy(const Group& group) const
    {
        double energy = 0.0;
        for (size_t i = 0; i < group.size(); ++i) {
            energy += func(spc.at(group, i));
            if (std::isnan(energy)) {
                break;
            }
        }
        return energy;
    }

  public:
    ParticleSelfEnergy(const Space& spc, std::function<double(const Particle&)> f)
        : spc(spc)
        , func(std::move(f))
    {
        name = "particle-self-energy";
    }
    double energy(const Change& change) override
    {
        double energy = 0.0;
        if (change.volume_change || change.everything || change.matter_change) {
            for (const auto& group : spc.groups) {
                energy += groupEnergy(group);
                if (!std::isfinite(energy)) {
                    break;
                }
            }
            return energy;
        }
        for (const auto& gc : change.groups) {
            const auto& group = spc.groups.at(gc.group_index);
            if (gc.all) {
                energy += groupEnergy(group);
            }
            else {
                for (auto index : gc.relative_atom_indices) {
                    energy += func(spc.at(group, index));
                }
            }
            if (!std::isfinite(energy)) {
                break;
            }
        }
        return energy;
    }
};

} // namespace oracle

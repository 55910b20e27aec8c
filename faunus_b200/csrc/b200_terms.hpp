// B200 adaptor terms: `Energy::EnergyTerm` subclasses that stand where the reference's
// `Energy::Nonbonded<…>`, `Energy::Ewald` and `ParticleSelfEnergy` stand in the Hamiltonian and
// forward energy(Change) / updateState / sync / init to the CUDA library through its C ABI
// (include/faunus_b200.h). No CPU fallback: if the device library cannot create a context the
// constructor throws.
//
// Protocol (src/montecarlo.cpp:151-175): trial.updateState(c) → trial.energy(c) → accepted.energy(c)
// → accepted.sync(trial) | trial.sync(accepted). The two Hamiltonians (accepted, trial) share ONE
// device context with two mirror slots; `updateState` pushes only the particles listed in the
// Change into the trial slot; `trial.energy` evaluates new and old in one fused launch and the
// following `accepted.energy` picks up the cached old value; `sync` is a device-to-device copy of
// the changed particles. Callers that skip updateState/sync (Widom, SystemEnergy) are served by
// re-uploading the changed group from the Space the term is bound to.
#pragma once
#include <array>
#include "../../include/faunus_b200.h"
#include "host/energyterm.hpp"
#include "host/analysis_rdf.hpp"
#include "host/montecarlo.hpp"
#include "host/potential_tables.hpp"
#include <map>
#include <mutex>

namespace fb {

inline void fbCheck(int rc, fb_ctx* ctx, const char* what)
{
    if (rc != FB_OK) {
        throw std::runtime_error(std::string(what) + ": " + fb_last_error(ctx));
    }
}

/** Flat view of a Change for the C ABI (keeps the index storage alive) */
struct FlatChange
{
    fb_change change{};
    std::vector<fb_group_change> groups;
    std::vector<std::vector<int>> indices;
    explicit FlatChange(const Change& c)
    {
        change.everything = c.everything;
        change.volume_change = c.volume_change;
        change.matter_change = c.matter_change;
        indices.resize(c.groups.size());
        for (size_t i = 0; i < c.groups.size(); ++i) {
            const auto& g = c.groups[i];
            indices[i].assign(g.relative_atom_indices.begin(), g.relative_atom_indices.end());
            fb_group_change f{};
            f.group_index = static_cast<int>(g.group_index);
            f.all = g.all;
            f.internal = g.internal;
            f.n_atoms = static_cast<int>(indices[i].size());
            f.atoms = indices[i].data();
            groups.push_back(f);
        }
        change.n_groups = static_cast<int>(groups.size());
        change.groups = groups.data();
    }
    /** content signature used to pair trial.energy with the following accepted.energy */
    std::string signature() const
    {
        std::string s = std::to_string(change.everything) + ":" + std::to_string(change.volume_change) + ":" +
                        std::to_string(change.matter_change);
        for (size_t i = 0; i < groups.size(); ++i) {
            s += "|" + std::to_string(groups[i].group_index) + "," + std::to_string(groups[i].all) + "," +
                 std::to_string(groups[i].internal);
            for (int a : indices[i]) {
                s += "," + std::to_string(a);
            }
        }
        return s;
    }
};

inline fb_group groupRecord(const Group& g)
{
    fb_group r{};
    r.begin = static_cast<int>(g.begin);
    r.size = static_cast<int>(g.size());
    r.capacity = static_cast<int>(g.capacity());
    r.molid = g.id;
    r.cm[0] = g.mass_center.x;
    r.cm[1] = g.mass_center.y;
    r.cm[2] = g.mass_center.z;
    return r;
}

/**
 * One device context shared by the accepted and trial instances of a non-bonded term (and its Ewald
 * sibling). Created by the first instance, joined by the second (registry keyed on the topology
 * the two States share, in construction order: accepted → slot 0, trial → slot 1).
 */
class DeviceContext
{
  public:
    fb_ctx* ctx = nullptr;
    PairTables tables;
    int attached = 0;
    // fused ΔU cache: filled by the trial term, consumed by the accepted term
    bool cache_valid = false;
    std::string cache_signature;
    double cached_old_energy = 0;
    // Ewald sibling: present? eligible for the fused fast path (no surface term)? old groups known?
    bool has_ewald = false;
    int ewald_policy = 0;
    bool ewald_fast_ok = false;
    bool ewald_have_old = false;
    // fast path: one staged small trial move (fb_trial_energy / fb_trial_commit)
    bool fast_enabled = true;
    bool fast_staged = false;
    bool fast_evaluated = false;
    bool fast_sync_for_ewald = false;
    fb_trial_move fast_move{};
    uint64_t fast_key = 0;
    double fast_u_new = 0, fast_u_old = 0, fast_ew_new = 0, fast_ew_old = 0;
    const Space* spaces[2] = {nullptr, nullptr}; //!< Space of the instance bound to each slot
    /**
     * A caller that mutates a Space, asks for energy(everything) and puts the Space back without ever calling
     * updateState or sync (VirtualVolumeMove, src/analysis.cpp:825-843) leaves the mirror with whatever it was
     * refreshed from last. Such a refresh marks the slot; whoever touches the mirror next brings it back in step.
     */
    bool unsynchronised[2] = {false, false};
    std::vector<size_t> unsynchronised_groups[2]; //!< … the same for single groups (VirtualTranslate, :2847-2860)
    /**
     * The particles of the slot's mirror already ARE those of its Space (a replica exchange imported the partner's
     * state on the device, fb_nccl_exchange_state, and the Space followed): the next full upload only sends the box
     * and the group records.
     */
    bool particles_current[2] = {false, false};

    /** cheap identity of a single-group Change (group, flags, indices) */
    static uint64_t changeKey(const Change& c)
    {
        if (c.everything || c.volume_change || c.groups.size() != 1) {
            return 0;
        }
        const auto& g = c.groups[0];
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
        mix(g.group_index + 1);
        mix((g.all ? 2 : 0) | (g.internal ? 1 : 0));
        for (auto i : g.relative_atom_indices) {
            mix(i + 7);
        }
        return h | 1ull;
    }

    void evaluateFast()
    {
        fbCheck(fb_trial_energy(ctx, &fast_move, &fast_u_new, &fast_u_old, &fast_ew_new, &fast_ew_old), ctx,
                "fb_trial_energy");
        fast_evaluated = true;
    }

    DeviceContext(const std::string& name, const Json& cfg, const Space& spc, int device)
    {
        const Topology& topo = *spc.topology;
        tables = buildPairTables(name, cfg, topo);
        fb_config fc{};
        fc.device = device;
        const auto& L = spc.geometry.getLength();
        for (int i = 0; i < 3; ++i) {
            fc.box[i] = L[i];
            fc.periodic[i] = spc.geometry.isPeriodic(i);
        }
        fc.n_atom_types = static_cast<int>(topo.atoms.size());
        fc.n_molecule_types = static_cast<int>(topo.molecules.size());
        std::vector<int> molflags, molnatoms;
        std::vector<const unsigned char*> exclusions;
        for (const auto& m : topo.molecules) {
            molflags.push_back((m.atomic ? FB_MOL_ATOMIC : 0) | (m.rigid ? FB_MOL_RIGID : 0) |
                               (m.compressible ? FB_MOL_COMPRESSIBLE : 0));
            molnatoms.push_back(static_cast<int>(m.atoms.size()));
            exclusions.push_back(m.excluded.empty() ? nullptr : m.excluded.data());
        }
        fc.molecule_flags = molflags.data();
        fc.molecule_natoms = molnatoms.data();
        fc.exclusions = exclusions.data();
        fc.g2g_cutoff_squared = tables.g2g_cutoff_squared.data();
        fc.kind = tables.kind;
        auto opt = [](const auto& v) { return v.empty() ? nullptr : v.data(); };
        fc.pair_flags = (tables.kind == potkind::FUNCTOR || tables.kind == potkind::SPLINED) ? tables.flags.data()
                                                                                            : nullptr;
        fc.lj_sigma2 = opt(tables.lj_s2);
        fc.lj_eps4 = opt(tables.lj_e4);
        fc.wca_sigma2 = opt(tables.wca_s2);
        fc.wca_eps4 = opt(tables.wca_e4);
        fc.hs_sigma2 = opt(tables.hs_s2);
        if (tables.has_coulomb) {
            fc.coulomb_bjerrum_length = tables.coulomb.bjerrum_length;
            fc.coulomb_cutoff = tables.coulomb.cutoff;
            fc.coulomb_kappa = tables.coulomb.kappa;
            fc.coulomb_n_knots = static_cast<int>(tables.coulomb.S.knots.size());
            fc.coulomb_knots = tables.coulomb.S.knots.data();
            fc.coulomb_coeffs = tables.coulomb.S.coeffs.data();
        }
        fc.plain_bjerrum_length = tables.plain_bjerrum_length;
        if (tables.kind == potkind::SPLINED) {
            fc.spline_offset = tables.sp_offset.data();
            fc.spline_knots = tables.sp_knots.data();
            fc.spline_coeffs = tables.sp_coeffs.data();
            fc.spline_rmin2 = tables.sp_rmin2.data();
            fc.spline_rmax2 = tables.sp_rmax2.data();
            fc.spline_hardsphere = tables.sp_hs.data();
        }
        const int rc = fb_create(&fc, &ctx);
        if (rc != FB_OK) {
            throw std::runtime_error(std::string("fb_create: ") + fb_last_error(nullptr));
        }
        if (tables.has_coulomb && (tables.kind == potkind::COULOMB_LJ || tables.kind == potkind::COULOMB_WCA)) {
            // the two kinds with pair forces (fb_nonbonded_force): S'(q) next to S(q)
            fbCheck(fb_set_force_table(ctx, static_cast<int>(tables.coulomb.dS.knots.size()), tables.coulomb.dS.knots.data(),
                                       tables.coulomb.dS.coeffs.data()),
                    ctx, "fb_set_force_table");
        }
    }
    ~DeviceContext() { fb_destroy(ctx); }
    DeviceContext(const DeviceContext&) = delete;
    DeviceContext& operator=(const DeviceContext&) = delete;

    /** full upload of a Space into a slot */
    void uploadSpace(int slot, const Space& spc)
    {
        if (particles_current[slot]) {
            particles_current[slot] = false;
            std::vector<fb_group> records;
            for (const auto& g : spc.groups) {
                records.push_back(groupRecord(g));
            }
            const auto& len = spc.geometry.getLength();
            const double lengths[3] = {len.x, len.y, len.z};
            fbCheck(fb_set_box(ctx, slot, lengths), ctx, "fb_set_box");
            fbCheck(fb_upload_groups(ctx, slot, records.data(), static_cast<int>(records.size())), ctx, "fb_upload_groups");
            cache_valid = false;
            return;
        }
        const size_t n = spc.particles.size();
        std::vector<double> xyzq(4 * n);
        std::vector<int> ids(n);
        for (size_t i = 0; i < n; ++i) {
            const auto& p = spc.particles[i];
            xyzq[4 * i] = p.pos.x;
            xyzq[4 * i + 1] = p.pos.y;
            xyzq[4 * i + 2] = p.pos.z;
            xyzq[4 * i + 3] = p.charge;
            ids[i] = p.id;
        }
        std::vector<fb_group> groups;
        for (const auto& g : spc.groups) {
            groups.push_back(groupRecord(g));
        }
        const auto& L = spc.geometry.getLength();
        const double box[3] = {L.x, L.y, L.z};
        fbCheck(fb_set_box(ctx, slot, box), ctx, "fb_set_box");
        fbCheck(fb_upload_space(ctx, slot, xyzq.data(), ids.data(), groups.data(), static_cast<int>(n),
                                static_cast<int>(groups.size())),
                ctx, "fb_upload_space");
        cache_valid = false;
    }

    /** the mirrors that were last refreshed from an unsynchronised Space := the Spaces as they are now */
    void resynchronise()
    {
        for (int slot = 0; slot < 2; ++slot) {
            if (unsynchronised[slot] && spaces[slot] != nullptr) {
                unsynchronised[slot] = false;
                unsynchronised_groups[slot].clear();
                uploadSpace(slot, *spaces[slot]);
            }
            if (!unsynchronised_groups[slot].empty() && spaces[slot] != nullptr) {
                Change refresh;
                for (const auto g : unsynchronised_groups[slot]) {
                    auto& record = refresh.groups.emplace_back();
                    record.group_index = g;
                    record.all = true;
                }
                unsynchronised_groups[slot].clear();
                uploadChange(slot, *spaces[slot], refresh);
            }
        }
    }

    /** push the groups/particles a partial Change lists from `spc` into `slot` */
    void uploadChange(int slot, const Space& spc, const Change& change)
    {
        for (const auto& gc : change.groups) {
            const auto& g = spc.groups.at(gc.group_index);
            const fb_group rec = groupRecord(g);
            std::vector<int> rel;
            if (gc.all || gc.relative_atom_indices.empty()) {
                rel.resize(g.capacity()); // whole group incl. inactive tail (Space::sync copies it too)
                std::iota(rel.begin(), rel.end(), 0);
            }
            else {
                rel.assign(gc.relative_atom_indices.begin(), gc.relative_atom_indices.end());
            }
            std::vector<double> xyzq(4 * rel.size());
            std::vector<int> ids(rel.size());
            for (size_t i = 0; i < rel.size(); ++i) {
                const auto& p = spc.at(g, rel[i]);
                xyzq[4 * i] = p.pos.x;
                xyzq[4 * i + 1] = p.pos.y;
                xyzq[4 * i + 2] = p.pos.z;
                xyzq[4 * i + 3] = p.charge;
                ids[i] = p.id;
            }
            fbCheck(fb_update_group(ctx, slot, static_cast<int>(gc.group_index), &rec, static_cast<int>(rel.size()),
                                    rel.data(), xyzq.data(), ids.data()),
                    ctx, "fb_update_group");
        }
        cache_valid = false;
    }
};

/** process-wide registry pairing the accepted and trial instances (see DeviceContext) */
inline std::map<std::pair<const Topology*, std::string>, std::weak_ptr<DeviceContext>>& deviceRegistry()
{
    static std::map<std::pair<const Topology*, std::string>, std::weak_ptr<DeviceContext>> registry;
    return registry;
}

/** CUDA device new contexts are created on; per thread (in-process replicas run one thread per GPU) */
inline int& defaultDevice()
{
    static thread_local int device = 0;
    return device;
}

inline std::mutex& deviceRegistryMutex()
{
    static std::mutex m;
    return m;
}

/** Replaces Energy::Nonbonded<PairEnergy<…>, GroupPairing<…>> (src/energy.h:1512-1598) */
class NonbondedB200 : public EnergyTerm
{
    const Space& spc;
    std::shared_ptr<DeviceContext> dev;
    int slot = 0;             //!< 0 accepted, 1 trial (construction order)
    bool state_pushed = false; //!< updateState() already pushed the pending change into our slot
    std::string pushed_signature;

  public:
    NonbondedB200(const std::string& key, const Json& j, Space& spc)
        : spc(spc)
    {
        name = "nonbonded";
        const auto reg_key = std::make_pair(spc.topology.get(), key + j.dump());
        std::lock_guard<std::mutex> lock(deviceRegistryMutex());
        auto& registry = deviceRegistry();
        auto it = registry.find(reg_key);
        if (it != registry.end()) {
            dev = it->second.lock();
        }
        if (!dev || dev->attached >= 2) {
            dev = std::make_shared<DeviceContext>(key, j, spc, defaultDevice());
            registry[reg_key] = dev;
        }
        slot = dev->attached++;
        dev->spaces[slot] = &spc;
        dev->uploadSpace(slot, spc);
    }
    ~NonbondedB200() override
    {
        if (dev) {
            dev->attached--;
        }
    }
    const std::shared_ptr<DeviceContext>& device() const { return dev; }
    int deviceSlot() const { return slot; }

    void init() override { dev->uploadSpace(slot, spc); }

    /**
     * Nonbonded::force (src/energy.h:1584-1597): adds the pair forces of the whole particle vector. Forces are asked
     * for outside the updateState / sync protocol (src/forcemove.cpp:124-125), so the mirror is refreshed first.
     */
    void force(std::vector<Point>& forces) override
    {
        static_assert(sizeof(Point) == 3 * sizeof(double), "forces travel as [n][3] doubles");
        if (forces.size() != spc.particles.size()) {
            throw std::runtime_error("the forces size must match the particle size");
        }
        dev->uploadSpace(slot, spc);
        fbCheck(fb_nonbonded_force(dev->ctx, slot, &forces[0].x), dev->ctx, "fb_nonbonded_force");
    }

    /**
     * Small move of one group (≤ 8 atoms, no size change) on the trial instance: nothing is pushed;
     * the particles are staged for the fused single-launch evaluation.
     */
    bool stageFastMove(const Change& change)
    {
        auto& d = *dev;
        if (!d.fast_enabled || state != MonteCarloState::TRIAL || d.attached != 2 || slot != 1) {
            return false;
        }
        const uint64_t key = DeviceContext::changeKey(change);
        if (key == 0 || change.matter_change) {
            return false;
        }
        if (d.has_ewald && !(d.ewald_fast_ok && d.ewald_have_old)) {
            return false;
        }
        const auto& gc = change.groups[0];
        const auto& g = spc.groups.at(gc.group_index);
        const auto& g_old = d.spaces[0]->groups.at(gc.group_index);
        if (g.size() != g_old.size()) {
            return false;
        }
        const bool whole = gc.relative_atom_indices.empty();
        const size_t n = whole ? g.size() : gc.relative_atom_indices.size();
        if (n < 1 || n > FB_FAST_ATOMS) {
            return false;
        }
        fb_trial_move& mv = d.fast_move;
        mv.group_index = static_cast<int>(gc.group_index);
        mv.n_atoms = static_cast<int>(n);
        for (size_t i = 0; i < n; ++i) {
            const size_t rel = whole ? i : gc.relative_atom_indices[i];
            if (rel >= g.size()) {
                return false;
            }
            const auto& p = spc.at(g, rel);
            mv.rel_index[i] = static_cast<int>(rel);
            mv.xyzq[i][0] = p.pos.x;
            mv.xyzq[i][1] = p.pos.y;
            mv.xyzq[i][2] = p.pos.z;
            mv.xyzq[i][3] = p.charge;
            mv.atom_id[i] = p.id;
        }
        mv.cm[0] = g.mass_center.x;
        mv.cm[1] = g.mass_center.y;
        mv.cm[2] = g.mass_center.z;
        mv.internal = gc.internal ? 1 : 0;
        mv.with_ewald = 0; // set by the Ewald sibling's updateState
        d.fast_staged = true;
        d.fast_evaluated = false;
        d.fast_key = key;
        d.cache_valid = false;
        return true;
    }

    /** trial Space was mutated by a move: push what the Change lists into our slot */
    void updateState(const Change& change) override
    {
        if (!change) {
            return;
        }
        dev->resynchronise();
        dev->fast_staged = false;
        if (stageFastMove(change)) {
            return;
        }
        if (change.everything || change.volume_change) {
            dev->uploadSpace(slot, spc);
        }
        else {
            dev->uploadChange(slot, spc, change);
        }
        state_pushed = true;
        pushed_signature = FlatChange(change).signature();
    }

    double energy(const Change& change) override
    {
        if (!change) {
            return 0.0;
        }
        if (change.everything || change.volume_change) {
            dev->unsynchronised[slot] = false; // refreshed below (or by updateState just now) anyway
        }
        dev->resynchronise();
        if (dev->fast_staged && dev->fast_key == DeviceContext::changeKey(change)) {
            if (state == MonteCarloState::TRIAL) {
                if (!dev->fast_evaluated) {
                    dev->evaluateFast();
                }
                return dev->fast_u_new;
            }
            if (dev->fast_evaluated) {
                return dev->fast_u_old;
            }
        }
        if (dev->fast_staged && state == MonteCarloState::TRIAL) {
            // a different change is evaluated on the trial state while a fast move is staged: materialise it
            dev->fast_staged = false;
            dev->uploadChange(slot, spc, change);
        }
        FlatChange flat(change);
        const std::string sig = flat.signature();
        const bool partial = !change.everything && !change.volume_change;
        // accepted instance right after the trial instance evaluated the same change: cached old energy
        if (partial && dev->cache_valid && state != MonteCarloState::TRIAL && dev->cache_signature == sig) {
            dev->cache_valid = false;
            return dev->cached_old_energy;
        }
        const bool pushed = state_pushed && pushed_signature == sig;
        state_pushed = false;
        if (!pushed) { // caller mutated our Space without updateState (Widom, SystemEnergy, tests)
            if (partial) {
                dev->uploadChange(slot, spc, change);
                for (const auto& gc : change.groups) { // the caller may put the group back without telling anybody
                    dev->unsynchronised_groups[slot].push_back(gc.group_index);
                }
            }
            else {
                dev->uploadSpace(slot, spc);
                dev->unsynchronised[slot] = true; // … or the whole Space
            }
        }
        double u = 0.0;
        if (partial && pushed && state == MonteCarloState::TRIAL && dev->attached == 2) {
            double u_old = 0.0;
            fbCheck(fb_nonbonded_delta(dev->ctx, slot, 1 - slot, &flat.change, &u, &u_old), dev->ctx,
                    "fb_nonbonded_delta");
            dev->cache_valid = true;
            dev->cache_signature = sig;
            dev->cached_old_energy = u_old;
        }
        else {
            fbCheck(fb_nonbonded_energy(dev->ctx, slot, &flat.change, &u), dev->ctx, "fb_nonbonded_energy");
        }
        return u;
    }

    /** our slot := other's slot for the changed particles (device-to-device) */
    void sync(EnergyTerm* other_term, const Change& change) override
    {
        auto* other = dynamic_cast<NonbondedB200*>(other_term);
        if (!other || other->dev != dev) {
            throw std::runtime_error("sync error");
        }
        dev->resynchronise();
        if (dev->fast_staged && dev->fast_key == DeviceContext::changeKey(change)) {
            // accepted.sync(trial) = accept, trial.sync(accepted) = reject (src/montecarlo.cpp:167-175)
            if (!dev->fast_evaluated) {
                dev->evaluateFast();
            }
            fbCheck(fb_trial_commit(dev->ctx, slot == 0 ? 1 : 0), dev->ctx, "fb_trial_commit");
            dev->fast_staged = false;
            dev->fast_evaluated = false;
            dev->fast_sync_for_ewald = dev->has_ewald;
            return;
        }
        dev->fast_staged = false;
        FlatChange flat(change);
        fbCheck(fb_sync(dev->ctx, slot, other->slot, &flat.change), dev->ctx, "fb_sync");
        dev->cache_valid = false;
        state_pushed = false;
    }

    /** NonbondedBase::particleParticleEnergy, src/energy.h:1500, 1531-1534 */
    double particleParticleEnergy(const Particle& particle1, const Particle& particle2)
    {
        const double a[4] = {particle1.pos.x, particle1.pos.y, particle1.pos.z, particle1.charge};
        const double b[4] = {particle2.pos.x, particle2.pos.y, particle2.pos.z, particle2.charge};
        const int ida = particle1.id, idb = particle2.id;
        double u = 0.0;
        dev->resynchronise();
        fbCheck(fb_particle_pair_energy(dev->ctx, slot, 1, a, &ida, b, &idb, &u), dev->ctx, "fb_particle_pair_energy");
        return u;
    }

    /** NonbondedBase::groupGroupEnergy, src/energy.h:1501, 1536-1545 (groups of this term's Space, by index) */
    double groupGroupEnergy(size_t group1, size_t group2)
    {
        double u = 0.0;
        dev->resynchronise();
        fbCheck(fb_group_group_energy(dev->ctx, slot, static_cast<int>(group1), static_cast<int>(group2), &u), dev->ctx,
                "fb_group_group_energy");
        return u;
    }

    void to_json(Json& j) const override
    {
        j["device"] = "B200 sm_100a";
        j["kind"] = dev->tables.kind;
        j["launches"] = static_cast<size_t>(fb_launch_count(dev->ctx));
    }
};

/** Replaces Energy::Ewald (src/energy.cpp:539-658); shares the context of its non-bonded sibling */
class EwaldB200 : public EnergyTerm
{
    const Space& spc;
    std::shared_ptr<DeviceContext> dev;
    std::shared_ptr<NonbondedB200> sibling;
    int slot;
    bool have_old = false; //!< `old_groups` set (after the first sync from the accepted instance)
    int n_kvectors = 0;

    void fullUpdate()
    {
        int K = 0;
        fbCheck(fb_ewald_update_box(dev->ctx, slot, &K), dev->ctx, "fb_ewald_update_box");
        n_kvectors = K;
        fbCheck(fb_ewald_update_full(dev->ctx, slot), dev->ctx, "fb_ewald_update_full");
    }

  public:
    EwaldB200(const Json& j, const Space& spc, std::shared_ptr<NonbondedB200> nonbonded)
        : spc(spc)
        , dev(nonbonded->device())
        , sibling(std::move(nonbonded))
        , slot(sibling->deviceSlot())
    {
        name = "ewald";
        fb_ewald_config cfg{};
        cfg.alpha = j.at("alpha").number();
        cfg.n_cutoff = j.contains("kcutoff") ? j.at("kcutoff").number() : j.at("ncutoff").number();
        cfg.kappa = j.value("kappa", 0.0);
        cfg.surface_dielectric_constant = j.value("epss", 0.0);
        cfg.bjerrum_length = pc::bjerrumLength(j.at("epsr").number());
        cfg.spherical_sum = j.value("spherical_sum", true);
        const std::string scheme = j.value("ipbc", false) ? "IPBC" : j.value("ewaldscheme", "PBC");
        if (scheme == "PBC") {
            cfg.policy = 0;
        }
        else if (scheme == "PBCEigen") {
            cfg.policy = 1;
        }
        else if (scheme == "IPBC" || scheme == "IPBCEigen") {
            cfg.policy = 2;
        }
        else {
            throw std::runtime_error("invalid `ewaldpolicy`");
        }
        fbCheck(fb_ewald_configure(dev->ctx, &cfg), dev->ctx, "fb_ewald_configure");
        dev->has_ewald = true;
        dev->ewald_policy = cfg.policy;
        dev->ewald_fast_ok = cfg.surface_dielectric_constant < 1.0; // tinfoil: no surface term to track
        init();
    }

    /** the mirror was uploaded by the sibling's init(); rebuild k-vectors and Q(k) */
    void init() override { fullUpdate(); }

    /**
     * Ewald::force (src/energy.cpp:596-629): surface + reciprocal-space force from the CURRENT positions and the Q(k)
     * this instance holds (the reference does not rebuild Q there either); overwrites `forces`, as the reference does.
     */
    void force(std::vector<Point>& forces) override
    {
        if (forces.size() != spc.particles.size()) {
            throw std::runtime_error("the forces size must match the particle size");
        }
        dev->uploadSpace(slot, spc);
        fbCheck(fb_ewald_force(dev->ctx, slot, &forces[0].x), dev->ctx, "fb_ewald_force");
    }

    void updateState(const Change& change) override
    {
        if (!change) {
            return;
        }
        if (dev->fast_staged && dev->fast_key == DeviceContext::changeKey(change)) {
            dev->fast_move.with_ewald = 1; // evaluated by the sibling's fused launch
            return;
        }
        // the sibling non-bonded term (earlier in the Hamiltonian) has already pushed the change
        if (!change.groups.empty() && have_old && !change.everything && !change.volume_change) {
            FlatChange flat(change);
            fbCheck(fb_ewald_update_partial(dev->ctx, slot, 1 - slot, &flat.change), dev->ctx,
                    "fb_ewald_update_partial");
        }
        else {
            fullUpdate();
        }
    }

    double energy(const Change& change) override
    {
        if (!change) {
            return 0.0;
        }
        if (dev->fast_staged && dev->fast_evaluated && dev->fast_move.with_ewald &&
            dev->fast_key == DeviceContext::changeKey(change)) {
            return state == MonteCarloState::TRIAL ? dev->fast_ew_new : dev->fast_ew_old;
        }
        FlatChange flat(change);
        double u = 0.0;
        fbCheck(fb_ewald_energy(dev->ctx, slot, &flat.change, &u), dev->ctx, "fb_ewald_energy");
        return u;
    }

    void sync(EnergyTerm* other_term, const Change& change) override
    {
        auto* other = dynamic_cast<EwaldB200*>(other_term);
        if (!other || other->dev != dev) {
            throw std::runtime_error("sync error");
        }
        if (!have_old && other->state == MonteCarloState::ACCEPTED) {
            have_old = true;
            dev->ewald_have_old = true;
        }
        if (dev->fast_sync_for_ewald) { // the sibling's fb_trial_commit swapped Q already
            dev->fast_sync_for_ewald = false;
            return;
        }
        FlatChange flat(change);
        fbCheck(fb_ewald_sync(dev->ctx, slot, other->slot, &flat.change), dev->ctx, "fb_ewald_sync");
    }

    void to_json(Json& j) const override
    {
        j["device"] = "B200 sm_100a";
        j["wavefunctions"] = n_kvectors;
    }
};

/**
 * ParticleSelfEnergy (src/externalpotential.cpp:94-126, 514-537) with the Coulomb scheme's
 * per-particle self energy lB·prefactor·q²/Rc (src/potentials.cpp:1599-1604). O(changed particles)
 * scalar work on the host; full sums are O(N) and only run with everything/volume changes.
 */
class ParticleSelfEnergyB200 : public EnergyTerm
{
    const Space& spc;
    double prefactor; //!< lB · self_prefactor / cutoff

    double groupEnergy(const Group& g) const
    {
        double e = 0;
        for (size_t i = 0; i < g.size(); ++i) {
            const double q = spc.at(g, i).charge;
            e += prefactor * q * q;
        }
        return e;
    }

  public:
    ParticleSelfEnergyB200(const Space& spc, const CoulombTable& c)
        : spc(spc)
        , prefactor(c.bjerrum_length * c.self_prefactor / c.cutoff)
    {
        name = "particle-self-energy";
    }
    double energy(const Change& change) override
    {
        double e = 0;
        if (change.volume_change || change.everything || change.matter_change) {
            for (const auto& g : spc.groups) {
                e += groupEnergy(g);
            }
            return e;
        }
        for (const auto& gc : change.groups) {
            const auto& g = spc.groups.at(gc.group_index);
            if (gc.all) {
                e += groupEnergy(g);
            }
            else {
                for (auto i : gc.relative_atom_indices) {
                    const double q = spc.at(g, i).charge;
                    e += prefactor * q * q;
                }
            }
        }
        return e;
    }
};

/** Hamiltonian term factory for the product build (order: self energy, nonbonded, ewald) */
inline bool b200TermFactory(Hamiltonian& h, Space& spc, const std::string& name, const Json& cfg)
{
    if (!isNonbondedName(name)) {
        return false;
    }
    auto nonbonded = std::make_shared<NonbondedB200>(name, cfg, spc);
    const auto& tables = nonbonded->device()->tables;
    if (tables.has_coulomb) { // CoulombGalore always defines a self energy functor (possibly zero)
        h.push_back(std::make_shared<ParticleSelfEnergyB200>(spc, tables.coulomb));
    }
    h.push_back(nonbonded);
    if (tables.ewald_json != nullptr) {
        h.push_back(std::make_shared<EwaldB200>(*tables.ewald_json, spc, nonbonded));
    }
    return true;
}

/**
 * Windowed evaluation of runs of `transrot` moves on the device (fb_batch_trial / fb_batch_commit):
 * the engine hands over a window of drawn proposals, one pass evaluates all of them against the
 * accepted state, and `energies()` assembles for each move — given the decisions on the earlier
 * ones — what `trial.energy(change)` and `accepted.energy(change)` return in the one-at-a-time
 * protocol (src/montecarlo.cpp:151-156; Hamiltonian::energy term order and early stop,
 * src/energy.cpp:1227-1241). Only additions of device results happen on the host.
 */
class B200WindowEvaluator : public WindowEvaluator
{
    MetropolisMonteCarlo& mc;
    std::shared_ptr<DeviceContext> dev;
    bool with_ewald = false;
    int window_capacity;
    fb_batch_result res{};
    std::vector<fb_batch_move> moves;
    std::vector<fb_batch_group_move> group_moves;
    bool group_mode = false;         //!< the window in flight holds rigid-molecule moves
    std::vector<int> first_atom;     //!< group mode: atoms [first_atom[m], first_atom[m + 1]) belong to move m
    int largest_molecule = 0;        //!< atoms in the largest molecular group
    std::vector<double> rec_change; //!< corrected raw reciprocal change Σ_k A_k(…) of each decided move
    // runs: windows of single-atom moves decided on the device (fb_run_submit), several windows per round trip
    int run_capacity = 0;            //!< 0: every window is walked on the host
    /** a run pays off from two windows on: a single window is walked on the host, which costs one kernel less */
    int run_threshold = FB_BATCH_MAX;
    bool run_mode = false;           //!< the evaluation in flight is a run
    std::vector<fb_run_move> run_moves;
    fb_run_config run_config{};
    fb_run_result run_res{};
    enum class Kind
    {
        SELF,
        NONBONDED,
        EWALD
    };
    std::vector<Kind> kinds; //!< Hamiltonian term order

  public:
    static constexpr double cancellation_limit = 1e4; //!< kT; larger pair terms in a correction → re-evaluate

    /** nullptr if the Hamiltonian holds anything but per-atom host terms (self energy, isobaric, container
     * overlap), one B200 non-bonded term and a tinfoil PBC Ewald term */
    static std::unique_ptr<B200WindowEvaluator> tryCreate(MetropolisMonteCarlo& mc, int capacity)
    {
        if (capacity < 1) {
            return nullptr;
        }
        std::vector<Kind> kinds;
        std::shared_ptr<NonbondedB200> nonbonded;
        const auto& trial_terms = mc.trial_state.pot->terms();
        const auto& terms = mc.state.pot->terms();
        if (terms.size() != trial_terms.size()) {
            return nullptr;
        }
        for (const auto& t : terms) {
            // host terms that only look at the atoms the Change lists (or at nothing for a particle move): their
            // energy(change) stays valid while OTHER proposals are pending in the trial Space
            if (std::dynamic_pointer_cast<ParticleSelfEnergyB200>(t) || std::dynamic_pointer_cast<Isobaric>(t) ||
                std::dynamic_pointer_cast<ContainerOverlap>(t)) {
                kinds.push_back(Kind::SELF);
            }
            else if (auto nb = std::dynamic_pointer_cast<NonbondedB200>(t)) {
                if (nonbonded) {
                    return nullptr;
                }
                nonbonded = nb;
                kinds.push_back(Kind::NONBONDED);
            }
            else if (std::dynamic_pointer_cast<EwaldB200>(t)) {
                kinds.push_back(Kind::EWALD);
            }
            else {
                return nullptr;
            }
        }
        if (!nonbonded || nonbonded->device()->attached != 2) {
            return nullptr;
        }
        const auto& d = *nonbonded->device();
        if (d.has_ewald && (!d.ewald_fast_ok || d.ewald_policy == 2)) {
            return nullptr; // surface term / IPBC: one move at a time
        }
        auto e = std::unique_ptr<B200WindowEvaluator>(new B200WindowEvaluator(mc, std::min(capacity, FB_BATCH_MAX)));
        e->dev = nonbonded->device();
        e->with_ewald = d.has_ewald;
        e->kinds = std::move(kinds);
        for (const auto& g : mc.state.spc->groups) {
            if (g.isMolecular()) {
                e->largest_molecule = std::max(e->largest_molecule, static_cast<int>(g.capacity()));
            }
        }
        return e;
    }

    int capacity() const override { return std::max(window_capacity, run_capacity); }

    /** let the device walk the windows of single-atom moves itself, up to `moves` (≤ FB_RUN_MAX) per round trip */
    void enableRuns(int moves, int min_moves = FB_BATCH_MAX + 1)
    {
        run_threshold = std::max(0, min_moves - 1);
        // the device walk adds [the caller's terms …, non-bonded, Ewald] in this order: the Hamiltonian must look so
        size_t i = 0;
        while (i < kinds.size() && kinds[i] == Kind::SELF) {
            ++i;
        }
        bool layout_ok = i < kinds.size() && kinds[i] == Kind::NONBONDED;
        ++i;
        if (with_ewald) {
            layout_ok = layout_ok && i < kinds.size() && kinds[i] == Kind::EWALD;
            ++i;
        }
        layout_ok = layout_ok && i == kinds.size();
        run_capacity = (layout_ok && window_capacity == FB_BATCH_MAX) ? std::max(0, std::min(moves, FB_RUN_MAX)) : 0;
    }
    int runCapacity() const { return run_capacity; }

    bool supports(WindowProposal::Kind kind) const override
    {
        return kind == WindowProposal::Kind::ATOM || (largest_molecule >= 1 && largest_molecule <= FB_FAST_ATOMS);
    }

    /** group mode: a window holds at most FB_BATCH_MAX ATOMS */
    int fit(const std::vector<WindowProposal>& window, int first, int ready) const override
    {
        if (ready > run_threshold && run_capacity > 0 && window[first].kind == WindowProposal::Kind::ATOM) {
            return std::min(ready, run_capacity);
        }
        int n = std::min(ready, window_capacity);
        for (int m = 1; m < n; ++m) { // walked on the host: a conditional proposal waits for the window after
            if (window[first + m].conditional && !window[first + m].applied) {
                n = m;
            }
        }
        if (n > 0 && window[first].kind == WindowProposal::Kind::GROUP) {
            const Space& trial = *mc.trial_state.spc;
            int atoms = 0;
            int m = 0;
            for (; m < n; ++m) {
                atoms += static_cast<int>(trial.groups.at(window[first + m].change.groups.at(0).group_index).size());
                if (atoms > FB_BATCH_MAX) {
                    break;
                }
            }
            n = std::max(1, m);
        }
        return n;
    }

    void submitGroups(const std::vector<WindowProposal>& all, int first, int n)
    {
        dev->resynchronise();
        const WindowProposal* window = all.data() + first;
        group_moves.resize(static_cast<size_t>(n));
        first_atom.assign(static_cast<size_t>(n) + 1, 0);
        const Space& trial = *mc.trial_state.spc;
        const Space& accepted = *mc.state.spc;
        for (int m = 0; m < n; ++m) {
            const auto& gc = window[m].change.groups.at(0);
            const auto& g = trial.groups.at(gc.group_index);
            const auto& g_old = accepted.groups.at(gc.group_index);
            fb_batch_group_move& mv = group_moves[m];
            if (static_cast<int>(g.size()) > FB_FAST_ATOMS || g.size() != g_old.size() || !gc.all || gc.internal) {
                throw std::runtime_error("windowed moltransrot: unexpected change record");
            }
            mv.group_index = static_cast<int>(gc.group_index);
            mv.n_atoms = static_cast<int>(g.size());
            first_atom[m + 1] = first_atom[m] + mv.n_atoms;
            for (int i = 0; i < mv.n_atoms; ++i) {
                const auto& p = trial.at(g, i);
                const auto& q = accepted.at(g_old, i);
                mv.atom_id[i] = p.id;
                mv.xyzq[i][0] = p.pos.x;
                mv.xyzq[i][1] = p.pos.y;
                mv.xyzq[i][2] = p.pos.z;
                mv.xyzq[i][3] = p.charge;
                mv.old_atom_id[i] = q.id;
                mv.old_xyzq[i][0] = q.pos.x;
                mv.old_xyzq[i][1] = q.pos.y;
                mv.old_xyzq[i][2] = q.pos.z;
                mv.old_xyzq[i][3] = q.charge;
            }
            mv.cm[0] = g.mass_center.x;
            mv.cm[1] = g.mass_center.y;
            mv.cm[2] = g.mass_center.z;
            mv.old_cm[0] = g_old.mass_center.x;
            mv.old_cm[1] = g_old.mass_center.y;
            mv.old_cm[2] = g_old.mass_center.z;
        }
        dev->fast_staged = false;
        dev->cache_valid = false;
        fbCheck(fb_batch_submit_groups(dev->ctx, n, group_moves.data(), with_ewald ? 1 : 0), dev->ctx,
                "fb_batch_submit_groups");
        rec_change.assign(static_cast<size_t>(n), 0.0);
    }

    static void fillMove(fb_batch_move& mv, const Change::GroupChange& gc, const Particle& particle, const Point& start,
                         const Point& trial)
    {
        mv.group_index = static_cast<int>(gc.group_index);
        mv.rel_index = static_cast<int>(gc.relative_atom_indices[0]);
        mv.atom_id = particle.id;
        mv.xyzq[0] = trial.x;
        mv.xyzq[1] = trial.y;
        mv.xyzq[2] = trial.z;
        mv.xyzq[3] = particle.charge;
        mv.old_atom_id = particle.id; // a translation: id and charge stay
        mv.old_xyzq[0] = start.x;
        mv.old_xyzq[1] = start.y;
        mv.old_xyzq[2] = start.z;
        mv.old_xyzq[3] = particle.charge;
    }

    /** single-atom proposals → fb_batch_move records (trial and accepted positions: the caller owns the Space) */
    void packMoves(const WindowProposal* window, int n)
    {
        moves.resize(static_cast<size_t>(n));
        const Space& trial = *mc.trial_state.spc;
        const Space& accepted = *mc.state.spc;
        for (int m = 0; m < n; ++m) {
            const auto& gc = window[m].change.groups.at(0);
            const auto& g = trial.groups.at(gc.group_index);
            const auto& p = trial.at(g, gc.relative_atom_indices.at(0));
            fb_batch_move& mv = moves[m];
            if (window[m].conditional && !window[m].applied) { // the variant "the earlier move on this atom is accepted"
                fillMove(mv, gc, p, window[m].alt_start[0], window[m].alt_new[0]);
                continue;
            }
            mv.group_index = static_cast<int>(gc.group_index);
            mv.rel_index = static_cast<int>(gc.relative_atom_indices[0]);
            mv.atom_id = p.id;
            mv.xyzq[0] = p.pos.x;
            mv.xyzq[1] = p.pos.y;
            mv.xyzq[2] = p.pos.z;
            mv.xyzq[3] = p.charge;
            const auto& q = accepted.at(accepted.groups.at(gc.group_index), gc.relative_atom_indices[0]);
            mv.old_atom_id = q.id;
            mv.old_xyzq[0] = q.pos.x;
            mv.old_xyzq[1] = q.pos.y;
            mv.old_xyzq[2] = q.pos.z;
            mv.old_xyzq[3] = q.charge;
        }
    }

    bool pipelined(const std::vector<WindowProposal>& window, int first, int ready) const override
    {
        return run_mode && run_capacity > 0 && ready > run_threshold && window[first].kind == WindowProposal::Kind::ATOM;
    }

    bool conditionals() const override { return run_capacity > 0; }

    /** the proposals of a run travel with their Metropolis uniform and the host terms' energies */
    void prepare(const std::vector<WindowProposal>& all, int first, int n) override
    {
        const WindowProposal* window = all.data() + first;
        packMoves(window, n);
        run_moves.resize(static_cast<size_t>(n));
        run_config = fb_run_config{};
        run_config.max_energy = mc.state.pot->maximumAllowedEnergy();
        run_config.cancellation_limit = cancellation_limit;
        const auto& trial_terms = mc.trial_state.pot->terms();
        const auto& terms = mc.state.pot->terms();
        // Hamiltonian::energy over the caller's own terms, which all precede the device terms (enableRuns): they look
        // at the atoms of the Change only, so their energies are valid while other proposals are pending
        auto leading_sum = [&](const std::vector<std::shared_ptr<EnergyTerm>>& list, EnergyTerm::MonteCarloState state,
                               const Change& change, bool& closed) {
            double sum = 0.0;
            closed = false;
            for (size_t i = 0; i < kinds.size() && kinds[i] == Kind::SELF; ++i) {
                list[i]->state = state;
                const double u = list[i]->energy(change);
                sum += u;
                if (u >= run_config.max_energy || std::isnan(u)) {
                    closed = true;
                    break;
                }
            }
            return sum;
        };
        for (int m = 0; m < n; ++m) {
            fb_run_move& r = run_moves[m];
            r.move = moves[m];
            r.uniform = window[m].uniform;
            r.depends_on = -1;
            r.alt = moves[m];
            r.alt_host_new = r.alt_host_old = 0.0;
            bool closed_new = false, closed_old = false;
            if (window[m].conditional && !window[m].applied) {
                // the proposal it depends on: earlier in the queue — in this run, or in the run in flight, whose
                // moves are the first `first` proposals of the queue
                int at = first + m - 1;
                while (at >= 0 && all[at].serial != window[m].dependency) {
                    --at;
                }
                if (at < 0) {
                    throw std::runtime_error("windowed evaluation: lost the proposal a conditional one depends on");
                }
                r.depends_on = at >= first ? at - first : (FB_RUN_DEP_PREVIOUS | at);
                const auto& gc = window[m].change.groups.at(0);
                auto& trial_particle = mc.trial_state.spc->at(mc.trial_state.spc->groups.at(gc.group_index),
                                                              gc.relative_atom_indices.at(0));
                auto& particle = mc.state.spc->at(mc.state.spc->groups.at(gc.group_index), gc.relative_atom_indices[0]);
                fillMove(r.alt, gc, trial_particle, window[m].alt_start[1], window[m].alt_new[1]);
                // the caller's terms see the Spaces: show them either outcome in turn
                const Point keep_trial = trial_particle.pos, keep = particle.pos;
                double host[2][2];
                bool closed[2][2];
                for (int v = 0; v < 2; ++v) {
                    trial_particle.pos = window[m].alt_new[v];
                    particle.pos = window[m].alt_start[v];
                    host[v][0] = leading_sum(trial_terms, mc.trial_state.pot->state, window[m].change, closed[v][0]);
                    host[v][1] = leading_sum(terms, mc.state.pot->state, window[m].change, closed[v][1]);
                }
                trial_particle.pos = keep_trial;
                particle.pos = keep;
                r.host_new = host[0][0];
                r.host_old = host[0][1];
                r.alt_host_new = host[1][0];
                r.alt_host_old = host[1][1];
                r.flags = (closed[0][0] ? FB_RUN_HOST_NEW_CLOSED : 0) | (closed[0][1] ? FB_RUN_HOST_OLD_CLOSED : 0) |
                          (closed[1][0] ? FB_RUN_ALT_HOST_NEW_CLOSED : 0) | (closed[1][1] ? FB_RUN_ALT_HOST_OLD_CLOSED : 0);
                continue;
            }
            r.host_new = leading_sum(trial_terms, mc.trial_state.pot->state, window[m].change, closed_new);
            r.host_old = leading_sum(terms, mc.state.pot->state, window[m].change, closed_old);
            r.flags = (closed_new ? FB_RUN_HOST_NEW_CLOSED : 0) | (closed_old ? FB_RUN_HOST_OLD_CLOSED : 0);
        }
    }

    void submitPrepared() override
    {
        dev->resynchronise();
        dev->fast_staged = false;
        dev->cache_valid = false;
        fbCheck(fb_run_submit(dev->ctx, static_cast<int>(run_moves.size()), run_moves.data(), with_ewald ? 1 : 0,
                              &run_config),
                dev->ctx, "fb_run_submit");
        run_mode = true;
    }

    void submit(const std::vector<WindowProposal>& window, int first, int n) override
    {
        group_mode = n > 0 && window[first].kind == WindowProposal::Kind::GROUP;
        run_mode = false;
        if (group_mode) {
            submitGroups(window, first, n);
            return;
        }
        if (run_capacity > 0 && n > run_threshold) {
            prepare(window, first, n);
            submitPrepared();
            return;
        }
        packMoves(window.data() + first, n);
        dev->resynchronise();
        dev->fast_staged = false;
        dev->cache_valid = false;
        fbCheck(fb_batch_submit(dev->ctx, n, moves.data(), with_ewald ? 1 : 0), dev->ctx, "fb_batch_submit");
        rec_change.assign(static_cast<size_t>(n), 0.0);
    }

    void wait() override
    {
        if (run_mode) {
            fbCheck(fb_run_wait(dev->ctx, &run_res), dev->ctx, "fb_run_wait");
        }
        else {
            fbCheck(fb_batch_wait(dev->ctx, &res), dev->ctx, "fb_batch_wait");
        }
    }

    int decision(int m) const override { return run_mode ? static_cast<int>(run_res.accepted[m]) : -1; }

    bool energies(int m, const std::vector<unsigned char>& accepted, const WindowProposal& proposal,
                  double& new_energy, double& old_energy) override
    {
        if (run_mode) { // decided on the device, in the same order with the same numbers
            new_energy = run_res.u_new[m];
            old_energy = run_res.u_old[m];
            return m < run_res.n_moves;
        }
        const size_t S = static_cast<size_t>(res.stride);
        double nb_new = res.u_new[m];
        double nb_old = res.u_old[m];
        double rec = 0.0;
        if (with_ewald && group_mode) { // the molecule's δ is the sum of its atoms' δ (fb_batch_submit_groups)
            for (int j = first_atom[m]; j < first_atom[m + 1]; ++j) {
                rec += res.rec_delta[j];
                for (int i = first_atom[m]; i < j; ++i) {
                    rec += 2.0 * res.rec_cross[static_cast<size_t>(j) * S + i];
                }
            }
        }
        else if (with_ewald) {
            rec = res.rec_delta[m];
        }
        double rec_running = res.rec_start;
        for (int a = 0; a < m; ++a) {
            if (!accepted[a]) {
                continue;
            }
            const size_t am = static_cast<size_t>(m) * S + a; // row m: contiguous over the earlier moves
            if (!(res.cross_max[am] < cancellation_limit)) {
                return false;
            }
            nb_new += res.cross_new[am];
            nb_old += res.cross_old[am];
            if (with_ewald) {
                if (group_mode) {
                    double cross = 0.0;
                    for (int j = first_atom[m]; j < first_atom[m + 1]; ++j) {
                        for (int i = first_atom[a]; i < first_atom[a + 1]; ++i) {
                            cross += res.rec_cross[static_cast<size_t>(j) * S + i];
                        }
                    }
                    rec += 2.0 * cross;
                }
                else {
                    rec += 2.0 * res.rec_cross[am];
                }
                rec_running += rec_change[a];
            }
        }
        rec_change[m] = rec;
        // Hamiltonian::energy on the trial and on the accepted state
        const double limit = mc.state.pot->maximumAllowedEnergy();
        auto total = [&](Hamiltonian& pot, bool is_trial) {
            double sum = 0.0;
            const auto& terms = pot.terms();
            for (size_t i = 0; i < terms.size(); ++i) {
                double u = 0.0;
                switch (kinds[i]) {
                case Kind::SELF:
                    terms[i]->state = pot.state;
                    u = terms[i]->energy(proposal.change);
                    break;
                case Kind::NONBONDED:
                    u = is_trial ? nb_new : nb_old;
                    break;
                case Kind::EWALD:
                    u = res.rec_prefactor * (is_trial ? rec_running + rec : rec_running);
                    break;
                }
                sum += u;
                if (u >= limit || std::isnan(u)) {
                    break;
                }
            }
            return sum;
        };
        new_energy = total(*mc.trial_state.pot, true);
        old_energy = total(*mc.state.pot, false);
        return true;
    }

    void commit(const std::vector<unsigned char>& accepted) override
    {
        if (run_mode) {
            return; // the device keeps its own books
        }
        fbCheck(fb_batch_commit(dev->ctx, static_cast<int>(accepted.size()), accepted.data()), dev->ctx,
                "fb_batch_commit");
    }

  private:
    B200WindowEvaluator(MetropolisMonteCarlo& mc, int capacity)
        : mc(mc)
        , window_capacity(capacity)
    {
    }
};

/**
 * Batched Widom insertion: all `ninsert` ghosts of one sample event are generated on the host in the
 * reference's RNG order (they never depend on energies), evaluated in ONE launch, and exp(−ΔU) is
 * accumulated on the host in the reference's sequential order (bit-compatible `Average`).
 * Replaces the loop of src/analysis.cpp:1255-1262; scalar terms (self energy, container overlap)
 * come from the host Hamiltonian terms.
 */
class WidomB200 : public WidomInsertion
{
    std::shared_ptr<NonbondedB200> nonbonded;
    std::vector<std::shared_ptr<EnergyTerm>> other_terms;

  public:
    WidomB200(const Json& j, MetropolisMonteCarlo& mc)
        : WidomInsertion(j, *mc.state.spc, *mc.state.pot, mc.rng.global)
    {
        const auto nb = mc.state.pot->find<NonbondedB200>();
        if (nb.size() != 1) {
            throw std::runtime_error("batched Widom needs exactly one B200 non-bonded term");
        }
        nonbonded = nb.front();
        for (const auto& t : mc.state.pot->terms()) {
            if (t != nonbonded && !std::dynamic_pointer_cast<EwaldB200>(t)) {
                other_terms.push_back(t);
            }
        }
    }

    /** ΔU of insertions [first, first + count) of the prepared event: ONE launch for the whole slice */
    void evaluateSlice(int first, int count, double* du_total) override
    {
        checkSlice(first, count);
        if (count == 0) {
            return;
        }
        nonbonded->device()->resynchronise();
        const Change& change = prepared.change;
        const size_t gi = change.groups.at(0).group_index;
        auto& group = spc.groups.at(gi);
        const int n_g = static_cast<int>(group.capacity());
        std::vector<double> xyzq(static_cast<size_t>(count) * n_g * 4), cm(static_cast<size_t>(count) * 3);
        std::vector<double> host_terms(static_cast<size_t>(count), 0.0), du(static_cast<size_t>(count));
        std::vector<int> ids(static_cast<size_t>(n_g));
        group.resize(group.capacity());
        // Ewald terms see a Q(k) without the ghost (Widom never calls updateState):
        // their energy is the same for every insertion and is evaluated once.
        double ewald_energy = 0.0;
        for (const auto& t : pot.find<EwaldB200>()) {
            ewald_energy += t->energy(change);
        }
        for (int b = 0; b < count; ++b) {
            updateGroup(group, prepared.ghosts[first + b]);
            for (int a = 0; a < n_g; ++a) {
                const auto& p = spc.at(group, a);
                const size_t o = (static_cast<size_t>(b) * n_g + a) * 4;
                xyzq[o] = p.pos.x;
                xyzq[o + 1] = p.pos.y;
                xyzq[o + 2] = p.pos.z;
                xyzq[o + 3] = p.charge;
                ids[a] = p.id;
            }
            cm[3 * b] = group.mass_center.x;
            cm[3 * b + 1] = group.mass_center.y;
            cm[3 * b + 2] = group.mass_center.z;
            for (const auto& t : other_terms) { // scalar host terms, in Hamiltonian order semantics
                t->state = pot.state;
                host_terms[b] += t->energy(change);
            }
        }
        const bool molecular = group.isMolecular();
        group.resize(0);
        auto& dev = *nonbonded->device();
        fbCheck(fb_widom_batch(dev.ctx, nonbonded->deviceSlot(), static_cast<int>(gi), n_g, count, xyzq.data(),
                               ids.data(), molecular ? cm.data() : nullptr, change.groups[0].internal ? 1 : 0,
                               du.data()),
                dev.ctx, "fb_widom_batch");
        for (int b = 0; b < count; ++b) {
            du_total[b] = host_terms[b] + du[b] + ewald_energy;
        }
    }
};

/**
 * `atomrdf` with the pair loop on the device: one fb_atom_rdf per sample on the accepted slot's mirror (the same
 * mirror the energy terms keep current), exact pair counts added to the histogram. Replaces
 * AtomRDF::sampleIdentical / sampleDifferent (src/analysis.cpp:1581-1600).
 */
class AtomRDFB200 : public AtomRDF
{
    std::shared_ptr<NonbondedB200> nonbonded;

    void count(int shard, int n_shards) override
    {
        const int n_bins = std::max(static_cast<int>(histogram.size()), binsForCell());
        histogram.resize(static_cast<size_t>(n_bins), 0ull);
        auto& dev = *nonbonded->device();
        dev.resynchronise();
        if (molecular) {
            fbCheck(fb_molecule_rdf(dev.ctx, nonbonded->deviceSlot(), id1, id2, dr, shard, n_shards, n_bins, histogram.data()),
                    dev.ctx, "fb_molecule_rdf");
            return;
        }
        fbCheck(fb_atom_rdf(dev.ctx, nonbonded->deviceSlot(), id1, id2, dr, slicedir, thickness, shard, n_shards, n_bins,
                            histogram.data()),
                dev.ctx, "fb_atom_rdf");
    }

  public:
    AtomRDFB200(const Json& j, MetropolisMonteCarlo& mc)
        : AtomRDF(j, *mc.state.spc)
    {
        const auto nb = mc.state.pot->find<NonbondedB200>();
        if (nb.size() != 1) {
            throw std::runtime_error("atomrdf on the device needs exactly one B200 non-bonded term");
        }
        nonbonded = nb.front();
    }
};

/**
 * The Space follows a packed mirror state (fb_export_state layout: box, group sizes, x y z q id of every particle):
 * volume (exchangeVolume, src/mpicontroller.cpp:231-246), group sizes (src/move.cpp:860-867), all particles incl.
 * inactive ones (ExchangeParticles::replace, src/mpicontroller.cpp:208-219), mass centres (src/move.cpp:879).
 */
inline void applyPackedState(Space& spc, const std::vector<double>& state, VolumeMethod method, Change& change)
{
    const double old_volume = spc.geometry.getVolume();
    const double new_volume = state[0] * state[1] * state[2];
    if (new_volume <= pc::epsilon_dbl) {
        throw std::runtime_error("tempering: invalid partner volume");
    }
    if (std::fabs(new_volume - old_volume) > pc::epsilon_dbl) {
        spc.geometry.setVolume(new_volume, method);
        change.volume_change = true;
    }
    const size_t n_groups = spc.groups.size();
    if (state.size() != 3 + n_groups + 5 * spc.particles.size()) {
        throw std::runtime_error("tempering: the partner's state has a different layout");
    }
    for (size_t g = 0; g < n_groups; ++g) {
        spc.groups[g].resize(static_cast<size_t>(state[3 + g]));
    }
    const double* p = state.data() + 3 + n_groups;
    for (auto& particle : spc.particles) {
        particle.pos = {p[0], p[1], p[2]};
        particle.charge = p[3];
        particle.id = static_cast<int>(p[4]);
        p += 5;
    }
    spc.updateMassCenters();
    change.everything = true;
}

/**
 * Replica exchange of the packed mirror through a communicator that only moves doubles between host buffers
 * (in-process replicas, tests): export → sendrecv → import on the device, the Space follows. The same steps as
 * NcclReplicaComm::exchangeState with the NVLink hop replaced by the communicator.
 */
inline bool exchangePackedStateThroughHost(DeviceContext& dev, int trial_slot, ReplicaComm& comm, Space& spc, int partner,
                                           VolumeMethod method, Change& change)
{
    dev.resynchronise();
    std::vector<double> state(fb_state_doubles(dev.ctx));
    // the ACCEPTED mirror is the authoritative copy of the state the two Spaces share at this point
    fbCheck(fb_export_state_host(dev.ctx, 1 - trial_slot, state.data()), dev.ctx, "fb_export_state_host");
    comm.sendrecvReplace(state.data(), state.size(), partner);
    fbCheck(fb_import_state_host(dev.ctx, trial_slot, state.data()), dev.ctx, "fb_import_state_host");
    applyPackedState(spc, state, method, change);
    dev.particles_current[trial_slot] = true; // updateState(everything) sends box + group records only
    return true;
}

/**
 * Replica communicator on the device library's own NCCL communicator (fb_nccl_*, one context per GPU / process):
 * the messages of the Temper move (src/move.cpp:844-968) without the launcher in the loop — the packed mirror of the
 * trial state goes GPU to GPU and is imported on the device, the 8-byte messages go through pinned staging. The
 * communicator is created when the first message is due (collective: every replica tempers in the same sweep).
 */
class NcclReplicaComm : public ReplicaComm
{
    std::array<char, 128> id{};
    int my_rank, n_ranks;
    std::shared_ptr<DeviceContext> dev;
    int trial_slot = 1;
    bool ready = false;

    fb_ctx* ctx()
    {
        if (!dev) {
            throw std::runtime_error("NCCL replica communicator: no device context bound");
        }
        if (!ready) {
            fbCheck(fb_nccl_init(dev->ctx, id.data(), my_rank, n_ranks), dev->ctx, "fb_nccl_init");
            ready = true;
        }
        return dev->ctx;
    }

  public:
    unsigned long exchanges = 0;

    NcclReplicaComm(const char unique_id[128], int rank, int size)
        : my_rank(rank)
        , n_ranks(size)
    {
        std::copy(unique_id, unique_id + 128, id.begin());
        if (size < 2 || rank < 0 || rank >= size) {
            throw std::runtime_error("NCCL replica communicator: bad rank / size");
        }
    }
    /** the device context of the simulation's non-bonded term and the slot of its TRIAL state */
    void bind(std::shared_ptr<DeviceContext> device, int slot_of_trial_state)
    {
        dev = std::move(device);
        trial_slot = slot_of_trial_state;
    }
    int rank() const override { return my_rank; }
    int size() const override { return n_ranks; }
    void barrier() override { (void)gather(0.0); }
    void sendrecvReplace(double* data, size_t n, int partner) override
    {
        fbCheck(fb_nccl_sendrecv_host(ctx(), data, n, partner), dev->ctx, "fb_nccl_sendrecv_host");
        exchanges++;
    }
    std::vector<double> gather(double value) override
    {
        std::vector<double> out(static_cast<size_t>(n_ranks), value);
        fbCheck(fb_nccl_allgather_host(ctx(), value, out.data()), dev->ctx, "fb_nccl_allgather_host");
        return out;
    }
    bool exchangeState(Space& spc, int partner, VolumeMethod method, Change& change) override
    {
        fb_ctx* c = ctx();
        dev->resynchronise();
        std::vector<double> state(fb_state_doubles(c));
        // the ACCEPTED mirror is the authoritative copy of the state the two Spaces share at this point (the one-move
        // fast path never writes trial positions into the trial mirror); the partner's state lands in the trial slot
        fbCheck(fb_nccl_exchange_state(c, 1 - trial_slot, trial_slot, partner, state.data()), c, "fb_nccl_exchange_state");
        exchanges++;
        applyPackedState(spc, state, method, change);
        dev->particles_current[trial_slot] = true; // updateState(everything) sends box + group records only
        return true;
    }
};

} // namespace fb

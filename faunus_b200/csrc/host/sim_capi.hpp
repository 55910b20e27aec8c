// Host-level C ABI over the MC engine (`MetropolisMonteCarlo`, `WidomInsertion`) used by the Python
// harness (ctypes), tests and bench. The same driver code is instantiated twice with different
// energy-term factories: `fbh_*` in libfaunus_b200.so (B200 adaptor terms) and `fo_*` in the oracle
// library (CPU restatement). Nothing here throws across the ABI: functions return a negative code
// and `<prefix>_last_error()` holds the message.
#pragma once
#include "montecarlo.hpp"
#include "analysis_rdf.hpp"
#include "analysis_virtual.hpp"
#include "replica_comm.hpp"
#include <cstring>
#include <mutex>
#include <thread>

namespace fb::capi {

inline thread_local std::string last_error;

template <class F> int guarded(F&& f)
{
    try {
        f();
        return 0;
    }
    catch (const std::exception& e) {
        last_error = e.what();
        return -1;
    }
    catch (...) {
        last_error = "unknown exception";
        return -1;
    }
}

struct Sim
{
    std::unique_ptr<ReplicaComm> comm; //!< must outlive mc (the temper move holds a reference)
    std::unique_ptr<MetropolisMonteCarlo> mc;
    std::vector<std::unique_ptr<WidomInsertion>> widoms;
    std::vector<std::unique_ptr<AtomRDF>> rdfs;
    std::vector<std::unique_ptr<VirtualVolumeMove>> virtual_volumes;
    std::vector<std::unique_ptr<VirtualTranslate>> virtual_translates;
    Change pending; //!< change of the manual trial-move protocol
};

using WidomFactory =
    std::function<std::unique_ptr<WidomInsertion>(const Json&, MetropolisMonteCarlo&)>;

inline std::unique_ptr<WidomInsertion> defaultWidom(const Json& j, MetropolisMonteCarlo& mc)
{
    return std::make_unique<WidomInsertion>(j, *mc.state.spc, *mc.state.pot, mc.rng.global);
}

inline Sim* create(const char* json_text, const TermFactory& factory, std::unique_ptr<ReplicaComm> comm = nullptr)
{
    Sim* sim = nullptr;
    const int rc = guarded([&] {
        const Json j = Json::parse(json_text);
        auto s = std::make_unique<Sim>();
        s->comm = std::move(comm);
        s->mc = std::make_unique<MetropolisMonteCarlo>(j, factory, s->comm.get());
        sim = s.release();
    });
    return rc == 0 ? sim : nullptr;
}

/** Per-replica result of an in-process tempering run */
struct LocalReplicaResult
{
    double energy = 0;
    double drift = 0;
    std::vector<double> xyzq;
    std::string info;
    std::string error;
};

/**
 * All replicas of a parallel-tempering run in ONE process, one thread per replica (LocalComm).
 * `configs` is a JSON array of per-replica input documents; `setup(rank)` runs in the replica's thread
 * before its simulation is created (e.g. selects the CUDA device).
 */
inline std::vector<LocalReplicaResult> runLocalReplicas(
    const Json& configs, int sweeps, const TermFactory& factory, const std::function<void(int)>& setup,
    const std::function<void(MetropolisMonteCarlo&, LocalComm&)>& after_create = nullptr)
{
    const int size = static_cast<int>(configs.size());
    auto exchange = std::make_shared<LocalExchange>(size);
    std::vector<LocalReplicaResult> results(static_cast<size_t>(size));
    std::vector<std::thread> threads;
    for (int r = 0; r < size; ++r) {
        threads.emplace_back([&, r] {
            auto comm = std::make_unique<LocalComm>(exchange, r);
            LocalComm* raw = comm.get();
            try {
                if (setup) {
                    setup(r);
                }
                // the constructor sets the process-global pc::temperature and converts units / builds tables from
                // it (the reference runs one process per replica): replicas are constructed one at a time
                static std::mutex construction;
                std::unique_ptr<MetropolisMonteCarlo> owner;
                {
                    std::lock_guard<std::mutex> lock(construction);
                    owner = std::make_unique<MetropolisMonteCarlo>(configs.at(r), factory, raw);
                }
                MetropolisMonteCarlo& mc = *owner;
                if (after_create) {
                    after_create(mc, *raw);
                }
                for (int i = 0; i < sweeps; ++i) {
                    mc.sweep();
                }
                auto& res = results[r];
                res.energy = mc.systemEnergy();
                res.drift = mc.relativeEnergyDrift();
                for (const auto& p : mc.state.spc->particles) {
                    res.xyzq.insert(res.xyzq.end(), {p.pos.x, p.pos.y, p.pos.z, p.charge});
                }
                Json jm = Json::array();
                for (const auto& m : mc.moves->all()) {
                    Json inner = Json::object();
                    m->to_json(inner);
                    Json w = Json::object();
                    w[m->name] = inner;
                    jm.push_back(w);
                }
                res.info = jm.dump();
            }
            catch (const std::exception& e) {
                results[r].error = e.what();
                raw->fail();
            }
        });
    }
    for (auto& t : threads) {
        t.join();
    }
    return results;
}

inline int copyOut(const std::string& s, char* buf, int len)
{
    if (buf == nullptr || len <= 0) {
        return static_cast<int>(s.size()) + 1;
    }
    const int n = std::min<int>(len - 1, static_cast<int>(s.size()));
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
    return static_cast<int>(s.size()) + 1;
}

/** Build a Change from a flat description (tests drive energy(change) directly with it) */
inline Change makeChange(int everything, int volume_change, int group_index, int all, int internal,
                         const int* indices, int n_indices)
{
    Change c;
    c.everything = everything != 0;
    c.volume_change = volume_change != 0;
    if (group_index >= 0) {
        auto& gc = c.groups.emplace_back();
        gc.group_index = static_cast<size_t>(group_index);
        gc.all = all != 0;
        gc.internal = internal != 0;
        for (int i = 0; i < n_indices; ++i) {
            gc.relative_atom_indices.push_back(static_cast<size_t>(indices[i]));
        }
    }
    return c;
}

inline State& pick(Sim& s, int which)
{
    return which == 0 ? s.mc->state : s.mc->trial_state;
}

} // namespace fb::capi

/**
 * Expands to the extern "C" entry points `<P>_sim_*` / `<P>_widom_*`.
 * FACTORY: expression yielding a `const fb::TermFactory&`; WIDOM: a `fb::capi::WidomFactory`.
 */
#define FB_DEFINE_SIM_CAPI(P, FACTORY, WIDOM)                                                                 \
    extern "C" {                                                                                              \
    __attribute__((visibility("default"))) const char* P##_last_error()                                      \
    {                                                                                                         \
        return fb::capi::last_error.c_str();                                                                 \
    }                                                                                                         \
    __attribute__((visibility("default"))) void* P##_sim_create(const char* json_text)                       \
    {                                                                                                         \
        return fb::capi::create(json_text, FACTORY);                                                         \
    }                                                                                                         \
    __attribute__((visibility("default"))) void* P##_sim_create_replica(const char* json_text,               \
                                                                        const fb_replica_callbacks* cb)      \
    {                                                                                                         \
        std::unique_ptr<fb::ReplicaComm> comm;                                                               \
        if (fb::capi::guarded([&] { comm = std::make_unique<fb::CallbackComm>(*cb); }) != 0) {               \
            return nullptr;                                                                                  \
        }                                                                                                    \
        return fb::capi::create(json_text, FACTORY, std::move(comm));                                        \
    }                                                                                                         \
    /* in-process tempering: configs = JSON array of inputs; returns JSON array of per-replica results */    \
    __attribute__((visibility("default"))) int P##_temper_run_local(const char* configs_json, int sweeps,    \
                                                                    char* buf, int len)                      \
    {                                                                                                         \
        int n = -1;                                                                                          \
        fb::capi::guarded([&] {                                                                              \
            const auto configs = fb::Json::parse(configs_json);                                              \
            const auto results = fb::capi::runLocalReplicas(configs, sweeps, FACTORY, P##_replica_setup);    \
            fb::Json out = fb::Json::array();                                                                \
            for (const auto& r : results) {                                                                  \
                fb::Json j = fb::Json::object();                                                             \
                j["energy"] = r.energy;                                                                      \
                j["drift"] = r.drift;                                                                        \
                j["xyzq"] = fb::Json::fromVector(r.xyzq);                                                    \
                j["moves"] = r.info.empty() ? fb::Json() : fb::Json::parse(r.info);                          \
                j["error"] = r.error;                                                                        \
                out.push_back(j);                                                                            \
            }                                                                                                \
            n = fb::capi::copyOut(out.dump(), buf, len);                                                     \
        });                                                                                                  \
        return n;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) void P##_sim_destroy(void* h)                                     \
    {                                                                                                         \
        delete static_cast<fb::capi::Sim*>(h);                                                               \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_restore(void* h, const char* state_json)              \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->mc->restore(fb::Json::parse(state_json)); });                      \
    }                                                                                                         \
    /* state files as the reference writes and reads them: .json or .ubj by suffix                           \
       (SaveState, src/analysis.cpp:640-682; --state, src/faunus.cpp:430-455) */                            \
    __attribute__((visibility("default"))) int P##_sim_save_state(void* h, const char* filename,             \
                                                                  int save_random)                           \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->mc->saveStateFile(filename, save_random != 0); });                 \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_load_state(void* h, const char* filename)             \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->mc->restoreFile(filename); });                                     \
    }                                                                                                         \
    /* JSON text -> UBJSON bytes and back (known-answer tests of the encoding); return the size needed */    \
    __attribute__((visibility("default"))) int P##_json_to_ubjson(const char* text, char* buf, int len)      \
    {                                                                                                         \
        int n = -1;                                                                                          \
        fb::capi::guarded([&] {                                                                              \
            const std::string bytes = fb::Json::parse(text).toUbjson();                                      \
            n = static_cast<int>(bytes.size());                                                              \
            if (buf != nullptr && len >= n) {                                                                \
                std::memcpy(buf, bytes.data(), bytes.size());                                                \
            }                                                                                                \
        });                                                                                                  \
        return n;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_ubjson_to_json(const char* bytes, int n_bytes, char* buf, \
                                                                  int len)                                   \
    {                                                                                                         \
        int n = -1;                                                                                          \
        fb::capi::guarded([&] {                                                                              \
            n = fb::capi::copyOut(fb::Json::fromUbjson(std::string(bytes, bytes + n_bytes)).dump(), buf,     \
                                  len);                                                                      \
        });                                                                                                  \
        return n;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_sweep(void* h, int n)                                 \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            for (int i = 0; i < n; ++i) {                                                                    \
                s->mc->sweep();                                                                              \
            }                                                                                                \
        });                                                                                                  \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_moves_per_sweep(void* h)                              \
    {                                                                                                         \
        return static_cast<int>(static_cast<fb::capi::Sim*>(h)->mc->moves->movesPerSweep());                 \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_system_energy(void* h, double* total, double* terms,  \
                                                                     int max_terms, int* n_terms)            \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            std::vector<double> per;                                                                         \
            *total = s->mc->systemEnergy(&per);                                                              \
            if (n_terms) {                                                                                   \
                *n_terms = static_cast<int>(per.size());                                                     \
            }                                                                                                \
            for (int i = 0; terms && i < max_terms && i < static_cast<int>(per.size()); ++i) {               \
                terms[i] = per[i];                                                                           \
            }                                                                                                \
        });                                                                                                  \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_energy(void* h, int which, int everything,            \
                                                              int volume_change, int group_index, int all,   \
                                                              int internal, const int* indices,              \
                                                              int n_indices, double* energy)                 \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            const auto c = fb::capi::makeChange(everything, volume_change, group_index, all, internal,       \
                                                indices, n_indices);                                         \
            *energy = fb::capi::pick(*s, which).pot->energy(c);                                              \
        });                                                                                                  \
    }                                                                                                         \
    /* Hamiltonian::force (src/energy.cpp:1162-1166) on a zeroed vector (src/forcemove.cpp:124-125) for      \
       term < 0, else EnergyTerm::force of that term alone; out[3 * n_particles], kT/Angstrom */            \
    __attribute__((visibility("default"))) int P##_sim_forces(void* h, int which, int term, double* out)     \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            auto& st = fb::capi::pick(*s, which);                                                            \
            std::vector<fb::Point> forces(st.spc->particles.size());                                         \
            if (term < 0) {                                                                                  \
                st.pot->force(forces);                                                                       \
            }                                                                                                \
            else {                                                                                           \
                st.pot->terms().at(static_cast<size_t>(term))->force(forces);                                \
            }                                                                                                \
            for (size_t i = 0; i < forces.size(); ++i) {                                                     \
                out[3 * i] = forces[i].x;                                                                    \
                out[3 * i + 1] = forces[i].y;                                                                \
                out[3 * i + 2] = forces[i].z;                                                                \
            }                                                                                                \
        });                                                                                                  \
    }                                                                                                         \
    /* manual trial move: displace atoms of one group in the TRIAL space, then the reference's call          \
       protocol updateState → trial.energy → accepted.energy (montecarlo.cpp:151-155) */                     \
    __attribute__((visibility("default"))) int P##_sim_trial_set(void* h, int group_index, int all,          \
                                                                 int internal, const int* indices,           \
                                                                 int n_indices, const double* xyz,           \
                                                                 double* u_new, double* u_old)               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            auto& spc = *s->mc->trial_state.spc;                                                             \
            auto& g = spc.groups.at(group_index);                                                            \
            for (int i = 0; i < n_indices; ++i) {                                                            \
                auto& p = spc.at(g, indices[i]);                                                             \
                p.pos = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};                                        \
            }                                                                                                \
            if (g.isMolecular()) {                                                                           \
                g.mass_center = spc.massCenter(g, -g.mass_center);                                           \
            }                                                                                                \
            s->pending = fb::capi::makeChange(0, 0, group_index, all, internal, all ? nullptr : indices,     \
                                              all ? 0 : n_indices);                                          \
            s->mc->trial_state.pot->updateState(s->pending);                                                 \
            *u_new = s->mc->trial_state.pot->energy(s->pending);                                             \
            *u_old = s->mc->state.pot->energy(s->pending);                                                   \
        });                                                                                                  \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_trial_commit(void* h, int accept)                     \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            if (accept) {                                                                                    \
                s->mc->state.sync(s->mc->trial_state, s->pending);                                           \
            }                                                                                                \
            else {                                                                                           \
                s->mc->trial_state.sync(s->mc->state, s->pending);                                           \
            }                                                                                                \
        });                                                                                                  \
    }                                                                                                         \
    /* A matter change made by the caller, carried through the MC protocol (MetropolisMonteCarlo::                \
       performExternalChange): `change_json` = {"groups": [{"index": g, "size": new number of active particles,   \
       "atoms": [relative indices], "pos": [[x, y, z] of each listed atom] (optional), "all", "internal", "dNatomic"}]};\
       mode 0 reject, 1 accept,                                                                                   \
       2 Metropolis. out = {u_new, u_old, bias, accepted}. */                                                     \
    __attribute__((visibility("default"))) int P##_sim_matter_change(void* h, const char* change_json, int mode, \
                                                                     double* out)                                \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            const auto j = fb::Json::parse(change_json);                                                     \
            fb::Change change;                                                                               \
            change.matter_change = true;                                                                     \
            for (const auto& g : j.at("groups").items()) {                                                   \
                auto& gc = change.groups.emplace_back();                                                     \
                gc.group_index = static_cast<size_t>(g.at("index").integer());                               \
                gc.all = g.value("all", false);                                                              \
                gc.internal = g.value("internal", false);                                                    \
                gc.dNatomic = g.value("dNatomic", false);                                                    \
                for (const auto& a : g.at("atoms").items()) {                                                \
                    gc.relative_atom_indices.push_back(static_cast<size_t>(a.integer()));                    \
                }                                                                                            \
                auto& group = s->mc->trial_state.spc->groups.at(gc.group_index);                             \
                if (const auto* pos = g.find("pos")) { /* where the listed particles appear */               \
                    size_t k = 0;                                                                            \
                    for (const auto& p : pos->items()) {                                                     \
                        auto& particle = s->mc->trial_state.spc->particles.at(                               \
                            group.begin + gc.relative_atom_indices.at(k++));                                 \
                        particle.pos = fb::pointFromJson(p);                                                 \
                        s->mc->trial_state.spc->geometry.boundary(particle.pos);                             \
                    }                                                                                        \
                }                                                                                            \
                group.resize(static_cast<size_t>(g.at("size").integer()));                                   \
                if (group.isMolecular() && !group.empty()) {                                                 \
                    group.mass_center = s->mc->trial_state.spc->massCenter(                                  \
                        group, -s->mc->trial_state.spc->at(group, 0).pos);                                   \
                }                                                                                            \
            }                                                                                                \
            std::sort(change.groups.begin(), change.groups.end());                                           \
            const auto r = s->mc->performExternalChange(change, mode);                                       \
            out[0] = r.new_energy;                                                                           \
            out[1] = r.old_energy;                                                                           \
            out[2] = r.bias;                                                                                 \
            out[3] = r.accepted ? 1.0 : 0.0;                                                                 \
        });                                                                                                  \
    }                                                                                                         \
    /* per-rank generators of a sharded analysis (every rank draws its OWN ghosts): the reference seeds the ranks  \
       of an MPI run individually (`random: {seed: hardware}`); here the seed is explicit and reproducible */    \
    __attribute__((visibility("default"))) int P##_sim_seed_global(void* h, unsigned seed)                   \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->mc->rng.global.engine.seed(seed); });                              \
    }                                                                                                         \
    __attribute__((visibility("default"))) double P##_sim_drift(void* h)                                     \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        double d = std::nan("");                                                                             \
        fb::capi::guarded([&] { d = s->mc->relativeEnergyDrift(); });                                        \
        return d;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) double P##_sim_initial_energy(void* h)                            \
    {                                                                                                         \
        return static_cast<fb::capi::Sim*>(h)->mc->initial_energy;                                           \
    }                                                                                                         \
    __attribute__((visibility("default"))) double P##_sim_sum_energy_changes(void* h)                        \
    {                                                                                                         \
        return static_cast<fb::capi::Sim*>(h)->mc->sum_of_energy_changes;                                    \
    }                                                                                                         \
    __attribute__((visibility("default"))) void P##_sim_trace_enable(void* h, int on)                        \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        s->mc->record_trace = on != 0;                                                                       \
        s->mc->trace.clear();                                                                                \
    }                                                                                                         \
    __attribute__((visibility("default"))) long P##_sim_trace_size(void* h)                                  \
    {                                                                                                         \
        return static_cast<long>(static_cast<fb::capi::Sim*>(h)->mc->trace.size());                          \
    }                                                                                                         \
    __attribute__((visibility("default"))) long P##_sim_trace_get(void* h, long offset, long n, double* du,  \
                                                                  double* u_new, double* u_old,              \
                                                                  int* accepted, int* move_id)               \
    {                                                                                                         \
        const auto& t = static_cast<fb::capi::Sim*>(h)->mc->trace;                                           \
        long k = 0;                                                                                          \
        for (long i = offset; i < static_cast<long>(t.size()) && k < n; ++i, ++k) {                          \
            if (du) du[k] = t[i].du;                                                                         \
            if (u_new) u_new[k] = t[i].u_new;                                                                \
            if (u_old) u_old[k] = t[i].u_old;                                                                \
            if (accepted) accepted[k] = t[i].accepted;                                                       \
            if (move_id) move_id[k] = t[i].move_id;                                                          \
        }                                                                                                    \
        return k;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_num_particles(void* h)                                \
    {                                                                                                         \
        return static_cast<int>(static_cast<fb::capi::Sim*>(h)->mc->state.spc->particles.size());            \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_num_groups(void* h)                                   \
    {                                                                                                         \
        return static_cast<int>(static_cast<fb::capi::Sim*>(h)->mc->state.spc->groups.size());               \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_get_particles(void* h, int which, double* xyzq,       \
                                                                     int* ids)                               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        const auto& spc = *fb::capi::pick(*s, which).spc;                                                    \
        for (size_t i = 0; i < spc.particles.size(); ++i) {                                                  \
            const auto& p = spc.particles[i];                                                                \
            xyzq[4 * i] = p.pos.x;                                                                           \
            xyzq[4 * i + 1] = p.pos.y;                                                                       \
            xyzq[4 * i + 2] = p.pos.z;                                                                       \
            xyzq[4 * i + 3] = p.charge;                                                                      \
            if (ids) ids[i] = p.id;                                                                          \
        }                                                                                                    \
        return static_cast<int>(spc.particles.size());                                                       \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_get_groups(void* h, int which, int* begin_size_cap_id,\
                                                                  double* cm)                                \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        const auto& spc = *fb::capi::pick(*s, which).spc;                                                    \
        for (size_t i = 0; i < spc.groups.size(); ++i) {                                                     \
            const auto& g = spc.groups[i];                                                                   \
            begin_size_cap_id[4 * i] = static_cast<int>(g.begin);                                            \
            begin_size_cap_id[4 * i + 1] = static_cast<int>(g.size());                                       \
            begin_size_cap_id[4 * i + 2] = static_cast<int>(g.capacity());                                   \
            begin_size_cap_id[4 * i + 3] = g.id;                                                             \
            if (cm) {                                                                                        \
                cm[3 * i] = g.mass_center.x;                                                                 \
                cm[3 * i + 1] = g.mass_center.y;                                                             \
                cm[3 * i + 2] = g.mass_center.z;                                                             \
            }                                                                                                \
        }                                                                                                    \
        return static_cast<int>(spc.groups.size());                                                          \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_state_json(void* h, char* buf, int len)               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::copyOut(s->mc->saveState().dump(), buf, len);                                        \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_sim_info_json(void* h, char* buf, int len)                \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        fb::Json j = fb::Json::object();                                                                     \
        fb::Json je;                                                                                         \
        s->mc->state.pot->to_json(je);                                                                       \
        j["energy"] = je;                                                                                    \
        fb::Json jm = fb::Json::array();                                                                     \
        for (const auto& m : s->mc->moves->all()) {                                                          \
            fb::Json inner = fb::Json::object();                                                             \
            m->to_json(inner);                                                                               \
            fb::Json w = fb::Json::object();                                                                 \
            w[m->name] = inner;                                                                              \
            jm.push_back(w);                                                                                 \
        }                                                                                                    \
        j["moves"] = jm;                                                                                     \
        fb::Json jt = fb::Json::array();                                                                     \
        for (const auto& t : s->mc->state.pot->terms()) {                                                    \
            fb::Json w = fb::Json::object();                                                                 \
            w["name"] = t->name;                                                                             \
            w["seconds"] = t->seconds;                                                                       \
            jt.push_back(w);                                                                                 \
        }                                                                                                    \
        j["term_seconds"] = jt;                                                                              \
        return fb::capi::copyOut(j.dump(), buf, len);                                                        \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_widom_create(void* h, const char* json_text)              \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int id = -1;                                                                                         \
        fb::capi::guarded([&] {                                                                              \
            s->widoms.push_back((WIDOM)(fb::Json::parse(json_text), *s->mc));                                \
            id = static_cast<int>(s->widoms.size()) - 1;                                                     \
        });                                                                                                  \
        return id;                                                                                           \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_widom_sample(void* h, int id, int nsamples)               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            for (int i = 0; i < nsamples; ++i) {                                                             \
                s->widoms.at(id)->sample();                                                                  \
            }                                                                                                \
        });                                                                                                  \
    }                                                                                                         \
    /* one sample event sharded over ranks: prepare (same ghosts on every rank) → evaluate a slice → collect all */ \
    __attribute__((visibility("default"))) int P##_widom_prepare(void* h, int id)                            \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int n = -1;                                                                                          \
        fb::capi::guarded([&] {                                                                              \
            auto& w = *s->widoms.at(id);                                                                     \
            n = w.prepare() ? w.preparedInsertions() : 0;                                                    \
        });                                                                                                  \
        return n;                                                                                            \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_widom_evaluate_slice(void* h, int id, int first,          \
                                                                        int count, double* du)               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->widoms.at(id)->evaluateSlice(first, count, du); });                \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_widom_collect(void* h, int id, const double* du, int n)   \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->widoms.at(id)->collectAll(du, n); });                              \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_widom_result(void* h, int id, double* sum_exp,            \
                                                                long* count, double* last_du, int max_du)    \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        const auto& w = *s->widoms.at(id);                                                                   \
        *sum_exp = w.sum_exp;                                                                                \
        *count = static_cast<long>(w.count);                                                                 \
        const int n = static_cast<int>(w.last_du.size());                                                    \
        for (int i = 0; last_du && i < n && i < max_du; ++i) {                                               \
            last_du[i] = w.last_du[i];                                                                       \
        }                                                                                                    \
        return n;                                                                                            \
    }                                                                                                         \
    }

/** `<P>_virtualvolume_*`: the virtual volume move analysis (host code only; energies from the Hamiltonian of `h`) */
#define FB_DEFINE_VIRTUALVOLUME_CAPI(P)                                                                       \
    extern "C" {                                                                                              \
    __attribute__((visibility("default"))) int P##_virtualvolume_create(void* h, const char* json_text)      \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int id = -1;                                                                                         \
        fb::capi::guarded([&] {                                                                              \
            s->virtual_volumes.push_back(std::make_unique<fb::VirtualVolumeMove>(                            \
                fb::Json::parse(json_text), *s->mc->state.spc, *s->mc->state.pot));                          \
            id = static_cast<int>(s->virtual_volumes.size()) - 1;                                            \
        });                                                                                                  \
        return id;                                                                                           \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_virtualvolume_sample(void* h, int id)                     \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->virtual_volumes.at(id)->sample(); });                              \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_virtualtranslate_create(void* h, const char* json_text)   \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int id = -1;                                                                                         \
        fb::capi::guarded([&] {                                                                              \
            s->virtual_translates.push_back(std::make_unique<fb::VirtualTranslate>(                          \
                fb::Json::parse(json_text), *s->mc->state.spc, *s->mc->state.pot, s->mc->rng.global));       \
            id = static_cast<int>(s->virtual_translates.size()) - 1;                                         \
        });                                                                                                  \
        return id;                                                                                           \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_virtualtranslate_sample(void* h, int id)                  \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->virtual_translates.at(id)->sample(); });                           \
    }                                                                                                         \
    /* out[0] = Σ exp(−ΔU), out[1] = samples, out[2] = last ΔU, out[3] = mean force / kT Å⁻¹ */              \
    __attribute__((visibility("default"))) int P##_virtualtranslate_result(void* h, int id, double out[4])   \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            const auto& v = *s->virtual_translates.at(id);                                                   \
            out[0] = v.sum_exp;                                                                              \
            out[1] = static_cast<double>(v.count);                                                           \
            out[2] = v.last_energy_change;                                                                   \
            out[3] = v.count > 0 ? v.meanForce() : 0.0;                                                      \
        });                                                                                                  \
    }                                                                                                         \
    /* out[0] = Σ exp(−ΔU), out[1] = samples, out[2] = last ΔU, out[3] = excess pressure / kT Å⁻³ */         \
    __attribute__((visibility("default"))) int P##_virtualvolume_result(void* h, int id, double out[4])      \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] {                                                                       \
            const auto& v = *s->virtual_volumes.at(id);                                                      \
            out[0] = v.sum_exp;                                                                              \
            out[1] = static_cast<double>(v.count);                                                           \
            out[2] = v.last_energy_change;                                                                   \
            out[3] = v.count > 0 ? v.excessPressure() : 0.0;                                                 \
        });                                                                                                  \
    }                                                                                                         \
    }

/**
 * `<P>_rdf_*`: the atomic radial distribution function analysis. RDF: callable (const fb::Json&, fb::capi::Sim&) →
 * std::unique_ptr<fb::AtomRDF> (the CPU pair loop in the oracle build, the device histogram in the B200 build).
 */
#define FB_DEFINE_RDF_CAPI(P, RDF)                                                                            \
    extern "C" {                                                                                              \
    __attribute__((visibility("default"))) int P##_rdf_create(void* h, const char* json_text)                \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int id = -1;                                                                                         \
        fb::capi::guarded([&] {                                                                              \
            s->rdfs.push_back((RDF)(fb::Json::parse(json_text), *s));                                        \
            id = static_cast<int>(s->rdfs.size()) - 1;                                                       \
        });                                                                                                  \
        return id;                                                                                           \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_rdf_sample(void* h, int id)                               \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->rdfs.at(id)->sample(); });                                         \
    }                                                                                                         \
    __attribute__((visibility("default"))) int P##_rdf_sample_shard(void* h, int id, int shard, int n_shards) \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        return fb::capi::guarded([&] { s->rdfs.at(id)->sampleShard(shard, n_shards); });                     \
    }                                                                                                         \
    /* returns the number of bins; fills r[i], pairs[i] (exact counts), g[i] for i < max */                  \
    __attribute__((visibility("default"))) int P##_rdf_result(void* h, int id, double* r,                    \
                                                              unsigned long long* pairs, double* g, int max) \
    {                                                                                                         \
        auto* s = static_cast<fb::capi::Sim*>(h);                                                            \
        int n = -1;                                                                                          \
        fb::capi::guarded([&] {                                                                              \
            const auto& f = *s->rdfs.at(id);                                                                 \
            n = static_cast<int>(f.size());                                                                  \
            const double total = f.total();                                                                  \
            for (int i = 0; i < n && i < max; ++i) {                                                         \
                if (r) {                                                                                     \
                    r[i] = f.distance(i);                                                                    \
                }                                                                                            \
                if (pairs) {                                                                                 \
                    pairs[i] = f.pairs(i);                                                                   \
                }                                                                                            \
                if (g) {                                                                                     \
                    g[i] = f.g(i, total);                                                                    \
                }                                                                                            \
            }                                                                                                \
        });                                                                                                  \
        return n;                                                                                            \
    }                                                                                                         \
    }

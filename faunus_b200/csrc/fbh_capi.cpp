// Host-level entry points of libfaunus_b200.so (`fbh_*`): the MC driver and Widom analysis running on
// the B200 adaptor terms. See host/sim_capi.hpp for the ABI and b200_terms.hpp for the adaptors.
#include "b200_terms.hpp"
#include "host/sim_capi.hpp"

namespace {
const fb::TermFactory b200_factory = fb::b200TermFactory;

std::unique_ptr<fb::WidomInsertion> makeWidom(const fb::Json& j, fb::MetropolisMonteCarlo& mc)
{
    if (j.value("batched", true)) {
        return std::make_unique<fb::WidomB200>(j, mc);
    }
    return fb::capi::defaultWidom(j, mc); // sequential reference order through energy(change)
}
} // namespace

/** in-process replicas: replica r runs on GPU r modulo the visible device count */
static void fbh_replica_setup(int rank)
{
    const int n = fb_device_count();
    fb::defaultDevice() = n > 0 ? rank % n : 0;
}

FB_DEFINE_SIM_CAPI(fbh, b200_factory, makeWidom)
FB_DEFINE_VIRTUALVOLUME_CAPI(fbh)
FB_DEFINE_RDF_CAPI(fbh, [](const fb::Json& j, fb::capi::Sim& s) -> std::unique_ptr<fb::AtomRDF> {
    return std::make_unique<fb::AtomRDFB200>(j, *s.mc);
})

/**
 * One replica of a parallel-tempering run whose Temper move talks over the device library's NCCL communicator
 * (fb_nccl_*): `unique_id` = the 128 bytes of fb_nccl_unique_id made by one rank and handed to all by the launcher.
 */
extern "C" __attribute__((visibility("default"))) void* fbh_sim_create_replica_nccl(const char* json_text,
                                                                                    const char* unique_id, int rank,
                                                                                    int size)
{
    std::unique_ptr<fb::NcclReplicaComm> comm;
    if (fb::capi::guarded([&] { comm = std::make_unique<fb::NcclReplicaComm>(unique_id, rank, size); }) != 0) {
        return nullptr;
    }
    fb::NcclReplicaComm* raw = comm.get();
    void* handle = fb::capi::create(json_text, b200_factory, std::move(comm));
    if (handle != nullptr) {
        auto* s = static_cast<fb::capi::Sim*>(handle);
        const int rc = fb::capi::guarded([&] {
            const auto terms = s->mc->trial_state.pot->find<fb::NonbondedB200>();
            if (terms.size() != 1) {
                throw std::runtime_error("tempering over NCCL needs exactly one B200 non-bonded term");
            }
            raw->bind(terms.front()->device(), terms.front()->deviceSlot());
        });
        if (rc != 0) {
            delete s;
            return nullptr;
        }
    }
    return handle;
}

/**
 * In-process tempering (one thread and one device context per replica) with the state exchange on the packed
 * device mirrors: fb_export_state → in-process hand-over → fb_import_state, the Spaces follow. The same host logic as
 * the NCCL communicator, testable on one GPU. Result format as fbh_temper_run_local.
 */
extern "C" __attribute__((visibility("default"))) int fbh_temper_run_local_packed(const char* configs_json, int sweeps,
                                                                                  char* buf, int len)
{
    int n = -1;
    fb::capi::guarded([&] {
        const auto configs = fb::Json::parse(configs_json);
        const auto results = fb::capi::runLocalReplicas(
            configs, sweeps, b200_factory, fbh_replica_setup, [](fb::MetropolisMonteCarlo& mc, fb::LocalComm& comm) {
                const auto terms = mc.trial_state.pot->find<fb::NonbondedB200>();
                if (terms.size() != 1) {
                    throw std::runtime_error("packed exchange needs exactly one B200 non-bonded term");
                }
                auto dev = terms.front()->device();
                const int slot = terms.front()->deviceSlot();
                comm.state_exchanger = [dev, slot, &comm](fb::Space& spc, int partner, fb::VolumeMethod method,
                                                          fb::Change& change) {
                    return fb::exchangePackedStateThroughHost(*dev, slot, comm, spc, partner, method, change);
                };
            });
        fb::Json out = fb::Json::array();
        for (const auto& r : results) {
            fb::Json j = fb::Json::object();
            j["energy"] = r.energy;
            j["drift"] = r.drift;
            j["xyzq"] = fb::Json::fromVector(r.xyzq);
            j["moves"] = r.info.empty() ? fb::Json() : fb::Json::parse(r.info);
            j["error"] = r.error;
            out.push_back(j);
        }
        n = fb::capi::copyOut(out.dump(), buf, len);
    });
    return n;
}

/** messages and bytes the replica has exchanged over NCCL so far */
extern "C" __attribute__((visibility("default"))) int fbh_sim_exchange_stats(void* h, double out[2])
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    out[0] = out[1] = 0;
    if (auto* comm = dynamic_cast<fb::NcclReplicaComm*>(s->comm.get())) {
        out[0] = static_cast<double>(comm->exchanges);
        const auto terms = s->mc->state.pot->find<fb::NonbondedB200>();
        if (!terms.empty()) {
            out[1] = static_cast<double>(fb_nccl_bytes_exchanged(terms.front()->device()->ctx));
        }
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) void fbh_set_device(int device)
{
    fb::defaultDevice() = device;
}

/** kernels launched so far by the non-bonded/Ewald context of a simulation */
extern "C" __attribute__((visibility("default"))) unsigned long long fbh_sim_launch_count(void* h)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    unsigned long long n = 0;
    for (const auto& t : s->mc->state.pot->find<fb::NonbondedB200>()) {
        n += fb_launch_count(t->device()->ctx);
    }
    return n;
}

/** S(q) table of a `coulomb` block as built by the product's host-side generator (tests compare it
 * bit for bit with the oracle's) */
extern "C" __attribute__((visibility("default"))) int fbh_coulomb_table(const char* coulomb_json, double temperature,
                                                                        double* knots, double* coeffs, int max_knots,
                                                                        double* lB, double* cutoff, double* kappa,
                                                                        double* self_prefactor)
{
    int n = -1;
    fb::capi::guarded([&] {
        fb::pc::temperature = temperature;
        const auto t = fb::makeCoulombTable(fb::Json::parse(coulomb_json));
        n = static_cast<int>(t.S.knots.size());
        for (int i = 0; i < n && i < max_knots; ++i) {
            knots[i] = t.S.knots[i];
        }
        for (int i = 0; i < 6 * (n - 1) && i < 6 * (max_knots - 1); ++i) {
            coeffs[i] = t.S.coeffs[i];
        }
        *lB = t.bjerrum_length;
        *cutoff = t.cutoff;
        *kappa = t.kappa;
        *self_prefactor = t.self_prefactor;
    });
    return n;
}

/** the same for S'(q), the table behind fb_nonbonded_force */
extern "C" __attribute__((visibility("default"))) int fbh_coulomb_force_table(const char* coulomb_json, double temperature,
                                                                              double* knots, double* coeffs, int max_knots)
{
    int n = -1;
    fb::capi::guarded([&] {
        fb::pc::temperature = temperature;
        const auto t = fb::makeCoulombTable(fb::Json::parse(coulomb_json));
        n = static_cast<int>(t.dS.knots.size());
        for (int i = 0; i < n && i < max_knots; ++i) {
            knots[i] = t.dS.knots[i];
        }
        for (int i = 0; i < 6 * (n - 1) && i < 6 * (max_knots - 1); ++i) {
            coeffs[i] = t.dS.coeffs[i];
        }
    });
    return n;
}

/** Host-side pair tables of one `energy` entry as JSON (mixing matrices, flags, spline ranges) */
extern "C" __attribute__((visibility("default"))) int fbh_pair_tables_json(const char* input_json,
                                                                           const char* nonbonded_name, char* buf,
                                                                           int len)
{
    int n = -1;
    fb::capi::guarded([&] {
        const auto j = fb::Json::parse(input_json);
        fb::pc::temperature = j.at("temperature").number();
        const auto topo = fb::topologyFromJson(j);
        const fb::Json* cfg = nullptr;
        for (const auto& e : j.at("energy").items()) {
            if (e.single().first == nonbonded_name) {
                cfg = &e.single().second;
            }
        }
        if (!cfg) {
            throw std::runtime_error("energy entry not found");
        }
        const auto t = fb::buildPairTables(nonbonded_name, *cfg, *topo);
        fb::Json out = fb::Json::object();
        out["kind"] = t.kind;
        out["n_types"] = t.n_types;
        std::vector<double> flags(t.flags.begin(), t.flags.end());
        out["flags"] = fb::Json::fromVector(flags);
        out["lj_s2"] = fb::Json::fromVector(t.lj_s2);
        out["lj_e4"] = fb::Json::fromVector(t.lj_e4);
        out["wca_s2"] = fb::Json::fromVector(t.wca_s2);
        out["wca_e4"] = fb::Json::fromVector(t.wca_e4);
        out["hs_s2"] = fb::Json::fromVector(t.hs_s2);
        out["plain_lB"] = t.plain_bjerrum_length;
        out["g2g_cutoff_squared"] = fb::Json::fromVector(t.g2g_cutoff_squared);
        out["sp_rmin2"] = fb::Json::fromVector(t.sp_rmin2);
        out["sp_rmax2"] = fb::Json::fromVector(t.sp_rmax2);
        out["sp_knots"] = fb::Json::fromVector(t.sp_knots);
        out["sp_coeffs"] = fb::Json::fromVector(t.sp_coeffs);
        std::vector<double> off(t.sp_offset.begin(), t.sp_offset.end());
        out["sp_offset"] = fb::Json::fromVector(off);
        n = fb::capi::copyOut(out.dump(), buf, len);
    });
    return n;
}

/** toggle CUDA-event timing of the hot kernels; read the accumulators (see fb_get_timing) */
extern "C" __attribute__((visibility("default"))) int fbh_sim_enable_timing(void* h, int on)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    for (const auto& t : s->mc->state.pot->find<fb::NonbondedB200>()) {
        fb_enable_timing(t->device()->ctx, on);
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) int fbh_sim_get_timing(void* h, double out[8])
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    for (int i = 0; i < 8; ++i) {
        out[i] = 0;
    }
    for (const auto& t : s->mc->state.pot->find<fb::NonbondedB200>()) {
        double v[8];
        fb_get_timing(t->device()->ctx, v);
        for (int i = 0; i < 8; ++i) {
            out[i] += v[i];
        }
    }
    return 0;
}

/** the fb_ctx of the (single) B200 non-bonded term of a simulation, for direct C-ABI calls (benchmarks, tests) */
extern "C" __attribute__((visibility("default"))) void* fbh_sim_ctx(void* h)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    const auto terms = s->mc->state.pot->find<fb::NonbondedB200>();
    return terms.empty() ? nullptr : static_cast<void*>(terms.front()->device()->ctx);
}

/**
 * Windowed evaluation of `transrot` runs (fb_batch_trial): capacity 0 switches it off (one move at a
 * time through updateState/energy/sync). Returns the capacity in effect (0 if the Hamiltonian is not
 * eligible: only self-energy, B200 non-bonded and tinfoil PBC Ewald terms can be windowed).
 */
extern "C" __attribute__((visibility("default"))) int fbh_sim_set_window(void* h, int capacity)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    int result = 0;
    fb::capi::guarded([&] {
        s->mc->window_evaluator = fb::B200WindowEvaluator::tryCreate(*s->mc, capacity);
        result = s->mc->window_evaluator ? s->mc->window_evaluator->capacity() : 0;
    });
    return result;
}

/**
 * Runs: the device walks the windows of single-atom moves itself (fb_run_submit), up to `moves` proposals per
 * host round trip (0: the host walks every window), used when at least `min_moves` proposals are ready (default 65:
 * a single window is walked on the host). Needs fbh_sim_set_window(64). Returns the run capacity in effect.
 */
extern "C" __attribute__((visibility("default"))) int fbh_sim_set_run(void* h, int moves, int min_moves)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    int result = 0;
    fb::capi::guarded([&] {
        auto* e = dynamic_cast<fb::B200WindowEvaluator*>(s->mc->window_evaluator.get());
        if (e != nullptr) {
            e->enableRuns(moves, min_moves);
            result = e->runCapacity();
        }
    });
    return result;
}

/** out[0..5] as fb_get_batch_timing; out[6] = host ms inside evaluate (launch + wait), out[7] = host ms of the
 * whole windowed sweep part (drawing proposals, evaluate, deciding); out[8] = host round trips, out[9] = runs */
extern "C" __attribute__((visibility("default"))) int fbh_sim_get_window_timing(void* h, double out[10])
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    for (int i = 0; i < 10; ++i) {
        out[i] = 0;
    }
    for (const auto& t : s->mc->state.pot->find<fb::NonbondedB200>()) {
        double v[8];
        fb_get_batch_timing(t->device()->ctx, v);
        for (int i = 0; i < 8; ++i) {
            out[i] += v[i];
        }
        out[8] += v[6];
        out[9] += v[7];
    }
    out[6] = 1e3 * s->mc->window_seconds_evaluate;
    out[7] = 1e3 * s->mc->window_seconds_total;
    return 0;
}

/**
 * This rank's share of the full-system non-bonded and reciprocal Ewald energy of the accepted state
 * (tile rows / k-vector slab `shard` of `n_shards`); the shares of all ranks add up to the terms
 * `systemEnergy` reports. out[0] = non-bonded, out[1] = reciprocal (0 without Ewald).
 */
extern "C" __attribute__((visibility("default"))) int fbh_system_energy_shard(void* h, int shard, int n_shards,
                                                                              double out[2])
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    return fb::capi::guarded([&] {
        const auto terms = s->mc->state.pot->find<fb::NonbondedB200>();
        if (terms.size() != 1) {
            throw std::runtime_error("exactly one B200 non-bonded term expected");
        }
        auto& t = *terms.front();
        t.device()->resynchronise();
        fb::fbCheck(fb_system_energy_shard(t.device()->ctx, t.deviceSlot(), shard, n_shards, &out[0], &out[1]),
                    t.device()->ctx, "fb_system_energy_shard");
    });
}

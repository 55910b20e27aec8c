// ORACLE — test infrastructure, not product code. Builds liboracle: the CPU restatement of the
// reference's energy path behind the same MC driver the product uses, exported as `fo_*`.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline / reference arm may load it.
//
// Term construction order follows the reference: for a `nonbonded*` entry the pair potential's
// self energy is pushed BEFORE the nonbonded term (src/energy.h:462-477) and an Ewald reciprocal
// term is appended when coulomb.type == "ewald" (src/energy.cpp:1134-1160, 1185-1198); name → class
// map: src/energy.cpp:1280-1327.
#include "../faunus_b200/csrc/host/sim_capi.hpp"
#include "ewald.hpp"
#include "nonbonded.hpp"
#include "rdf.hpp"
#include "window_shadow.hpp"

namespace oracle {

template <class TPot> bool addNonbonded(Hamiltonian& h, Space& spc, const Json& cfg)
{
    auto term = std::make_shared<Nonbonded<TPot>>(cfg, spc);
    if (auto self = term->pairPotential().selfEnergy()) {
        h.push_back(std::make_shared<ParticleSelfEnergy>(spc, self));
    }
    h.push_back(term);
    return true;
}

/** src/energy.cpp:1134-1160 */
inline void addEwald(Hamiltonian& h, Space& spc, const Json& j)
{
    const Json* coulomb = nullptr;
    if (const auto* def = j.find("default")) {
        for (const auto& i : def->items()) {
            if (const auto* c = i.find("coulomb")) {
                coulomb = c;
                break;
            }
        }
    }
    else if (const auto* c = j.find("coulomb")) {
        coulomb = c;
    }
    if (coulomb && coulomb->value("type", "") == "ewald") {
        h.push_back(std::make_shared<Ewald>(*coulomb, spc));
    }
}

inline bool termFactory(Hamiltonian& h, Space& spc, const std::string& name, const Json& cfg)
{
    bool ok = false;
    if (name == "nonbonded_coulomblj" || name == "nonbonded_newcoulomblj") {
        ok = addNonbonded<CoulombLJ>(h, spc, cfg);
    }
    else if (name == "nonbonded_coulombwca") {
        ok = addNonbonded<CoulombWCA>(h, spc, cfg);
    }
    else if (name == "nonbonded_pm" || name == "nonbonded_coulombhs") {
        ok = addNonbonded<PrimitiveModel>(h, spc, cfg);
    }
    else if (name == "nonbonded_pmwca") {
        ok = addNonbonded<PrimitiveModelWCA>(h, spc, cfg);
    }
    else if (name == "nonbonded" || name == "nonbonded_exact") {
        ok = addNonbonded<FunctorPotential>(h, spc, cfg);
    }
    else if (name == "nonbonded_splined" || name == "nonbonded_cached") {
        // NonbondedCached (src/energy.h:1614-1758) returns the energies of Nonbonded<SplinedPotential>; its
        // group-group cache changes how often pairs are summed, not what the sums are
        ok = addNonbonded<SplinedPotential>(h, spc, cfg);
    }
    if (ok) {
        addEwald(h, spc, cfg);
    }
    return ok;
}

static const TermFactory factory = termFactory;

} // namespace oracle

static void fo_replica_setup(int) {}

FB_DEFINE_SIM_CAPI(fo, oracle::factory, fb::capi::defaultWidom)
FB_DEFINE_VIRTUALVOLUME_CAPI(fo)
FB_DEFINE_RDF_CAPI(fo, [](const fb::Json& j, fb::capi::Sim& s) -> std::unique_ptr<fb::AtomRDF> {
    return std::make_unique<oracle::AtomRDFCpu>(j, *s.mc->state.spc);
})

#define FO_API extern "C" __attribute__((visibility("default")))

/**
 * Test hook: run the windowed engine of `h` against a CPU stand-in for the device (window_shadow.hpp) built from the
 * same input `json_text`; `moves` proposals per evaluation. Only for inputs whose moves are all `transrot`.
 */
FO_API int fo_sim_set_shadow_window(void* h, const char* json_text, int moves)
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    return fb::capi::guarded([&] {
        auto shadow = std::make_unique<fb::MetropolisMonteCarlo>(fb::Json::parse(json_text), oracle::factory, nullptr);
        s->mc->window_evaluator = std::make_unique<oracle::ShadowWindowEvaluator>(*s->mc, std::move(shadow), moves);
    });
}

/** out[0] = conditional proposals shipped, out[1] = evaluations queued behind another one, out[2] = evaluations */
FO_API int fo_sim_shadow_stats(void* h, double out[3])
{
    auto* s = static_cast<fb::capi::Sim*>(h);
    auto* e = dynamic_cast<oracle::ShadowWindowEvaluator*>(s->mc->window_evaluator.get());
    if (e == nullptr) {
        return -1;
    }
    out[0] = static_cast<double>(e->conditional_proposals);
    out[1] = static_cast<double>(e->queued_evaluations);
    out[2] = static_cast<double>(e->evaluations);
    return 0;
}

/** Andrea spline of a named test function; pins src/tabulate.h:313-365 */
FO_API int fo_andrea_test(double utol, double ftol, double xmin, double xmax, double* knots, int max_knots,
                          double* coeffs, int max_coeffs, int* n_coeffs)
{
    oracle::Andrea spline;
    spline.setTolerance(utol, ftol);
    const auto d = spline.generate([](double x) { return 0.5 * x * std::sin(x) + 2; }, xmin, xmax);
    for (size_t i = 0; i < d.r2.size() && static_cast<int>(i) < max_knots; ++i) {
        knots[i] = d.r2[i];
    }
    for (size_t i = 0; i < d.c.size() && static_cast<int>(i) < max_coeffs; ++i) {
        coeffs[i] = d.c[i];
    }
    *n_coeffs = static_cast<int>(d.c.size());
    return static_cast<int>(d.r2.size());
}

FO_API double fo_andrea_test_eval(double utol, double ftol, double xmin, double xmax, double x)
{
    oracle::Andrea spline;
    spline.setTolerance(utol, ftol);
    const auto d = spline.generate([](double t) { return 0.5 * t * std::sin(t) + 2; }, xmin, xmax);
    return oracle::Andrea::eval(d, x);
}

/** Splined Coulomb S(q) table for a `coulomb` JSON block at temperature T; returns #knots */
FO_API int fo_coulomb_table(const char* coulomb_json, double temperature, double* knots, double* coeffs,
                            int max_knots, double* lB, double* cutoff, double* kappa, double* self_prefactor)
{
    int n = -1;
    fb::capi::guarded([&] {
        fb::pc::temperature = temperature;
        oracle::NewCoulombGalore pot;
        fb::Topology topo;
        pot.from_json(fb::Json::parse(coulomb_json), topo);
        n = static_cast<int>(pot.table.r2.size());
        for (int i = 0; i < n && i < max_knots; ++i) {
            knots[i] = pot.table.r2[i];
        }
        for (int i = 0; i < 6 * (n - 1) && i < 6 * (max_knots - 1); ++i) {
            coeffs[i] = pot.table.c[i];
        }
        *lB = pot.bjerrum_length;
        *cutoff = pot.scheme.cutoff;
        *kappa = pot.scheme.kappa;
        *self_prefactor = pot.scheme.self_prefactor;
    });
    return n;
}

/** the same for S'(q), the table behind the pair force; returns #knots */
FO_API int fo_coulomb_force_table(const char* coulomb_json, double temperature, double* knots, double* coeffs, int max_knots)
{
    int n = -1;
    fb::capi::guarded([&] {
        fb::pc::temperature = temperature;
        oracle::NewCoulombGalore pot;
        fb::Topology topo;
        pot.from_json(fb::Json::parse(coulomb_json), topo);
        n = static_cast<int>(pot.slope_table.r2.size());
        for (int i = 0; i < n && i < max_knots; ++i) {
            knots[i] = pot.slope_table.r2[i];
        }
        for (int i = 0; i < 6 * (n - 1) && i < 6 * (max_knots - 1); ++i) {
            coeffs[i] = pot.slope_table.c[i];
        }
    });
    return n;
}

/** Mass-centre cutoffs² [n_mol²] parsed from a non-bonded block (`cutoff_g2g`); src/energy.cpp:1897-1933, doctest :1955-2002 */
FO_API int fo_group_cutoffs(const char* input_json, const char* nonbonded_json, double* cutoff_squared, int max_values)
{
    int n = -1;
    fb::capi::guarded([&] {
        const auto j = fb::Json::parse(input_json);
        auto topo = fb::topologyFromJson(j);
        fb::Geometry geometry = fb::Geometry::fromJson(fb::Json::parse(R"({"type":"cuboid","length":100})"));
        oracle::GroupCutoff cutoff(geometry);
        cutoff.from_json(fb::Json::parse(nonbonded_json), *topo);
        const int n_mol = static_cast<int>(topo->molecules.size());
        n = n_mol * n_mol;
        for (int a = 0; a < n_mol; ++a) {
            for (int b = 0; b < n_mol; ++b) {
                if (a * n_mol + b < max_values) {
                    cutoff_squared[a * n_mol + b] = cutoff.cutoffSquared(a, b);
                }
            }
        }
    });
    return n;
}

/** Pair energy u(a,b,r) of an `energy`-entry style potential for two atom types (functor tests) */
FO_API int fo_pair_energy(const char* input_json, const char* nonbonded_name, int id_a, int id_b, const double* r,
                          int n, double* u)
{
    return fb::capi::guarded([&] {
        const auto j = fb::Json::parse(input_json);
        fb::pc::temperature = j.at("temperature").number();
        auto topo = fb::topologyFromJson(j);
        fb::Space spc;
        spc.topology = topo;
        spc.geometry = fb::Geometry::fromJson(fb::Json::parse(R"({"type":"cuboid","length":1e9})"));
        fb::Particle a = topo->makeParticle(id_a);
        fb::Particle b = topo->makeParticle(id_b);
        const fb::Json* cfg = nullptr;
        for (const auto& e : j.at("energy").items()) {
            if (e.single().first == nonbonded_name) {
                cfg = &e.single().second;
            }
        }
        if (!cfg) {
            throw std::runtime_error("energy entry not found");
        }
        const std::string name = nonbonded_name;
        auto run = [&](auto pot) {
            pot.from_json(*cfg, *topo);
            for (int i = 0; i < n; ++i) {
                u[i] = pot(a, b, r[i] * r[i]);
            }
        };
        if (name == "nonbonded_coulomblj") {
            run(oracle::CoulombLJ());
        }
        else if (name == "nonbonded_coulombwca") {
            run(oracle::CoulombWCA());
        }
        else if (name == "nonbonded_pm") {
            run(oracle::PrimitiveModel());
        }
        else if (name == "nonbonded_pmwca") {
            run(oracle::PrimitiveModelWCA());
        }
        else if (name == "nonbonded") {
            run(oracle::FunctorPotential());
        }
        else if (name == "nonbonded_splined" || name == "nonbonded_cached") {
            run(oracle::SplinedPotential());
        }
        else {
            throw std::runtime_error("unknown nonbonded name");
        }
    });
}

/**
 * Pair force on a due to b (kT/Å) for n distance vectors b → a (xyz[3n]) of an `energy`-entry style potential; the
 * known-answer values of src/potentials.cpp:399, 661, 777-778, 1612-1626 go through here. Returns −1 with the
 * reference's message for a potential without forces (PairPotential::force, src/potentials.cpp:246-251).
 */
FO_API int fo_pair_force(const char* input_json, const char* nonbonded_name, int id_a, int id_b, const double* xyz, int n,
                         double* f)
{
    return fb::capi::guarded([&] {
        const auto j = fb::Json::parse(input_json);
        fb::pc::temperature = j.at("temperature").number();
        auto topo = fb::topologyFromJson(j);
        fb::Particle a = topo->makeParticle(id_a);
        fb::Particle b = topo->makeParticle(id_b);
        const fb::Json* cfg = nullptr;
        for (const auto& e : j.at("energy").items()) {
            if (e.single().first == nonbonded_name) {
                cfg = &e.single().second;
            }
        }
        if (!cfg) {
            throw std::runtime_error("energy entry not found");
        }
        const std::string name = nonbonded_name;
        auto run = [&](auto pot) {
            pot.from_json(*cfg, *topo);
            for (int i = 0; i < n; ++i) {
                const fb::Point r(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
                const fb::Point force = pot.force(a, b, r.squaredNorm(), r);
                f[3 * i] = force.x;
                f[3 * i + 1] = force.y;
                f[3 * i + 2] = force.z;
            }
        };
        if (name == "nonbonded_coulomblj") {
            run(oracle::CoulombLJ());
        }
        else if (name == "nonbonded_coulombwca") {
            run(oracle::CoulombWCA());
        }
        else if (name == "nonbonded_pm") {
            run(oracle::PrimitiveModel());
        }
        else if (name == "nonbonded_pmwca") {
            run(oracle::PrimitiveModelWCA());
        }
        else {
            throw std::logic_error("Force computation not implemented for this setup!");
        }
    });
}

/**
 * Known-answer harness for the Ewald policies (src/energy.cpp:74-99, 249-305): particles xyzq in a
 * cubic box, returns K and {self, surface, reciprocal} energies.
 */
FO_API int fo_ewald_kat(const char* ewald_json, double temperature, double box, const double* xyzq, int n,
                        double* self_surface_reciprocal, double* lB)
{
    int K = -1;
    fb::capi::guarded([&] {
        fb::pc::temperature = temperature;
        auto topo = std::make_shared<fb::Topology>();
        fb::AtomData atom;
        atom.name = "A";
        atom.id = 0;
        topo->atoms.push_back(atom);
        fb::MoleculeData mol;
        mol.name = "M";
        mol.id = 0;
        mol.atomic = true;
        mol.atoms = {0};
        topo->molecules.push_back(mol);
        fb::Space spc;
        spc.topology = topo;
        spc.geometry = fb::Geometry::fromJson(
            fb::Json::parse("{\"type\":\"cuboid\",\"length\":" + std::to_string(box) + "}"));
        fb::ParticleVector pv;
        for (int i = 0; i < n; ++i) {
            fb::Particle p;
            p.id = 0;
            p.pos = {xyzq[4 * i], xyzq[4 * i + 1], xyzq[4 * i + 2]};
            p.charge = xyzq[4 * i + 3];
            pv.push_back(p);
        }
        spc.addGroup(0, pv);
        oracle::EwaldData data(fb::Json::parse(ewald_json));
        oracle::ewald::updateBox(data, spc.geometry.getLength());
        oracle::ewald::updateComplex(data, spc);
        fb::Change c;
        c.everything = true;
        self_surface_reciprocal[0] = oracle::ewald::selfEnergy(data, c, spc);
        self_surface_reciprocal[1] = oracle::ewald::surfaceEnergy(data, c, spc);
        self_surface_reciprocal[2] = oracle::ewald::reciprocalEnergy(data);
        *lB = data.bjerrum_length;
        K = static_cast<int>(data.k_vectors.size());
    });
    return K;
}

FO_API void fo_set_parallel_ewald_init(int on)
{
    oracle::ewald::parallel_full_update = on != 0;
}

/**
 * Large synthetic systems (N = 1e5): full Q(k) updates and the Σ_{i<j} of one big atomic group over all host
 * threads, and repeated evaluations of an unchanged configuration answered from the stored result. The per-move
 * path (partial update, group2all / groupInternal(index), reciprocal sum) is untouched. Mode 2 (N = 1e6) also
 * drops the all-pairs sum of a big atomic group: system energies and the drift are then meaningless, per-move
 * energies and decisions are not affected.
 */
FO_API void fo_set_large_system_mode(int on)
{
    oracle::ewald::parallel_full_update = on != 0;
    oracle::ewald::memoize_full_update = on != 0;
    oracle::parallel_group_internal = on != 0;
    oracle::skip_group_internal = on == 2; // N = 1e6: no Σ_{i<j} over the start configuration at all
    if (!on) {
        oracle::ewald::full_update_memo = {};
        oracle::group_internal_memo = {};
    }
}

/** torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline sets its thread count explicitly */
FO_API void fo_set_openmp_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) {
        omp_set_num_threads(n);
    }
#else
    (void)n;
#endif
}

FO_API int fo_openmp_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

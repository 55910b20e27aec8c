/*
 * faunus_b200.h — C ABI of the B200 (sm_100a) energy-evaluation library, libfaunus_b200.so.
 *
 * These are the entry points a Faunus-side `Energy::EnergyTerm` adaptor binds in place of the
 * reference's CPU loops (INTEGRATION.md shows the adaptor). Plain pointers and sizes only, no C++
 * or torch types. All functions return 0 on success or a negative fb_status; the message is
 * available from fb_last_error(). Nothing throws across this boundary. A context owns all of its
 * device memory and one CUDA stream; callers own the host buffers they pass. A context is not
 * thread-safe (the reference drives everything from one host thread).
 *
 * State model: a context holds TWO device-resident structure-of-arrays mirrors of `Space`
 * ("slots"), matching the reference's accepted and trial `State` (src/montecarlo.h:52-64).
 * Slot 0 mirrors the accepted Space, slot 1 the trial Space. Only changed particles cross PCIe.
 *
 * Reference interface each entry point replaces (file:line in mlund/faunus):
 *   fb_create / fb_destroy ......... Nonbonded<>::Nonbonded + PairEnergy::from_json  src/energy.h:1525-1534, 470-477
 *                                    (pair-potential tables: src/potentials.cpp:672-703, 956-977, 1628-1699)
 *   fb_upload_space ................ EnergyTerm::init() with Space contents           src/externalpotential.h:38, src/space.h:92-373
 *   fb_update_group ................ Space::updateParticles / trial Space mutation    src/space.h:193-217, src/move.cpp:225-240
 *   fb_set_box ..................... Space::scaleVolume → Chameleon::setVolume        src/space.cpp:249-302
 *   fb_sync ........................ Space::sync + EnergyTerm::sync                   src/space.cpp:199-240, src/energy.cpp:1254-1268
 *   fb_nonbonded_energy ............ Nonbonded::energy(Change) → GroupPairing::accumulate  src/energy.h:1561-1577, 1447-1475
 *   fb_nonbonded_delta ............. the pair of calls trial.energy / accepted.energy  src/montecarlo.cpp:154-155
 *   fb_ewald_* ..................... Energy::Ewald + PolicyIonIon                      src/energy.cpp:28-59, 133-247, 466-531, 539-658
 *   fb_batch_trial / fb_batch_commit a window of consecutive MetropolisMonteCarlo::performMove calls
 *                                    (updateState + energy(trial) + energy(accepted) + sync each)     src/montecarlo.cpp:139-187
 *   fb_widom_batch ................. WidomInsertion::_sample insertion loop            src/analysis.cpp:1243-1265
 *   fb_export_state/fb_import_state  MPI::ExchangeParticles / exchangeGroupSizes       src/mpicontroller.cpp:192-219, src/move.cpp:860-881
 *   fb_nonbonded_force ............. Nonbonded::force                                   src/energy.h:1584-1597
 *   fb_ewald_force ................. Ewald::force                                       src/energy.cpp:596-629
 *   fb_nccl_* ...................... the MPI messages of move::ParallelTempering        src/move.cpp:844-968, src/mpicontroller.cpp:69-259
 */
#ifndef FAUNUS_B200_H
#define FAUNUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_ctx fb_ctx;

typedef enum
{
    FB_OK = 0,
    FB_ERR_INVALID = -1, /* bad argument / unsupported configuration */
    FB_ERR_CUDA = -2,    /* CUDA runtime error (no device, launch failure, ...) */
    FB_ERR_STATE = -3,   /* call order violated (e.g. energy before upload) */
    FB_ERR_NOMEM = -4
} fb_status;

/* Pair-potential flavour, one per `nonbonded*` name (src/energy.cpp:1284-1327) */
typedef enum
{
    FB_POT_COULOMB_LJ = 0,  /* nonbonded_coulomblj : splined CoulombGalore + Lennard-Jones   */
    FB_POT_COULOMB_WCA = 1, /* nonbonded_coulombwca: splined CoulombGalore + WCA             */
    FB_POT_PM = 2,          /* nonbonded_pm        : plain Coulomb + hard sphere             */
    FB_POT_PMWCA = 3,       /* nonbonded_pmwca     : plain Coulomb + WCA                     */
    FB_POT_FUNCTOR = 4,     /* nonbonded           : per-(type,type) sum of terms, see flags */
    FB_POT_SPLINED = 5      /* nonbonded_splined   : per-(type,type) Andrea table in r^2     */
} fb_potential_kind;

/* Per type-pair term flags for FB_POT_FUNCTOR / exact part of FB_POT_SPLINED */
enum
{
    FB_TERM_COULOMB_SPLINED = 1, /* lB zz/r S(r/Rc) exp(-kappa r), r < Rc (src/potentials.h:591-598) */
    FB_TERM_COULOMB_PLAIN = 2,   /* lB zz/sqrt(r2)                         (src/potentials.h:472-476) */
    FB_TERM_LJ = 4,              /* 4 eps ((s2/r2)^6 - (s2/r2)^3)          (src/potentials.h:42-49)   */
    FB_TERM_WCA = 8,             /* LJ + 1/4 cut at r2 > s2 2^(1/3)         (src/potentials.h:151-160) */
    FB_TERM_HARDSPHERE = 16      /* r2 < s2 ? inf : 0                       (src/potentials.h:203-207) */
};

/* Molecule-kind flags (src/molecule.h:233-236) */
enum
{
    FB_MOL_ATOMIC = 1,
    FB_MOL_RIGID = 2,
    FB_MOL_COMPRESSIBLE = 4
};

typedef struct
{
    int device; /* CUDA device ordinal */

    /* geometry: orthogonal cell, per-axis periodicity (cuboid 1,1,1; slit 1,1,0; sphere 0,0,0) */
    double box[3];
    int periodic[3];

    /* atom / molecule kinds */
    int n_atom_types;
    int n_molecule_types;
    const int* molecule_flags;              /* [n_molecule_types] FB_MOL_* */
    const int* molecule_natoms;             /* [n_molecule_types] atoms per molecule (exclusion matrix edge) */
    const unsigned char* const* exclusions; /* [n_molecule_types] natoms*natoms 0/1 matrix or NULL */
    const double* g2g_cutoff_squared;       /* [n_molecule_types^2] mass-centre cutoff^2 (DBL_MAX = none) */

    /* pair potential */
    int kind;                    /* fb_potential_kind */
    const uint32_t* pair_flags;  /* [n_atom_types^2] FB_TERM_* (FUNCTOR, SPLINED); NULL otherwise */
    const double* lj_sigma2;     /* [n_atom_types^2] sigma_ij^2   or NULL */
    const double* lj_eps4;       /* [n_atom_types^2] 4 eps_ij/kT  or NULL */
    const double* wca_sigma2;    /* idem for WCA */
    const double* wca_eps4;
    const double* hs_sigma2;     /* [n_atom_types^2] hard-sphere sigma_ij^2 or NULL */

    /* splined Coulomb (CoulombGalore): u = lB zz / r * S(r/Rc) * exp(-kappa r) for r < Rc */
    double coulomb_bjerrum_length;
    double coulomb_cutoff;
    double coulomb_kappa;
    int coulomb_n_knots;          /* Andrea table of S(q), q in [0,1] */
    const double* coulomb_knots;  /* [n_knots] */
    const double* coulomb_coeffs; /* [6 (n_knots-1)] */

    /* plain Coulomb (FB_POT_PM, FB_POT_PMWCA, FB_TERM_COULOMB_PLAIN) */
    double plain_bjerrum_length;

    /* per-pair splines in r^2 (FB_POT_SPLINED), concatenated over the n_atom_types^2 pairs */
    const int* spline_offset;     /* [n_atom_types^2 + 1] first knot of each pair table */
    const double* spline_knots;   /* [spline_offset[n^2]] */
    const double* spline_coeffs;  /* 6 per interval, interval k of pair p at 6*(spline_offset[p] - p + k) */
    const double* spline_rmin2;   /* [n_atom_types^2] */
    const double* spline_rmax2;   /* [n_atom_types^2] */
    const unsigned char* spline_hardsphere; /* [n_atom_types^2] infinite below rmin */
} fb_config;

/* One group record: index range into the particle arrays + mass centre (src/group.h:60-177) */
typedef struct
{
    int begin;    /* first particle slot */
    int size;     /* active particles */
    int capacity; /* active + inactive */
    int molid;    /* molecule type */
    double cm[3]; /* mass centre (molecular groups) */
} fb_group;

/* One changed group of a Change record (src/space.h:40-52) */
typedef struct
{
    int group_index;
    int all;          /* Change::GroupChange::all */
    int internal;     /* Change::GroupChange::internal */
    int n_atoms;      /* number of relative_atom_indices (0 with all) */
    const int* atoms; /* relative_atom_indices */
} fb_group_change;

/* Change record (src/space.h:29-75); matter_change is outside the hot-path scope */
typedef struct
{
    int everything;
    int volume_change;
    int n_groups;
    const fb_group_change* groups;
    /* Change::matter_change (src/space.h:29-75): the number of active particles of the listed groups differs
     * between the two states; the energy is that of the pairs with a listed ACTIVE particle
     * (GroupPairing::accumulateSpeciation, src/energy.h:1390-1435) */
    int matter_change;
} fb_change;

/* Ewald reciprocal space parameters (EwaldData, src/energy.cpp:28-59) */
typedef struct
{
    double alpha;
    double n_cutoff;
    double kappa;
    double surface_dielectric_constant; /* epss; < 1 means tinfoil */
    double bjerrum_length;
    int spherical_sum;
    int policy; /* 0 = PBC, 1 = PBCEigen (reference full-update quirk), 2 = IPBC */
} fb_ewald_config;

/* ---- life cycle ------------------------------------------------------------------------- */
int fb_create(const fb_config* config, fb_ctx** ctx);
void fb_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx); /* ctx may be NULL: error of the last failed fb_create */
int fb_device_count(void);                    /* number of visible CUDA devices, 0 if none */

/* ---- device-resident Space mirror ------------------------------------------------------- */
/* full upload of one slot: xyzq[4*n_particles] (x,y,z,charge), atom_id[n_particles] */
int fb_upload_space(fb_ctx* ctx, int slot, const double* xyzq, const int* atom_id, const fb_group* groups,
                    int n_particles, int n_groups);
/* incremental update of one group: its record plus n_atoms particles at relative indices rel_index
 * (rel_index == NULL: the first n_atoms particles of the group) */
int fb_update_group(fb_ctx* ctx, int slot, int group_index, const fb_group* record, int n_atoms,
                    const int* rel_index, const double* xyzq, const int* atom_id);
int fb_set_box(fb_ctx* ctx, int slot, const double box[3]);
/* dst slot := src slot for what `change` lists (Space::sync direction); device-to-device */
int fb_sync(fb_ctx* ctx, int dst_slot, int src_slot, const fb_change* change);
int fb_download_space(fb_ctx* ctx, int slot, double* xyzq, int* atom_id, fb_group* groups);

/* ---- non-bonded energy -------------------------------------------------------------------- */
/* energy of the pairs `change` touches, evaluated on one slot (Nonbonded::energy) */
int fb_nonbonded_energy(fb_ctx* ctx, int slot, const fb_change* change, double* energy);
/* the same for two slots in ONE pass over the particles: u_new on slot_new, u_old on slot_old */
int fb_nonbonded_delta(fb_ctx* ctx, int slot_new, int slot_old, const fb_change* change, double* u_new,
                       double* u_old);

/* NonbondedBase::particleParticleEnergy (src/energy.h:1500, 1531-1534): pair energies of n explicit particle pairs
 * (x, y, z, q and atom id each) in the cell of `slot`; NonbondedBase::groupGroupEnergy (:1501, 1536-1545): all active
 * particles of two different groups, zero beyond the mass-centre cutoff of two molecular groups */
int fb_particle_pair_energy(fb_ctx* ctx, int slot, int n_pairs, const double* a_xyzq, const int* a_id,
                            const double* b_xyzq, const int* b_id, double* energy);
int fb_group_group_energy(fb_ctx* ctx, int slot, int group1, int group2, double* energy);

/* ---- forces (EnergyTerm::force, src/externalpotential.h:40; Hamiltonian::force, src/energy.cpp:1162-1166) ------
 * forces[3 * n_particles] in kT/Angstrom, one vector per particle slot of the uploaded space (the reference sizes the
 * vector by spc.particles, src/energy.h:1588).
 * fb_set_force_table: Andrea table of S'(q), q in [0,1] — the derivative of the CoulombGalore short-range function
 *   whose own table went into fb_config (coulomb_knots / coulomb_coeffs); knots[n_knots], coeffs[6 (n_knots - 1)].
 * fb_nonbonded_force: Nonbonded::force (src/energy.h:1584-1597) — ADDS, to every particle, the pair forces of all
 *   other particles of the vector (active or not, no group rules: the reference's stub, reproduced). Pair forces
 *   exist for FB_POT_COULOMB_LJ and FB_POT_COULOMB_WCA (src/potentials.h:32-40, 173-184, 600-606); every other kind
 *   returns FB_ERR_INVALID with the message of PairPotential::force (src/potentials.cpp:246-251).
 * fb_ewald_force: Ewald::force (src/energy.cpp:596-629) — OVERWRITES forces with the surface + reciprocal-space
 *   force from the Q(k) of `slot` (as the reference does: `(*force) = ...`). */
int fb_set_force_table(fb_ctx* ctx, int n_knots, const double* knots, const double* coeffs);
int fb_nonbonded_force(fb_ctx* ctx, int slot, double* forces);
int fb_ewald_force(fb_ctx* ctx, int slot, double* forces);

/* share `shard` of `n_shards` of the full-system energy of a slot, for one evaluation spread over several
 * GPUs that hold the same Space (GroupPairingPolicy::all, src/energy.h:1290-1326: tile rows are dealt
 * round robin; PolicyIonIon::updateComplex + reciprocalEnergy, src/energy.cpp:191-206, 524-531: a slab of
 * k-vectors, Q(k) rebuilt from the positions and not stored). The shares of all shards add up to
 * fb_nonbonded_energy(everything) and to the reciprocal part of fb_ewald_energy. */
int fb_system_energy_shard(fb_ctx* ctx, int slot, int shard, int n_shards, double* nonbonded, double* reciprocal);

/* ---- fast path: one launch per small trial move -------------------------------------------- */
/* A trial move of up to FB_FAST_ATOMS particles of one group without size change
 * (AtomicTranslateRotate, src/move.cpp:267-293; TranslateRotate of small rigid molecules, :1670-1689).
 * Both slots must mirror the accepted state when fb_trial_energy is called; the trial particles
 * travel in the kernel parameters. fb_trial_energy = updateState + energy(trial) + energy(accepted)
 * of the non-bonded term and, with `with_ewald`, of the Ewald term, in ONE kernel launch.
 * fb_trial_commit = the following sync: accept is applied lazily by the next launch, reject is free. */
#define FB_FAST_ATOMS 8
typedef struct
{
    int group_index;
    int n_atoms;                     /* 1..FB_FAST_ATOMS */
    int rel_index[FB_FAST_ATOMS];    /* relative atom indices within the group */
    double xyzq[FB_FAST_ATOMS][4];   /* trial position and charge */
    int atom_id[FB_FAST_ATOMS];
    double cm[3];                    /* trial mass centre of the group */
    int internal;                    /* Change::GroupChange::internal */
    int with_ewald;                  /* include the reciprocal-space partial update + energy */
} fb_trial_move;
/* ewald_new / ewald_old: surface-free reciprocal energies 2 pi lB / V * sum_k A_k |Q_k|^2 (NULL allowed
 * without with_ewald) */
int fb_trial_energy(fb_ctx* ctx, const fb_trial_move* move, double* u_new, double* u_old, double* ewald_new,
                    double* ewald_old);
int fb_trial_commit(fb_ctx* ctx, int accept);

/* ---- windowed path: a run of consecutive single-atom trial moves in one pass ------------------ */
/* The proposals of `transrot` (AtomicTranslateRotate, src/move.cpp:267-293) do not depend on energies,
 * so the caller can draw a window of up to FB_BATCH_MAX consecutive moves on DISTINCT atoms of atomic
 * groups ahead of time. fb_batch_trial evaluates all of them against the window-start state (both
 * slots must mirror the accepted state) and returns, besides the per-move energies, the exact
 * corrections that apply when an EARLIER move of the window has been accepted:
 *
 *   u_new(m | accepted set A) = u_new[m] + sum_{a in A, a < m} cross_new[m * stride + a]
 *   u_old(m | A)              = u_old[m] + sum_{a in A, a < m} cross_old[m * stride + a]
 *   dU_rec(m | A)             = rec_prefactor * (rec_delta[m] + 2 sum_{a in A, a < m} rec_cross[m * stride + a])
 *
 * (nonbonded energy of the moved atom with all other active particles, src/energy.h:1182-1195 + 885-914;
 * reciprocal Ewald energy change of the partial update, src/energy.cpp:219-247 + 524-531.) The caller
 * decides the moves in order (Metropolis on the host, reference RNG order) and reports the outcome with
 * fb_batch_commit; undecided moves (n_decided < n_moves) are simply dropped and may be re-submitted.
 * cross_max[m * stride + a] is the largest |pair energy| entering cross_new/old: if it is huge or
 * infinite the caller should stop the window at move m and re-evaluate it in the next one
 * (cancellation). The result arrays live in pinned memory owned by the context and stay valid until
 * the next fb_batch_trial. */
#define FB_BATCH_MAX 64
typedef struct
{
    int group_index;
    int rel_index;  /* relative atom index within the group */
    int atom_id;        /* atom type at the trial position */
    double xyzq[4];     /* trial position and charge */
    int old_atom_id;    /* the atom as it is in the accepted Space (what both slots hold for it) */
    double old_xyzq[4];
} fb_batch_move;
typedef struct
{
    int n_moves;
    int stride;              /* row length of the [m][a] matrices (16, 32 or 64): one row per move m */
    const double* u_new;     /* [n_moves] */
    const double* u_old;     /* [n_moves] */
    const double* rec_delta; /* [n_moves] sum_k A_k (2 Re(conj Q_k d_k) + |d_k|^2), 0 without Ewald */
    const double* cross_new; /* [stride * stride], row m holds the entries of the earlier moves a < m */
    const double* cross_old;
    const double* cross_max;
    const double* rec_cross; /* sum_k A_k Re(conj d_a,k d_m,k) */
    double rec_start;        /* sum_k A_k |Q_k|^2 of the window-start state */
    double rec_prefactor;    /* 2 pi lB / V */
    int n_atoms;             /* group mode: moved atoms in the window (rec_delta/rec_cross are per ATOM); else n_moves */
} fb_batch_result;
int fb_batch_trial(fb_ctx* ctx, int n_moves, const fb_batch_move* moves, int with_ewald, fb_batch_result* result);
/* the same in two halves, so that the caller can draw the next proposals while the device works:
 * fb_batch_submit queues the launches and returns, fb_batch_wait blocks for the results */
int fb_batch_submit(fb_ctx* ctx, int n_moves, const fb_batch_move* moves, int with_ewald);
int fb_batch_wait(fb_ctx* ctx, fb_batch_result* result);
/* Group mode: a window of rigid-body moves of whole molecular groups (Move::TranslateRotate::_move,
 * src/move.cpp:1623-1689; Change {group, all = true, internal = false}), at most 8 atoms per group and 64 atoms per
 * window, distinct groups. u_new/u_old[m] = energy of moved group m with every other group (group2all with the
 * mass-centre cutoff, src/energy.h:1155-1163, 979-989, 761-768), cross_*[m * stride + a] = the same group-group
 * differences between the moves of the window. The k-space results stay PER ATOM, atoms numbered in the order
 * given (move 0's atoms first): for a move holding atoms i..j
 *   rec_delta(move) = sum_i rec_delta[i] + 2 sum_{i<j in move} rec_cross[j * stride + i],
 *   rec_cross(m, a) = sum_{i in a, j in m} rec_cross[j * stride + i].
 * Decide and fb_batch_commit as in atomic mode (accepted[] is per move). */
typedef struct
{
    int group_index;
    int n_atoms;              /* active size of the group, 1..8 */
    int atom_id[8];
    double xyzq[8][4];        /* trial positions and charges */
    double cm[3];             /* trial mass centre */
    int old_atom_id[8];
    double old_xyzq[8][4];    /* the group as it is in the accepted Space */
    double old_cm[3];
} fb_batch_group_move;
int fb_batch_submit_groups(fb_ctx* ctx, int n_moves, const fb_batch_group_move* moves, int with_ewald);
/* Runs: windows decided on the device (the batched sequential-MC driver). For single-atom moves
 * everything the in-order walk needs besides the device results is known when the proposals are drawn: the
 * Metropolis uniform (always drawn, src/montecarlo.cpp:17-34) and the energies of the caller's own Hamiltonian
 * terms. A run is up to FB_RUN_MAX proposals on DISTINCT atoms; the device evaluates it window by window
 * (64 moves), walks each window in order — corrected energies, Hamiltonian sum in term order with the reference's
 * early exit (src/energy.cpp:1227-1247), getEnergyChange (src/montecarlo.cpp:193-209), Metropolis; as a fixed-point
 * iteration over all moves of the window at once, same results — and feeds the accepted moves to the next window
 * without a host round trip. fb_run_wait returns the outcome of every move
 * (all moves of the run are decided when it returns); the caller replays it into its Space. Nothing to commit. */
#define FB_RUN_MAX 1024
/* The caller's Hamiltonian must be [its own terms ..., the non-bonded term, the Ewald term (with_ewald)]: the walk
 * adds host_new / host_old (the caller's in-order sum of its own terms), the pair energy and the reciprocal energy
 * in this order and stops after a term >= max_energy or NaN like Hamiltonian::energy (src/energy.cpp:1238-1244). */
#define FB_RUN_HOST_NEW_CLOSED 1 /* the caller's own sum already ended early (trial state) */
#define FB_RUN_HOST_OLD_CLOSED 2 /* ... accepted state */
#define FB_RUN_ALT_HOST_NEW_CLOSED 4
#define FB_RUN_ALT_HOST_OLD_CLOSED 8
#define FB_RUN_DEP_PREVIOUS (1 << 30)
/* A proposal on an atom that ONE earlier, still undecided proposal already moves may travel too: it starts where
 * that move leaves the atom, so the caller supplies both variants — `move`, host_new, host_old if the earlier move
 * (`depends_on`: its index in this run, or FB_RUN_DEP_PREVIOUS | its index in the run this one is queued behind) is
 * accepted, `alt` (+ alt_host_*) if it is rejected. The device picks the variant once that move is decided and
 * never puts the two into one window. depends_on = -1: an ordinary proposal. */
typedef struct
{
    fb_batch_move move;
    double uniform;   /* the Metropolis uniform drawn for this move */
    double host_new;  /* in-order sum of the caller's own Hamiltonian terms, trial state */
    double host_old;  /* ... accepted state */
    int flags;        /* FB_RUN_*_CLOSED */
    int depends_on;
    fb_batch_move alt;
    double alt_host_new, alt_host_old;
} fb_run_move;
typedef struct
{
    double max_energy;         /* the sum over terms stops after a term >= this or NaN */
    double cancellation_limit; /* see cross_max above */
} fb_run_config;
typedef struct
{
    int n_moves;
    int n_windows;                  /* windows the device needed */
    int n_rounds;                   /* rounds of the fixed-point walk, summed over the windows */
    const unsigned char* accepted;  /* [n_moves] */
    const double* u_new;            /* [n_moves] Hamiltonian energy of the move in the trial state at its turn */
    const double* u_old;            /* [n_moves] ... in the accepted state */
} fb_run_result;
int fb_run_submit(fb_ctx* ctx, int n_moves, const fb_run_move* moves, int with_ewald, const fb_run_config* config);
int fb_run_wait(fb_ctx* ctx, fb_run_result* result);
/* pair_sums_ahead != 0: inside a run the pair sums of a window are evaluated one window ahead (beside the k-space
 * kernel and the walk of the window before) and corrected for the moves accepted since — the same exact identity
 * as the cross terms inside a window; 0 (default): every window evaluates its pair sums itself. Off by default: the
 * k-space kernel owns the register file, so the kernel running ahead only finds room in the gaps, where it delays
 * the walk (measured: +1 % moves/s at N = 1e5, -12 % at N = 2304). */
/* bit 1 of the same argument (value 2) switches the CUDA-graph replay of a run's window launches off: by default the
 * launch sequence of a run (4 kernels per window on two streams) is captured the second time it comes up with the same
 * arguments and replayed afterwards — same kernels, same order, same results, fewer microseconds between them. */
int fb_configure_runs(fb_ctx* ctx, int pair_sums_ahead);
/* since creation: out[0] = runs, out[1] = windows in runs, out[2] = rounds of the fixed-point walk, out[3] = moves */
int fb_get_run_stats(const fb_ctx* ctx, double out[4]);
/* accepted[m] != 0 for the accepted ones among the first n_decided moves of the last window */
int fb_batch_commit(fb_ctx* ctx, int n_decided, const unsigned char* accepted);
/* The pair part of a window goes through a device cell list (cell edge = box / floor(box / cutoff), 27
 * neighbour cells per position; src/celllistimpl.h:223-236 for the geometry convention) when the system is
 * all-atomic, every pair term has a cutoff, the cell is periodic in x, y, z with at least 3 cells per axis
 * and there are at least `min_particles` particle slots (default 200000 — below that the brute-force pass over
 * the L2-resident positions is as fast; 0: whenever eligible; < 0: never).
 * The reference itself is brute force (src/energy.h:1182-1195); sums differ in order only. */
int fb_configure_cells(fb_ctx* ctx, int min_particles);
/* test hook: initial bucket capacity of the cell list (doubled by the library whenever a bucket runs full) */
int fb_debug_set_cell_capacity(fb_ctx* ctx, int capacity);
/* timing enabled: out[0..2] = ms in the pair / k-space / other (commit, phase tables, final sums) kernels of
 * the windowed path (kernels serialised while timing), out[3] = windows, out[4] = moves evaluated,
 * out[5] = ms from the first to the last kernel of every window (always accumulated), out[6] = host round trips
 * (fb_batch_wait + fb_run_wait), out[7] = runs */
int fb_get_batch_timing(const fb_ctx* ctx, double out[8]);
/* timing enabled: the k-space share out[1] of fb_get_batch_timing split into out[0] = ms in windowFrontKernel (phase
 * tables, commit of the previous window) and out[1] = ms in windowKspaceKernel */
int fb_get_kspace_timing(const fb_ctx* ctx, double out[2]);

/* ---- pair distance histogram ------------------------------------------------------------------ */
/* One sample of AtomRDF (src/analysis.cpp:1556-1600) on the slot's mirror: all pairs of active atoms of the types
 * atom_id1, atom_id2 (i < j when the types are equal), minimum-image distance vector (src/geometry.h:429-458),
 * bin floor(r / dr) (src/aux/equidistant_table.h:32-40); slice_dir (may be NULL) and thickness as `slicedir` /
 * `thickness` of the analysis. The pair counts are ADDED to counts[n_bins] (n_bins <= 12288; an error if a
 * distance falls beyond the last bin). Exact integer counts, independent of any summation order — so the work
 * shards over GPUs without a second thought: shard `shard` of `n_shards` takes every n_shards-th row of 256 first
 * particles, the histograms of the shards add up (an integer all-reduce) to the unsharded one. */
int fb_atom_rdf(fb_ctx* ctx, int slot, int atom_id1, int atom_id2, double dr, const int* slice_dir, double thickness,
                int shard, int n_shards, int n_bins, unsigned long long* counts);

/* MoleculeRDF (src/analysis.cpp:1607-1658): the same for the mass centres of the active molecular groups of the kinds
 * molid1, molid2, distance = sqrt(Geometry::sqdist) (src/geometry.h:460-470), no slices. */
int fb_molecule_rdf(fb_ctx* ctx, int slot, int molid1, int molid2, double dr, int shard, int n_shards, int n_bins,
                    unsigned long long* counts);

/* ---- Ewald reciprocal space --------------------------------------------------------------- */
int fb_ewald_configure(fb_ctx* ctx, const fb_ewald_config* config);
/* k-vectors and A_k for the slot's current box (PolicyIonIon::updateBox); returns K in *n_kvectors */
int fb_ewald_update_box(fb_ctx* ctx, int slot, int* n_kvectors);
/* Q(k) = sum_j q_j exp(i k.r_j) over all active particles (updateComplex, full; src/energy.cpp:191-217). PBC / PBCEigen:
 * a complex matrix product [X.Y] x [Z] on the FP64 tensor path (ewaldFullGemmKernel, fb_fullq.cuh), sums in a fixed order;
 * with timing enabled fb_last_kernel_ms() is the device time of the rebuild. FAUNUS_B200_FULLQ=cells (at fb_create)
 * selects the older one-block-per-k-cell kernel for comparisons. */
int fb_ewald_update_full(fb_ctx* ctx, int slot);
/* Q_new(k) = Q_old(k) + sum_moved [q e^{ik.r}]_new - [q e^{ik.r}]_old (updateComplex, partial) */
int fb_ewald_update_partial(fb_ctx* ctx, int slot_new, int slot_old, const fb_change* change);
/* surface + reciprocal energy of a slot (Ewald::energy) */
int fb_ewald_energy(fb_ctx* ctx, int slot, const fb_change* change, double* energy);
/* Ewald::sync: Q (and, for everything/volume changes, k-vectors and A_k) dst := src */
int fb_ewald_sync(fb_ctx* ctx, int dst_slot, int src_slot, const fb_change* change);
/* debug/tests, no device needed: the k-vectors of `config` in `box` in the storage order of fb_ewald_update_box (integer
 * triplets nxyz[3K]) and the tile layout of the full-Q matrix product (fb_fullq.cuh): index[K] = tile * 2048 + row * 64 +
 * column, tiles[4 T] = {nx of the first row, ny + ceil(n_cutoff) of the first row, nz + ceil(n_cutoff) of the first column,
 * column groups of 8}, order[T] (heaviest tiles first), column_first_tile[columns + 1] (slab boundaries). NULL skips;
 * FB_ERR_INVALID if max_k / max_tiles are too small (n_k, n_tiles are set). */
int fb_debug_fullq_layout(const fb_ewald_config* config, const double box[3], int max_k, int max_tiles, int* nxyz, int* index,
                          int* tiles, int* order, int* column_first_tile, int* n_k, int* n_tiles, int* n_columns);
/* debug/tests: copy out K complex numbers (re,im interleaved), k-vectors [3K], A_k [K] (NULL skips) */
int fb_ewald_download(fb_ctx* ctx, int slot, double* q_re_im, double* kvectors, double* aks);

/* ---- Widom: B independent ghost insertions per launch ------------------------------------- */
/* ghost_xyzq[B*n_ghost_atoms*4], ghost_atom_id[n_ghost_atoms], ghost_cm[B*3] (molecular ghosts;
 * NULL for atomic), internal != 0 adds the ghost's own pair energy; du[B] = non-bonded energy of
 * each ghost with all active particles of `slot` */
int fb_widom_batch(fb_ctx* ctx, int slot, int ghost_group_index, int n_ghost_atoms, int n_insertions,
                   const double* ghost_xyzq, const int* ghost_atom_id, const double* ghost_cm, int internal,
                   double* du);

/* ---- replica exchange (parallel tempering) ------------------------------------------------ */
/* packed device-side state of a slot: [box(3) | group sizes (G) | x,y,z,q,id per particle (5N)]
 * as doubles; fb_state_doubles gives the length. The buffers are DEVICE pointers so that the
 * exchange can go GPU to GPU (NCCL send/recv or peer copy) without touching the host. */
size_t fb_state_doubles(const fb_ctx* ctx);
int fb_export_state(fb_ctx* ctx, int slot, double* device_buffer);
int fb_import_state(fb_ctx* ctx, int slot, const double* device_buffer);
/* host-side view of the same packing (tests, gloo) */
int fb_export_state_host(fb_ctx* ctx, int slot, double* host_buffer);
int fb_import_state_host(fb_ctx* ctx, int slot, const double* host_buffer);
/* group records (sizes, mass centres) of a slot whose particles are already in place (after fb_import_state: the
 * mass centres of molecular groups follow the reference's minimum-image rule, src/geometry.h:504-527, and are
 * computed by the caller's Space); begin / capacity / molid must be those of the uploaded space */
int fb_upload_groups(fb_ctx* ctx, int slot, const fb_group* groups, int n_groups);

/* NCCL between the contexts of the replicas (one context per GPU / process), replacing the reference's MPI
 * point-to-point messages of the Temper move (src/move.cpp:860-923, src/mpicontroller.cpp:192-259). libnccl is
 * opened at run time. fb_nccl_unique_id: 128 bytes made by ONE rank and distributed by the launcher (any channel);
 * fb_nccl_init: collective over the `size` contexts. */
int fb_nccl_unique_id(char out[128]);
int fb_nccl_init(fb_ctx* ctx, const char id[128], int rank, int size);
int fb_nccl_finalize(fb_ctx* ctx);
/* packed state of `send_slot` (the accepted state: the authoritative mirror) → partner, the partner's → `recv_slot`
 * (the trial state): ncclSend/ncclRecv on the device buffers, import on the device; host_received (fb_state_doubles
 * doubles, may be NULL) gets a copy of what arrived so that the caller's Space can follow. Both partners call it
 * with each other's rank. */
int fb_nccl_exchange_state(fb_ctx* ctx, int send_slot, int recv_slot, int partner, double* host_received);
/* n doubles in place with `partner` (MPI_Sendrecv_replace of the 8-byte energy change, src/move.cpp:905-923) */
int fb_nccl_sendrecv_host(fb_ctx* ctx, double* data, size_t n, int partner);
/* one double per rank to every rank, out[size] (checkRandomEngineState, src/mpicontroller.cpp:253-259; a barrier too) */
int fb_nccl_allgather_host(fb_ctx* ctx, double value, double* out);
unsigned long long fb_nccl_bytes_exchanged(const fb_ctx* ctx);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* number of kernels this context has launched so far */
unsigned long long fb_launch_count(const fb_ctx* ctx);
/* raw CUDA stream (cudaStream_t) of the context, for event timing by the caller */
void* fb_stream(const fb_ctx* ctx);
/* device time in milliseconds spent in the kernels launched by the most recent energy call
 * (cudaEvent pair on the context's stream); enabled with fb_enable_timing */
int fb_enable_timing(fb_ctx* ctx, int on);
double fb_last_kernel_ms(const fb_ctx* ctx);
/* accumulated since creation (timing enabled): out[0] = ms in the moved-set pair kernel, out[1] = its
 * launches, out[2] = ms in the Ewald partial-update kernel, out[3] = its launches, out[4] = ms in the
 * full-energy kernel (+ ordered sum), out[5] = its launches, out[6] = ms in Widom kernels, out[7] = launches */
int fb_get_timing(const fb_ctx* ctx, double out[8]);
/* dependent-free DFMA microbenchmark on `device`: sustained FP64 FMA throughput in TFLOP/s */
int fb_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* FAUNUS_B200_H */

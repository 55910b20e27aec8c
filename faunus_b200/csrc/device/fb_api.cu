// Host side of the C ABI declared in include/faunus_b200.h: context, device-resident Space mirror
// (two slots), lowering of `Change` records to launch descriptors, kernel launches.
// One CUDA stream per context; results come back through mapped pinned memory.
#include "../../../include/faunus_b200.h"
#include <map>
#include <tuple>
#include <queue>
#include "fb_kernels.cuh"
#include "fb_batch.cuh"
#include "fb_kspace.cuh"
#include "fb_stream.cuh"
#include "fb_fullq.cuh"
#include "fb_cells.cuh"
#include "fb_run.cuh"
#include "fb_rdf.cuh"
#include "fb_force.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace fbdev;

#define FB_API extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_create_error;

struct CudaError
{
    std::string msg;
};

#define CUDA_CHECK(expr)                                                                                      \
    do {                                                                                                      \
        cudaError_t err__ = (expr);                                                                           \
        if (err__ != cudaSuccess) {                                                                           \
            throw CudaError{std::string(#expr) + ": " + cudaGetErrorString(err__)};                          \
        }                                                                                                     \
    } while (0)

template <class T> struct DeviceBuffer
{
    T* ptr = nullptr;
    size_t count = 0;
    void alloc(size_t n)
    {
        release();
        if (n > 0) {
            CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
        }
        count = n;
    }
    void ensure(size_t n)
    {
        if (n > count) {
            alloc(n + n / 4);
        }
    }
    void upload(const T* src, size_t n, cudaStream_t s)
    {
        ensure(n);
        if (n > 0) {
            CUDA_CHECK(cudaMemcpyAsync(ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
        }
    }
    void uploadVector(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
    void release()
    {
        if (ptr) {
            cudaFree(ptr);
            ptr = nullptr;
        }
        count = 0;
    }
    ~DeviceBuffer() { release(); }
};

template <class T> struct PinnedBuffer
{
    T* ptr = nullptr;
    size_t count = 0;
    void ensure(size_t n)
    {
        if (n > count) {
            if (ptr) {
                cudaFreeHost(ptr);
                ptr = nullptr;
            }
            CUDA_CHECK(cudaHostAlloc(&ptr, (n + n / 4) * sizeof(T), cudaHostAllocDefault));
            count = n + n / 4;
        }
    }
    ~PinnedBuffer()
    {
        if (ptr) {
            cudaFreeHost(ptr);
        }
    }
};

struct Slot
{
    DeviceBuffer<double4> posq;
    DeviceBuffer<int> atom_id;
    DeviceBuffer<int> gid;
    DeviceBuffer<double4> gcm;
    DeviceBuffer<int> gsize;
    double box[3] = {0, 0, 0};
    std::vector<fb_group> groups; //!< host shadow of the group table
    bool uploaded = false;
    // Ewald
    DeviceBuffer<double4> kA;
    DeviceBuffer<int4> kn; //!< integer triplets (nx, ny, nz) of the k-vectors, same order as kA
    DeviceBuffer<double> ksq; //!< sqrt(A_k)
    DeviceBuffer<int> cell_start; //!< [n_cells + 1] first k of each 4×4×4 cell of integer triplets (storage order)
    int n_cells = 0;
    // tiles of the full rebuild of Q(k) (fb_fullq.cuh)
    DeviceBuffer<int4> gemm_tiles; //!< [n_gemm_tiles] {nx, y table index, first column group, column groups}, storage order
    DeviceBuffer<int> gemm_order;  //!< [n_gemm_tiles] the tiles heaviest first
    DeviceBuffer<int> gemm_row_groups; //!< [n_gemm_tiles] column groups per nx of the tile (8 bits each: first | (last + 1) << 4)
    DeviceBuffer<int> gemm_index;  //!< [K] tile · 2048 + row · 64 + column of every k-vector
    int n_gemm_tiles = 0;
    std::vector<int> gemm_tile_first_k;      //!< [n_gemm_tiles + 1] first k-vector of the tile's column of cells
    std::vector<int> gemm_column_first_tile; //!< [columns + 1] slabs of the sharded energy are ranges of columns
    // work units of the window k-space kernel (fb_kspace.cuh): halves of the cells, 32 k-slots each
    DeviceBuffer<double2> aks;            //!< [K] {A_k, √A_k}
    DeviceBuffer<int4> unit_info;         //!< [n_units] {first k of the cell, x, y, z table index of the first slot}
    DeviceBuffer<unsigned char> unit_map; //!< [n_units][32] slot → index inside the cell's storage range, 255: none
    DeviceBuffer<double> unit_sa;         //!< [n_units][32] √A_k in slot layout, 0 for empty slots
    int n_units = 0;
    // static schedule of the units over the blocks of windowKspaceKernel (balanced by cost on the host: a result never
    // depends on which block happens to run first)
    DeviceBuffer<unsigned char> unit_steps; //!< [n_units] bit s: slot column s = 4·jj + l holds a k-vector for some x-index
    DeviceBuffer<int> sched_first;          //!< [n_sched_blocks + 1]
    DeviceBuffer<int> sched_units;          //!< [n_units] units of block b: sched_units[sched_first[b] … sched_first[b + 1])
    int n_sched_blocks = 0;
    // work items of the commit (windowFrontKernel): two z-adjacent cells = up to four units that share table entries
    DeviceBuffer<int4> item_units; //!< [n_items] unit of (cell 0, h 0), (cell 0, h 1), (cell 1, h 0), (cell 1, h 1); −1: none
    DeviceBuffer<int4> item_base;  //!< [n_items] table index of x, y (h = 0), z of cell 0, z of cell 1
    int n_items = 0;
    std::vector<int> perm; //!< storage index → index in the reference's k-vector order
    DeviceBuffer<double2> Q;
    int K = 0;
    double ewald_box[3] = {0, 0, 0};
    bool rec_valid = false; //!< rec_sum holds Σ A_k |Q_k|² of the current Q
    double rec_sum = 0;
};

} // namespace

struct fb_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    unsigned long long launches = 0;

    int n_slots = 0;
    int n_groups = 0;
    int periodic[3] = {1, 1, 1};
    Slot slot[2];
    DeviceBuffer<int> gbegin, gcap, ginfo;
    std::vector<int> molecule_flags;

    // potential tables
    PotParams P{};
    DeviceBuffer<unsigned> d_flags;
    DeviceBuffer<double> d_lj_s2, d_lj_e4, d_wca_s2, d_wca_e4, d_hs_s2, d_knots, d_coef, d_g2g;
    DeviceBuffer<int> d_lut, d_sp_offset, d_excl_offset, d_excl_natoms;
    DeviceBuffer<double> d_sp_knots, d_sp_coef, d_sp_rmin2, d_sp_rmax2;
    DeviceBuffer<unsigned char> d_sp_hs, d_excl;

    // reductions and results
    DeviceBuffer<double> partials;
    DeviceBuffer<unsigned> ticket;
    double* h_result = nullptr; //!< mapped pinned, 8 doubles
    double* d_result = nullptr; //!< device alias of h_result

    // staging
    PinnedBuffer<int> h_list;
    DeviceBuffer<int> d_list;
    PinnedBuffer<int> h_upd_slot, h_upd_id;
    PinnedBuffer<double4> h_upd_posq;
    DeviceBuffer<int> d_upd_slot, d_upd_id;
    DeviceBuffer<double4> d_upd_posq;
    DeviceBuffer<double4> d_ghost, d_ghost_cm;
    DeviceBuffer<int> d_ghost_id;
    DeviceBuffer<double> d_widom_partial, d_widom_du;
    PinnedBuffer<double> h_widom_du;
    PinnedBuffer<double4> h_ghost;
    DeviceBuffer<double> d_state, d_state_recv, d_xchg;
    PinnedBuffer<double> h_state, h_xchg;
    // replica exchange over NCCL (fb_nccl.inl)
    void* nccl_comm = nullptr;
    int nccl_rank = 0, nccl_size = 0;
    unsigned long long bytes_exchanged = 0;

    // Ewald
    bool ewald_configured = false;
    fb_ewald_config ewald{};

    // fast path (fb_trial_energy / fb_trial_commit)
    Overlay commit{};          //!< accepted move not yet written to the mirrors
    bool has_commit = false;
    Overlay trial{};
    bool trial_active = false;
    bool trial_with_ewald = false;
    double trial_rec_sum = 0;
    double sequence = 0;

    // timing
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0;
    int timing_class = -1;    //!< which accumulator the current event pair feeds
    double acc_ms[4] = {0, 0, 0, 0};
    double acc_launches[4] = {0, 0, 0, 0};

    int max_blocks = 148 * 4;
    int n_sm = 148;

    // windowed speculative evaluation (fb_batch_trial / fb_batch_commit), see fb_batch.cuh
    struct Batch
    {
        DeviceBuffer<BatchInput> d_in[2];
        DeviceBuffer<BatchInput> d_ahead[2];  //!< runs: the window predicted to follow (pair sums one window ahead)
        DeviceBuffer<double> d_pair_fix;      //!< … and what the accepted moves of the window before change in them
        DeviceBuffer<int> d_pair_redo;
        bool prepair = false;                 //!< fb_configure_runs (off: measured +1 % at S1, −12 % on examples/bulk)
        DeviceBuffer<double2> d_table[2];
        PinnedBuffer<BatchInput> h_in;
        DeviceBuffer<double> d_pair_partials, d_r_partials, d_g_partials, d_e_partials, d_result;
        DeviceBuffer<double2> d_kq; //!< [n_units][32] √A_k·Q_k of the window-start state (windowFrontKernel)
        DeviceBuffer<unsigned> d_tail_ticket; //!< windowTailKernel: blocks done (the last one walks the window)
        PinnedBuffer<double> h_result;
        int parity = 0;
        int last_n = 0;            //!< moves of the most recent window (0: none evaluated)
        int last_with_ewald = 0;
        int last_slots[kBatchMax] = {};
        CommitList pending{};      //!< accepted moves (atoms) of the previous window not yet on the device
        CommitList pending_moves{}; //!< group mode: the accepted moves (for the mass centres)
        int last_groups = 0;       //!< group mode: moves of the most recent window (0: atomic mode)
        int last_first[kBatchMax] = {}, last_natoms[kBatchMax] = {};
        bool has_pending = false;
        bool pending_with_ewald = false;
        bool q_dirty = false;      //!< slot 0's Q(k) is ahead of slot 1's
        bool rec_known = false;    //!< rec_sum is Σ A_k|Q_k|² of slot 0's current Q(k)
        bool last_rec_fresh = false; //!< the last window recomputed that sum on the device
        bool in_flight = false;      //!< fb_batch_submit done, fb_batch_wait pending
        int flight_n = 0, flight_stride = 0, flight_with_ewald = 0, flight_atoms = 0, flight_groups = 0;
        bool flight_timing = false;
        bool kspace_unit_configured = false; //!< dynamic shared memory opt-in of windowKspaceKernel done
        // device cell list of slot 0 for the pair part of a window (fb_cells.cuh)
        int cell_min_particles = 200000; //!< use the cell list from this many particle slots on (< 0: never)
        bool cells_valid = false;
        bool cells_used = false;        //!< the window in flight went through the cell list
        int cell_cap = 0;
        bool cell_cap_forced = false;
        int cell_dims[3] = {0, 0, 0};
        DeviceBuffer<int> d_cell_count, d_cell_bucket, d_cell_overflow;
        std::vector<fb_batch_move> last_moves; //!< proposals of the window in flight (for a brute-force re-run)
        // runs of windows decided on the device (fb_run.cuh)
        struct RunBlock
        {
            RunHeader header;
            RunMove moves[kRunMax];
        };
        struct RunBack
        {
            RunState state;
            double overflow; //!< result[2] of the last window: a cell bucket ran full
            RunOutput out[kRunMax];
        };
        /** a run in flight; two of them so that the next one can be queued behind the one the device works on */
        struct RunSlot
        {
            PinnedBuffer<RunBlock> h_run;
            DeviceBuffer<RunBlock> d_run;
            PinnedBuffer<RunBack> h_back;
            DeviceBuffer<RunBack> d_back;
            cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_done = nullptr;
            // copies beside the kernels (run_in_stream / run_out_stream): input uploaded, first kernel of the run through
            // (the chain kernel reads the header of the run before it: that slot's next upload waits for it), windows through
            cudaEvent_t ev_in = nullptr, ev_started = nullptr, ev_tail = nullptr;
            int id = 0;            //!< stamp of its proposals in run_stamp
            int n = 0, stride = 0, with_ewald = 0, steps_launched = 0;
            int last_parity = 0;   //!< window buffer of the last window launched for it
            bool chained = false;  //!< queued while its predecessor was still in flight (started by runChainKernel)
        };
        RunSlot run[2];
        bool gap_stats = false; //!< FAUNUS_B200_GAP_STATS: measure the device time between chained runs, report at fb_destroy
        double gap_ms_total = 0.0;
        long gap_count = 0;
        int run_head = 0;      //!< slot of the oldest run in flight
        int runs_in_flight = 0;
        std::vector<int> run_stamp; //!< per particle slot: the run that last touched it (distinctness check) …
        std::vector<int> run_last;  //!< … and its latest move there
        int run_id = 0;
        bool run_decide_configured = false;
        bool tail_configured[6] = {false, false, false, false, false, false};
        // CUDA graphs of the window launches of a run (launchRunGraph): the 4 kernels × steps windows on two streams
        // are captured the second time the same launch sequence comes up and replayed from then on
        struct RunGraph
        {
            int seen = 0;
            cudaGraphExec_t exec = nullptr;
            long launches = 0;
        };
        std::map<unsigned long long, RunGraph> run_graphs;
        bool run_graphs_enabled = true;
        long run_graph_replays = 0;
        std::vector<unsigned char> run_accepted;
        std::vector<double> run_u_new, run_u_old;
        double run_steps = 0, run_count = 0, run_rounds = 0, run_moves = 0;
        double round_trips = 0; //!< waits for the device (windows walked on the host + runs)
        bool force_brute = false;
        double rec_sum = 0;
        PhaseGeometry geo{};
        cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; //!< [5]: between the front and the k-space kernel
        cudaStream_t pair_stream = nullptr; //!< the pair kernel runs beside the k-space kernels
        // A run's input goes up and its decisions come back on streams of their own: on the context's stream the three
        // copies sat between the last kernel of a run and the first of the next (≈ 25 µs of an idle GPU per run)
        cudaStream_t run_in_stream = nullptr, run_out_stream = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        double acc_ms[3] = {0, 0, 0}; //!< pair, ewald, other (commit + phase + finish)
        double acc_front_ms = 0;      //!< the windowFrontKernel part of acc_ms[1]
        double acc_total_ms = 0;      //!< first to last kernel of every window (all modes)
        double windows = 0, moves = 0;
    } batch;
    double pair_cut2 = 0; //!< no pair energy beyond this r² (+inf when some term has no cutoff)
    DeviceBuffer<unsigned long long> rdf_hist; //!< fb_atom_rdf
    DeviceBuffer<int> rdf_flag;
    DeviceBuffer<double4> rdf_list[2];
    DeviceBuffer<int> rdf_n;
    bool rdf_configured = false;
    // forces (fb_force.cuh)
    DeviceBuffer<double> d_force_knots, d_force_coef; //!< Andrea table of S'(q) (fb_set_force_table)
    int force_nk = 0;
    int full_pair_path = 0; //!< 0: FP32-screened full pair sum (fullScreenKernel), 1: all-FP64 fullStreamKernel (FAUNUS_B200_FULLPAIR=fp64)
    int full_q_path = 0; //!< 0: matrix product (fb_fullq.cuh), 1: one block per k-cell (FAUNUS_B200_FULLQ=cells; comparisons)
    DeviceBuffer<double2> fullq_steps;   //!< full rebuild of Q(k): [3 N] weights and x, y unit phases, [N] z unit phases
    DeviceBuffer<double> fullq_partials; //!< full rebuild of Q(k): [tiles][particle ranges][2][32][64] shares
    DeviceBuffer<double2> q_partials; //!< sharded reciprocal energy: [cells of the slab][particle splits][64] shares of Q(k)
    DeviceBuffer<double> d_forces; //!< [n_slots][3]
    DeviceBuffer<double> d_force_shares; //!< [j ranges][n_slots][3]
    PinnedBuffer<double> h_forces;
};

namespace {

constexpr int kMaxPartialBlocks = 4096;

SlotView makeView(fb_ctx* c, int s)
{
    Slot& sl = c->slot[s];
    SlotView v{};
    v.posq = sl.posq.ptr;
    v.atom_id = sl.atom_id.ptr;
    v.gid = sl.gid.ptr;
    v.gcm = sl.gcm.ptr;
    v.gsize = sl.gsize.ptr;
    v.gbegin = c->gbegin.ptr;
    v.gcap = c->gcap.ptr;
    v.ginfo = c->ginfo.ptr;
    for (int i = 0; i < 3; ++i) {
        v.len[i] = sl.box[i];
        v.half[i] = 0.5 * sl.box[i];
        v.len_or_zero[i] = c->periodic[i] ? sl.box[i] : 0.0;
    }
    v.n_slots = c->n_slots;
    v.n_groups = c->n_groups;
    return v;
}

EwaldView makeEwaldView(fb_ctx* c, int s)
{
    EwaldView e{};
    e.kA = c->slot[s].kA.ptr;
    e.Q = c->slot[s].Q.ptr;
    e.K = c->slot[s].K;
    e.policy = c->ewald.policy;
    return e;
}

void launched(fb_ctx* c, const char* what);

/**
 * Cutoff² for the FP32 screening of the distance test (batchPairScreenKernel, widomScreenKernel, …): the true
 * cutoff² enlarged by a bound of the FP32 rounding error of a minimum-image r² — coordinates and box lengths to
 * 2⁻²⁴ relative, two subtractions per component (≤ 8·2⁻²⁴·L per component), three products and two sums — so that no
 * pair inside the true cutoff is lost
 */
float screeningCutoffFor(const fb_ctx* c, double lmax)
{
    const double eps = 5.9604644775390625e-08; // 2⁻²⁴
    const double rc = std::sqrt(c->pair_cut2);
    const double ec = 8.0 * eps * lmax;
    const double widened = (c->pair_cut2 + 2.0 * std::sqrt(3.0) * rc * ec + 3.0 * ec * ec + 8.0 * eps * c->pair_cut2) * (1.0 + 1e-6);
    return std::nextafter(static_cast<float>(widened), std::numeric_limits<float>::infinity());
}

float screeningCutoff(const fb_ctx* c, int s)
{
    const double* box = c->slot[s].box;
    return screeningCutoffFor(c, std::max(box[0], std::max(box[1], box[2])));
}

/** Write a lazily accepted fast-path move into both mirrors before any other kind of access */
void flushBatch(fb_ctx* c);

void flushPending(fb_ctx* c)
{
    flushBatch(c);
    if (c->has_commit) {
        applyCommitKernel<<<1, 32, 0, c->stream>>>(makeView(c, 0), makeView(c, 1), c->commit);
        launched(c, "applyCommitKernel");
        c->has_commit = false;
    }
    c->trial_active = false;
}

void checkSlot(fb_ctx* c, int s, bool need_upload = true)
{
    if (s < 0 || s > 1) {
        throw CudaError{"slot must be 0 (accepted) or 1 (trial)"};
    }
    if (need_upload && !c->slot[s].uploaded) {
        throw CudaError{"slot has no uploaded Space (call fb_upload_space first)"};
    }
}

enum TimingClass
{
    TIME_PAIR = 0,
    TIME_EWALD = 1,
    TIME_FULL = 2,
    TIME_WIDOM = 3
};

void beginTiming(fb_ctx* c, int timing_class = -1)
{
    if (c->timing) {
        c->timing_class = timing_class;
        CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    }
}

/** Wait for the stream; collects kernel time when timing is on */
void finish(fb_ctx* c)
{
    if (c->timing) {
        CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->timing) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->last_ms = ms;
        if (c->timing_class >= 0) {
            c->acc_ms[c->timing_class] += ms;
            c->acc_launches[c->timing_class] += 1;
            c->timing_class = -1;
        }
    }
}

void launched(fb_ctx* c, const char* what)
{
    c->launches++;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        throw CudaError{std::string(what) + " launch: " + cudaGetErrorString(err)};
    }
}

template <class F> int guarded(fb_ctx* c, F&& f)
{
    try {
        if (c == nullptr) {
            return FB_ERR_INVALID;
        }
        CUDA_CHECK(cudaSetDevice(c->device));
        f();
        return FB_OK;
    }
    catch (const CudaError& e) {
        c->last_error = e.msg;
        return (e.msg.find("cuda") == 0) ? FB_ERR_CUDA : FB_ERR_INVALID;
    }
    catch (const std::exception& e) {
        c->last_error = e.what();
        return FB_ERR_INVALID;
    }
}

/**
 * Lower a partial Change to the moved-atom list (GroupPairing::accumulate dispatch,
 * src/energy.h:1447-1475 and accumulateGroup :1347-1377): one group → by the number of indices
 * (1, none = all, subset); several groups → all active atoms of each, no internal pairs.
 * `for_ewald` follows PolicyIonIon::updateComplex (src/energy.cpp:228-235): `all` → iota(max size).
 */
MovedDesc lowerChange(fb_ctx* c, int slot_a, int slot_b, const fb_change* change, bool for_ewald)
{
    if (change->n_groups <= 0) {
        throw CudaError{"change lists no groups"};
    }
    if (change->n_groups > kMaxMovedGroups) {
        throw CudaError{"too many changed groups in one Change (max 16)"};
    }
    MovedDesc md{};
    md.n_groups = change->n_groups;
    std::vector<int> slots, gpos;
    const auto& ga = c->slot[slot_a].groups;
    const auto& gb = c->slot[slot_b].groups;
    const bool multi = change->n_groups > 1 && !for_ewald;
    md.all_moved = 1;
    for (int t = 0; t < change->n_groups; ++t) {
        const fb_group_change& gc = change->groups[t];
        if (gc.group_index < 0 || gc.group_index >= c->n_groups) {
            throw CudaError{"group index out of range"};
        }
        md.groups[t] = gc.group_index;
        const fb_group& g = ga[gc.group_index];
        bool take_all;
        int count;
        if (for_ewald) {
            take_all = gc.all != 0 && c->ewald.policy != 2;
            count = std::max(g.size, gb[gc.group_index].size);
        }
        else {
            take_all = multi || gc.n_atoms == 0;
            count = g.size;
        }
        if (take_all) {
            for (int i = 0; i < count; ++i) {
                slots.push_back(g.begin + i);
                gpos.push_back(t);
            }
        }
        else {
            md.all_moved = 0;
            for (int i = 0; i < gc.n_atoms; ++i) {
                if (gc.atoms[i] < 0 || gc.atoms[i] >= g.capacity) {
                    throw CudaError{"relative atom index out of range"};
                }
                slots.push_back(g.begin + gc.atoms[i]);
                gpos.push_back(t);
            }
        }
    }
    md.internal = (!multi && change->groups[0].internal) ? 1 : 0;
    md.n_moved = static_cast<int>(slots.size());
    if (md.n_moved <= kInlineMoved) {
        md.list = nullptr;
        for (int i = 0; i < md.n_moved; ++i) {
            md.inline_slot[i] = slots[i];
            md.inline_gpos[i] = gpos[i];
        }
    }
    else {
        CUDA_CHECK(cudaStreamSynchronize(c->stream)); // staging buffer reuse
        c->h_list.ensure(2 * slots.size());
        std::copy(slots.begin(), slots.end(), c->h_list.ptr);
        std::copy(gpos.begin(), gpos.end(), c->h_list.ptr + slots.size());
        c->d_list.upload(c->h_list.ptr, 2 * slots.size(), c->stream);
        md.list = c->d_list.ptr;
    }
    return md;
}

/**
 * Change::matter_change on ONE slot (the two states differ in the number of active particles, so they are
 * evaluated separately): the ACTIVE listed atoms of every changed group — flagged kMovedCross — and, for a group
 * changed as a whole (`all`), its other active atoms, which only take part in the pairs inside the group
 * (accumulateSpeciation, src/energy.h:1390-1435: group2groups with the static groups, group2group between changed
 * groups, groupInternal(group) / groupInternal(group, index)). A group without an active listed atom adds nothing.
 */
MovedDesc lowerMatterChange(fb_ctx* c, int s, const fb_change* change)
{
    if (change->n_groups <= 0 || change->n_groups > kMaxMovedGroups) {
        throw CudaError{"a matter change lists 1..16 groups"};
    }
    MovedDesc md{};
    md.n_groups = change->n_groups;
    md.matter = 1;
    md.internal = 1;
    md.all_moved = 0;
    std::vector<int> slots, gpos;
    const auto& groups = c->slot[s].groups;
    for (int t = 0; t < change->n_groups; ++t) {
        const fb_group_change& gc = change->groups[t];
        if (gc.group_index < 0 || gc.group_index >= c->n_groups) {
            throw CudaError{"group index out of range"};
        }
        if (t > 0 && gc.group_index <= change->groups[t - 1].group_index) {
            throw CudaError{"the groups of a Change must be sorted and distinct"};
        }
        md.groups[t] = gc.group_index;
        const fb_group& g = groups[gc.group_index];
        std::vector<char> listed(static_cast<size_t>(std::max(g.size, 0)), 0);
        bool any = false;
        for (int i = 0; i < gc.n_atoms; ++i) {
            if (gc.atoms[i] < 0 || gc.atoms[i] >= g.capacity) {
                throw CudaError{"relative atom index out of range"};
            }
            if (gc.atoms[i] < g.size && !listed[gc.atoms[i]]) { // active particles only
                listed[gc.atoms[i]] = 1;
                any = true;
            }
        }
        if (!any) {
            continue;
        }
        for (int i = 0; i < g.size; ++i) {
            if (listed[i]) {
                slots.push_back(g.begin + i);
                gpos.push_back(t | kMovedCross);
            }
            else if (gc.all) {
                slots.push_back(g.begin + i);
                gpos.push_back(t);
            }
        }
    }
    md.n_moved = static_cast<int>(slots.size());
    if (md.n_moved <= kInlineMoved) {
        md.list = nullptr;
        for (int i = 0; i < md.n_moved; ++i) {
            md.inline_slot[i] = slots[i];
            md.inline_gpos[i] = gpos[i];
        }
    }
    else {
        CUDA_CHECK(cudaStreamSynchronize(c->stream)); // staging buffer reuse
        c->h_list.ensure(2 * slots.size());
        std::copy(slots.begin(), slots.end(), c->h_list.ptr);
        std::copy(gpos.begin(), gpos.end(), c->h_list.ptr + slots.size());
        c->d_list.upload(c->h_list.ptr, 2 * slots.size(), c->stream);
        md.list = c->d_list.ptr;
    }
    return md;
}

int gridFor(fb_ctx* c, int n, int block)
{
    return std::max(1, std::min((n + block - 1) / block, c->max_blocks));
}

template <bool FUSED> void launchMoved(fb_ctx* c, const SlotView& A, const SlotView& B, const MovedDesc& md)
{
    const int grid = gridFor(c, c->n_slots, kBlock);
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        movedEnergyKernel<K, FUSED><<<grid, kBlock, 0, c->stream>>>(A, B, c->P, md, c->partials.ptr,          \
                                                                   c->ticket.ptr, c->d_result);               \
        break;
    switch (c->P.kind) {
        FB_CASE(POT_COULOMB_LJ)
        FB_CASE(POT_COULOMB_WCA)
        FB_CASE(POT_PM)
        FB_CASE(POT_PMWCA)
        FB_CASE(POT_FUNCTOR)
        FB_CASE(POT_SPLINED)
    default:
        throw CudaError{"unknown potential kind"};
    }
#undef FB_CASE
    launched(c, "movedEnergyKernel");
}

template <int KIND> void launchFullStream(fb_ctx* c, const SlotView& V, dim3 grid, int shard, int n_shards)
{
    if (std::isinf(c->pair_cut2)) {
        fullStreamKernel<KIND, true><<<grid, kStreamThreads, 0, c->stream>>>(V, c->P, c->pair_cut2, 0, shard, n_shards,
                                                                             c->partials.ptr);
    }
    else if (c->full_pair_path == 0 && c->n_slots < (1 << 26)) { // finite cutoff: FP32 screening, FP64 candidates
        const double lmax = std::max(V.len[0], std::max(V.len[1], V.len[2]));
        fullScreenKernel<KIND><<<grid, kStreamThreads, 0, c->stream>>>(V, c->P, c->pair_cut2, screeningCutoffFor(c, lmax), shard,
                                                                      n_shards, c->partials.ptr);
    }
    else {
        fullStreamKernel<KIND, false><<<grid, kStreamThreads, 0, c->stream>>>(V, c->P, c->pair_cut2, 0, shard,
                                                                              n_shards, c->partials.ptr);
    }
}

void launchFull(fb_ctx* c, const SlotView& V, int volume_predicate, int shard = 0, int n_shards = 1)
{
    if (!c->P.any_molecular) { // all groups atomic: no mass-centre cutoffs, exclusions or rigid bodies to honour
        const int n_itiles = (c->n_slots + kStreamVariants - 1) / kStreamVariants;
        // rows of 64 particles × every gy-th chunk of the particles j: enough blocks to fill the machine, and — when the
        // rows are dealt over several GPUs — short enough ones (row r streams N − 64 r particles: with one block per row the
        // first rows of a rank alone took 1–1.6 ms of the 0.54 ms an eighth of the pairs should). ≈ 64 blocks per SM: the
        // rows are a triangle, short blocks even it out (S1, one GPU, measured: gy 1 → 2.95 ms, 4 → 2.59, 8 → 2.53)
        static const int gy_forced = [] { // (experiments: FAUNUS_B200_FULL_GY)
            const char* v = std::getenv("FAUNUS_B200_FULL_GY");
            return v != nullptr ? std::atoi(v) : 0;
        }();
        const int gy = gy_forced > 0
                           ? gy_forced
                           : std::max(1, std::min(16, (64 * c->n_sm * n_shards + n_itiles - 1) / std::max(1, n_itiles)));
        const dim3 grid(n_itiles, gy);
        const size_t npart = static_cast<size_t>(n_itiles) * gy;
        c->partials.ensure(std::max<size_t>(npart, 4 * kMaxPartialBlocks));
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        launchFullStream<K>(c, V, grid, shard, n_shards);                                                     \
        break;
        switch (c->P.kind) {
            FB_CASE(POT_COULOMB_LJ)
            FB_CASE(POT_COULOMB_WCA)
            FB_CASE(POT_PM)
            FB_CASE(POT_PMWCA)
            FB_CASE(POT_FUNCTOR)
            FB_CASE(POT_SPLINED)
        default:
            throw CudaError{"unknown potential kind"};
        }
#undef FB_CASE
        launched(c, "fullStreamKernel");
        orderedSumKernel<<<1, 1024, 0, c->stream>>>(c->partials.ptr, npart, 1, c->d_result);
        launched(c, "orderedSumKernel");
        return;
    }
    const int nt = (c->n_slots + kTile - 1) / kTile;
    const size_t npart = static_cast<size_t>(nt) * (nt + 1) / 2;
    c->partials.ensure(std::max<size_t>(npart, 4 * kMaxPartialBlocks));
    dim3 grid(nt, nt);
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        fullEnergyKernel<K><<<grid, kTile, 0, c->stream>>>(V, c->P, volume_predicate, shard, n_shards,        \
                                                           c->partials.ptr);                                  \
        break;
    switch (c->P.kind) {
        FB_CASE(POT_COULOMB_LJ)
        FB_CASE(POT_COULOMB_WCA)
        FB_CASE(POT_PM)
        FB_CASE(POT_PMWCA)
        FB_CASE(POT_FUNCTOR)
        FB_CASE(POT_SPLINED)
    default:
        throw CudaError{"unknown potential kind"};
    }
#undef FB_CASE
    launched(c, "fullEnergyKernel");
    orderedSumKernel<<<1, 1024, 0, c->stream>>>(c->partials.ptr, npart, 1, c->d_result);
    launched(c, "orderedSumKernel");
}

/** Q(k) of the k-cells [cell_begin, cell_end) of a slot from its positions (ewaldFullCellKernel) */
void launchFullQ(fb_ctx* c, int s, int cell_begin, int cell_end, bool store_q, double* e_partials)
{
    Slot& sl = c->slot[s];
    PhaseGeometry geo{};
    geo.ncc = static_cast<int>(std::ceil(c->ewald.n_cutoff));
    geo.table_stride = 0;
    for (int i = 0; i < 3; ++i) {
        geo.len[i] = sl.ewald_box[i];
    }
    static thread_local int configured_device = -1; // (a thread drives one context at a time)
    if (configured_device != c->device) {
        CUDA_CHECK(cudaFuncSetAttribute(ewaldFullCellKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(FullQSmem))));
        configured_device = c->device;
    }
    // a slab of few cells (one of several GPUs): also split the particles, so that the slab still fills the machine.
    // 8192 particles per share whatever the number of GPUs: the same shares, added in the same order, everywhere.
    constexpr int kSplitSize = 8192;
    const int n_cells = cell_end - cell_begin;
    const int n_splits = (c->n_slots + kSplitSize - 1) / kSplitSize;
    if (!store_q && e_partials != nullptr && n_splits > 1 && n_cells < 4 * c->n_sm) {
        c->q_partials.ensure(static_cast<size_t>(n_cells) * n_splits * kTileK);
        ewaldFullCellKernel<<<dim3(n_cells, n_splits), kBlock, sizeof(FullQSmem), c->stream>>>(
            makeView(c, s), makeEwaldView(c, s), sl.kn.ptr, sl.cell_start.ptr, cell_begin, geo, 0, e_partials, kSplitSize,
            c->q_partials.ptr);
        launched(c, "ewaldFullCellKernel");
        ewaldCellEnergyKernel<<<n_cells, kTileK, 0, c->stream>>>(makeEwaldView(c, s), sl.cell_start.ptr, cell_begin, n_splits,
                                                                c->q_partials.ptr, e_partials);
        launched(c, "ewaldCellEnergyKernel");
        return;
    }
    ewaldFullCellKernel<<<n_cells, kBlock, sizeof(FullQSmem), c->stream>>>(
        makeView(c, s), makeEwaldView(c, s), sl.kn.ptr, sl.cell_start.ptr, cell_begin, geo, store_q ? 1 : 0, e_partials);
    launched(c, "ewaldFullCellKernel");
}

/** what buildGemmLayout makes of a list of k-vectors (host side of fb_fullq.cuh) */
struct GemmLayout
{
    std::vector<int4> tiles;            //!< {nx, y table index, z table index of the first column, column groups}
    std::vector<int> row_groups;        //!< per tile, 8 bits per nx: first column group with a k-vector | (last + 1) << 4
    std::vector<int> order;             //!< the tiles heaviest first
    std::vector<int> index;             //!< [K] tile · 2048 + row · 64 + column of every k-vector
    std::vector<int> tile_first_k;      //!< [tiles + 1] first k-vector of the tile's column of cells
    std::vector<int> column_first_tile; //!< [columns + 1]
};

/**
 * Tiles of ewaldFullGemmKernel for the k-vectors `kn` (cell by cell, as stored): 4 nx × 8 ny rows — two y-adjacent columns of
 * cells, a contiguous range of k-vectors — and windows of at most 8 column groups of 8 nz from the first to the last nz
 * that holds a k-vector of the tile. Tiles are numbered in storage order of their k-vectors, so a range of tiles is a
 * range of k-vectors (the slabs of the sharded energy); `order` lists them heaviest first (they start first).
 * Pure host code: fb_debug_fullq_layout hands it to the CPU tests.
 */
GemmLayout buildGemmLayout(const std::vector<int4>& kn, int ncc)
{
    GemmLayout out;
    if (kn.empty()) {
        return out;
    }
    struct Column
    {
        int lo = std::numeric_limits<int>::max(), hi = -1; // z table indices (nz + ncc)
        int first_k = 0, end_k = 0, first_tile = 0;
    };
    std::vector<Column> columns; // (nx / 4, y index / 8), in storage order
    auto key = [&](const int4& n) { return std::make_pair(n.x >> 2, (n.y + ncc) >> 3); };
    for (size_t i = 0; i < kn.size(); ++i) {
        if (i == 0 || key(kn[i]) != key(kn[i - 1])) {
            columns.emplace_back();
            columns.back().first_k = static_cast<int>(i);
        }
        Column& col = columns.back();
        col.lo = std::min(col.lo, kn[i].z + ncc);
        col.hi = std::max(col.hi, kn[i].z + ncc);
        col.end_k = static_cast<int>(i) + 1;
    }
    out.index.resize(kn.size());
    for (Column& col : columns) {
        col.first_tile = static_cast<int>(out.tiles.size());
        const int4 n0 = kn[col.first_k];
        for (int z0 = col.lo; z0 <= col.hi; z0 += kGemmCols) { // (the windows start at the column's first nz, not on a grid)
            // (the k-vectors of ONE tile are contiguous only if the column has a single window; a slab is a range of
            // columns, which always is)
            out.tile_first_k.push_back(col.first_k);
            out.tiles.push_back(make_int4(n0.x & ~3, (n0.y + ncc) & ~7, z0, std::min(8, (col.hi - z0 + 8) / 8)));
        }
        for (int i = col.first_k; i < col.end_k; ++i) {
            const int4 n = kn[i];
            const int iy = n.y + ncc, iz = n.z + ncc;
            const int t = col.first_tile + (iz - col.lo) / kGemmCols;
            out.index[i] = t * (kGemmRows * kGemmCols) + (8 * (n.x & 3) + (iy & 7)) * kGemmCols + (iz - out.tiles[t].z);
        }
    }
    out.tile_first_k.push_back(static_cast<int>(kn.size()));
    { // column groups every nx of a tile needs (the warps of ewaldFullGemmKernel skip the others)
        std::vector<int> lo(4 * out.tiles.size(), 16), hi(4 * out.tiles.size(), 0);
        for (size_t i = 0; i < kn.size(); ++i) {
            const int t = out.index[i] / (kGemmRows * kGemmCols);
            const int row = 4 * t + (kn[i].x & 3);
            const int g = (out.index[i] % kGemmCols) / 8;
            lo[row] = std::min(lo[row], g);
            hi[row] = std::max(hi[row], g + 1);
        }
        out.row_groups.assign(out.tiles.size(), 0);
        for (size_t row = 0; row < lo.size(); ++row) {
            if (hi[row] > 0) {
                out.row_groups[row / 4] |= (lo[row] | (hi[row] << 4)) << (8 * (row & 3));
            }
        }
    }
    // a slab boundary must not fall between the windows of one column: boundaries are moved to the column's first tile
    for (const Column& col : columns) {
        out.column_first_tile.push_back(col.first_tile);
    }
    out.column_first_tile.push_back(static_cast<int>(out.tiles.size()));
    out.order.resize(out.tiles.size());
    for (size_t t = 0; t < out.order.size(); ++t) {
        out.order[t] = static_cast<int>(t);
    }
    std::stable_sort(out.order.begin(), out.order.end(), [&](int a, int b) { return out.tiles[a].w > out.tiles[b].w; });
    return out;
}

void buildGemmTiles(fb_ctx* c, Slot& sl, const std::vector<int4>& kn)
{
    GemmLayout layout = buildGemmLayout(kn, static_cast<int>(std::ceil(c->ewald.n_cutoff)));
    sl.n_gemm_tiles = static_cast<int>(layout.tiles.size());
    sl.gemm_tile_first_k.swap(layout.tile_first_k);
    sl.gemm_column_first_tile.swap(layout.column_first_tile);
    if (sl.n_gemm_tiles == 0) {
        return;
    }
    sl.gemm_tiles.upload(layout.tiles.data(), layout.tiles.size(), c->stream);
    sl.gemm_order.upload(layout.order.data(), layout.order.size(), c->stream);
    sl.gemm_row_groups.upload(layout.row_groups.data(), layout.row_groups.size(), c->stream);
    sl.gemm_index.upload(layout.index.data(), layout.index.size(), c->stream);
    CUDA_CHECK(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
}

/**
 * Q(k) of a slot from its positions as the complex matrix product of fb_fullq.cuh: all tiles → Q(k) (e_partials null), or
 * the tiles [tile_begin, tile_end) of a slab → partial sums of Σ A_k |Q_k|² over its k-vectors, Q not stored; returns the
 * number of partial sums. The particle ranges depend on the number of particles and of ALL tiles only — not on the device,
 * not on the number of slabs: ≈ 16 blocks per SM of a 148-SM part for the full rebuild (measured: 6 → 2.47 ms, 12 → 2.30, 24 → 2.28 at S1), of 8 such
 * parts for the slabs.
 */
int launchFullQGemm(fb_ctx* c, int s, int tile_begin = 0, int tile_end = -1, double* e_partials = nullptr)
{
    Slot& sl = c->slot[s];
    const bool slab = e_partials != nullptr;
    tile_end = tile_end < 0 ? sl.n_gemm_tiles : tile_end;
    PhaseGeometry geo{};
    geo.ncc = static_cast<int>(std::ceil(c->ewald.n_cutoff));
    geo.table_stride = 0;
    for (int i = 0; i < 3; ++i) {
        geo.len[i] = sl.ewald_box[i];
    }
    const int n = std::max(1, c->n_slots);
    static const int blocks_per_sm = [] { // (experiments: FAUNUS_B200_FULLQ_BLOCKS)
        const char* v = std::getenv("FAUNUS_B200_FULLQ_BLOCKS");
        return v != nullptr && std::atoi(v) > 0 ? std::atoi(v) : 16;
    }();
    int n_ranges = std::max(1, ((slab ? 8 : 1) * blocks_per_sm * 148 + sl.n_gemm_tiles - 1) / sl.n_gemm_tiles);
    n_ranges = std::min(n_ranges, std::max(1, n / 256));
    int range_size = (n + n_ranges - 1) / n_ranges;
    range_size = (range_size + kGemmChunk - 1) / kGemmChunk * kGemmChunk;
    n_ranges = (n + range_size - 1) / range_size;
    c->fullq_partials.ensure(static_cast<size_t>(tile_end - tile_begin) * n_ranges * kGemmShare);
    static thread_local int configured_device = -1; // (a thread drives one context at a time)
    if (configured_device != c->device) {
        CUDA_CHECK(cudaFuncSetAttribute(ewaldFullGemmKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(FullGemmSmem))));
        CUDA_CHECK(cudaFuncSetAttribute(ewaldFullGemmKernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(FullGemmSmem))));
        CUDA_CHECK(cudaFuncSetAttribute(ewaldFullGemmKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(sizeof(FullGemmSmem))));
        configured_device = c->device;
    }
    const dim3 grid(tile_end - tile_begin, n_ranges);
    const int* order = slab ? nullptr : sl.gemm_order.ptr;
    c->fullq_steps.ensure(4 * static_cast<size_t>(n));
    double2* steps = c->fullq_steps.ptr;
    double2* zsteps = steps + 3 * static_cast<size_t>(n);
    ewaldStepPhaseKernel<<<(n + 255) / 256, 256, 0, c->stream>>>(makeView(c, s), geo, c->ewald.policy == 1 ? 1 : 0, steps, zsteps);
    launched(c, "ewaldStepPhaseKernel");
    if (c->ewald.policy == 2) { // IPBC: the real product of the cosines
        ewaldFullGemmKernel<2><<<grid, kGemmThreads, sizeof(FullGemmSmem), c->stream>>>(
            c->n_slots, steps, zsteps, sl.gemm_tiles.ptr, sl.gemm_row_groups.ptr, order, tile_begin, geo, range_size,
            c->fullq_partials.ptr);
    }
    else if (c->ewald.policy == 1) {
        ewaldFullGemmKernel<1><<<grid, kGemmThreads, sizeof(FullGemmSmem), c->stream>>>(
            c->n_slots, steps, zsteps, sl.gemm_tiles.ptr, sl.gemm_row_groups.ptr, order, tile_begin, geo, range_size,
            c->fullq_partials.ptr);
    }
    else {
        ewaldFullGemmKernel<0><<<grid, kGemmThreads, sizeof(FullGemmSmem), c->stream>>>(
            c->n_slots, steps, zsteps, sl.gemm_tiles.ptr, sl.gemm_row_groups.ptr, order, tile_begin, geo, range_size,
            c->fullq_partials.ptr);
    }
    launched(c, "ewaldFullGemmKernel");
    if (!slab) {
        ewaldFullGatherKernel<<<(sl.K + 255) / 256, 256, 0, c->stream>>>(makeEwaldView(c, s), sl.gemm_index.ptr, n_ranges,
                                                                        c->fullq_partials.ptr);
        launched(c, "ewaldFullGatherKernel");
        return 0;
    }
    const int k_begin = sl.gemm_tile_first_k[tile_begin], k_end = sl.gemm_tile_first_k[tile_end];
    const int blocks = (k_end - k_begin + 255) / 256;
    ewaldFullGatherEnergyKernel<<<blocks, 256, 0, c->stream>>>(makeEwaldView(c, s), sl.gemm_index.ptr, tile_begin, n_ranges,
                                                              c->fullq_partials.ptr, k_begin, k_end, e_partials);
    launched(c, "ewaldFullGatherEnergyKernel");
    return blocks;
}

/** PolicyIonIon::updateBox / PolicyIonIonIPBC::updateBox, src/energy.cpp:133-186, 356-412 */
void generateKVectors(const fb_ewald_config& cfg, const double box[3], std::vector<double4>& kA,
                      std::vector<int4>& kn)
{
    kA.clear();
    kn.clear();
    const bool ipbc = cfg.policy == 2;
    const int ncc = static_cast<int>(std::ceil(cfg.n_cutoff));
    const double pi = 3.141592653589793238462643383279502884;
    const double lmax = std::max(box[0], std::max(box[1], box[2]));
    const double check_k2_zero = 0.1 * std::pow(2 * pi / lmax, 2);
    const long k_vector_size = static_cast<long>(2 * ncc + 1) * (2 * ncc + 1) * (2 * ncc + 1) - 1;
    if (k_vector_size == 0) {
        kA.push_back(make_double4(1, 0, 0, 0));
        kn.push_back(make_int4(0, 0, 0, 0));
        return;
    }
    const double nc2 = cfg.n_cutoff * cfg.n_cutoff;
    const double kappa2 = cfg.kappa * cfg.kappa;
    const int start_value = ipbc ? 0 : 1;
    for (int nx = 0; nx <= ncc; nx++) {
        const double dnx2 = static_cast<double>(nx * nx);
        const double xfactor = (nx > 0) ? 2.0 : 1.0;
        for (int ny = -ncc * start_value; ny <= ncc; ny++) {
            const double dny2 = static_cast<double>(ny * ny);
            const double yfactor = (ny > 0) ? 2.0 : 1.0;
            for (int nz = -ncc * start_value; nz <= ncc; nz++) {
                double factor = xfactor;
                if (ipbc) {
                    factor = xfactor * yfactor;
                    if (nz > 0) {
                        factor *= 2;
                    }
                }
                const double kx = 2 * pi * nx / box[0];
                const double ky = 2 * pi * ny / box[1];
                const double kz = 2 * pi * nz / box[2];
                const double k2 = kx * kx + ky * ky + kz * kz + kappa2;
                if (k2 < check_k2_zero) {
                    continue;
                }
                if (cfg.spherical_sum) {
                    const double dnz2 = static_cast<double>(nz * nz);
                    if ((dnx2 + dny2 + dnz2) / nc2 > 1) {
                        continue;
                    }
                }
                kA.push_back(make_double4(kx, ky, kz, factor * std::exp(-k2 / (4 * cfg.alpha * cfg.alpha)) / k2));
                kn.push_back(make_int4(nx, ny, nz, 0));
            }
        }
    }
}

template <class T> void copyTable(DeviceBuffer<T>& dst, const T* src, size_t n, cudaStream_t s, const T** out)
{
    if (src == nullptr) {
        *out = nullptr;
        return;
    }
    dst.upload(src, n, s);
    *out = dst.ptr;
}

} // namespace

// =================================================================================================
// life cycle
// =================================================================================================
FB_API int fb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

FB_API const char* fb_last_error(const fb_ctx* ctx)
{
    return ctx ? ctx->last_error.c_str() : g_create_error.c_str();
}

FB_API int fb_create(const fb_config* cfg, fb_ctx** out)
{
    if (cfg == nullptr || out == nullptr) {
        g_create_error = "null argument";
        return FB_ERR_INVALID;
    }
    *out = nullptr;
    fb_ctx* c = nullptr;
    try {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            g_create_error = "no CUDA device: libfaunus_b200 has no CPU fallback";
            return FB_ERR_CUDA;
        }
        if (cfg->device < 0 || cfg->device >= ndev) {
            g_create_error = "device ordinal out of range";
            return FB_ERR_INVALID;
        }
        if (cfg->n_atom_types <= 0 || cfg->n_molecule_types <= 0) {
            g_create_error = "need at least one atom type and one molecule type";
            return FB_ERR_INVALID;
        }
        c = new fb_ctx();
        c->device = cfg->device;
        c->batch.gap_stats = std::getenv("FAUNUS_B200_GAP_STATS") != nullptr;
        if (const char* v = std::getenv("FAUNUS_B200_FULLPAIR")) {
            c->full_pair_path = std::strcmp(v, "fp64") == 0 ? 1 : 0;
        }
        if (const char* v = std::getenv("FAUNUS_B200_FULLQ")) {
            c->full_q_path = std::strcmp(v, "cells") == 0 ? 1 : 0;
        }
        CUDA_CHECK(cudaSetDevice(c->device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&c->ev0));
        CUDA_CHECK(cudaEventCreate(&c->ev1));
        cudaDeviceProp prop{};
        CUDA_CHECK(cudaGetDeviceProperties(&prop, c->device));
        c->max_blocks = std::min(prop.multiProcessorCount * 4, kMaxPartialBlocks);
        c->n_sm = prop.multiProcessorCount;
        for (auto& e : c->batch.ev) {
            CUDA_CHECK(cudaEventCreate(&e));
        }
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->batch.pair_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->batch.ev_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->batch.ev_join, cudaEventDisableTiming));
        for (auto& r : c->batch.run) {
            CUDA_CHECK(cudaEventCreate(&r.ev_begin));
            CUDA_CHECK(cudaEventCreate(&r.ev_end));
            CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_done, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_in, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_started, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&r.ev_tail, cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->batch.run_in_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->batch.run_out_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaHostAlloc(&c->h_result, 8 * sizeof(double), cudaHostAllocMapped));
        CUDA_CHECK(cudaHostGetDevicePointer(&c->d_result, c->h_result, 0));
        c->partials.alloc(4 * kMaxPartialBlocks);
        c->ticket.alloc(1);
        CUDA_CHECK(cudaMemsetAsync(c->ticket.ptr, 0, sizeof(unsigned), c->stream));

        for (int i = 0; i < 3; ++i) {
            c->periodic[i] = cfg->periodic[i];
            c->slot[0].box[i] = c->slot[1].box[i] = cfg->box[i];
        }
        const size_t T2 = static_cast<size_t>(cfg->n_atom_types) * cfg->n_atom_types;
        const size_t M = static_cast<size_t>(cfg->n_molecule_types);
        PotParams& P = c->P;
        P.kind = cfg->kind;
        P.n_types = cfg->n_atom_types;
        P.n_mol = cfg->n_molecule_types;
        cudaStream_t s = c->stream;
        copyTable(c->d_flags, reinterpret_cast<const unsigned*>(cfg->pair_flags), T2, s, &P.flags);
        copyTable(c->d_lj_s2, cfg->lj_sigma2, T2, s, &P.lj_s2);
        copyTable(c->d_lj_e4, cfg->lj_eps4, T2, s, &P.lj_e4);
        copyTable(c->d_wca_s2, cfg->wca_sigma2, T2, s, &P.wca_s2);
        copyTable(c->d_wca_e4, cfg->wca_eps4, T2, s, &P.wca_e4);
        copyTable(c->d_hs_s2, cfg->hs_sigma2, T2, s, &P.hs_s2);
        P.lB = cfg->coulomb_bjerrum_length;
        P.Rc = cfg->coulomb_cutoff;
        P.invRc = cfg->coulomb_cutoff > 0 ? 1.0 / cfg->coulomb_cutoff : 0.0;
        P.kappa = cfg->coulomb_kappa;
        P.lB_plain = cfg->plain_bjerrum_length;
        P.nk = cfg->coulomb_n_knots;
        const bool needs_spline_coulomb =
            cfg->kind == FB_POT_COULOMB_LJ || cfg->kind == FB_POT_COULOMB_WCA ||
            (cfg->pair_flags && std::any_of(cfg->pair_flags, cfg->pair_flags + T2,
                                            [](uint32_t f) { return f & FB_TERM_COULOMB_SPLINED; }));
        if (needs_spline_coulomb) {
            if (cfg->coulomb_n_knots < 2 || !cfg->coulomb_knots || !cfg->coulomb_coeffs) {
                throw CudaError{"splined Coulomb requires the S(q) table"};
            }
            const int nk = cfg->coulomb_n_knots;
            copyTable(c->d_knots, cfg->coulomb_knots, static_cast<size_t>(nk), s, &P.knots);
            // pad the coefficient table with one block so that index nk-1 is addressable
            std::vector<double> coef(cfg->coulomb_coeffs, cfg->coulomb_coeffs + 6 * (nk - 1));
            coef.resize(6 * nk, 0.0);
            c->d_coef.uploadVector(coef, s);
            P.coef = c->d_coef.ptr;
            // bucket start index: (#knots < b/nlut) − 1, clamped to 0
            const int nlut = 256;
            std::vector<int> lut(nlut);
            for (int b = 0; b < nlut; ++b) {
                const double x = static_cast<double>(b) / nlut;
                int cnt = 0;
                while (cnt < nk && cfg->coulomb_knots[cnt] < x) {
                    ++cnt;
                }
                lut[b] = std::max(0, cnt - 1);
            }
            c->d_lut.uploadVector(lut, s);
            P.lut = c->d_lut.ptr;
            P.nlut = nlut;
        }
        const bool functor_like = cfg->kind == FB_POT_FUNCTOR || cfg->kind == FB_POT_SPLINED;
        if (functor_like && cfg->pair_flags == nullptr) {
            throw CudaError{"functor / splined potentials require pair_flags"};
        }
        auto need = [&](const void* p, const char* what) {
            if (p == nullptr) {
                throw CudaError{std::string("missing table: ") + what};
            }
        };
        if (cfg->kind == FB_POT_COULOMB_LJ) {
            need(cfg->lj_sigma2, "lj_sigma2");
            need(cfg->lj_eps4, "lj_eps4");
        }
        if (cfg->kind == FB_POT_COULOMB_WCA || cfg->kind == FB_POT_PMWCA) {
            need(cfg->wca_sigma2, "wca_sigma2");
            need(cfg->wca_eps4, "wca_eps4");
        }
        if (cfg->kind == FB_POT_PM) {
            need(cfg->hs_sigma2, "hs_sigma2");
        }
        if (functor_like) {
            for (size_t t = 0; t < T2; ++t) {
                const uint32_t f = cfg->pair_flags[t];
                if ((f & FB_TERM_LJ) && (!cfg->lj_sigma2 || !cfg->lj_eps4)) {
                    throw CudaError{"pair_flags request LJ without tables"};
                }
                if ((f & FB_TERM_WCA) && (!cfg->wca_sigma2 || !cfg->wca_eps4)) {
                    throw CudaError{"pair_flags request WCA without tables"};
                }
                if ((f & FB_TERM_HARDSPHERE) && !cfg->hs_sigma2) {
                    throw CudaError{"pair_flags request hard spheres without table"};
                }
            }
        }
        if (cfg->kind == FB_POT_SPLINED) {
            need(cfg->spline_offset, "spline_offset");
            need(cfg->spline_knots, "spline_knots");
            need(cfg->spline_coeffs, "spline_coeffs");
            const int total = cfg->spline_offset[T2];
            copyTable(c->d_sp_offset, cfg->spline_offset, T2 + 1, s, &P.sp_offset);
            copyTable(c->d_sp_knots, cfg->spline_knots, static_cast<size_t>(total), s, &P.sp_knots);
            // re-block coefficients per knot index: interval k of pair p lives at 6*(offset[p]+k)
            std::vector<double> coef(6 * static_cast<size_t>(total), 0.0);
            size_t src = 0;
            for (size_t p = 0; p < T2; ++p) {
                const int first = cfg->spline_offset[p];
                const int nk = cfg->spline_offset[p + 1] - first;
                for (int k = 0; k + 1 < nk; ++k) {
                    for (int i = 0; i < 6; ++i) {
                        coef[6 * (static_cast<size_t>(first) + k) + i] = cfg->spline_coeffs[src++];
                    }
                }
            }
            c->d_sp_coef.uploadVector(coef, s);
            P.sp_coef = c->d_sp_coef.ptr;
            copyTable(c->d_sp_rmin2, cfg->spline_rmin2, T2, s, &P.sp_rmin2);
            copyTable(c->d_sp_rmax2, cfg->spline_rmax2, T2, s, &P.sp_rmax2);
            copyTable(c->d_sp_hs, cfg->spline_hardsphere, T2, s, &P.sp_hs);
        }
        { // r² beyond which every pair energy is exactly zero (pre-filter of the windowed pair kernel)
            const double inf = std::numeric_limits<double>::infinity();
            const double wca_factor = 1.2599210498948732;
            auto table_max = [&](const double* t, double scale) {
                double m = 0.0;
                for (size_t i = 0; i < T2; ++i) {
                    m = std::max(m, t[i] * scale);
                }
                return m;
            };
            const double rc2 = cfg->coulomb_cutoff * cfg->coulomb_cutoff;
            double cut = 0.0;
            switch (cfg->kind) {
            case FB_POT_COULOMB_WCA:
                cut = std::max(rc2, table_max(cfg->wca_sigma2, wca_factor));
                break;
            case FB_POT_FUNCTOR:
                for (size_t t = 0; t < T2; ++t) {
                    const uint32_t f = cfg->pair_flags[t];
                    if (f & (FB_TERM_LJ | FB_TERM_COULOMB_PLAIN)) {
                        cut = inf;
                    }
                    if (f & FB_TERM_COULOMB_SPLINED) {
                        cut = std::max(cut, rc2);
                    }
                    if (f & FB_TERM_WCA) {
                        cut = std::max(cut, cfg->wca_sigma2[t] * wca_factor);
                    }
                    if (f & FB_TERM_HARDSPHERE) {
                        cut = std::max(cut, cfg->hs_sigma2[t]);
                    }
                }
                break;
            case FB_POT_SPLINED:
                cut = table_max(cfg->spline_rmax2, 1.0);
                break;
            default: // Lennard-Jones and plain Coulomb have no cutoff
                cut = inf;
            }
            c->pair_cut2 = cut * (1.0 + 1e-9);
        }
        // molecules
        need(cfg->molecule_flags, "molecule_flags");
        c->molecule_flags.assign(cfg->molecule_flags, cfg->molecule_flags + M);
        std::vector<double> g2g(M * M, DBL_MAX);
        if (cfg->g2g_cutoff_squared) {
            g2g.assign(cfg->g2g_cutoff_squared, cfg->g2g_cutoff_squared + M * M);
        }
        c->d_g2g.uploadVector(g2g, s);
        P.g2g_cut2 = c->d_g2g.ptr;
        std::vector<int> excl_offset(M, -1), excl_natoms(M, 0);
        std::vector<unsigned char> excl;
        for (size_t m = 0; m < M; ++m) {
            const int n = cfg->molecule_natoms ? cfg->molecule_natoms[m] : 0;
            excl_natoms[m] = n;
            if (cfg->exclusions && cfg->exclusions[m] && n > 0) {
                excl_offset[m] = static_cast<int>(excl.size());
                excl.insert(excl.end(), cfg->exclusions[m], cfg->exclusions[m] + static_cast<size_t>(n) * n);
            }
        }
        if (excl.empty()) {
            excl.push_back(0);
        }
        c->d_excl_offset.uploadVector(excl_offset, s);
        c->d_excl_natoms.uploadVector(excl_natoms, s);
        c->d_excl.uploadVector(excl, s);
        P.excl_offset = c->d_excl_offset.ptr;
        P.excl_natoms = c->d_excl_natoms.ptr;
        P.excl = c->d_excl.ptr;
        P.any_molecular = 1; // refined at upload time
        CUDA_CHECK(cudaStreamSynchronize(s));
        *out = c;
        return FB_OK;
    }
    catch (const CudaError& e) {
        g_create_error = e.msg;
        delete c;
        return FB_ERR_CUDA;
    }
    catch (const std::exception& e) {
        g_create_error = e.what();
        delete c;
        return FB_ERR_INVALID;
    }
}

FB_API void fb_destroy(fb_ctx* c)
{
    if (c == nullptr) {
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
    }
    fb_nccl_finalize(c);
    if (c->h_result) {
        cudaFreeHost(c->h_result);
    }
    if (c->ev0) {
        cudaEventDestroy(c->ev0);
    }
    if (c->ev1) {
        cudaEventDestroy(c->ev1);
    }
    for (auto& e : c->batch.ev) {
        if (e) {
            cudaEventDestroy(e);
        }
    }
    if (c->batch.ev_fork) {
        cudaEventDestroy(c->batch.ev_fork);
    }
    if (c->batch.ev_join) {
        cudaEventDestroy(c->batch.ev_join);
    }
    for (auto& r : c->batch.run) {
        for (cudaEvent_t e : {r.ev_begin, r.ev_end, r.ev_done}) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
    }
    if (c->batch.gap_stats && c->batch.gap_count > 0) {
        std::fprintf(stderr, "faunus_b200: %ld chained runs, %.2f us between the last window of a run and the first of the next\n",
                     c->batch.gap_count, 1e3 * c->batch.gap_ms_total / static_cast<double>(c->batch.gap_count));
    }
    for (cudaStream_t extra : {c->batch.run_in_stream, c->batch.run_out_stream}) {
        if (extra) {
            cudaStreamSynchronize(extra);
            cudaStreamDestroy(extra);
        }
    }
    for (auto& r : c->batch.run) {
        for (cudaEvent_t e : {r.ev_in, r.ev_started, r.ev_tail}) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
    }
    if (c->batch.pair_stream) {
        cudaStreamSynchronize(c->batch.pair_stream);
        for (auto& [key, g] : c->batch.run_graphs) {
            if (g.exec) {
                cudaGraphExecDestroy(g.exec);
            }
        }
        cudaStreamDestroy(c->batch.pair_stream);
    }
    cudaStream_t s = c->stream;
    delete c;
    if (s) {
        cudaStreamDestroy(s);
    }
}

FB_API unsigned long long fb_launch_count(const fb_ctx* c)
{
    return c ? c->launches : 0;
}

FB_API void* fb_stream(const fb_ctx* c)
{
    return c ? static_cast<void*>(c->stream) : nullptr;
}

FB_API int fb_enable_timing(fb_ctx* c, int on)
{
    if (!c) {
        return FB_ERR_INVALID;
    }
    c->timing = on != 0;
    return FB_OK;
}

FB_API double fb_last_kernel_ms(const fb_ctx* c)
{
    return c ? c->last_ms : 0.0;
}

FB_API int fb_get_timing(const fb_ctx* c, double out[8])
{
    if (!c || !out) {
        return FB_ERR_INVALID;
    }
    for (int i = 0; i < 4; ++i) {
        out[2 * i] = c->acc_ms[i];
        out[2 * i + 1] = c->acc_launches[i];
    }
    return FB_OK;
}

FB_API int fb_measure_fp64_peak(int device, double* tflops)
{
    if (!tflops || cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return FB_ERR_CUDA;
    }
    cudaDeviceProp prop{};
    double* d = nullptr;
    cudaEvent_t e0, e1;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaMalloc(&d, 64) != cudaSuccess) {
        cudaGetLastError();
        return FB_ERR_CUDA;
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8;
    const int iters = 200000;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfmaPeakKernel<<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * iters * 256.0 * blocks;
        if (rep > 0 && ms > 0) {
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_CUDA;
}

// =================================================================================================
// Space mirror
// =================================================================================================
FB_API int fb_upload_space(fb_ctx* c, int s, const double* xyzq, const int* atom_id, const fb_group* groups,
                           int n_particles, int n_groups)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s, false);
        if (n_particles <= 0 || n_groups <= 0 || !xyzq || !atom_id || !groups) {
            throw CudaError{"empty space"};
        }
        if (c->n_slots != 0 && (c->n_slots != n_particles || c->n_groups != n_groups)) {
            throw CudaError{"particle / group count differs from the first upload"};
        }
        int expect = 0;
        bool any_molecular = false;
        for (int g = 0; g < n_groups; ++g) {
            if (groups[g].begin != expect || groups[g].size > groups[g].capacity || groups[g].size < 0) {
                throw CudaError{"groups must tile the particle array"};
            }
            if (groups[g].molid < 0 || groups[g].molid >= c->P.n_mol) {
                throw CudaError{"molecule id out of range"};
            }
            expect += groups[g].capacity;
            any_molecular |= !(c->molecule_flags[groups[g].molid] & FB_MOL_ATOMIC);
        }
        if (expect != n_particles) {
            throw CudaError{"groups must tile the particle array"};
        }
        for (int i = 0; i < n_particles; ++i) {
            if (atom_id[i] < 0 || atom_id[i] >= c->P.n_types) {
                throw CudaError{"atom id out of range"};
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->n_slots = n_particles;
        c->n_groups = n_groups;
        c->P.any_molecular = any_molecular ? 1 : 0;
        Slot& sl = c->slot[s];
        sl.groups.assign(groups, groups + n_groups);
        std::vector<int> gbegin(n_groups), gcap(n_groups), ginfo(n_groups), gsize(n_groups);
        std::vector<double4> gcm(n_groups);
        for (int g = 0; g < n_groups; ++g) {
            gbegin[g] = groups[g].begin;
            gcap[g] = groups[g].capacity;
            gsize[g] = groups[g].size;
            ginfo[g] = (groups[g].molid << 8) | c->molecule_flags[groups[g].molid];
            gcm[g] = make_double4(groups[g].cm[0], groups[g].cm[1], groups[g].cm[2], 0.0);
        }
        c->gbegin.uploadVector(gbegin, c->stream);
        c->gcap.uploadVector(gcap, c->stream);
        c->ginfo.uploadVector(ginfo, c->stream);
        sl.gsize.uploadVector(gsize, c->stream);
        sl.gcm.uploadVector(gcm, c->stream);
        sl.posq.upload(reinterpret_cast<const double4*>(xyzq), n_particles, c->stream);
        sl.atom_id.upload(atom_id, n_particles, c->stream);
        sl.gid.ensure(n_particles);
        sl.uploaded = true;
        buildGidKernel<<<n_groups, 128, 0, c->stream>>>(makeView(c, s));
        launched(c, "buildGidKernel");
        CUDA_CHECK(cudaStreamSynchronize(c->stream)); // host vectors above go out of scope
        sl.rec_valid = false;
    });
}

FB_API int fb_upload_groups(fb_ctx* c, int s, const fb_group* groups, int n_groups)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!groups || n_groups != c->n_groups) {
            throw CudaError{"group count differs from the uploaded space"};
        }
        Slot& sl = c->slot[s];
        std::vector<int> gsize(n_groups);
        std::vector<double4> gcm(n_groups);
        for (int g = 0; g < n_groups; ++g) {
            if (groups[g].begin != sl.groups[g].begin || groups[g].capacity != sl.groups[g].capacity ||
                groups[g].molid != sl.groups[g].molid || groups[g].size < 0 || groups[g].size > groups[g].capacity) {
                throw CudaError{"group records must keep the layout of the uploaded space"};
            }
            gsize[g] = groups[g].size;
            gcm[g] = make_double4(groups[g].cm[0], groups[g].cm[1], groups[g].cm[2], 0.0);
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        sl.groups.assign(groups, groups + n_groups);
        sl.gsize.uploadVector(gsize, c->stream);
        sl.gcm.uploadVector(gcm, c->stream);
        buildGidKernel<<<n_groups, 128, 0, c->stream>>>(makeView(c, s));
        launched(c, "buildGidKernel");
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        sl.rec_valid = false;
        c->batch.cells_valid = false;
    });
}

FB_API int fb_update_group(fb_ctx* c, int s, int group_index, const fb_group* record, int n_atoms,
                           const int* rel_index, const double* xyzq, const int* atom_id)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (group_index < 0 || group_index >= c->n_groups || !record) {
            throw CudaError{"group index out of range"};
        }
        Slot& sl = c->slot[s];
        fb_group& g = sl.groups[group_index];
        if (record->begin != g.begin || record->capacity != g.capacity || record->molid != g.molid ||
            record->size < 0 || record->size > g.capacity) {
            throw CudaError{"group record does not match the uploaded layout"};
        }
        if (n_atoms < 0 || n_atoms > g.capacity || (n_atoms > 0 && (!xyzq || !atom_id))) {
            throw CudaError{"bad particle update"};
        }
        g = *record;
        InlineUpdate upd{};
        int n_staged = 0;
        auto slot_of = [&](int i) {
            const int rel = rel_index ? rel_index[i] : i;
            if (rel < 0 || rel >= g.capacity) {
                throw CudaError{"relative atom index out of range"};
            }
            return g.begin + rel;
        };
        if (n_atoms <= kInlineUpdate) {
            upd.n = n_atoms;
            for (int i = 0; i < n_atoms; ++i) {
                upd.slot[i] = slot_of(i);
                upd.id[i] = atom_id[i];
                upd.posq[i] = make_double4(xyzq[4 * i], xyzq[4 * i + 1], xyzq[4 * i + 2], xyzq[4 * i + 3]);
            }
        }
        else {
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            n_staged = n_atoms;
            c->h_upd_slot.ensure(n_atoms);
            c->h_upd_id.ensure(n_atoms);
            c->h_upd_posq.ensure(n_atoms);
            for (int i = 0; i < n_atoms; ++i) {
                c->h_upd_slot.ptr[i] = slot_of(i);
                c->h_upd_id.ptr[i] = atom_id[i];
                c->h_upd_posq.ptr[i] =
                    make_double4(xyzq[4 * i], xyzq[4 * i + 1], xyzq[4 * i + 2], xyzq[4 * i + 3]);
            }
            c->d_upd_slot.upload(c->h_upd_slot.ptr, n_atoms, c->stream);
            c->d_upd_id.upload(c->h_upd_id.ptr, n_atoms, c->stream);
            c->d_upd_posq.upload(c->h_upd_posq.ptr, n_atoms, c->stream);
        }
        const int work = std::max(g.capacity, n_atoms);
        const int grid = std::max(1, std::min((work + 255) / 256, 64));
        updateGroupKernel<<<grid, 256, 0, c->stream>>>(
            makeView(c, s), group_index, g.begin, g.capacity, g.size,
            make_double4(g.cm[0], g.cm[1], g.cm[2], 0.0), upd, n_staged, c->d_upd_slot.ptr, c->d_upd_id.ptr,
            c->d_upd_posq.ptr);
        launched(c, "updateGroupKernel");
        sl.rec_valid = sl.rec_valid; // Q is unaffected until fb_ewald_update_*
    });
}

FB_API int fb_set_box(fb_ctx* c, int s, const double box[3])
{
    return guarded(c, [&] {
        checkSlot(c, s, false);
        for (int i = 0; i < 3; ++i) {
            c->slot[s].box[i] = box[i];
        }
    });
}

FB_API int fb_sync(fb_ctx* c, int dst, int src, const fb_change* change)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, dst, false);
        checkSlot(c, src);
        if (dst == src || !change) {
            throw CudaError{"bad sync arguments"};
        }
        Slot& d = c->slot[dst];
        Slot& sr = c->slot[src];
        if (change->everything || change->volume_change) {
            for (int i = 0; i < 3; ++i) {
                d.box[i] = sr.box[i];
            }
        }
        if (change->everything || !d.uploaded) {
            const size_t n = c->n_slots;
            const size_t g = c->n_groups;
            d.posq.ensure(n);
            d.atom_id.ensure(n);
            d.gid.ensure(n);
            d.gcm.ensure(g);
            d.gsize.ensure(g);
            CUDA_CHECK(cudaMemcpyAsync(d.posq.ptr, sr.posq.ptr, n * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(d.atom_id.ptr, sr.atom_id.ptr, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(d.gid.ptr, sr.gid.ptr, n * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(d.gcm.ptr, sr.gcm.ptr, g * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(d.gsize.ptr, sr.gsize.ptr, g * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
            d.groups = sr.groups;
            d.uploaded = true;
            return;
        }
        if (change->n_groups == 0) {
            return;
        }
        if (change->n_groups > kMaxMovedGroups) {
            throw CudaError{"too many changed groups in one Change (max 16)"};
        }
        SyncDesc sd{};
        sd.n_groups = change->n_groups;
        std::vector<int> slots;
        for (int t = 0; t < change->n_groups; ++t) {
            const fb_group_change& gc = change->groups[t];
            if (gc.group_index < 0 || gc.group_index >= c->n_groups) {
                throw CudaError{"group index out of range"};
            }
            sd.group[t] = gc.group_index;
            sd.whole[t] = gc.all ? 1 : 0;
            const fb_group& g = sr.groups[gc.group_index];
            if (!gc.all) {
                for (int i = 0; i < gc.n_atoms; ++i) {
                    if (gc.atoms[i] < 0 || gc.atoms[i] >= g.capacity) {
                        throw CudaError{"relative atom index out of range"};
                    }
                    slots.push_back(g.begin + gc.atoms[i]);
                }
            }
            d.groups[gc.group_index] = g;
        }
        sd.n_slots = static_cast<int>(slots.size());
        int work = sd.n_slots;
        for (int t = 0; t < sd.n_groups; ++t) {
            work = std::max(work, sr.groups[sd.group[t]].capacity);
        }
        if (sd.n_slots <= kInlineMoved) {
            sd.slots = nullptr;
            std::copy(slots.begin(), slots.end(), sd.inline_slots);
        }
        else {
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            c->h_list.ensure(slots.size());
            std::copy(slots.begin(), slots.end(), c->h_list.ptr);
            c->d_list.upload(c->h_list.ptr, slots.size(), c->stream);
            sd.slots = c->d_list.ptr;
        }
        const int grid = std::max(1, std::min((work + 255) / 256, 64));
        syncGroupsKernel<<<grid, 256, 0, c->stream>>>(makeView(c, dst), makeView(c, src), sd);
        launched(c, "syncGroupsKernel");
    });
}

FB_API int fb_download_space(fb_ctx* c, int s, double* xyzq, int* atom_id, fb_group* groups)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        Slot& sl = c->slot[s];
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (xyzq) {
            CUDA_CHECK(cudaMemcpy(xyzq, sl.posq.ptr, c->n_slots * sizeof(double4), cudaMemcpyDeviceToHost));
        }
        if (atom_id) {
            CUDA_CHECK(cudaMemcpy(atom_id, sl.atom_id.ptr, c->n_slots * sizeof(int), cudaMemcpyDeviceToHost));
        }
        if (groups) {
            std::vector<double4> gcm(c->n_groups);
            std::vector<int> gsize(c->n_groups);
            CUDA_CHECK(cudaMemcpy(gcm.data(), sl.gcm.ptr, c->n_groups * sizeof(double4), cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(gsize.data(), sl.gsize.ptr, c->n_groups * sizeof(int), cudaMemcpyDeviceToHost));
            for (int g = 0; g < c->n_groups; ++g) {
                groups[g] = sl.groups[g];
                groups[g].size = gsize[g];
                groups[g].cm[0] = gcm[g].x;
                groups[g].cm[1] = gcm[g].y;
                groups[g].cm[2] = gcm[g].z;
            }
        }
    });
}

// =================================================================================================
// non-bonded energy
// =================================================================================================
FB_API int fb_nonbonded_energy(fb_ctx* c, int s, const fb_change* change, double* energy)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!change || !energy) {
            throw CudaError{"null argument"};
        }
        if (change->everything || change->volume_change) {
            beginTiming(c, TIME_FULL);
            launchFull(c, makeView(c, s), (!change->everything && change->volume_change) ? 1 : 0);
        }
        else {
            if (change->n_groups == 0) {
                *energy = 0.0;
                return;
            }
            const MovedDesc md = change->matter_change ? lowerMatterChange(c, s, change) : lowerChange(c, s, s, change, false);
            if (md.n_moved == 0) {
                *energy = 0.0;
                return;
            }
            const SlotView V = makeView(c, s);
            beginTiming(c, TIME_PAIR);
            launchMoved<false>(c, V, V, md);
        }
        finish(c);
        *energy = c->h_result[0];
    });
}

FB_API int fb_nonbonded_delta(fb_ctx* c, int s_new, int s_old, const fb_change* change, double* u_new,
                              double* u_old)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s_new);
        checkSlot(c, s_old);
        if (!change || !u_new || !u_old || s_new == s_old) {
            throw CudaError{"bad arguments"};
        }
        if (change->everything || change->volume_change) {
            const int pred = (!change->everything && change->volume_change) ? 1 : 0;
            beginTiming(c, TIME_FULL);
            launchFull(c, makeView(c, s_new), pred);
            finish(c);
            *u_new = c->h_result[0];
            beginTiming(c, TIME_FULL);
            launchFull(c, makeView(c, s_old), pred);
            finish(c);
            *u_old = c->h_result[0];
            return;
        }
        if (change->n_groups == 0) {
            *u_new = *u_old = 0.0;
            return;
        }
        if (change->matter_change) { // the active sets of the two states differ: one pass per state
            double* out[2] = {u_new, u_old};
            const int slots[2] = {s_new, s_old};
            for (int k = 0; k < 2; ++k) {
                const MovedDesc one = lowerMatterChange(c, slots[k], change);
                *out[k] = 0.0;
                if (one.n_moved > 0) {
                    const SlotView V = makeView(c, slots[k]);
                    beginTiming(c, TIME_PAIR);
                    launchMoved<false>(c, V, V, one);
                    finish(c);
                    *out[k] = c->h_result[0];
                }
            }
            return;
        }
        const MovedDesc md = lowerChange(c, s_new, s_old, change, false);
        if (md.n_moved == 0) {
            *u_new = *u_old = 0.0;
            return;
        }
        beginTiming(c, TIME_PAIR);
        launchMoved<true>(c, makeView(c, s_new), makeView(c, s_old), md);
        finish(c);
        *u_new = c->h_result[0];
        *u_old = c->h_result[1];
    });
}

FB_API int fb_particle_pair_energy(fb_ctx* c, int s, int n, const double* a_xyzq, const int* a_id, const double* b_xyzq,
                                   const int* b_id, double* energy)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (n <= 0 || !a_xyzq || !a_id || !b_xyzq || !b_id || !energy) {
            throw CudaError{"bad arguments"};
        }
        for (int i = 0; i < n; ++i) {
            if (a_id[i] < 0 || a_id[i] >= c->P.n_types || b_id[i] < 0 || b_id[i] >= c->P.n_types) {
                throw CudaError{"atom id out of range"};
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->d_ghost.ensure(2 * static_cast<size_t>(n));
        c->d_ghost_id.ensure(2 * static_cast<size_t>(n));
        c->d_widom_du.ensure(static_cast<size_t>(n));
        c->h_widom_du.ensure(static_cast<size_t>(n));
        CUDA_CHECK(cudaMemcpyAsync(c->d_ghost.ptr, a_xyzq, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->d_ghost.ptr + n, b_xyzq, n * sizeof(double4), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->d_ghost_id.ptr, a_id, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->d_ghost_id.ptr + n, b_id, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        const SlotView V = makeView(c, s);
#define FB_CASE(K)                                                                                              \
    case K:                                                                                                     \
        particlePairKernel<K><<<gridFor(c, n, 128), 128, 0, c->stream>>>(V, c->P, n, c->d_ghost.ptr, c->d_ghost_id.ptr, \
                                                                         c->d_ghost.ptr + n, c->d_ghost_id.ptr + n,     \
                                                                         c->d_widom_du.ptr);                            \
        break;
        switch (c->P.kind) {
            FB_CASE(POT_COULOMB_LJ)
            FB_CASE(POT_COULOMB_WCA)
            FB_CASE(POT_PM)
            FB_CASE(POT_PMWCA)
            FB_CASE(POT_FUNCTOR)
            FB_CASE(POT_SPLINED)
        default:
            throw CudaError{"unknown potential kind"};
        }
#undef FB_CASE
        launched(c, "particlePairKernel");
        CUDA_CHECK(cudaMemcpyAsync(c->h_widom_du.ptr, c->d_widom_du.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        std::memcpy(energy, c->h_widom_du.ptr, n * sizeof(double));
    });
}

FB_API int fb_group_group_energy(fb_ctx* c, int s, int group1, int group2, double* energy)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (group1 < 0 || group1 >= c->n_groups || group2 < 0 || group2 >= c->n_groups || group1 == group2 || !energy) {
            throw CudaError{"two different groups of the uploaded space expected"};
        }
        const SlotView V = makeView(c, s);
#define FB_CASE(K)                                                                             \
    case K:                                                                                    \
        groupPairKernel<K><<<1, kBlock, 0, c->stream>>>(V, c->P, group1, group2, c->d_result); \
        break;
        switch (c->P.kind) {
            FB_CASE(POT_COULOMB_LJ)
            FB_CASE(POT_COULOMB_WCA)
            FB_CASE(POT_PM)
            FB_CASE(POT_PMWCA)
            FB_CASE(POT_FUNCTOR)
            FB_CASE(POT_SPLINED)
        default:
            throw CudaError{"unknown potential kind"};
        }
#undef FB_CASE
        launched(c, "groupPairKernel");
        finish(c);
        *energy = c->h_result[0];
    });
}

// =================================================================================================
// Forces (fb_force.cuh)
// =================================================================================================
FB_API int fb_set_force_table(fb_ctx* c, int n_knots, const double* knots, const double* coeffs)
{
    return guarded(c, [&] {
        flushPending(c);
        if (n_knots < 2 || !knots || !coeffs) {
            throw CudaError{"an Andrea table of S'(q) with at least two knots expected"};
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->d_force_knots.upload(knots, static_cast<size_t>(n_knots), c->stream);
        c->d_force_coef.upload(coeffs, 6 * static_cast<size_t>(n_knots - 1), c->stream);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->force_nk = n_knots;
    });
}

FB_API int fb_nonbonded_force(fb_ctx* c, int s, double* forces)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!forces) {
            throw CudaError{"null argument"};
        }
        if (c->P.kind != POT_COULOMB_LJ && c->P.kind != POT_COULOMB_WCA) {
            // PairPotential::force, src/potentials.cpp:246-251
            throw CudaError{"Force computation not implemented for this setup!"};
        }
        if (c->force_nk < 2) {
            throw CudaError{"fb_set_force_table has not been called"};
        }
        const int n = c->n_slots;
        c->d_forces.ensure(3 * static_cast<size_t>(n));
        c->h_forces.ensure(3 * static_cast<size_t>(n));
        const SlotView V = makeView(c, s);
        ForceTable T{c->force_nk, c->d_force_knots.ptr, c->d_force_coef.ptr};
        // particles j per share: 2048, or N/64 rounded up to whole tiles if that is more (at most 64 shares) — a function of
        // N alone, so the sums do not depend on the machine
        const int kForceChunk = std::max(2048, ((n + 63) / 64 + kForceBlock - 1) / kForceBlock * kForceBlock);
        const int n_ranges = (n + kForceChunk - 1) / kForceChunk;
        c->d_force_shares.ensure(3 * static_cast<size_t>(n) * n_ranges);
        const dim3 grid((n + kForceBlock - 1) / kForceBlock, n_ranges);
        if (c->P.kind == POT_COULOMB_LJ) {
            nonbondedForceKernel<POT_COULOMB_LJ><<<grid, kForceBlock, 0, c->stream>>>(V, c->P, T, kForceChunk, c->d_force_shares.ptr);
        }
        else {
            nonbondedForceKernel<POT_COULOMB_WCA><<<grid, kForceBlock, 0, c->stream>>>(V, c->P, T, kForceChunk, c->d_force_shares.ptr);
        }
        forceSumKernel<<<(3 * n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->d_force_shares.ptr, 3 * n, n_ranges, c->d_forces.ptr);
        launched(c, "nonbondedForceKernel");
        CUDA_CHECK(cudaMemcpyAsync(c->h_forces.ptr, c->d_forces.ptr, 3 * static_cast<size_t>(n) * sizeof(double),
                                   cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        for (size_t i = 0; i < 3 * static_cast<size_t>(n); ++i) {
            forces[i] += c->h_forces.ptr[i]; // forces[i] += f, src/energy.h:1593
        }
    });
}

FB_API int fb_ewald_force(fb_ctx* c, int s, double* forces)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!forces) {
            throw CudaError{"null argument"};
        }
        Slot& sl = c->slot[s];
        if (sl.K <= 0) {
            throw CudaError{"no k-vectors"};
        }
        const int n = c->n_slots;
        c->d_forces.ensure(3 * static_cast<size_t>(n));
        c->h_forces.ensure(3 * static_cast<size_t>(n));
        const SlotView V = makeView(c, s);
        {
            const int grid = gridFor(c, n, kBlock);
            c->partials.ensure(3 * static_cast<size_t>(grid));
            dipoleAllKernel<<<grid, kBlock, 0, c->stream>>>(V, c->partials.ptr, c->ticket.ptr, c->d_result + 5);
            launched(c, "dipoleAllKernel");
        }
        const double pi = 3.141592653589793238462643383279502884;
        const double volume = sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2];
        const double surface = 1.0 / (2.0 * c->ewald.surface_dielectric_constant + 1.0);
        const double scale = -4.0 * pi / volume * c->ewald.bjerrum_length;
        const int grid = (n + kEwaldForceParticles - 1) / kEwaldForceParticles;
        ewaldForceKernel<<<grid, kEwaldForceParticles * kEwaldForceLanes, 0, c->stream>>>(V, makeEwaldView(c, s), c->d_result + 5,
                                                                                          surface, scale, c->d_forces.ptr);
        launched(c, "ewaldForceKernel");
        CUDA_CHECK(cudaMemcpyAsync(c->h_forces.ptr, c->d_forces.ptr, 3 * static_cast<size_t>(n) * sizeof(double),
                                   cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        std::memcpy(forces, c->h_forces.ptr, 3 * static_cast<size_t>(n) * sizeof(double)); // (*force) = …, src/energy.cpp:610
    });
}

/** share `shard` of `n_shards` of the full-system non-bonded and reciprocal energies (multi-GPU system energy) */
namespace {

/** the pair-distance histogram of atoms of two types, or of the mass centres of molecular groups of two kinds */
int pairRdf(fb_ctx* c, int s, bool molecular, int atom_id1, int atom_id2, double dr, const int* slice_dir, double thickness,
            int shard, int n_shards, int n_bins, unsigned long long* counts)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (n_shards < 1 || shard < 0 || shard >= n_shards) {
            throw CudaError{"fb_atom_rdf: bad shard arguments"};
        }
        const int n_kinds = molecular ? c->P.n_mol : c->P.n_types;
        if (!counts || !(dr > 0.0) || n_bins < 1 || n_bins > kRdfMaxBins || atom_id1 < 0 || atom_id1 >= n_kinds ||
            atom_id2 < 0 || atom_id2 >= n_kinds) {
            throw CudaError{"fb_atom_rdf / fb_molecule_rdf: bad arguments (1..12288 bins, dr > 0, known types)"};
        }
        if (molecular && ((c->molecule_flags[atom_id1] | c->molecule_flags[atom_id2]) & FB_MOL_ATOMIC)) {
            throw CudaError{"fb_molecule_rdf: molecular groups required"};
        }
        c->rdf_hist.ensure(static_cast<size_t>(n_bins));
        c->rdf_flag.ensure(1);
        CUDA_CHECK(cudaMemsetAsync(c->rdf_hist.ptr, 0, sizeof(unsigned long long) * n_bins, c->stream));
        CUDA_CHECK(cudaMemsetAsync(c->rdf_flag.ptr, 0, sizeof(int), c->stream));
        // the particles of the two types, compacted (all threads and all inner iterations of the tiles do work)
        const int n_items = molecular ? c->n_groups : c->n_slots;
        c->rdf_list[0].ensure(static_cast<size_t>(n_items));
        c->rdf_list[1].ensure(static_cast<size_t>(n_items));
        c->rdf_n.ensure(2);
        CUDA_CHECK(cudaMemsetAsync(c->rdf_n.ptr, 0, 2 * sizeof(int), c->stream));
        const bool identical = atom_id1 == atom_id2;
        const int tiles = (n_items + kRdfTile - 1) / kRdfTile; // upper bound: blocks beyond the lists return
        const size_t smem = sizeof(unsigned int) * static_cast<size_t>(n_bins);
        if (!c->rdf_configured) {
            CUDA_CHECK(cudaFuncSetAttribute(atomRdfKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(sizeof(unsigned int) * kRdfMaxBins)));
            c->rdf_configured = true;
        }
        const int sx = slice_dir ? slice_dir[0] : 0, sy = slice_dir ? slice_dir[1] : 0, sz = slice_dir ? slice_dir[2] : 0;
        beginTiming(c, TIME_FULL);
        if (molecular) {
            moleculeRdfCompactKernel<<<1, kRdfCompactThreads, 0, c->stream>>>(makeView(c, s), atom_id1, atom_id2,
                                                                              c->rdf_list[0].ptr, c->rdf_list[1].ptr,
                                                                              c->rdf_n.ptr);
            launched(c, "moleculeRdfCompactKernel");
        }
        else {
            atomRdfCompactKernel<<<1, kRdfCompactThreads, 0, c->stream>>>(makeView(c, s), atom_id1, atom_id2,
                                                                          c->rdf_list[0].ptr, c->rdf_list[1].ptr,
                                                                          c->rdf_n.ptr);
            launched(c, "atomRdfCompactKernel");
        }
        atomRdfKernel<<<dim3((tiles + n_shards - 1) / n_shards, tiles), kRdfTile, smem, c->stream>>>(
            makeView(c, s), c->rdf_list[0].ptr, c->rdf_list[1].ptr, c->rdf_n.ptr, identical, molecular, 1.0 / dr, sx, sy, sz,
            thickness,
            n_bins, shard, n_shards, c->rdf_hist.ptr, c->rdf_flag.ptr);
        launched(c, "atomRdfKernel");
        std::vector<unsigned long long> host(static_cast<size_t>(n_bins));
        int flag = 0;
        CUDA_CHECK(cudaMemcpyAsync(host.data(), c->rdf_hist.ptr, sizeof(unsigned long long) * n_bins, cudaMemcpyDeviceToHost,
                                   c->stream));
        CUDA_CHECK(cudaMemcpyAsync(&flag, c->rdf_flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        if (flag != 0) {
            throw CudaError{"fb_atom_rdf: a distance fell beyond the last bin"};
        }
        for (int b = 0; b < n_bins; ++b) {
            counts[b] += host[b];
        }
    });
}

} // namespace

FB_API int fb_atom_rdf(fb_ctx* c, int s, int atom_id1, int atom_id2, double dr, const int* slice_dir, double thickness,
                       int shard, int n_shards, int n_bins, unsigned long long* counts)
{
    return pairRdf(c, s, false, atom_id1, atom_id2, dr, slice_dir, thickness, shard, n_shards, n_bins, counts);
}

FB_API int fb_molecule_rdf(fb_ctx* c, int s, int molid1, int molid2, double dr, int shard, int n_shards, int n_bins,
                           unsigned long long* counts)
{
    return pairRdf(c, s, true, molid1, molid2, dr, nullptr, 0.0, shard, n_shards, n_bins, counts);
}

FB_API int fb_system_energy_shard(fb_ctx* c, int s, int shard, int n_shards, double* nonbonded, double* reciprocal)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!nonbonded || !reciprocal || n_shards < 1 || shard < 0 || shard >= n_shards) {
            throw CudaError{"bad shard arguments"};
        }
        beginTiming(c, TIME_FULL);
        launchFull(c, makeView(c, s), 0, shard, n_shards);
        finish(c);
        *nonbonded = c->h_result[0];
        *reciprocal = 0.0;
        Slot& sl = c->slot[s];
        if (c->ewald_configured && sl.K > 0) {
            double sum = 0.0;
            if (sl.n_gemm_tiles > 0 && c->full_q_path == 0) { // a slab of tile columns (all three policies)
                const int n_columns = static_cast<int>(sl.gemm_column_first_tile.size()) - 1;
                // boundaries where the k-vectors divide evenly: the first column that starts at or after K·r / n
                auto boundary = [&](int r) {
                    const long long want = static_cast<long long>(sl.K) * r / n_shards;
                    int col = 0;
                    while (col < n_columns && sl.gemm_tile_first_k[sl.gemm_column_first_tile[col]] < want) {
                        ++col;
                    }
                    return r >= n_shards ? sl.n_gemm_tiles : sl.gemm_column_first_tile[col];
                };
                const int tile_begin = boundary(shard), tile_end = boundary(shard + 1);
                if (tile_end > tile_begin) {
                    const int k_blocks = (sl.gemm_tile_first_k[tile_end] - sl.gemm_tile_first_k[tile_begin] + 255) / 256;
                    c->partials.ensure(static_cast<size_t>(k_blocks));
                    const int grid = launchFullQGemm(c, s, tile_begin, tile_end, c->partials.ptr);
                    orderedSumKernel<<<1, 1024, 0, c->stream>>>(c->partials.ptr, static_cast<size_t>(grid), 1, c->d_result);
                    launched(c, "orderedSumKernel");
                    CUDA_CHECK(cudaStreamSynchronize(c->stream));
                    sum = c->h_result[0];
                }
            }
            else if (c->ewald.policy != 2 && sl.n_cells > 0) { // a slab of k-cells
                const int cell_begin = static_cast<int>(static_cast<long long>(sl.n_cells) * shard / n_shards);
                const int cell_end = static_cast<int>(static_cast<long long>(sl.n_cells) * (shard + 1) / n_shards);
                if (cell_end > cell_begin) {
                    const int grid = cell_end - cell_begin;
                    c->partials.ensure(static_cast<size_t>(grid));
                    launchFullQ(c, s, cell_begin, cell_end, false, c->partials.ptr);
                    orderedSumKernel<<<1, 1024, 0, c->stream>>>(c->partials.ptr, static_cast<size_t>(grid), 1, c->d_result);
                    launched(c, "orderedSumKernel");
                    CUDA_CHECK(cudaStreamSynchronize(c->stream));
                    sum = c->h_result[0];
                }
            }
            else {
                const int k_begin = static_cast<int>(static_cast<long long>(sl.K) * shard / n_shards);
                const int k_end = static_cast<int>(static_cast<long long>(sl.K) * (shard + 1) / n_shards);
                if (k_end > k_begin) {
                    const int grid = (k_end - k_begin + kEwaldBlock - 1) / kEwaldBlock;
                    c->partials.ensure(static_cast<size_t>(grid));
                    ewaldSlabEnergyKernel<<<grid, kEwaldBlock, 0, c->stream>>>(makeView(c, s), makeEwaldView(c, s),
                                                                              k_begin, k_end, c->partials.ptr);
                    launched(c, "ewaldSlabEnergyKernel");
                    orderedSumKernel<<<1, 1024, 0, c->stream>>>(c->partials.ptr, static_cast<size_t>(grid), 1, c->d_result);
                    launched(c, "orderedSumKernel");
                    CUDA_CHECK(cudaStreamSynchronize(c->stream));
                    sum = c->h_result[0];
                }
            }
            const double pi = 3.141592653589793238462643383279502884;
            const double volume = sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2];
            *reciprocal = 2 * pi * sum * c->ewald.bjerrum_length / volume;
        }
    });
}

// =================================================================================================
// fast path
// =================================================================================================
namespace {

constexpr unsigned long long kSentinel = 0xFFF8DEADBEEF5A5Aull; //!< a NaN payload arithmetic never produces

void armResults(fb_ctx* c, int n)
{
    volatile unsigned long long* r = reinterpret_cast<volatile unsigned long long*>(c->h_result);
    for (int i = 0; i < n; ++i) {
        r[i] = kSentinel;
    }
    __sync_synchronize();
}

/** Spin until the kernel has overwritten all `n` armed result slots (mapped pinned memory) */
void waitResults(fb_ctx* c, int n)
{
    volatile unsigned long long* r = reinterpret_cast<volatile unsigned long long*>(c->h_result);
    unsigned long spins = 0;
    while (true) {
        bool done = true;
        for (int i = 0; i < n; ++i) {
            done = done && (r[i] != kSentinel);
        }
        if (done) {
            return;
        }
        __builtin_ia32_pause();
        if ((++spins & 0xFFFFFul) == 0) {
            const cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess) {
                bool ok = true;
                for (int i = 0; i < n; ++i) {
                    ok = ok && (r[i] != kSentinel);
                }
                if (!ok) {
                    throw CudaError{"cuda: trial kernel finished without publishing its result"};
                }
                return;
            }
            if (q != cudaErrorNotReady) {
                throw CudaError{std::string("cuda: trial kernel failed: ") + cudaGetErrorString(q)};
            }
        }
    }
}

template <int K>
void launchTrial(fb_ctx* c, int grid, int n_pair_blocks, int internal, const EwaldView& Ecur, const EwaldView& Eout)
{
    trialMoveKernel<K><<<grid, kBlock, 0, c->stream>>>(makeView(c, 0), makeView(c, 1), c->P,
                                                      c->has_commit ? c->commit : Overlay{0, -1, {}, {}, {}, {}},
                                                      c->trial, internal, Ecur, Eout, n_pair_blocks,
                                                      c->partials.ptr, c->ticket.ptr, c->d_result);
}

} // namespace

FB_API int fb_trial_energy(fb_ctx* c, const fb_trial_move* mv, double* u_new, double* u_old, double* ewald_new,
                           double* ewald_old)
{
    return guarded(c, [&] {
        flushBatch(c);
        checkSlot(c, 0);
        checkSlot(c, 1);
        if (!mv || !u_new || !u_old) {
            throw CudaError{"null argument"};
        }
        if (mv->group_index < 0 || mv->group_index >= c->n_groups || mv->n_atoms < 1 || mv->n_atoms > kFastAtoms) {
            throw CudaError{"trial move outside the fast path (1..8 atoms of one group)"};
        }
        const fb_group& g = c->slot[0].groups[mv->group_index];
        Overlay t{};
        t.n = mv->n_atoms;
        t.group = mv->group_index;
        for (int i = 0; i < mv->n_atoms; ++i) {
            if (mv->rel_index[i] < 0 || mv->rel_index[i] >= g.size) {
                throw CudaError{"relative atom index out of range (fast path moves active atoms only)"};
            }
            if (mv->atom_id[i] < 0 || mv->atom_id[i] >= c->P.n_types) {
                throw CudaError{"atom id out of range"};
            }
            t.slot[i] = g.begin + mv->rel_index[i];
            t.id[i] = mv->atom_id[i];
            t.posq[i] = make_double4(mv->xyzq[i][0], mv->xyzq[i][1], mv->xyzq[i][2], mv->xyzq[i][3]);
        }
        t.cm = make_double4(mv->cm[0], mv->cm[1], mv->cm[2], 0.0);
        c->trial = t;
        c->trial_with_ewald = mv->with_ewald != 0;
        EwaldView Ecur{}, Eout{};
        int n_k_blocks = 0;
        if (mv->with_ewald) {
            if (!c->ewald_configured || c->slot[0].K <= 0 || c->slot[1].K != c->slot[0].K) {
                throw CudaError{"Ewald state of the two slots is not aligned for the fast path"};
            }
            if (!ewald_new || !ewald_old) {
                throw CudaError{"null argument"};
            }
            if (!c->slot[0].rec_valid) { // Σ A_k|Q_k|² of the accepted state, once
                Slot& sl = c->slot[0];
                const int grid = (sl.K + kBlock - 1) / kBlock;
                c->partials.ensure(static_cast<size_t>(grid));
                ewaldEnergyKernel<<<grid, kBlock, 0, c->stream>>>(makeEwaldView(c, 0), c->partials.ptr, c->ticket.ptr,
                                                                 c->d_result + 4);
                launched(c, "ewaldEnergyKernel");
                CUDA_CHECK(cudaStreamSynchronize(c->stream));
                sl.rec_sum = c->h_result[4];
                sl.rec_valid = true;
            }
            Ecur = makeEwaldView(c, 0);
            Eout = makeEwaldView(c, 1);
            n_k_blocks = (Ecur.K + kBlock - 1) / kBlock;
        }
        // one resident wave: split the block budget between the pair part and the k-space part by work
        int n_pair_blocks = (c->n_slots + kBlock - 1) / kBlock;
        if (n_pair_blocks + n_k_blocks > c->max_blocks) {
            const double share = static_cast<double>(n_pair_blocks) / (n_pair_blocks + n_k_blocks);
            int pair_budget = std::max(1, static_cast<int>(share * c->max_blocks));
            if (n_k_blocks > 0) {
                pair_budget = std::min(pair_budget, c->max_blocks - 1);
            }
            n_pair_blocks = std::min(n_pair_blocks, pair_budget);
            n_k_blocks = std::min(n_k_blocks, c->max_blocks - n_pair_blocks);
        }
        const int grid = n_pair_blocks + n_k_blocks;
        c->partials.ensure(static_cast<size_t>(3) * grid);
        if (!c->timing) {
            armResults(c, 3);
        }
        beginTiming(c, TIME_PAIR);
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        launchTrial<K>(c, grid, n_pair_blocks, mv->internal, Ecur, Eout);                                     \
        break;
        switch (c->P.kind) {
            FB_CASE(POT_COULOMB_LJ)
            FB_CASE(POT_COULOMB_WCA)
            FB_CASE(POT_PM)
            FB_CASE(POT_PMWCA)
            FB_CASE(POT_FUNCTOR)
            FB_CASE(POT_SPLINED)
        default:
            throw CudaError{"unknown potential kind"};
        }
#undef FB_CASE
        launched(c, "trialMoveKernel");
        c->has_commit = false; // the launch materialises the pending commit
        if (c->timing) {
            finish(c);
        }
        else {
            waitResults(c, 3);
        }
        *u_new = c->h_result[0];
        *u_old = c->h_result[1];
        if (mv->with_ewald) {
            const double pi = 3.141592653589793238462643383279502884;
            const Slot& sl = c->slot[0];
            const double pref = 2 * pi * c->ewald.bjerrum_length / (sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2]);
            c->trial_rec_sum = c->h_result[2];
            *ewald_new = pref * c->trial_rec_sum;
            *ewald_old = pref * sl.rec_sum;
        }
        c->trial_active = true;
    });
}

FB_API int fb_trial_commit(fb_ctx* c, int accept)
{
    return guarded(c, [&] {
        if (!c->trial_active) {
            throw CudaError{"no evaluated trial move to commit"};
        }
        c->trial_active = false;
        if (!accept) {
            return; // nothing was written anywhere
        }
        c->commit = c->trial;
        c->has_commit = true;
        for (int s = 0; s < 2; ++s) {
            fb_group& g = c->slot[s].groups[c->trial.group];
            g.cm[0] = c->trial.cm.x;
            g.cm[1] = c->trial.cm.y;
            g.cm[2] = c->trial.cm.z;
        }
        if (c->trial_with_ewald) { // Q_out becomes the accepted structure factor: swap the two buffers
            std::swap(c->slot[0].Q.ptr, c->slot[1].Q.ptr);
            std::swap(c->slot[0].Q.count, c->slot[1].Q.count);
            c->slot[0].rec_sum = c->trial_rec_sum;
            c->slot[0].rec_valid = true;
            c->slot[1].rec_valid = false;
        }
    });
}

// =================================================================================================
// Ewald
// =================================================================================================
FB_API int fb_ewald_configure(fb_ctx* c, const fb_ewald_config* cfg)
{
    return guarded(c, [&] {
        if (!cfg || cfg->alpha <= 0 || cfg->policy < 0 || cfg->policy > 2) {
            throw CudaError{"bad Ewald configuration"};
        }
        c->ewald = *cfg;
        c->ewald_configured = true;
    });
}

FB_API int fb_ewald_update_box(fb_ctx* c, int s, int* n_kvectors)
{
    return guarded(c, [&] {
        checkSlot(c, s, false);
        if (!c->ewald_configured) {
            throw CudaError{"Ewald not configured"};
        }
        Slot& sl = c->slot[s];
        flushBatch(c);
        std::vector<double4> kA;
        std::vector<int4> kn;
        generateKVectors(c->ewald, sl.box, kA, kn);
        { // store cell by cell (see fb_batch.cuh); the reference order is kept in `perm` for downloads
            const int ncc = static_cast<int>(std::ceil(c->ewald.n_cutoff));
            const long nc = (2 * ncc) / 4 + 1;
            auto cell_of = [&](const int4& n) {
                return (static_cast<long>(n.x >> 2) * nc + ((n.y + ncc) >> 2)) * nc + ((n.z + ncc) >> 2);
            };
            std::vector<int> perm(kA.size());
            for (size_t i = 0; i < perm.size(); ++i) {
                perm[i] = static_cast<int>(i);
            }
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return cell_of(kn[a]) < cell_of(kn[b]); });
            std::vector<double4> kA2(kA.size());
            std::vector<int4> kn2(kn.size());
            std::vector<int> cell_start;
            for (size_t i = 0; i < perm.size(); ++i) {
                kA2[i] = kA[perm[i]];
                kn2[i] = kn[perm[i]];
                if (i == 0 || cell_of(kn2[i]) != cell_of(kn2[i - 1])) {
                    cell_start.push_back(static_cast<int>(i));
                }
            }
            cell_start.push_back(static_cast<int>(perm.size()));
            kA.swap(kA2);
            kn.swap(kn2);
            sl.perm.swap(perm);
            sl.n_cells = static_cast<int>(cell_start.size()) - 1;
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            sl.cell_start.upload(cell_start.data(), cell_start.size(), c->stream);
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        sl.kA.upload(kA.data(), kA.size(), c->stream);
        sl.kn.upload(kn.data(), kn.size(), c->stream);
        buildGemmTiles(c, sl, kn);
        std::vector<double> ksq(kA.size());
        for (size_t i = 0; i < kA.size(); ++i) {
            ksq[i] = std::sqrt(kA[i].w);
        }
        sl.ksq.upload(ksq.data(), ksq.size(), c->stream);
        { // units of the window k-space kernel: the y-rows {0, 1} and {2, 3} of every cell that hold k-vectors
            const int ncc = static_cast<int>(std::ceil(c->ewald.n_cutoff));
            std::vector<double2> aks(kA.size());
            for (size_t i = 0; i < kA.size(); ++i) {
                aks[i] = make_double2(kA[i].w, ksq[i]);
            }
            std::vector<int4> unit_info;
            std::vector<unsigned char> unit_map;
            std::vector<double> unit_sa;
            std::vector<int> cell_start_host;
            std::vector<int4> cell_units; // per cell: unit of h = 0, of h = 1 (−1: none), table index of x, y (the z one: bz)
            std::vector<int4> cell_key;   // per cell: cell coordinates and the z table index
            for (size_t i = 0; i < kn.size(); ++i) {
                if (i == 0 || (kn[i].x >> 2) != (kn[i - 1].x >> 2) || ((kn[i].y + ncc) >> 2) != ((kn[i - 1].y + ncc) >> 2) ||
                    ((kn[i].z + ncc) >> 2) != ((kn[i - 1].z + ncc) >> 2)) {
                    cell_start_host.push_back(static_cast<int>(i));
                }
            }
            cell_start_host.push_back(static_cast<int>(kn.size()));
            for (size_t cell = 0; cell + 1 < cell_start_host.size(); ++cell) {
                const int p0 = cell_start_host[cell];
                const int len = cell_start_host[cell + 1] - p0;
                const int bx = kn[p0].x & ~3;
                const int by = (kn[p0].y + ncc) & ~3; // table indices (offset by ncc)
                const int bz = (kn[p0].z + ncc) & ~3;
                int unit_of_half[2] = {-1, -1};
                for (int h = 0; h < 2; ++h) {
                    unsigned char map[32];
                    double sa[32] = {};
                    std::fill(map, map + 32, static_cast<unsigned char>(255));
                    bool any = false;
                    for (int t = 0; t < len; ++t) {
                        const int4 v = kn[p0 + t];
                        const int li = v.x - bx, lj = (v.y + ncc) - by, ll = (v.z + ncc) - bz;
                        if ((lj >> 1) == h) {
                            map[8 * li + 4 * (lj & 1) + ll] = static_cast<unsigned char>(t);
                            sa[8 * li + 4 * (lj & 1) + ll] = ksq[p0 + t];
                            any = true;
                        }
                    }
                    if (any) {
                        unit_of_half[h] = static_cast<int>(unit_info.size());
                        unit_info.push_back(make_int4(p0, bx, by + 2 * h, bz));
                        unit_map.insert(unit_map.end(), map, map + 32);
                        unit_sa.insert(unit_sa.end(), sa, sa + 32);
                    }
                }
                cell_units.push_back(make_int4(unit_of_half[0], unit_of_half[1], bx, by));
                cell_key.push_back(make_int4(bx >> 2, by >> 2, bz >> 2, bz));
            }
            // commit items: pairs of cells that follow each other in z (the cells are stored z fastest)
            std::vector<int4> item_units, item_base;
            for (size_t cell = 0; cell < cell_units.size();) {
                const int4 u0 = cell_units[cell];
                const int4 k0 = cell_key[cell];
                int4 u1 = make_int4(-1, -1, 0, 0);
                int z1 = k0.w;
                size_t used = 1;
                if (cell + 1 < cell_units.size()) {
                    const int4 k1 = cell_key[cell + 1];
                    if (k1.x == k0.x && k1.y == k0.y && k1.z == k0.z + 1) {
                        u1 = cell_units[cell + 1];
                        z1 = k1.w;
                        used = 2;
                    }
                }
                item_units.push_back(make_int4(u0.x, u0.y, u1.x, u1.y));
                item_base.push_back(make_int4(u0.z, u0.w, k0.w, z1));
                cell += used;
            }
            sl.n_items = static_cast<int>(item_units.size());
            if (sl.n_items > 0) {
                sl.item_units.upload(item_units.data(), item_units.size(), c->stream);
                sl.item_base.upload(item_base.data(), item_base.size(), c->stream);
            }
            sl.n_units = static_cast<int>(unit_info.size());
            sl.aks.upload(aks.data(), aks.size(), c->stream);
            // Schedule of the units over the 2·SM blocks of the persistent kernel. A unit costs its phases (fixed) plus
            // one Gram step per slot column (jj, l) that holds a k-vector for some x-index — units cut by the sphere
            // have fewer (8.7 % of all steps at K = 57 950). Longest processing time first onto the least loaded block
            // (ties: lowest block), then every block walks its units in index order: 2134 units over 296 blocks are
            // 7 or 8 units each dealt round robin (the kernel takes as long as 8 full ones), ≈ 7.2 units' worth so.
            std::vector<unsigned char> unit_steps(unit_info.size(), 0);
            std::vector<int> sched_first, sched_units;
            {
                const int n_units = static_cast<int>(unit_info.size());
                std::vector<int> cost(n_units);
                for (int u = 0; u < n_units; ++u) {
                    unsigned mask = 0;
                    for (int slot = 0; slot < 32; ++slot) {
                        if (unit_map[static_cast<size_t>(u) * 32 + slot] != 255) {
                            mask |= 1u << (slot & 7); // slot = 8 xi + 4 jj + l
                        }
                    }
                    unit_steps[u] = static_cast<unsigned char>(mask);
                    cost[u] = 5 + 2 * __builtin_popcount(mask);
                }
                const int n_blocks = std::max(1, std::min(n_units, 2 * c->n_sm));
                std::vector<int> order(n_units);
                for (int u = 0; u < n_units; ++u) {
                    order[u] = u;
                }
                std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
                std::vector<std::vector<int>> of_block(n_blocks);
                std::priority_queue<std::pair<long, int>, std::vector<std::pair<long, int>>, std::greater<>> least;
                for (int b = 0; b < n_blocks; ++b) {
                    least.emplace(0L, b);
                }
                for (int u : order) {
                    auto [load, b] = least.top();
                    least.pop();
                    of_block[b].push_back(u);
                    least.emplace(load + cost[u], b);
                }
                sched_first.push_back(0);
                for (auto& units : of_block) {
                    std::sort(units.begin(), units.end());
                    sched_units.insert(sched_units.end(), units.begin(), units.end());
                    sched_first.push_back(static_cast<int>(sched_units.size()));
                }
                sl.n_sched_blocks = n_units > 0 ? n_blocks : 0;
            }
            if (sl.n_units > 0) {
                sl.unit_info.upload(unit_info.data(), unit_info.size(), c->stream);
                sl.unit_map.upload(unit_map.data(), unit_map.size(), c->stream);
                sl.unit_sa.upload(unit_sa.data(), unit_sa.size(), c->stream);
                sl.unit_steps.upload(unit_steps.data(), unit_steps.size(), c->stream);
                sl.sched_first.upload(sched_first.data(), sched_first.size(), c->stream);
                sl.sched_units.upload(sched_units.data(), sched_units.size(), c->stream);
            }
            CUDA_CHECK(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
        }
        sl.Q.ensure(kA.size());
        sl.K = static_cast<int>(kA.size());
        for (int i = 0; i < 3; ++i) {
            sl.ewald_box[i] = sl.box[i];
        }
        sl.rec_valid = false;
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (n_kvectors) {
            *n_kvectors = sl.K;
        }
    });
}

FB_API int fb_ewald_update_full(fb_ctx* c, int s)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        Slot& sl = c->slot[s];
        if (sl.K <= 0) {
            throw CudaError{"no k-vectors (call fb_ewald_update_box first)"};
        }
        beginTiming(c);
        if (sl.n_gemm_tiles > 0 && c->full_q_path == 0) { // PBC / PBCEigen: complex matrix product; IPBC: the real one
            launchFullQGemm(c, s);
        }
        else if (c->ewald.policy != 2 && sl.n_cells > 0) { // factorised phases, one block per k-cell
            launchFullQ(c, s, 0, sl.n_cells, true, nullptr);
        }
        else {
            const int grid = (sl.K + kEwaldBlock - 1) / kEwaldBlock;
            ewaldFullKernel<<<grid, kEwaldBlock, 0, c->stream>>>(makeView(c, s), makeEwaldView(c, s));
            launched(c, "ewaldFullKernel");
        }
        if (c->timing) {
            finish(c); // fb_last_kernel_ms: the rebuild alone
        }
        sl.rec_valid = false;
    });
}

FB_API int fb_ewald_update_partial(fb_ctx* c, int s_new, int s_old, const fb_change* change)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s_new);
        checkSlot(c, s_old);
        if (!change || s_new == s_old) {
            throw CudaError{"bad arguments"};
        }
        Slot& sn = c->slot[s_new];
        Slot& so = c->slot[s_old];
        if (sn.K <= 0 || sn.K != so.K) {
            throw CudaError{"k-vector sets of the two slots differ; use a full update"};
        }
        const MovedDesc md = lowerChange(c, s_new, s_old, change, true);
        const int grid = (sn.K + kBlock - 1) / kBlock;
        if (grid > kMaxPartialBlocks * 4) {
            c->partials.ensure(static_cast<size_t>(grid));
        }
        beginTiming(c, TIME_EWALD);
        ewaldPartialKernel<<<grid, kBlock, 0, c->stream>>>(makeView(c, s_new), makeView(c, s_old),
                                                           makeEwaldView(c, s_new), makeEwaldView(c, s_old), md,
                                                           c->partials.ptr, c->ticket.ptr, c->d_result + 4);
        launched(c, "ewaldPartialKernel");
        finish(c);
        sn.rec_sum = c->h_result[4];
        sn.rec_valid = true;
    });
}

FB_API int fb_ewald_energy(fb_ctx* c, int s, const fb_change* change, double* energy)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!energy) {
            throw CudaError{"null argument"};
        }
        const bool empty = !change || (!change->everything && !change->volume_change && change->n_groups == 0);
        if (empty) {
            *energy = 0.0;
            return;
        }
        Slot& sl = c->slot[s];
        if (sl.K <= 0) {
            throw CudaError{"no k-vectors"};
        }
        const double pi = 3.141592653589793238462643383279502884;
        const double volume = sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2];
        beginTiming(c);
        bool pending = false;
        if (!sl.rec_valid) {
            const int grid = (sl.K + kBlock - 1) / kBlock;
            c->partials.ensure(static_cast<size_t>(grid));
            ewaldEnergyKernel<<<grid, kBlock, 0, c->stream>>>(makeEwaldView(c, s), c->partials.ptr, c->ticket.ptr,
                                                             c->d_result + 4);
            launched(c, "ewaldEnergyKernel");
            pending = true;
        }
        const bool surface = c->ewald.surface_dielectric_constant >= 1.0;
        if (surface) {
            const int grid = gridFor(c, c->n_slots, kBlock);
            dipoleKernel<<<grid, kBlock, 0, c->stream>>>(makeView(c, s), c->partials.ptr, c->ticket.ptr,
                                                        c->d_result + 5);
            launched(c, "dipoleKernel");
            pending = true;
        }
        if (pending) {
            finish(c);
        }
        if (!sl.rec_valid) {
            sl.rec_sum = c->h_result[4];
            sl.rec_valid = true;
        }
        double u = 2 * pi * sl.rec_sum * c->ewald.bjerrum_length / volume;
        if (surface) {
            const double qr2 = c->h_result[5] * c->h_result[5] + c->h_result[6] * c->h_result[6] +
                               c->h_result[7] * c->h_result[7];
            u += 2.0 * pi / ((2.0 * c->ewald.surface_dielectric_constant + 1.0) * volume) * qr2 *
                 c->ewald.bjerrum_length;
        }
        *energy = u;
    });
}

FB_API int fb_ewald_sync(fb_ctx* c, int dst, int src, const fb_change* change)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, dst, false);
        checkSlot(c, src, false);
        if (dst == src || !change) {
            throw CudaError{"bad sync arguments"};
        }
        Slot& d = c->slot[dst];
        Slot& s = c->slot[src];
        if (s.K <= 0) {
            throw CudaError{"source slot has no k-vectors"};
        }
        if (change->everything || change->volume_change || d.K != s.K) {
            d.kA.ensure(s.K);
            CUDA_CHECK(cudaMemcpyAsync(d.kA.ptr, s.kA.ptr, s.K * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
            d.kn.ensure(s.K);
            CUDA_CHECK(cudaMemcpyAsync(d.kn.ptr, s.kn.ptr, s.K * sizeof(int4), cudaMemcpyDeviceToDevice, c->stream));
            d.ksq.ensure(s.K);
            CUDA_CHECK(cudaMemcpyAsync(d.ksq.ptr, s.ksq.ptr, s.K * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            d.cell_start.ensure(s.n_cells + 1);
            CUDA_CHECK(cudaMemcpyAsync(d.cell_start.ptr, s.cell_start.ptr, (s.n_cells + 1) * sizeof(int),
                                       cudaMemcpyDeviceToDevice, c->stream));
            d.n_cells = s.n_cells;
            d.n_gemm_tiles = s.n_gemm_tiles;
            if (s.n_gemm_tiles > 0) {
                d.gemm_tiles.ensure(s.n_gemm_tiles);
                CUDA_CHECK(cudaMemcpyAsync(d.gemm_tiles.ptr, s.gemm_tiles.ptr, s.n_gemm_tiles * sizeof(int4),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.gemm_order.ensure(s.n_gemm_tiles);
                CUDA_CHECK(cudaMemcpyAsync(d.gemm_order.ptr, s.gemm_order.ptr, s.n_gemm_tiles * sizeof(int),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.gemm_row_groups.ensure(s.n_gemm_tiles);
                CUDA_CHECK(cudaMemcpyAsync(d.gemm_row_groups.ptr, s.gemm_row_groups.ptr, s.n_gemm_tiles * sizeof(int),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.gemm_tile_first_k = s.gemm_tile_first_k;
                d.gemm_column_first_tile = s.gemm_column_first_tile;
                d.gemm_index.ensure(s.K);
                CUDA_CHECK(cudaMemcpyAsync(d.gemm_index.ptr, s.gemm_index.ptr, s.K * sizeof(int), cudaMemcpyDeviceToDevice,
                                           c->stream));
            }
            d.aks.ensure(s.K);
            CUDA_CHECK(cudaMemcpyAsync(d.aks.ptr, s.aks.ptr, s.K * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
            d.n_units = s.n_units;
            if (s.n_units > 0) {
                d.unit_info.ensure(s.n_units);
                CUDA_CHECK(cudaMemcpyAsync(d.unit_info.ptr, s.unit_info.ptr, s.n_units * sizeof(int4),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.unit_map.ensure(static_cast<size_t>(s.n_units) * 32);
                CUDA_CHECK(cudaMemcpyAsync(d.unit_map.ptr, s.unit_map.ptr, static_cast<size_t>(s.n_units) * 32,
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.n_items = s.n_items;
                d.item_units.ensure(s.n_items);
                CUDA_CHECK(cudaMemcpyAsync(d.item_units.ptr, s.item_units.ptr, s.n_items * sizeof(int4),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.item_base.ensure(s.n_items);
                CUDA_CHECK(cudaMemcpyAsync(d.item_base.ptr, s.item_base.ptr, s.n_items * sizeof(int4),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.unit_sa.ensure(static_cast<size_t>(s.n_units) * 32);
                CUDA_CHECK(cudaMemcpyAsync(d.unit_sa.ptr, s.unit_sa.ptr, static_cast<size_t>(s.n_units) * 32 * sizeof(double),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.unit_steps.ensure(static_cast<size_t>(s.n_units));
                CUDA_CHECK(cudaMemcpyAsync(d.unit_steps.ptr, s.unit_steps.ptr, static_cast<size_t>(s.n_units),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.n_sched_blocks = s.n_sched_blocks;
                d.sched_first.ensure(static_cast<size_t>(s.n_sched_blocks) + 1);
                CUDA_CHECK(cudaMemcpyAsync(d.sched_first.ptr, s.sched_first.ptr, (static_cast<size_t>(s.n_sched_blocks) + 1) * sizeof(int),
                                           cudaMemcpyDeviceToDevice, c->stream));
                d.sched_units.ensure(static_cast<size_t>(s.n_units));
                CUDA_CHECK(cudaMemcpyAsync(d.sched_units.ptr, s.sched_units.ptr, static_cast<size_t>(s.n_units) * sizeof(int),
                                           cudaMemcpyDeviceToDevice, c->stream));
            }
            d.perm = s.perm;
            d.K = s.K;
            for (int i = 0; i < 3; ++i) {
                d.ewald_box[i] = s.ewald_box[i];
            }
        }
        d.Q.ensure(s.K);
        CUDA_CHECK(cudaMemcpyAsync(d.Q.ptr, s.Q.ptr, s.K * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
        d.rec_valid = s.rec_valid;
        d.rec_sum = s.rec_sum;
    });
}

FB_API int fb_debug_fullq_layout(const fb_ewald_config* cfg, const double box[3], int max_k, int max_tiles, int* nxyz,
                                 int* index, int* tiles, int* order, int* column_first_tile, int* n_k, int* n_tiles,
                                 int* n_columns)
{
    if (!cfg || !box || !n_k || !n_tiles || !n_columns) {
        return FB_ERR_INVALID;
    }
    try {
        std::vector<double4> kA;
        std::vector<int4> kn;
        generateKVectors(*cfg, box, kA, kn);
        const int ncc = static_cast<int>(std::ceil(cfg->n_cutoff));
        const long nc = (2 * ncc) / 4 + 1; // the storage order of fb_ewald_update_box: cell by cell, stable
        auto cell_of = [&](const int4& n) {
            return (static_cast<long>(n.x >> 2) * nc + ((n.y + ncc) >> 2)) * nc + ((n.z + ncc) >> 2);
        };
        std::stable_sort(kn.begin(), kn.end(), [&](const int4& a, const int4& b) { return cell_of(a) < cell_of(b); });
        const GemmLayout layout = buildGemmLayout(kn, ncc);
        *n_k = static_cast<int>(kn.size());
        *n_tiles = static_cast<int>(layout.tiles.size());
        *n_columns = static_cast<int>(layout.column_first_tile.size()) - 1;
        if (*n_k > max_k || *n_tiles > max_tiles) {
            return FB_ERR_INVALID;
        }
        for (int i = 0; i < *n_k; ++i) {
            if (nxyz) {
                nxyz[3 * i] = kn[i].x;
                nxyz[3 * i + 1] = kn[i].y;
                nxyz[3 * i + 2] = kn[i].z;
            }
            if (index) {
                index[i] = layout.index[i];
            }
        }
        for (int t = 0; t < *n_tiles; ++t) {
            if (tiles) {
                tiles[4 * t] = layout.tiles[t].x;
                tiles[4 * t + 1] = layout.tiles[t].y;
                tiles[4 * t + 2] = layout.tiles[t].z;
                tiles[4 * t + 3] = layout.tiles[t].w;
            }
            if (order) {
                order[t] = layout.order[t];
            }
        }
        if (column_first_tile) {
            for (int col = 0; col <= *n_columns; ++col) {
                column_first_tile[col] = layout.column_first_tile[col];
            }
        }
        return FB_OK;
    }
    catch (const std::exception&) {
        return FB_ERR_INVALID;
    }
}

FB_API int fb_ewald_download(fb_ctx* c, int s, double* q_re_im, double* kvectors, double* aks)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s, false);
        Slot& sl = c->slot[s];
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        // the device stores the k-vectors cell by cell; hand them out in the reference's order
        if (q_re_im) {
            std::vector<double2> Q(sl.K);
            CUDA_CHECK(cudaMemcpy(Q.data(), sl.Q.ptr, sl.K * sizeof(double2), cudaMemcpyDeviceToHost));
            for (int i = 0; i < sl.K; ++i) {
                const int k = sl.perm.empty() ? i : sl.perm[i];
                q_re_im[2 * k] = Q[i].x;
                q_re_im[2 * k + 1] = Q[i].y;
            }
        }
        if (kvectors || aks) {
            std::vector<double4> kA(sl.K);
            CUDA_CHECK(cudaMemcpy(kA.data(), sl.kA.ptr, sl.K * sizeof(double4), cudaMemcpyDeviceToHost));
            for (int i = 0; i < sl.K; ++i) {
                const int k = sl.perm.empty() ? i : sl.perm[i];
                if (kvectors) {
                    kvectors[3 * k] = kA[i].x;
                    kvectors[3 * k + 1] = kA[i].y;
                    kvectors[3 * k + 2] = kA[i].z;
                }
                if (aks) {
                    aks[k] = kA[i].w;
                }
            }
        }
    });
}

// =================================================================================================
// Widom
// =================================================================================================
FB_API int fb_widom_batch(fb_ctx* c, int s, int ghost_group, int n_ghost_atoms, int n_insertions,
                          const double* ghost_xyzq, const int* ghost_atom_id, const double* ghost_cm, int internal,
                          double* du)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (ghost_group < 0 || ghost_group >= c->n_groups || n_insertions <= 0 || !ghost_xyzq ||
            !ghost_atom_id || !du) {
            throw CudaError{"bad Widom arguments"};
        }
        if (n_ghost_atoms <= 0 || n_ghost_atoms > kWidomMaxAtoms) {
            throw CudaError{"Widom ghosts of 1..8 atoms are supported"};
        }
        for (int a = 0; a < n_ghost_atoms; ++a) {
            if (ghost_atom_id[a] < 0 || ghost_atom_id[a] >= c->P.n_types) {
                throw CudaError{"ghost atom id out of range"};
            }
        }
        const size_t natoms = static_cast<size_t>(n_insertions) * n_ghost_atoms;
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->d_ghost.upload(reinterpret_cast<const double4*>(ghost_xyzq), natoms, c->stream);
        c->d_ghost_id.upload(ghost_atom_id, n_ghost_atoms, c->stream);
        const double4* d_cm = nullptr;
        const bool molecular = !(c->molecule_flags[c->slot[s].groups[ghost_group].molid] & FB_MOL_ATOMIC);
        if (molecular) {
            if (!ghost_cm) {
                throw CudaError{"molecular ghosts need mass centres"};
            }
            c->h_ghost.ensure(n_insertions);
            for (int b = 0; b < n_insertions; ++b) {
                c->h_ghost.ptr[b] = make_double4(ghost_cm[3 * b], ghost_cm[3 * b + 1], ghost_cm[3 * b + 2], 0);
            }
            c->d_ghost_cm.upload(c->h_ghost.ptr, n_insertions, c->stream);
            d_cm = c->d_ghost_cm.ptr;
        }
        if (!molecular) { // atomic ghosts: the streaming kernel (fb_stream.cuh)
            const int n_variants = n_insertions * n_ghost_atoms;
            const bool dense = std::isinf(c->pair_cut2);
            // finite cutoff: FP32 screening, particle ranges of kWidomSplit as the second grid dimension
            const int n_split = dense ? 1 : std::max(1, (c->n_slots + kWidomSplit - 1) / kWidomSplit);
            c->d_widom_partial.ensure(static_cast<size_t>(n_variants) * n_split);
            c->d_widom_du.ensure(n_insertions);
            c->h_widom_du.ensure(n_insertions);
            const SlotView V = makeView(c, s);
            const int grid = (n_variants + kStreamVariants - 1) / kStreamVariants;
            const float cut2_screen = dense ? 0.0f : screeningCutoff(c, s);
            beginTiming(c, TIME_WIDOM);
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        if (dense) {                                                                                          \
            widomStreamKernel<K, true><<<grid, kStreamThreads, 0, c->stream>>>(                               \
                V, c->P, ghost_group, n_ghost_atoms, n_variants, c->d_ghost.ptr, c->d_ghost_id.ptr, c->pair_cut2, \
                c->d_widom_partial.ptr);                                                                      \
            launched(c, "widomStreamKernel");                                                                 \
        }                                                                                                     \
        else {                                                                                                \
            widomScreenKernel<K><<<dim3(grid, n_split), kStreamThreads, 0, c->stream>>>(                      \
                V, c->P, ghost_group, n_ghost_atoms, n_variants, c->d_ghost.ptr, c->d_ghost_id.ptr, c->pair_cut2, \
                cut2_screen, c->d_widom_partial.ptr);                                                         \
            launched(c, "widomScreenKernel");                                                                 \
        }                                                                                                     \
        widomStreamFinishKernel<K><<<(n_insertions + 127) / 128, 128, 0, c->stream>>>(                        \
            V, c->P, n_ghost_atoms, n_insertions, c->d_ghost.ptr, c->d_ghost_id.ptr, internal,                \
            c->d_widom_partial.ptr, n_split, c->d_widom_du.ptr);                                              \
        launched(c, "widomStreamFinishKernel");                                                               \
        break;
            switch (c->P.kind) {
                FB_CASE(POT_COULOMB_LJ)
                FB_CASE(POT_COULOMB_WCA)
                FB_CASE(POT_PM)
                FB_CASE(POT_PMWCA)
                FB_CASE(POT_FUNCTOR)
                FB_CASE(POT_SPLINED)
            default:
                throw CudaError{"unknown potential kind"};
            }
#undef FB_CASE
            CUDA_CHECK(cudaMemcpyAsync(c->h_widom_du.ptr, c->d_widom_du.ptr, n_insertions * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->stream));
            finish(c);
            std::memcpy(du, c->h_widom_du.ptr, n_insertions * sizeof(double));
            return;
        }
        const int bx = (n_insertions + kWidomBlock - 1) / kWidomBlock;
        const int max_split = (c->n_slots + kWidomChunk - 1) / kWidomChunk;
        int n_split = std::max(1, std::min(max_split, (c->max_blocks * 2 + bx - 1) / bx));
        int j_per_split = (c->n_slots + n_split - 1) / n_split;
        j_per_split = ((j_per_split + kWidomChunk - 1) / kWidomChunk) * kWidomChunk;
        n_split = (c->n_slots + j_per_split - 1) / j_per_split;
        c->d_widom_partial.ensure(static_cast<size_t>(n_split) * n_insertions);
        c->d_widom_du.ensure(n_insertions);
        c->h_widom_du.ensure(n_insertions);
        const SlotView V = makeView(c, s);
        dim3 grid(bx, n_split);
        beginTiming(c, TIME_WIDOM);
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        widomKernel<K><<<grid, kWidomBlock, 0, c->stream>>>(V, c->P, ghost_group, n_ghost_atoms, n_insertions, \
                                                            c->d_ghost.ptr, c->d_ghost_id.ptr, d_cm,          \
                                                            j_per_split, c->d_widom_partial.ptr);             \
        launched(c, "widomKernel");                                                                           \
        widomFinishKernel<K><<<(n_insertions + 127) / 128, 128, 0, c->stream>>>(                              \
            V, c->P, ghost_group, n_ghost_atoms, n_insertions, n_split, c->d_ghost.ptr, c->d_ghost_id.ptr,    \
            internal, c->d_widom_partial.ptr, c->d_widom_du.ptr);                                             \
        launched(c, "widomFinishKernel");                                                                     \
        break;
        switch (c->P.kind) {
            FB_CASE(POT_COULOMB_LJ)
            FB_CASE(POT_COULOMB_WCA)
            FB_CASE(POT_PM)
            FB_CASE(POT_PMWCA)
            FB_CASE(POT_FUNCTOR)
            FB_CASE(POT_SPLINED)
        default:
            throw CudaError{"unknown potential kind"};
        }
#undef FB_CASE
        CUDA_CHECK(cudaMemcpyAsync(c->h_widom_du.ptr, c->d_widom_du.ptr, n_insertions * sizeof(double),
                                   cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        std::memcpy(du, c->h_widom_du.ptr, n_insertions * sizeof(double));
    });
}

// =================================================================================================
// replica exchange packing
// =================================================================================================
FB_API size_t fb_state_doubles(const fb_ctx* c)
{
    return c ? 3 + static_cast<size_t>(c->n_groups) + 5 * static_cast<size_t>(c->n_slots) : 0;
}

FB_API int fb_export_state(fb_ctx* c, int s, double* device_buffer)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!device_buffer) {
            throw CudaError{"null buffer"};
        }
        packStateKernel<<<gridFor(c, c->n_slots, 256), 256, 0, c->stream>>>(makeView(c, s), device_buffer);
        launched(c, "packStateKernel");
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

namespace {
/**
 * Packed state (fb_export_state layout) in device memory → slot `s`: particles, group sizes, box. `host_copy`:
 * pinned memory that already receives the same buffer on the context's stream (saves the read-back of the header),
 * or nullptr. The mass centres of molecular groups are NOT part of the packed state (they need the atom masses and
 * the reference's minimum-image rule, src/geometry.h:504-527): the caller follows up with fb_upload_groups. A box that
 * differs from the slot's drops the slot's k-vector tables (fb_ewald_update_box rebuilds them). Synchronises the
 * stream.
 */
void importPackedState(fb_ctx* c, int s, const double* device_buffer, const double* host_copy)
{
    Slot& sl = c->slot[s];
    std::vector<double> head(3 + c->n_groups);
    if (host_copy == nullptr) {
        CUDA_CHECK(cudaMemcpyAsync(head.data(), device_buffer, head.size() * sizeof(double), cudaMemcpyDeviceToHost,
                                   c->stream));
    }
    unpackStateKernel<<<gridFor(c, c->n_slots, 256), 256, 0, c->stream>>>(makeView(c, s), device_buffer);
    launched(c, "unpackStateKernel");
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (host_copy != nullptr) {
        std::copy(host_copy, host_copy + head.size(), head.begin());
    }
    bool box_changed = false;
    for (int i = 0; i < 3; ++i) {
        box_changed = box_changed || sl.box[i] != head[i];
    }
    if (box_changed) {
        const int rc = fb_set_box(c, s, head.data());
        if (rc != FB_OK) {
            throw CudaError{c->last_error};
        }
        sl.K = 0; // k-vectors and A_k belong to the old box: fb_ewald_update_box has to follow
    }
    std::vector<int> gsize(c->n_groups);
    for (int g = 0; g < c->n_groups; ++g) {
        sl.groups[g].size = static_cast<int>(head[3 + g]);
        gsize[g] = sl.groups[g].size;
    }
    sl.gsize.uploadVector(gsize, c->stream);
    buildGidKernel<<<c->n_groups, 128, 0, c->stream>>>(makeView(c, s));
    launched(c, "buildGidKernel");
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    sl.rec_valid = false;
    c->batch.cells_valid = false;
}
} // namespace

FB_API int fb_import_state(fb_ctx* c, int s, const double* device_buffer)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, s);
        if (!device_buffer) {
            throw CudaError{"null buffer"};
        }
        importPackedState(c, s, device_buffer, nullptr);
    });
}

FB_API int fb_export_state_host(fb_ctx* c, int s, double* host_buffer)
{
    return guarded(c, [&] {
        const size_t n = fb_state_doubles(c);
        c->d_state.ensure(n);
        const int rc = fb_export_state(c, s, c->d_state.ptr);
        if (rc != FB_OK) {
            throw CudaError{c->last_error};
        }
        CUDA_CHECK(cudaMemcpy(host_buffer, c->d_state.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
    });
}

FB_API int fb_import_state_host(fb_ctx* c, int s, const double* host_buffer)
{
    return guarded(c, [&] {
        const size_t n = fb_state_doubles(c);
        c->d_state.ensure(n);
        CUDA_CHECK(cudaMemcpy(c->d_state.ptr, host_buffer, n * sizeof(double), cudaMemcpyHostToDevice));
        const int rc = fb_import_state(c, s, c->d_state.ptr);
        if (rc != FB_OK) {
            throw CudaError{c->last_error};
        }
    });
}

#include "fb_batch_api.inl"
#include "fb_nccl.inl"

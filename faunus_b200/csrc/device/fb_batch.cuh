// Speculative evaluation of a WINDOW of sequential single-atom trial moves in one pass (sm_100a).
//
// The reference evaluates one trial move at a time (src/montecarlo.cpp:139-187): u_new = energy of the
// moved atom at its trial position with every other particle, u_old the same at the old position
// (GroupPairingPolicy::group2all + groupInternal, src/energy.h:1182-1195, 885-914) and, with Ewald,
// the reciprocal energy before/after the partial Q(k) update (src/energy.cpp:219-247, 524-531).
// A window of B such moves on DISTINCT atoms is evaluated here against the window-start state S0:
//
//   pair     u_new[m], u_old[m]     = Σ_{j≠p_m} u(trial_m | old_m , r_j(S0))           batchPairKernel / batchPairCellKernel
//   cross    C_new[a][m], C_old[a][m] = u(x_m, new_a) − u(x_m, old_a)                   batchPairFinishKernel
//   k-space  R[m]  = Σ_k A_k (2 Re(conj(Q_k) δ_m,k) + |δ_m,k|²)                        windowKspaceKernel (fb_kspace.cuh)
//            G[a][m] = Σ_k A_k Re(conj(δ_a,k) δ_m,k),  δ_m,k = q (e^{ik·new_m} − e^{ik·old_m})
//
// so that the energies of move m in the state where the accepted moves a < m have been applied are
//   u_x[m] + Σ_{a accepted} C_x[a][m]          and     ΔU_rec = pref (R[m] + 2 Σ_{a accepted} G[a][m]),
// exact identities — only the summation order differs from the one-move-at-a-time evaluation. The
// caller walks the window in order, decides each move, and tells the next launch which were accepted
// (batchPrepKernel writes their positions into both mirrors, windowFrontKernel adds their δ to Q(k);
// batchCommitKernel does both when windowed evaluation is left).
//
// e^{ik·r} is factorised into per-axis phase tables e^{i 2π n x/L} (k = 2π n/L on an orthogonal box),
// built once per window by windowFrontKernel: complex products per (k, position) instead of a sincos.
#pragma once
#include "fb_kernels.cuh"

// Programmatic dependent launch between the kernels of the front → k-space → tail → front chain (switch FB_PDL): a
// kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled as soon as every block of
// the kernel before it has passed FB_LAUNCH_DEPENDENTS (or exited), and must not touch that kernel's results before
// FB_GRID_DEPENDENCY_WAIT, which returns when the kernel before it has completed and its writes are visible. Both are
// no-ops for a kernel launched the ordinary way. MEASURED AND LEFT OFF: parity holds (76 run / window tests), but S1 drops from
// 8.98e5 to 8.56e5 moves/s — the blocks launched early sit on registers and shared memory while they wait, and the pair
// kernel on the other stream, which the k-space chain overlaps with, is what they take them from.
#ifndef FB_PDL
#define FB_PDL 0
#endif
#if FB_PDL
#define FB_GRID_DEPENDENCY_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define FB_LAUNCH_DEPENDENTS() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define FB_GRID_DEPENDENCY_WAIT()
#define FB_LAUNCH_DEPENDENTS()
#endif

namespace fbdev {

constexpr int kBatchMax = 64;    //!< moves per window

struct CommitList
{
    int n;
    int index[kBatchMax]; //!< indices into the PREVIOUS window
};

/** Description of one window: copied from the host as one block, or assembled on the device (fb_run.cuh) */
struct BatchInput
{
    int n;
    int with_ewald;
    int slot[kBatchMax];
    int id[kBatchMax];
    double4 pnew[kBatchMax];
    int idold[kBatchMax];    //!< the atoms as they are in the accepted state (supplied by the caller, who owns
    double4 pold[kBatchMax]; //!< the Space: no gather from the mirror on the critical path)
    // group mode (rigid-molecule moves, `moltransrot`): the n atoms above belong to n_groups moved groups, the
    // atoms of a move are contiguous; n_groups == 0: every atom is a move of its own (`transrot`)
    int n_groups;
    int move_first[kBatchMax];
    int move_natoms[kBatchMax];
    int move_group[kBatchMax];
    double4 cm_new[kBatchMax];
    double4 cm_old[kBatchMax];
    // what became of the PREVIOUS window: its accepted atoms (positions → mirrors, δ → Q(k)) and, in group mode,
    // its accepted moves (mass centres)
    CommitList commit;
    CommitList commit_moves;
    // runs (fb_run.cuh): first move of the window in its run; pair_ready: the pair sums of exactly this window were
    // evaluated ahead, against the state BEFORE the commit list above was applied (batchPairFixKernel corrects them)
    int first;
    int pair_ready;
};

/** Device-resident working set of one window */
struct BatchBuffers
{
    BatchInput* in;
    double4* pold;   //!< [kBatchMax] positions at window start (points into *in)
    int* idold;      //!< [kBatchMax]
    double2* table;  //!< [2·kBatchMax][table_stride] phase factors, variant 2m = new, 2m+1 = old
};

struct PhaseGeometry
{
    int ncc;          //!< ceil(n_cutoff): n_x ∈ [0, ncc], n_y, n_z ∈ [−ncc, ncc]
    int table_stride; //!< entries per position: (ncc+1) + 2(2ncc+1)
    double len[3];    //!< box lengths
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

/**
 * (a < cut) | (b < cut) | (c < cut) | (d < cut) as four DSETP with a chained predicate; written in PTX
 * because nvcc otherwise builds a 30-instruction min() tree for it. NaN compares false.
 */
__device__ __forceinline__ bool anyBelow(double a, double b, double c, double d, double cut)
{
    int p;
    asm("{\n\t.reg .pred q;\n\t"
        "setp.lt.f64 q, %1, %5;\n\t"
        "setp.lt.or.f64 q, %2, %5, q;\n\t"
        "setp.lt.or.f64 q, %3, %5, q;\n\t"
        "setp.lt.or.f64 q, %4, %5, q;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(p)
        : "d"(a), "d"(b), "d"(c), "d"(d), "d"(cut));
    return p != 0;
}

/**
 * Position (x, y, z) and fold flags of variant `v` from shared memory through 32-bit shared addresses that
 * the caller computes ONCE: with ordinary indexing nvcc re-derives the shared window base (S2UR
 * SR_CgaCtaId + 6 uniform instructions) in every iteration of the pair loops.
 */
__device__ __forceinline__ void loadVariant(unsigned var_addr, unsigned fold_addr, int v, double4& a, int& fold)
{
    const unsigned at = var_addr + static_cast<unsigned>(v) * 32u;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "r"(at));
    asm("ld.shared.f64 %0, [%1+16];" : "=d"(a.z) : "r"(at));
    asm("ld.shared.s32 %0, [%1];" : "=r"(fold) : "r"(fold_addr + static_cast<unsigned>(v) * 4u));
}

/** q e^{ik·r} for k = 2π(nx, ny, nz)/L from the phase table of one position (the x entries carry the charge) */
__device__ __forceinline__ double2 tablePhase(const double2* __restrict__ t, const PhaseGeometry& g, int nx, int ny,
                                              int nz)
{
    const double2 ex = __ldg(t + nx);
    const double2 ey = __ldg(t + (g.ncc + 1) + (ny + g.ncc));
    const double2 ez = __ldg(t + (g.ncc + 1) + (2 * g.ncc + 1) + (nz + g.ncc));
    return cmul(cmul(ex, ey), ez);
}

// ------------------------------------------------------------------------------------------------
// commit of the previous window's accepted moves: positions into both mirrors, δ into Q(k), and the
// reciprocal sum Σ_k A_k |Q_k|² of the resulting state (per-block partials)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
    batchCommitKernel(SlotView M0, SlotView M1, EwaldView E, const int4* __restrict__ kn, BatchBuffers prev,
                      PhaseGeometry geo, CommitList commit, int with_ewald, double* __restrict__ e_partials)
{
    __shared__ double scratch[kBlock / 32];
    __shared__ int s_table[kBatchMax]; //!< first table entry of the accepted move's new position
    if (threadIdx.x < commit.n) {
        const int m = commit.index[threadIdx.x];
        const double4 p = prev.in->pnew[m];
        s_table[threadIdx.x] = 2 * m * geo.table_stride;
        if (blockIdx.x == 0) {
            const int s = prev.in->slot[m];
            const int id = prev.in->id[m];
            M0.posq[s] = p;
            M0.atom_id[s] = id;
            M1.posq[s] = p;
            M1.atom_id[s] = id;
        }
    }
    if (!with_ewald) {
        return;
    }
    __syncthreads();
    const int ncommit = commit.n;
    const double2* __restrict__ table = prev.table;
    double e = 0.0;
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < E.K; k += gridDim.x * kBlock) {
        const int4 n = __ldg(kn + k);
        double2 Q = E.Q[k];
#pragma unroll 4
        for (int a = 0; a < ncommit; ++a) {
            const double2* t = table + s_table[a];
            const double2 en = tablePhase(t, geo, n.x, n.y, n.z);
            const double2 eo = tablePhase(t + geo.table_stride, geo, n.x, n.y, n.z);
            Q.x += en.x - eo.x;
            Q.y += en.y - eo.y;
        }
        if (ncommit > 0) {
            E.Q[k] = Q;
        }
        e += E.kA[k].w * (Q.x * Q.x + Q.y * Q.y);
    }
    const double s = blockSum<kBlock>(e, scratch);
    if (threadIdx.x == 0) {
        e_partials[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// window set-up: old positions from the (committed) mirror and the per-axis phase tables
// ------------------------------------------------------------------------------------------------
/**
 * The previous window's accepted trial positions go into both mirrors (pair stream, before the pair kernel);
 * in group mode also the mass centres of the accepted groups (`moves` lists move indices).
 */
__global__ void __launch_bounds__(kBatchMax)
    batchPrepKernel(SlotView M0, SlotView M1, BatchBuffers cur, BatchBuffers prev, int* __restrict__ redo = nullptr)
{
    if (redo != nullptr && threadIdx.x == 0) {
        *redo = 0; // raised by batchPairFixKernel, read by the gated pair kernel and the pair sums
    }
    const CommitList& commit = cur.in->commit;
    const CommitList& moves = cur.in->commit_moves;
    if (static_cast<int>(threadIdx.x) < commit.n) {
        const int m = commit.index[threadIdx.x];
        const int s = prev.in->slot[m];
        const double4 p = prev.in->pnew[m];
        const int id = prev.in->id[m];
        M0.posq[s] = p;
        M0.atom_id[s] = id;
        M1.posq[s] = p;
        M1.atom_id[s] = id;
    }
    if (static_cast<int>(threadIdx.x) < moves.n) {
        const int g = moves.index[threadIdx.x];
        const int group = prev.in->move_group[g];
        M0.gcm[group] = prev.in->cm_new[g];
        M1.gcm[group] = prev.in->cm_new[g];
    }
}

/** leaving windowed mode (group mode): the mass centres of the accepted groups, list by value */
__global__ void __launch_bounds__(kBatchMax) batchCommitGroupsKernel(SlotView M0, SlotView M1, BatchBuffers prev, CommitList moves)
{
    if (static_cast<int>(threadIdx.x) < moves.n) {
        const int g = moves.index[threadIdx.x];
        const int group = prev.in->move_group[g];
        M0.gcm[group] = prev.in->cm_new[g];
        M1.gcm[group] = prev.in->cm_new[g];
    }
}

/**
 * entry t of the per-axis phase tables e^{i 2π n x / L} of the 2n positions of a window; the x-axis entries carry the
 * CHARGE of the position (q e^{i k_x x}), so a product of one entry per axis is q e^{ik·r} and nobody downstream
 * needs the charges
 */
__device__ __forceinline__ void phaseTableEntry(const BatchInput* in, double2* __restrict__ table, const PhaseGeometry& geo,
                                                int t)
{
    const int variant = t / geo.table_stride;
    const int e = t - variant * geo.table_stride;
    const int m = variant >> 1;
    const double4 p = (variant & 1) ? in->pold[m] : in->pnew[m];
    int axis, nn;
    if (e <= geo.ncc) {
        axis = 0;
        nn = e;
    }
    else if (e < (geo.ncc + 1) + (2 * geo.ncc + 1)) {
        axis = 1;
        nn = e - (geo.ncc + 1) - geo.ncc;
    }
    else {
        axis = 2;
        nn = e - (geo.ncc + 1) - (2 * geo.ncc + 1) - geo.ncc;
    }
    const double x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    // k component exactly as the k-vector table has it: 2π n / L (src/energy.cpp:158-160)
    const double kc = 2.0 * 3.141592653589793238462643383279502884 * static_cast<double>(nn) / geo.len[axis];
    double sn, cs;
    sincos(kc * x, &sn, &cs);
    const double q = axis == 0 ? p.w : 1.0;
    table[t] = make_double2(q * cs, q * sn);
}

// ------------------------------------------------------------------------------------------------
// pair part: lane ↔ particle j (two per thread, held in registers for the whole window), the block
// walks the 2n move variants (warp-uniform data from shared memory). Pairs beyond `cut2` contribute
// exactly zero, so a variant is reduced across the warp only when some lane found a pair in range.
// The minimum-image fold of an axis is skipped for variants whose position is at least the cutoff away
// from both cell faces on that axis (then no pair in range folds, and an unfolded r² ≥ the folded one
// keeps every other pair out of range) — positions are inside the cell (Geometry::boundary).
// Inactive particles are staged with NaN coordinates, so their r² compares false.
// ------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 128;
constexpr int kPairPerThread = 2;
constexpr int kPairChunk = kPairThreads * kPairPerThread; //!< particles per block
constexpr int kPairVariantsPerBlock = 32;                 //!< variant range of one block (grid.y covers the window)

constexpr int kPairQueue = 256; //!< in-range candidates a warp collects before it evaluates them

template <int KIND, bool DENSE>
__global__ void __launch_bounds__(kPairThreads)
    batchPairKernel(SlotView M0, PotParams P, BatchBuffers cur, double cut2, int stride,
                    double* __restrict__ partials /*[gridDim.x][2·stride]*/, const int* __restrict__ redo = nullptr,
                    SlotView M1 = SlotView{}, BatchBuffers prev = BatchBuffers{}, int apply_commit = 0)
{
    // apply_commit: the accepted moves of the previous window are not in the mirrors yet (no batchPrepKernel ran):
    // every block writes those that fall into ITS particle range, into both mirrors, before it loads the range (the
    // blocks of the other variant ranges write the same values)
    if (apply_commit) {
        const CommitList& commit = cur.in->commit;
        if (static_cast<int>(threadIdx.x) < commit.n) {
            const int m = commit.index[threadIdx.x];
            const int s = prev.in->slot[m];
            if (s >= static_cast<int>(blockIdx.x) * kPairChunk && s < static_cast<int>(blockIdx.x + 1) * kPairChunk) {
                const double4 p = prev.in->pnew[m];
                const int id = prev.in->id[m];
                M0.posq[s] = p;
                M0.atom_id[s] = id;
                M1.posq[s] = p;
                M1.atom_id[s] = id;
            }
        }
        __syncthreads();
    }
    // redo != nullptr: the sums of this window may have been evaluated ahead (fb_run.cuh); then there is nothing to
    // do unless the correction met a cancellation (*redo)
    if (redo != nullptr && cur.in->pair_ready && *redo == 0) {
        return;
    }
    // DENSE: some term has no cutoff, every pair is evaluated in place. Otherwise the few pairs in range are
    // queued per warp (ballot order: deterministic) and evaluated afterwards with all lanes busy.
    __shared__ double4 s_pos[DENSE ? 1 : kPairChunk];
    __shared__ int s_id[DENSE ? 1 : kPairChunk];
    __shared__ double s_qr[DENSE ? 1 : kPairThreads / 32][DENSE ? 1 : kPairQueue];
    __shared__ unsigned short s_qe[DENSE ? 1 : kPairThreads / 32][DENSE ? 1 : kPairQueue];
    __shared__ double4 s_var[2 * kBatchMax];
    __shared__ int s_vid[2 * kBatchMax];
    __shared__ int s_vslot[2 * kBatchMax];
    __shared__ int s_vfold[2 * kBatchMax];
    __shared__ double s_acc[kPairThreads / 32][2 * kBatchMax];

    const int n = cur.in->n;
    const int nv = 2 * n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const double hx = M0.half[0], hy = M0.half[1], hz = M0.half[2];
    const double lx = M0.len_or_zero[0], ly = M0.len_or_zero[1], lz = M0.len_or_zero[2];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    for (int v = threadIdx.x; v < nv; v += kPairThreads) {
        const int m = v >> 1;
        double4 a;
        int id;
        if (v & 1) {
            a = cur.pold[m];
            id = cur.idold[m];
        }
        else {
            a = cur.in->pnew[m];
            id = cur.in->id[m];
        }
        s_var[v] = a;
        s_vid[v] = id;
        s_vslot[v] = cur.in->slot[m];
        // fold needed on an axis unless the variant keeps the cutoff distance from both faces
        const double rc = sqrt(cut2); // +inf when some term has no cutoff
        int fold = 0;
        fold |= (lx > 0.0 && !(fabs(a.x) + rc < hx)) ? 1 : 0;
        fold |= (ly > 0.0 && !(fabs(a.y) + rc < hy)) ? 2 : 0;
        fold |= (lz > 0.0 && !(fabs(a.z) + rc < hz)) ? 4 : 0;
        s_vfold[v] = fold;
    }
    for (int v = threadIdx.x; v < (kPairThreads / 32) * 2 * kBatchMax; v += kPairThreads) {
        (&s_acc[0][0])[v] = 0.0;
    }

    // this thread's particles: j0 = base + tid, j1 = base + 128 + tid (coalesced)
    const int base = blockIdx.x * kPairChunk;
    double4 p[kPairPerThread];
    int pid[kPairPerThread];
    int pj[kPairPerThread];
#pragma unroll
    for (int t = 0; t < kPairPerThread; ++t) {
        const int j = base + t * kPairThreads + threadIdx.x;
        pj[t] = j;
        pid[t] = 0;
        p[t] = make_double4(nan, 0, 0, 0);
        if (j < M0.n_slots) {
            p[t] = M0.posq[j];
            pid[t] = M0.atom_id[j];
            if (M0.gid[j] < 0) {
                p[t].x = nan;
            }
        }
        if (!DENSE) {
            s_pos[t * kPairThreads + threadIdx.x] = p[t];
            s_id[t * kPairThreads + threadIdx.x] = pid[t];
        }
    }
    __syncthreads();

    int queued = 0; // warp-uniform
    // evaluate the queued candidates of this warp: lane ↔ entry, then lane ↔ variant for the ordered sums
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0xffu;
            const int jl = ent >> 8;
            const double4 a = s_var[v];
            s_qr[warp][e] = pairEnergy<KIND>(P, s_vid[v], s_id[jl], a.w, s_pos[jl].w, s_qr[warp][e]);
        }
        __syncwarp();
        const int myv = blockIdx.y * kPairVariantsPerBlock + lane;
        double acc = s_acc[warp][myv]; // one running sum per variant: independent of when the queue is flushed
        for (int e = 0; e < queued; ++e) {
            if (static_cast<int>(s_qe[warp][e] & 0xffu) == myv) {
                acc += s_qr[warp][e];
            }
        }
        s_acc[warp][myv] = acc;
        __syncwarp();
        queued = 0;
    };

    // two variants per iteration (four independent r² chains per lane); blockIdx.y selects the variant range
    const unsigned var_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_var[0]));
    const unsigned fold_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_vfold[0]));
    const int v_begin = blockIdx.y * kPairVariantsPerBlock;
    const int v_end = min(nv, v_begin + kPairVariantsPerBlock);
    for (int v0 = v_begin; v0 < v_end; v0 += 2) {
        double4 a[2];
        int fold[2];
        double r2[2][kPairPerThread];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            loadVariant(var_addr, fold_addr, min(v0 + u, v_end - 1), a[u], fold[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double dx[kPairPerThread], dy[kPairPerThread], dz[kPairPerThread];
#pragma unroll
            for (int t = 0; t < kPairPerThread; ++t) {
                dx[t] = a[u].x - p[t].x;
                dy[t] = a[u].y - p[t].y;
                dz[t] = a[u].z - p[t].z;
            }
            if (fold[u] & 1) {
#pragma unroll
                for (int t = 0; t < kPairPerThread; ++t) {
                    const double ad = fabs(dx[t]);
                    dx[t] = (ad > hx) ? ad - lx : dx[t];
                }
            }
            if (fold[u] & 2) {
#pragma unroll
                for (int t = 0; t < kPairPerThread; ++t) {
                    const double ad = fabs(dy[t]);
                    dy[t] = (ad > hy) ? ad - ly : dy[t];
                }
            }
            if (fold[u] & 4) {
#pragma unroll
                for (int t = 0; t < kPairPerThread; ++t) {
                    const double ad = fabs(dz[t]);
                    dz[t] = (ad > hz) ? ad - lz : dz[t];
                }
            }
#pragma unroll
            for (int t = 0; t < kPairPerThread; ++t) {
                r2[u][t] = dx[t] * dx[t] + dy[t] * dy[t] + dz[t] * dz[t];
            }
        }
        static_assert(kPairPerThread == 2, "anyBelow takes the four r² of one iteration");
        const bool any_in = anyBelow(r2[0][0], r2[0][1], r2[1][0], r2[1][1], cut2);
        if (!DENSE) {
            if (__any_sync(0xffffffffu, any_in)) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int v = v0 + u;
                    if (v >= v_end) {
                        continue;
                    }
                    const int vslot = s_vslot[v];
#pragma unroll
                    for (int t = 0; t < kPairPerThread; ++t) {
                        const bool in = r2[u][t] < cut2 && pj[t] != vslot;
                        const unsigned mask = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int at = queued + __popc(mask & ((1u << lane) - 1u));
                            s_qe[warp][at] = static_cast<unsigned short>(v | ((t * kPairThreads + threadIdx.x) << 8));
                            s_qr[warp][at] = r2[u][t];
                        }
                        queued += __popc(mask);
                    }
                }
                __syncwarp();
                if (queued > kPairQueue - 2 * kPairPerThread * 32) {
                    flush();
                }
            }
        }
        else if (__any_sync(0xffffffffu, any_in)) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = v0 + u;
                bool in_u = false;
#pragma unroll
                for (int t = 0; t < kPairPerThread; ++t) {
                    in_u = in_u || (r2[u][t] < cut2);
                }
                if (v < v_end && __any_sync(0xffffffffu, in_u)) {
                    double e = 0.0;
                    const int vid = s_vid[v];
                    const int vslot = s_vslot[v];
#pragma unroll
                    for (int t = 0; t < kPairPerThread; ++t) {
                        if (r2[u][t] < cut2 && pj[t] != vslot) {
                            e += pairEnergy<KIND>(P, vid, pid[t], s_var[v].w, p[t].w, r2[u][t]);
                        }
                    }
                    e = warpSum(e);
                    if (lane == 0) {
                        s_acc[warp][v] = e;
                    }
                }
            }
        }
    }
    if (!DENSE) {
        flush();
    }
    __syncthreads();
    const int v_store_end = (blockIdx.y + 1 == gridDim.y) ? 2 * stride : v_begin + kPairVariantsPerBlock;
    for (int v = v_begin + threadIdx.x; v < v_store_end; v += kPairThreads) {
        double s = 0.0;
        if (v < nv) {
#pragma unroll
            for (int w = 0; w < kPairThreads / 32; ++w) {
                s += s_acc[w][v];
            }
        }
        partials[static_cast<size_t>(blockIdx.x) * (2 * stride) + v] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// pair part with a finite cutoff: FP32 screening, FP64 evaluation.
// At S1 99.9 % of the 2·B·N pairs of a window lie beyond the cutoff and contribute exactly zero; all the FP64 pipe did
// for them was to find that out (7 of its instructions per pair, and the pipe is what the k-space kernel lives on).
// Here the distance test runs in FP32 on the FMA pipes against a cutoff enlarged by a rigorous bound of the FP32
// rounding error (`cut2_screen`, set by the host): every pair inside the true cutoff passes, a few just outside pass
// too. Candidates are queued per warp in ballot order and evaluated in FP64 from the double-precision positions with
// the reference's minimum-image arithmetic and the exact test r² < cut² — so the energies, and the order in which
// they are added, are those of the FP64 kernel bit for bit. Four particles per lane in registers (12 floats).
// ------------------------------------------------------------------------------------------------
constexpr int kScreenPerThread = 4;
constexpr int kScreenChunk = kPairThreads * kScreenPerThread; //!< particles per block
constexpr int kScreenQueue = 384;                             //!< candidates a warp collects before it evaluates them

template <int KIND>
__global__ void __launch_bounds__(kPairThreads)
    batchPairScreenKernel(SlotView M0, PotParams P, BatchBuffers cur, double cut2, float cut2_screen, int stride,
                          double* __restrict__ partials /*[gridDim.x][2·stride]*/, SlotView M1, BatchBuffers prev,
                          int apply_commit)
{
    __shared__ double s_qr[kPairThreads / 32][kScreenQueue];
    __shared__ unsigned short s_qe[kPairThreads / 32][kScreenQueue]; //!< variant | particle of the block << 7
    __shared__ double4 s_var[2 * kBatchMax];
    __shared__ float4 s_varf[2 * kBatchMax]; //!< x, y, z in FP32, w: fold flags (bit pattern)
    __shared__ int s_vid[2 * kBatchMax];
    __shared__ int s_vslot[2 * kBatchMax];
    __shared__ double s_acc[kPairThreads / 32][2 * kBatchMax];

    // the accepted moves of the previous window that fall into this block's particle range → both mirrors
    if (apply_commit) {
        const CommitList& commit = cur.in->commit;
        if (static_cast<int>(threadIdx.x) < commit.n) {
            const int m = commit.index[threadIdx.x];
            const int s = prev.in->slot[m];
            if (s >= static_cast<int>(blockIdx.x) * kScreenChunk && s < static_cast<int>(blockIdx.x + 1) * kScreenChunk) {
                const double4 p = prev.in->pnew[m];
                const int id = prev.in->id[m];
                M0.posq[s] = p;
                M0.atom_id[s] = id;
                M1.posq[s] = p;
                M1.atom_id[s] = id;
            }
        }
        __syncthreads();
    }
    const int n = cur.in->n;
    const int nv = 2 * n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const float hx = static_cast<float>(M0.half[0]), hy = static_cast<float>(M0.half[1]), hz = static_cast<float>(M0.half[2]);
    const float lx = static_cast<float>(M0.len_or_zero[0]), ly = static_cast<float>(M0.len_or_zero[1]),
                lz = static_cast<float>(M0.len_or_zero[2]);
    const float nanf_ = __int_as_float(0x7fc00000);

    for (int v = threadIdx.x; v < nv; v += kPairThreads) {
        const int m = v >> 1;
        double4 a;
        int id;
        if (v & 1) {
            a = cur.pold[m];
            id = cur.idold[m];
        }
        else {
            a = cur.in->pnew[m];
            id = cur.in->id[m];
        }
        s_var[v] = a;
        s_vid[v] = id;
        s_vslot[v] = cur.in->slot[m];
        // fold needed on an axis unless the variant keeps the (screening) cutoff distance from both faces
        const float rc = sqrtf(cut2_screen) * 1.0001f;
        int fold = 0;
        fold |= (lx > 0.0f && !(fabsf(static_cast<float>(a.x)) + rc < hx)) ? 1 : 0;
        fold |= (ly > 0.0f && !(fabsf(static_cast<float>(a.y)) + rc < hy)) ? 2 : 0;
        fold |= (lz > 0.0f && !(fabsf(static_cast<float>(a.z)) + rc < hz)) ? 4 : 0;
        s_varf[v] = make_float4(static_cast<float>(a.x), static_cast<float>(a.y), static_cast<float>(a.z), __int_as_float(fold));
    }
    for (int v = threadIdx.x; v < (kPairThreads / 32) * 2 * kBatchMax; v += kPairThreads) {
        (&s_acc[0][0])[v] = 0.0;
    }
    // this thread's particles: j = base + t·128 + tid (coalesced); inactive ones are NaN and never in range
    const int base = blockIdx.x * kScreenChunk;
    float px[kScreenPerThread], py[kScreenPerThread], pz[kScreenPerThread];
#pragma unroll
    for (int t = 0; t < kScreenPerThread; ++t) {
        const int jl = t * kPairThreads + threadIdx.x;
        const int j = base + jl;
        double4 p = make_double4(0, 0, 0, 0);
        bool active = false;
        if (j < M0.n_slots) {
            p = M0.posq[j];
            active = M0.gid[j] >= 0;
        }
        px[t] = active ? static_cast<float>(p.x) : nanf_;
        py[t] = static_cast<float>(p.y);
        pz[t] = static_cast<float>(p.z);
    }
    __syncthreads();

    int queued = 0; // warp-uniform
    // evaluate the queued candidates of this warp in FP64: lane ↔ entry, then lane ↔ variant for the ordered sums
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0x7fu;
            const int jl = ent >> 7;
            const double4 a = s_var[v];
            const double4 b = M0.posq[base + jl]; // the candidates are few: double-precision positions from L2
            const double r2 = minImageR2(M0, a.x, a.y, a.z, b.x, b.y, b.z);
            double u = 0.0;
            if (r2 < cut2 && base + jl != s_vslot[v]) {
                u = pairEnergy<KIND>(P, s_vid[v], M0.atom_id[base + jl], a.w, b.w, r2);
            }
            s_qr[warp][e] = u;
        }
        __syncwarp();
        const int myv = blockIdx.y * kPairVariantsPerBlock + lane;
        double acc = s_acc[warp][myv]; // one running sum per variant: independent of when the queue is flushed
        for (int e = 0; e < queued; ++e) {
            if (static_cast<int>(s_qe[warp][e] & 0x7fu) == myv) {
                acc += s_qr[warp][e];
            }
        }
        s_acc[warp][myv] = acc;
        __syncwarp();
        queued = 0;
    };

    const int v_begin = blockIdx.y * kPairVariantsPerBlock;
    const int v_end = min(nv, v_begin + kPairVariantsPerBlock);
    for (int v0 = v_begin; v0 < v_end; v0 += 2) {
        float r2[2][kScreenPerThread];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float4 a = s_varf[min(v0 + u, v_end - 1)];
            const int fold = __float_as_int(a.w);
            float dx[kScreenPerThread], dy[kScreenPerThread], dz[kScreenPerThread];
#pragma unroll
            for (int t = 0; t < kScreenPerThread; ++t) {
                dx[t] = a.x - px[t];
                dy[t] = a.y - py[t];
                dz[t] = a.z - pz[t];
            }
            if (fold & 1) {
#pragma unroll
                for (int t = 0; t < kScreenPerThread; ++t) {
                    const float ad = fabsf(dx[t]);
                    dx[t] = (ad > hx) ? ad - lx : dx[t];
                }
            }
            if (fold & 2) {
#pragma unroll
                for (int t = 0; t < kScreenPerThread; ++t) {
                    const float ad = fabsf(dy[t]);
                    dy[t] = (ad > hy) ? ad - ly : dy[t];
                }
            }
            if (fold & 4) {
#pragma unroll
                for (int t = 0; t < kScreenPerThread; ++t) {
                    const float ad = fabsf(dz[t]);
                    dz[t] = (ad > hz) ? ad - lz : dz[t];
                }
            }
#pragma unroll
            for (int t = 0; t < kScreenPerThread; ++t) {
                r2[u][t] = fmaf(dz[t], dz[t], fmaf(dy[t], dy[t], dx[t] * dx[t]));
            }
        }
        bool any_in = false;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int t = 0; t < kScreenPerThread; ++t) {
                any_in = any_in || (r2[u][t] < cut2_screen);
            }
        }
        if (__any_sync(0xffffffffu, any_in)) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = v0 + u;
                if (v >= v_end) {
                    continue;
                }
#pragma unroll
                for (int t = 0; t < kScreenPerThread; ++t) {
                    const bool in = r2[u][t] < cut2_screen;
                    const unsigned mask = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int at = queued + __popc(mask & ((1u << lane) - 1u));
                        s_qe[warp][at] = static_cast<unsigned short>(v | ((t * kPairThreads + threadIdx.x) << 7));
                    }
                    queued += __popc(mask);
                }
            }
            __syncwarp();
            if (queued > kScreenQueue - 2 * kScreenPerThread * 32) {
                flush();
            }
        }
    }
    flush();
    __syncthreads();
    const int v_store_end = (blockIdx.y + 1 == gridDim.y) ? 2 * stride : v_begin + kPairVariantsPerBlock;
    for (int v = v_begin + threadIdx.x; v < v_store_end; v += kPairThreads) {
        double s = 0.0;
        if (v < nv) {
#pragma unroll
            for (int w = 0; w < kPairThreads / 32; ++w) {
                s += s_acc[w][v];
            }
        }
        partials[static_cast<size_t>(blockIdx.x) * (2 * stride) + v] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// group mode: rigid-molecule moves (TranslateRotate, src/move.cpp:1670-1689: Change {group, all, internal =
// false}). Energy of the moved group with every other group (group2all → group2group with the mass-centre
// cutoff, src/energy.h:1155-1163, 979-989, 761-768). One block per (move, new | old); threads ↔ other groups.
// ------------------------------------------------------------------------------------------------
constexpr int kGroupMoveAtoms = 8;

template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchPairGroupKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, double* __restrict__ result)
{
    __shared__ double4 s_atom[kGroupMoveAtoms];
    __shared__ int s_id[kGroupMoveAtoms];
    __shared__ double scratch[kBlock / 32];
    const int g = blockIdx.x >> 1;
    const bool is_old = blockIdx.x & 1;
    const int first = cur.in->move_first[g];
    const int na = cur.in->move_natoms[g];
    const int my_group = cur.in->move_group[g];
    if (static_cast<int>(threadIdx.x) < na) {
        s_atom[threadIdx.x] = is_old ? cur.pold[first + threadIdx.x] : cur.in->pnew[first + threadIdx.x];
        s_id[threadIdx.x] = is_old ? cur.idold[first + threadIdx.x] : cur.in->id[first + threadIdx.x];
    }
    const double4 cm = is_old ? cur.in->cm_old[g] : cur.in->cm_new[g];
    const int my_info = M0.ginfo[my_group];
    __syncthreads();
    double e = 0.0;
    for (int gg = threadIdx.x; gg < M0.n_groups; gg += kBlock) {
        const int size = M0.gsize[gg];
        if (gg == my_group || size == 0) {
            continue;
        }
        const int info = M0.ginfo[gg];
        if (!((info | my_info) & MOL_ATOMIC)) { // GroupCutoff::cut: both molecular
            const double4 c2 = M0.gcm[gg];
            const double r2 = minImageR2(M0, cm.x, cm.y, cm.z, c2.x, c2.y, c2.z);
            if (r2 >= __ldg(P.g2g_cut2 + (my_info >> 8) * P.n_mol + (info >> 8))) {
                continue;
            }
        }
        const int begin = M0.gbegin[gg];
        for (int j = begin; j < begin + size; ++j) {
            const double4 pj = M0.posq[j];
            const int idj = M0.atom_id[j];
            for (int i = 0; i < na; ++i) {
                const double4 a = s_atom[i];
                e += pairEnergy<KIND>(P, s_id[i], idj, a.w, pj.w, minImageR2(M0, a.x, a.y, a.z, pj.x, pj.y, pj.z));
            }
        }
    }
    const double sum = blockSum<kBlock>(e, scratch);
    if (threadIdx.x == 0) {
        result[8 + (is_old ? stride : 0) + g] = sum;
    }
}

/** group-group energy U(m_x, a_y) of two moved groups (x, y ∈ {new, old}) with the mass-centre cutoff; one warp */
template <int KIND>
__device__ __forceinline__ double groupPairEnergy(const SlotView& M0, const PotParams& P, const BatchBuffers& cur, int m,
                                                  bool m_new, int a, bool a_new, int lane)
{
    const int info_m = M0.ginfo[cur.in->move_group[m]];
    const int info_a = M0.ginfo[cur.in->move_group[a]];
    if (!((info_m | info_a) & MOL_ATOMIC)) {
        const double4 cm = m_new ? cur.in->cm_new[m] : cur.in->cm_old[m];
        const double4 ca = a_new ? cur.in->cm_new[a] : cur.in->cm_old[a];
        if (minImageR2(M0, cm.x, cm.y, cm.z, ca.x, ca.y, ca.z) >= __ldg(P.g2g_cut2 + (info_m >> 8) * P.n_mol + (info_a >> 8))) {
            return 0.0;
        }
    }
    const int nm = cur.in->move_natoms[m], na = cur.in->move_natoms[a];
    const int fm = cur.in->move_first[m], fa = cur.in->move_first[a];
    double e = 0.0;
    for (int t = lane; t < nm * na; t += 32) {
        const int i = fm + t / na;
        const int j = fa + t % na;
        const double4 pi = m_new ? cur.in->pnew[i] : cur.pold[i];
        const int idi = m_new ? cur.in->id[i] : cur.idold[i];
        const double4 pj = a_new ? cur.in->pnew[j] : cur.pold[j];
        const int idj = a_new ? cur.in->id[j] : cur.idold[j];
        e += pairEnergy<KIND>(P, idi, idj, pi.w, pj.w, minImageR2(M0, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z));
    }
    return warpSum(e); // valid in lane 0
}

/** group mode: cross terms between the moves of a window, one warp per (a, m); zero padding of the pair sums */
template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchPairGroupFinishKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, double* __restrict__ result)
{
    const int ng = cur.in->n_groups;
    const int S = stride;
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    double* u = result + 8;
    double* cross = result + 8 + 3 * S;
    if (w == 0 && lane == 0) {
        result[2] = 0.0;
    }
    if (w < 2 * S) {
        const int m = w % S;
        if (m >= ng && lane == 0) {
            u[(w / S) * S + m] = 0.0;
        }
    }
    else if (w < 2 * S + S * S) {
        const int t = w - 2 * S;
        const int a = t / S;
        const int m = t % S;
        double cn = 0.0, co = 0.0, cmax = 0.0;
        if (a < m && m < ng) {
            const double t1 = groupPairEnergy<KIND>(M0, P, cur, m, true, a, true, lane);
            const double t2 = groupPairEnergy<KIND>(M0, P, cur, m, true, a, false, lane);
            const double t3 = groupPairEnergy<KIND>(M0, P, cur, m, false, a, true, lane);
            const double t4 = groupPairEnergy<KIND>(M0, P, cur, m, false, a, false, lane);
            cn = t1 - t2;
            co = t3 - t4;
            cmax = fmax(fmax(fabs(t1), fabs(t2)), fmax(fabs(t3), fabs(t4)));
            if (t1 != t1 || t2 != t2 || t3 != t3 || t4 != t4) {
                cmax = __longlong_as_double(0x7ff0000000000000LL);
            }
        }
        if (lane == 0) {
            const int tt = m * S + a; // stored [m][a]
            cross[tt] = cn;
            cross[S * S + tt] = co;
            cross[2 * S * S + tt] = cmax;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k-vector cells (used by the full rebuild of Q(k), fb_stream.cuh; the window kernels are in fb_kspace.cuh)
// ------------------------------------------------------------------------------------------------
constexpr int kTileK = 64; //!< k-vectors per cell

/**
 * The k-vectors are stored cell by cell: a cell is the 4×4×4 block of integer triplets with the same
 * (nx >> 2, (ny + ncc) >> 2, (nz + ncc) >> 2), ≤ 64 k-vectors that need only 12 phase-table entries per
 * position. One block per cell stages those entries in shared memory once.
 */
constexpr int kCellEntries = 12;

struct CellBase
{
    int nx, ny, nz; //!< smallest triplet component of the cell
};

__device__ __forceinline__ CellBase cellBase(const int4& first, int ncc)
{
    CellBase b;
    b.nx = first.x & ~3;
    b.ny = ((first.y + ncc) & ~3) - ncc;
    b.nz = ((first.z + ncc) & ~3) - ncc;
    return b;
}

/** s_tab[v][0..4) = e^{i kx x}, [4..8) = e^{i ky y}, [8..12) = e^{i kz z} of the cell's triplet range */
__device__ __forceinline__ void stageCellTables(double2 (*s_tab)[kCellEntries], const double2* __restrict__ table,
                                                const int* table_of_variant, int n_variants, const CellBase& base,
                                                const PhaseGeometry& geo)
{
    for (int e = threadIdx.x; e < n_variants * kCellEntries; e += blockDim.x) {
        const int v = e / kCellEntries;
        const int t = e % kCellEntries;
        const int i = t & 3;
        int offset;
        if (t < 4) {
            offset = min(base.nx + i, geo.ncc);
        }
        else if (t < 8) {
            offset = (geo.ncc + 1) + min(base.ny + i, geo.ncc) + geo.ncc;
        }
        else {
            offset = (geo.ncc + 1) + (2 * geo.ncc + 1) + min(base.nz + i, geo.ncc) + geo.ncc;
        }
        const int first = table_of_variant ? table_of_variant[v] : v * geo.table_stride;
        s_tab[v][t] = __ldg(table + first + offset);
    }
}

__device__ __forceinline__ double2 cellPhase(const double2* t, int li, int lj, int ll)
{
    return cmul(cmul(t[li], t[4 + lj]), t[8 + ll]);
}

// cp.async (LDGSTS) helpers of the k-space kernels (fb_kspace.cuh)
__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem)
{
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cpAsync8(void* smem, const void* gmem)
{
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;\n" ::); }

// ------------------------------------------------------------------------------------------------
// final ordered sums + the pair cross terms
// result layout (doubles), S = stride:
//   [0] Σ_k A_k|Q_k|² at window start   [1] n
//   [8 + 0·S ..) u_new   [8 + 1·S ..) u_old   [8 + 2·S ..) R
//   [8 + 3·S + 0·S² ..) C_new   [+1·S²) C_old   [+2·S²) C_max   [+3·S²) G      each stored [m][a], a < m
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t batchResultDoubles(int stride)
{
    return 8 + 3 * static_cast<size_t>(stride) + 4 * static_cast<size_t>(stride) * stride;
}

/** Σ over `rows` of column `col` of a row-major [rows][ld] array by one warp (fixed order); valid in lane 0 */
__device__ __forceinline__ double warpColumnSum(const double* __restrict__ a, int rows, size_t ld, int col, int lane)
{
    double s = 0.0;
#pragma unroll 8
    for (int b = lane; b < rows; b += 32) {
        s += __ldcg(a + static_cast<size_t>(b) * ld + col);
    }
    return warpSum(s);
}

/**
 * Runs: the pair sums of this window were evaluated one window ahead (batchPairKernel on the predicted window,
 * overlapping the k-space kernel of the window before), i.e. against the state in which the accepted moves of the
 * window before — this window's commit list — were not applied yet. One warp per variant v adds up
 *     fix[v] = Σ_a [ u(x_v, new_a) − u(x_v, old_a) ]      (a: accepted moves of the previous window, in order)
 * — the same identity as the cross terms inside a window. A pair energy ≥ `limit` (or not finite) in there would
 * cancel badly: then *redo is raised and the gated batchPairKernel evaluates the window from scratch.
 */
template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchPairFixKernel(SlotView M0, PotParams P, BatchBuffers cur, BatchBuffers prev, double limit,
                       double* __restrict__ fix /*[2·kBatchMax]*/, int* __restrict__ redo)
{
    __shared__ double s_delta[kBlock / 32][kBatchMax];
    if (!cur.in->pair_ready) {
        return;
    }
    const int n = cur.in->n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int v = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    if (v >= 2 * n) {
        return;
    }
    const int m = v >> 1;
    const double4 x = (v & 1) ? cur.pold[m] : cur.in->pnew[m];
    const int id = (v & 1) ? cur.idold[m] : cur.in->id[m];
    const int nc = cur.in->commit.n;
    bool bad = false;
    for (int k = lane; k < nc; k += 32) {
        const int a = cur.in->commit.index[k];
        const double4 pn = prev.in->pnew[a];
        const double4 po = prev.pold[a];
        const double un = pairEnergy<KIND>(P, id, prev.in->id[a], x.w, pn.w, minImageR2(M0, x.x, x.y, x.z, pn.x, pn.y, pn.z));
        const double uo = pairEnergy<KIND>(P, id, prev.idold[a], x.w, po.w, minImageR2(M0, x.x, x.y, x.z, po.x, po.y, po.z));
        bad = bad || !(fabs(un) < limit) || !(fabs(uo) < limit);
        s_delta[warp][k] = un - uo;
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, bad) && lane == 0) {
        atomicOr(redo, 1);
    }
    if (lane == 0) {
        double s = 0.0;
        for (int k = 0; k < nc; ++k) {
            s += s_delta[warp][k];
        }
        fix[v] = s;
    }
}

/** pair side, one warp (number `w`) per output: 2S pair sums and the S² cross entries (4 pair energies each, a < m) */
template <int KIND>
__device__ __forceinline__ void pairFinishWarp(const SlotView& M0, const PotParams& P, const BatchBuffers& cur, int stride,
                                               int n_pair_blocks, const double* __restrict__ pair_partials, int sums_done,
                                               const int* __restrict__ cell_overflow, double* __restrict__ result,
                                               const double* __restrict__ fix, const int* __restrict__ redo, int w)
{
    const int n = cur.in->n;
    const int S = stride;
    const int lane = threadIdx.x & 31;
    double* u = result + 8;
    double* cross = result + 8 + 3 * S;
    if (w == 0 && lane == 0) { // the cell list ran out of bucket space: the caller re-runs the window brute force
        result[2] = (cell_overflow != nullptr && *cell_overflow != 0) ? 1.0 : 0.0;
    }
    if (w < 2 * S) { // variant order new/old interleaved → split
        double s = 0.0;
        if (w < 2 * n && !sums_done) {
            s = warpColumnSum(pair_partials, n_pair_blocks, 2 * static_cast<size_t>(S), w, lane);
            if (fix != nullptr && cur.in->pair_ready && *redo == 0) { // sums taken one window ahead: what the accepted
                s += fix[w];                                         // moves of the window before changed
            }
        }
        if (lane == 0 && !(sums_done && w < 2 * n)) { // with the cell list the sums are already in place
            u[(w & 1) * S + (w >> 1)] = s;
        }
    }
    else if (w < 2 * S + S * S) {
        const int t = w - 2 * S;
        const int a = t / S;
        const int m = t % S;
        double cn = 0.0, co = 0.0, cmax = 0.0;
        if (a < m && m < n) {
            // how the energies of move m change when the earlier move a has been accepted: lanes 0..3 take
            // the four pair energies u(new_m|old_m , new_a|old_a)
            double term = 0.0;
            if (lane < 4) {
                const bool m_new = lane < 2;
                const bool a_new = (lane & 1) == 0;
                const double4 pm = m_new ? cur.in->pnew[m] : cur.pold[m];
                const int idm = m_new ? cur.in->id[m] : cur.idold[m];
                const double4 pa = a_new ? cur.in->pnew[a] : cur.pold[a];
                const int ida = a_new ? cur.in->id[a] : cur.idold[a];
                term = pairEnergy<KIND>(P, idm, ida, pm.w, pa.w, minImageR2(M0, pm.x, pm.y, pm.z, pa.x, pa.y, pa.z));
            }
            const double t1 = __shfl_sync(0xffffffffu, term, 0);
            const double t2 = __shfl_sync(0xffffffffu, term, 1);
            const double t3 = __shfl_sync(0xffffffffu, term, 2);
            const double t4 = __shfl_sync(0xffffffffu, term, 3);
            cn = t1 - t2;
            co = t3 - t4;
            cmax = fmax(fmax(fabs(t1), fabs(t2)), fmax(fabs(t3), fabs(t4)));
            if (t1 != t1 || t2 != t2 || t3 != t3 || t4 != t4) {
                cmax = __longlong_as_double(0x7ff0000000000000LL);
            }
        }
        if (lane == 0) {
            const int tt = m * S + a; // stored [m][a]: the caller reads one row per move
            cross[tt] = cn;
            cross[S * S + tt] = co;
            cross[2 * S * S + tt] = cmax;
        }
    }
}

#ifndef FB_CROSS_PACKED
#define FB_CROSS_PACKED 1
#endif
/**
 * Cross terms, 16 per warp: the lane pair (2p, 2p + 1) of warp `wc` takes the pair of moves t = 16 wc + p = a S + m; the even
 * lane evaluates move m at its TRIAL position against the trial and the old position of move a, the odd lane move m at its
 * OLD position — the same four pair energies, the same two differences and the same maximum as pairFinishWarp (which used 4
 * of the 32 lanes of a warp per pair: 4096 warps, 128 blocks of the tail kernel, a second wave; now 256 warps, 8 blocks).
 */
template <int KIND>
__device__ __forceinline__ void pairCrossPacked(const SlotView& M0, const PotParams& P, const BatchBuffers& cur, int stride,
                                                double* __restrict__ result, int wc)
{
    const int n = cur.in->n;
    const int S = stride;
    const int lane = threadIdx.x & 31;
    double* cross = result + 8 + 3 * S;
    const int t = 16 * wc + (lane >> 1);
    if (t >= S * S) {
        return; // (whole lane pairs leave together; S² is a multiple of 16)
    }
    const int a = t / S;
    const int m = t % S;
    const bool m_new = (lane & 1) == 0;
    double diff = 0.0, largest = 0.0;
    bool invalid = false;
    const bool wanted = a < m && m < n;
    if (wanted) {
        const double4 pm = m_new ? cur.in->pnew[m] : cur.pold[m];
        const int idm = m_new ? cur.in->id[m] : cur.idold[m];
        const double4 pn = cur.in->pnew[a];
        const double4 po = cur.pold[a];
        const double with_new = pairEnergy<KIND>(P, idm, cur.in->id[a], pm.w, pn.w, minImageR2(M0, pm.x, pm.y, pm.z, pn.x, pn.y, pn.z));
        const double with_old = pairEnergy<KIND>(P, idm, cur.idold[a], pm.w, po.w, minImageR2(M0, pm.x, pm.y, pm.z, po.x, po.y, po.z));
        diff = with_new - with_old;
        largest = fmax(fabs(with_new), fabs(with_old));
        invalid = with_new != with_new || with_old != with_old;
    }
    const double other_largest = __shfl_xor_sync(0xffffffffu, largest, 1);
    const bool other_invalid = __shfl_xor_sync(0xffffffffu, invalid ? 1 : 0, 1) != 0;
    double cmax = fmax(largest, other_largest);
    if (invalid || other_invalid) {
        cmax = __longlong_as_double(0x7ff0000000000000LL);
    }
    const int tt = m * S + a; // stored [m][a]: the caller reads one row per move
    if (m_new) {
        cross[tt] = diff;
        cross[2 * S * S + tt] = wanted ? cmax : 0.0;
    }
    else {
        cross[S * S + tt] = diff;
    }
}

#ifndef FB_CROSS_EARLY
#define FB_CROSS_EARLY 0
#endif
/**
 * The cross terms of a window (S² entries, four pair energies each) need nothing but the window's own trial and old
 * positions — not the pair sums, not the mirror. In runs they are taken at the START of the window on the pair stream,
 * beside the front kernel, instead of by 128 more blocks of the tail kernel behind the k-space kernel: the tail is left
 * with 131 + 4 blocks, one wave (it was 263 blocks of 1024 threads on 148 SMs). MEASURED AND SWITCHED OFF (FB_CROSS_EARLY 0):
 * 8.71e5 against 8.98e5 moves/s at S1 — the extra kernel competes with the front kernel, which is on the critical path,
 * for the start of the window, and the tail's second wave was not what bounds it.
 */
template <int KIND>
__global__ void __launch_bounds__(kBlock) windowCrossKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, double* __restrict__ result)
{
    const int task = static_cast<int>((blockIdx.x * kBlock + threadIdx.x) >> 5);
    if (task < stride * stride) {
        pairFinishWarp<KIND>(M0, P, cur, stride, 0, nullptr, 0, nullptr, result, nullptr, nullptr, 2 * stride + task);
    }
}

template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchPairFinishKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, int n_pair_blocks,
                          const double* __restrict__ pair_partials, int sums_done, const int* __restrict__ cell_overflow,
                          double* __restrict__ result, const double* __restrict__ fix = nullptr,
                          const int* __restrict__ redo = nullptr)
{
    pairFinishWarp<KIND>(M0, P, cur, stride, n_pair_blocks, pair_partials, sums_done, cell_overflow, result, fix, redo,
                         static_cast<int>((blockIdx.x * kBlock + threadIdx.x) >> 5));
}

/**
 * k-space side: ordered sums of the per-block partials R[rows][S], G[rows][S²], E[rows] → the result block.
 * Coalesced: a block takes 32 adjacent columns (lane ↔ column), its 32 warps take the rows w, w + 32, … (all loads
 * of a warp in flight at once), the 32 partial sums per column are added in warp order. (One warp per column
 * striding over the rows — the first version — touched a different 32-byte sector with every load: 39 MB of L2
 * traffic for 9.8 MB of data, 9 µs.)
 */
constexpr int kFinishThreads = 1024;

__host__ __device__ inline int kspaceFinishGrid(int stride) { return stride * stride / 32 + (stride + 31) / 32 + 1; }

__device__ __forceinline__ void kspaceFinishBlock(const BatchBuffers& cur, int stride, int with_ewald, int n_rows, int n_e_rows,
                                                  const double* __restrict__ r_partials,
                                                  const double* __restrict__ g_partials,
                                                  const double* __restrict__ e_partials, double* __restrict__ result,
                                                  int block)
{
    __shared__ double s_part[kFinishThreads / 32][33];
    const int n = cur.in->n;
    const int S = stride;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g_blocks = S * S / 32;
    const int r_blocks = (S + 31) / 32;
    const double* src;
    size_t ld;
    int col;
    bool wanted; // is this output defined (else it is written as zero)?
    if (block < g_blocks) {
        col = block * 32 + lane; // = a·S + m
        src = g_partials;
        ld = static_cast<size_t>(S) * S;
        const int a = col / S, m = col - a * S;
        wanted = a < m && m < n;
    }
    else if (block < g_blocks + r_blocks) {
        col = (block - g_blocks) * 32 + lane;
        src = r_partials;
        ld = static_cast<size_t>(S);
        wanted = col < n;
    }
    else {
        col = 0;
        src = e_partials;
        ld = 1;
        wanted = true;
    }
    double s = 0.0;
    if (block >= g_blocks + r_blocks) {
        // Σ A_k|Q_k|²: one partial per unit (thousands of rows, one column): thread ↔ rows t, t + 1024, …
        if (with_ewald) {
            for (int row = threadIdx.x; row < n_e_rows; row += kFinishThreads) {
                s += __ldcg(src + row);
            }
        }
    }
    else if (with_ewald && wanted) {
#pragma unroll 10
        for (int row = warp; row < n_rows; row += kFinishThreads / 32) {
            s += __ldcg(src + static_cast<size_t>(row) * ld + col);
        }
    }
    s_part[warp][lane] = s;
    __syncthreads();
    if (warp != 0) {
        return;
    }
    double total = 0.0;
#pragma unroll
    for (int w = 0; w < kFinishThreads / 32; ++w) {
        total += s_part[w][lane];
    }
    double* u = result + 8;
    double* cross = result + 8 + 3 * S;
    if (block < g_blocks) {
        const int a = col / S, m = col - a * S;
        cross[3 * S * S + m * S + a] = wanted ? total : 0.0; // stored [m][a]
    }
    else if (block < g_blocks + r_blocks) {
        if (col < S) {
            u[2 * S + col] = wanted ? total : 0.0;
        }
    }
    else {
        total = warpSum(total); // the 32 lane sums in the fixed order of the shuffle tree
        if (lane == 0) {
            result[0] = (with_ewald && n_e_rows > 0) ? total : 0.0;
            result[1] = static_cast<double>(n);
        }
    }
}

__global__ void __launch_bounds__(kFinishThreads)
    batchKspaceFinishKernel(BatchBuffers cur, int stride, int with_ewald, int n_rows, int n_e_rows,
                            const double* __restrict__ r_partials, const double* __restrict__ g_partials,
                            const double* __restrict__ e_partials, double* __restrict__ result)
{
    kspaceFinishBlock(cur, stride, with_ewald, n_rows, n_e_rows, r_partials, g_partials, e_partials, result, blockIdx.x);
}

inline int pairFinishBlocks(int stride) { return (2 * stride + stride * stride + kFinishThreads / 32 - 1) / (kFinishThreads / 32); }

/**
 * Everything between the evaluation of a window and its walk in ONE launch: blocks [0, F) the ordered sums of the
 * k-space partials, the others one warp per pair output (pair sums, cross terms). As two kernels on two streams the
 * pair side could not start before the persistent k-space kernel had drained (it owns every register file): 11 µs
 * in series.
 */
template <int KIND>
__global__ void __launch_bounds__(kFinishThreads)
    windowFinishKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, int with_ewald, int n_rows, int n_e_rows,
                       const double* __restrict__ r_partials, const double* __restrict__ g_partials,
                       const double* __restrict__ e_partials, int n_pair_blocks, const double* __restrict__ pair_partials,
                       int sums_done, const int* __restrict__ cell_overflow, double* __restrict__ result,
                       const double* __restrict__ fix, const int* __restrict__ redo)
{
    const int k_blocks = kspaceFinishGrid(stride);
    if (static_cast<int>(blockIdx.x) < k_blocks) {
        kspaceFinishBlock(cur, stride, with_ewald, n_rows, n_e_rows, r_partials, g_partials, e_partials, result, blockIdx.x);
    }
    else {
        pairFinishWarp<KIND>(M0, P, cur, stride, n_pair_blocks, pair_partials, sums_done, cell_overflow, result, fix, redo,
                             static_cast<int>(((blockIdx.x - k_blocks) * kFinishThreads + threadIdx.x) >> 5));
    }
}

} // namespace fbdev

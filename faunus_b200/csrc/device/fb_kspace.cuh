// k-space part of a window (sm_100a): R[m], G[a][m], Σ_k A_k |Q_k|² and the commit of the previous window's
// accepted moves into Q(k) — the quantities defined at the top of fb_batch.cuh.
//
// Work unit = HALF of a 4×4×4 cell of integer triplets (the two y-rows ly ∈ {2h, 2h+1}): 32 k-slots
// ks = 8·i + 4·jj + l that need only 10 phase-table entries per position (4 x, 2 y, 4 z). Per-unit data is kept in
// SLOT layout ([unit][32], zero where the sphere cut leaves no k-vector), so nothing below tests for validity.
//
// windowFrontKernel   blocks [0, P): the per-axis phase tables of the 2n positions of the window (sincos; the x
//                     entries carry the charge); the other blocks: ΔQ of the accepted moves of the previous window
//                     as a small complex matrix product on the FP64 tensor path, one block per item (two z-adjacent
//                     cells = four units; a warp per 32 positions), lane ↔ slot; Q(k) updated in place, √A_k·Q_k in slot layout for the
//                     kernel below, Σ A_k |Q_k|².
// windowKspaceKernel  persistent, block b walks the units b, b + grid, …  Thread ↔ (move m = tid / 4, x-index
//                     i = tid % 4); warp w owns the moves 8w … 8w+7. A thread forms the 8 phases e^{ik·r} of its
//                     x-index from registers — xy_j = X_i·Y_j once per y, then xy_j·Z_l — for the trial and the old
//                     position of ITS move: 14 FP64 instructions per (k, move) (lane ↔ k needed 28 and six LDS.128).
//                     The 16 doubles √A_k·δ_m,k it ends up with ARE the m8n8k4 fragments of the Gram update: with
//                     lane = 4·(m % 8) + i, lane (g, t) of warp w holds D[8w + g][column (t, s)] — the A fragment
//                     A[g][t] of row block w and the B fragment B[t][g] of column block w for reduction step s (the
//                     order of the reduction columns is free). The diagonal tile of a warp needs no memory at all;
//                     for an off-diagonal tile the partner warp's fragments come from shared memory, stored lane by
//                     lane (conflict-free LDS.128, two steps per load). Tiles are dealt as a circulant: warp w takes
//                     (w, w) and (w, w+d mod 8), d = 1, 2, 3 (and d = 4 for w < 4): 4.5 DMMA per fragment load.
//                     Σ_k A_k |δ_m,k|² is the diagonal of G; R keeps only the 2 Re(conj(Q) δ) part.
//                     Pipeline: every WARP stages its own tables (cp.async, no block barrier), the fragments are
//                     double-buffered, so there is ONE block barrier per unit and a warp that is through with its
//                     tiles starts on the phases of the next unit while the others still feed the tensor pipe.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

constexpr int kUnitSlots = 32;    //!< k-slots of a unit: ks = 8·i + 4·jj + l
constexpr int kUnitEntries = 10;  //!< phase-table entries per position: X0..3, Y0..1, Z0..3
constexpr int kKsGroupThreads = 256;           //!< one work group: 64 moves × 4 x-indices
#ifndef FB_KS_GROUPS
#define FB_KS_GROUPS 1
#endif
constexpr int kKsGroups = FB_KS_GROUPS;        //!< 1: two blocks per SM; 2: one block of two groups per SM whose sums are added
                                               //!< before they leave (half the partial rows; measured: k-space kernel
                                               //!< +1.4 µs, tail −0.9 µs per window at S1 — kept as a switch, off)
constexpr int kKsThreads = kKsGroups * kKsGroupThreads;
constexpr int kKsWarps = kKsGroupThreads / 32; //!< warps of a group

/** skewed slot index: the four x-indices of a warp land in different bank groups */
__device__ __forceinline__ int unitSlotIndex(int i, int s) { return 9 * i + s; }

/** table index of entry t (X0..3, Y0..1, Z0..3) of a unit; info = {first k of the cell, x, y, z index of the first slot} */
__device__ __forceinline__ int unitTableOffset(const int4& info, int t, int ncc)
{
    const int base = t < 4 ? 0 : (t < 6 ? ncc + 1 : (ncc + 1) + (2 * ncc + 1));
    const int first = t < 4 ? info.y : (t < 6 ? info.z : info.w);
    const int local = t < 4 ? t : (t < 6 ? t - 4 : t - 6);
    const int limit = t < 4 ? ncc : 2 * ncc;
    return base + min(first + local, limit); // entries beyond the table belong to slots without a k-vector
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// front kernel: phase tables of this window + commit of the previous one
// ------------------------------------------------------------------------------------------------
constexpr int kFrontThreads = 128;
constexpr int kFrontWarps = kFrontThreads / 32;
constexpr int kItemEntries = 16; //!< table entries per position and item: X0..3, Y0..3, Z0..3 of cell 0, Z0..3 of cell 1
constexpr int kItemStride = 17;  //!< … stored with a stride of 272 B: the four positions of a step in different banks

inline int frontPhaseBlocks(const PhaseGeometry& geo) { return (2 * kBatchMax * geo.table_stride + kFrontThreads - 1) / kFrontThreads; }

#ifndef FB_FRONT_BLOCKS_PER_SM
#define FB_FRONT_BLOCKS_PER_SM 4
#endif
struct FrontSmem
{
    double2 stage[kFrontWarps][32][kItemStride]; //!< [warp][position of the warp's batch][entry]
#if FB_FRONT_BLOCKS_PER_SM <= 4
    double2 part[kFrontWarps][4][32];             //!< [warp][unit of the item][slot]: ΔQ share of the warp's positions
#endif
};

/**
 * Blocks [0, P): phase tables. Blocks [P, …): the commit of the previous window as a small complex matrix product on
 * the FP64 tensor path, ONE BLOCK PER ITEM = two z-adjacent cells = up to four units (two y-halves × two cells).
 * With the accepted positions p (trial: +, old: −; the x entries carry the charges) and a unit's slots
 * ks = 8 xi + 4 jj + l,
 *
 *     ΔQ[xi, jj, l] = Σ_p  ± X_p[xi] Y_p[jj] · Z_p[l]
 *
 * is C = A·B with A[row = 2 xi + jj][p] = ± X_p[xi] Y_p[jj] (one complex product per lane and step) and
 * B[p][l] = Z_p[l]. m8n8k4 takes 4 positions per step: lane (g, t) supplies A[g][p = 4 step + t] and, as column
 * g = 2 l + c of the real 4 × 8 matrices B' = [Z_re | Z_im] and B'' = [−Z_im | Z_re], the two B fragments, so that
 * Re(A)·B' + Im(A)·B'' leaves (Re, Im) of ΔQ[row g][l = t] in the two accumulator registers of lane (g, t) —
 * lane ↔ slot ks = lane: no reduction between lanes, and the order of the sum over p is fixed by the instruction.
 * The two y-halves of a cell share X and Z, the two cells share X and Y: 16 table entries per position serve 128
 * slots (a unit alone needs 10 for 32; the L2 sector rate set the time of the first versions). Warp w takes the
 * positions 32 w … 32 w + 31 (its own cp.async batch, 8 steps; Re(A)·B' and Im(A)·B'' in separate accumulators: a
 * dependent chain of 8 tensor instructions instead of 58), the four shares are added in warp order.
 *
 * @param aks        [K] {A_k, √A_k}, storage order
 * @param unit_info  [n_units] {first k of the unit's cell, …}
 * @param unit_map   [n_units][32] index of each slot's k-vector inside the cell's storage range, 255: none
 * @param item_units [n_items] unit of (cell 0, h 0), (cell 0, h 1), (cell 1, h 0), (cell 1, h 1); −1: none
 * @param item_base  [n_items] table index of x, y (h = 0), z of cell 0, z of cell 1
 * @param kq         [n_units][32] out: √A_k · Q_k of the state this window starts from (0 for empty slots)
 * @param e_partials [n_items] out: Σ A_k |Q_k|² over the item
 */
__global__ void __launch_bounds__(kFrontThreads, FB_FRONT_BLOCKS_PER_SM)
    windowFrontKernel(EwaldView E, const double2* __restrict__ aks, const int4* __restrict__ unit_info,
                      const unsigned char* __restrict__ unit_map, const int4* __restrict__ item_units,
                      const int4* __restrict__ item_base, int n_phase_blocks, BatchBuffers cur, BatchBuffers prev,
                      PhaseGeometry geo, double2* __restrict__ kq, double* __restrict__ e_partials)
{
    __shared__ FrontSmem sm;
    __shared__ int s_table[2 * kBatchMax]; //!< first table entry of every accepted position (trial, old, trial, …)
    const int tid = threadIdx.x;
    FB_GRID_DEPENDENCY_WAIT(); // the window description and the commit list are the tail kernel's
    FB_LAUNCH_DEPENDENTS();
    if (static_cast<int>(blockIdx.x) < n_phase_blocks) {
        const int total = 2 * cur.in->n * geo.table_stride;
        const int t = blockIdx.x * kFrontThreads + tid;
        if (t < total) {
            phaseTableEntry(cur.in, cur.table, geo, t);
        }
        return;
    }
    const int item = blockIdx.x - n_phase_blocks;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int4 units = __ldg(item_units + item);
    const int4 base = __ldg(item_base + item);
    // requested first (warp 0 finishes the item): slot ks = lane of the item's four units
    int slot_k[4] = {-1, -1, -1, -1};
    double2 slot_q[4], slot_a[4];
    if (warp == 0) {
        const int uu[4] = {units.x, units.y, units.z, units.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (uu[i] >= 0) {
                const int t = __ldg(unit_map + static_cast<size_t>(uu[i]) * kUnitSlots + lane);
                if (t != 255) {
                    slot_k[i] = __ldg(unit_info + uu[i]).x + t;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (slot_k[i] >= 0) {
                slot_q[i] = E.Q[slot_k[i]];
                slot_a[i] = __ldg(aks + slot_k[i]);
            }
        }
    }
    const CommitList& commit = cur.in->commit;
    const int ncommit = min(commit.n, kBatchMax);
    if (tid < 2 * ncommit) {
        s_table[tid] = (2 * commit.index[tid >> 1] + (tid & 1)) * geo.table_stride;
    }
    __syncthreads();
    const int n_pos = 2 * ncommit;
    const int left = n_pos - 32 * warp; // positions of this warp's batch
    double re[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}; // Re(A)·B' of the four units
    double im[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}; // Im(A)·B''
    if (left > 0) {
        const int g = lane >> 2; // row of A = 2 xi + jj; column of B', B'' = 2 l + c
        const int t = lane & 3;  // position inside the step; column pair of C = l
        const int ex = g >> 1, ey = 4 + (g & 1), ez = 8 + (g >> 1); // entries this lane multiplies (second half / cell: + 2, + 4)
        const bool imag = g & 1;
        const int y_base = geo.ncc + 1, z_base = (geo.ncc + 1) + (2 * geo.ncc + 1);
        // the batch's 32 × 16 table entries: half a warp per position, lane ↔ entry, so that every 64-byte run of four
        // entries is fetched as two whole sectors (one lane per position asked for every sector twice, and the L2
        // sector rate — 1.2 M requests per window — was the kernel time)
        {
            const int e = lane & 15;
            const int kind = e >> 2, i = e & 3; // X, Y, Z of cell 0, Z of cell 1
            const int first = kind == 0 ? base.x : (kind == 1 ? base.y : (kind == 2 ? base.z : base.w));
            const int offset = (kind == 0 ? 0 : (kind == 1 ? y_base : z_base)) + min(first + i, kind == 0 ? geo.ncc : 2 * geo.ncc);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int q = 2 * j + (lane >> 4);
                if (q < left) {
                    cpAsync16(&sm.stage[warp][q][e], prev.table + s_table[32 * warp + q] + offset);
                }
            }
        }
        cpAsyncCommit();
        cpAsyncWaitAll();
        __syncwarp();
        const double2(*st)[kItemStride] = sm.stage[warp];
        const int steps = min(8, (left + 3) >> 2);
        for (int r = 0; r < steps; ++r) {
            const int q = 4 * r + t;
            double2 a0 = make_double2(0.0, 0.0), a1 = a0, z0 = a0, z1 = a0;
            if (q < left) {
                double2 x = st[q][ex];
                if (q & 1) { // old position: subtracted (batches start at even positions)
                    x.x = -x.x;
                    x.y = -x.y;
                }
                a0 = cmul(x, st[q][ey]);
                a1 = cmul(x, st[q][ey + 2]);
                z0 = st[q][ez];
                z1 = st[q][ez + 4];
            }
            // B' = [Z_re | Z_im], B'' = [−Z_im | Z_re]
            const double b1 = imag ? z0.y : z0.x, b2 = imag ? z0.x : -z0.y;
            const double c1 = imag ? z1.y : z1.x, c2 = imag ? z1.x : -z1.y;
            dmma884(re[0][0], re[0][1], a0.x, b1);
            dmma884(re[1][0], re[1][1], a1.x, b1);
            dmma884(re[2][0], re[2][1], a0.x, c1);
            dmma884(re[3][0], re[3][1], a1.x, c1);
            dmma884(im[0][0], im[0][1], a0.y, b2);
            dmma884(im[1][0], im[1][1], a1.y, b2);
            dmma884(im[2][0], im[2][1], a0.y, c2);
            dmma884(im[3][0], im[3][1], a1.y, c2);
        }
    }
#if FB_FRONT_BLOCKS_PER_SM <= 4
    double2(*part)[4][32] = sm.part;
#else
    // more blocks per SM: the shares go where the warp's own staged tables were (34.8 kB per block instead of 43 kB)
    __syncwarp();
    double2(*part)[4][32] = reinterpret_cast<double2(*)[4][32]>(&sm.stage[0][0][0]);
    __syncthreads(); // every warp is through with its tables
#endif
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        part[warp][i][lane] = make_double2(re[i][0] + im[i][0], re[i][1] + im[i][1]);
    }
    __syncthreads();
    if (warp != 0) {
        return;
    }
    double e = 0.0;
    const int uu[4] = {units.x, units.y, units.z, units.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (uu[i] >= 0) {
            double2 out = make_double2(0.0, 0.0);
            if (slot_k[i] >= 0) {
                double2 Q = slot_q[i];
                if (ncommit > 0) {
                    double2 dq = part[0][i][lane];
#pragma unroll
                    for (int w = 1; w < kFrontWarps; ++w) { // the shares of the four batches, in order
                        dq.x += part[w][i][lane].x;
                        dq.y += part[w][i][lane].y;
                    }
                    Q.x += dq.x;
                    Q.y += dq.y;
                    E.Q[slot_k[i]] = Q; // only this block touches the item's k-vectors
                }
                out = make_double2(slot_a[i].y * Q.x, slot_a[i].y * Q.y);
                e += slot_a[i].x * (Q.x * Q.x + Q.y * Q.y);
            }
            kq[static_cast<size_t>(uu[i]) * kUnitSlots + lane] = out;
        }
    }
    e = warpSum(e); // fixed shuffle tree
    if (lane == 0) {
        e_partials[item] = e;
    }
}

// ------------------------------------------------------------------------------------------------
// the persistent k-space kernel
// ------------------------------------------------------------------------------------------------
/** what ONE warp stages for a unit: the tables of its 8 moves and the unit's √A_k·Q_k, √A_k */
struct KspaceWarpStage
{
    double2 tab[8][2][kUnitEntries]; //!< [move of the warp][trial | old]; 320 B per move ≡ 64 (mod 128)
    double2 kq[36];                  //!< √A_k · Q_k at the skewed slot index
    double sa[36];                   //!< √A_k
};

struct KspaceSmem
{
    double2 frag[2][kKsWarps][8][32]; //!< [unit parity][warp][step pair][lane] = √A_k δ (re, im) of one k-slot
    KspaceWarpStage stage[kKsWarps];
};

/**
 * @param unit_info  [n_units] {first k of the unit's cell, x, y, z table index of the first slot}
 * @param unit_sa    [n_units][32] √A_k in slot layout (0 for empty slots)
 * @param kq         [n_units][32] √A_k · Q_k in slot layout (windowFrontKernel)
 * @param r_partials [grid][stride]   2 Σ_k A_k Re(conj(Q_k) δ_m,k) + Σ_k A_k |δ_m,k|² over the block's units
 * @param g_partials [grid][stride²]  entries [a][m], a < m: Σ_k A_k Re(conj(δ_a,k) δ_m,k)
 */
__global__ void __launch_bounds__(kKsThreads, 2 / kKsGroups)
    windowKspaceKernel(const int4* __restrict__ unit_info, const double* __restrict__ unit_sa, const double2* __restrict__ kq,
                       const unsigned char* __restrict__ unit_steps, const int* __restrict__ sched_first,
                       const int* __restrict__ sched_units, int n_sched_blocks, BatchBuffers cur, PhaseGeometry geo, int stride,
                       double* __restrict__ r_partials, double* __restrict__ g_partials)
{
    extern __shared__ __align__(16) unsigned char ks_smem_raw[];
    // Two independent work groups of 256 threads per block (named barriers, own shared memory), each with its own
    // list of units — "virtual block" 2·blockIdx + group of the schedule. Their sums are added (group 0 + group 1) before
    // they leave the block: one row of partials per SM instead of two, half the bytes the tail kernel reads back.
    const int group = kKsGroups == 1 ? 0 : threadIdx.x >> 8;
    KspaceSmem& sm = reinterpret_cast<KspaceSmem*>(ks_smem_raw)[group];

    const int tid = threadIdx.x & (kKsGroupThreads - 1);
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int block = kKsGroups * blockIdx.x + group;
    FB_GRID_DEPENDENCY_WAIT(); // phase tables and √A_k·Q_k are the front kernel's
    FB_LAUNCH_DEPENDENTS();
    const int n = cur.in->n;
    const int n_active_warps = (n + 7) >> 3;
    const bool warp_active = warp < n_active_warps;
    const bool block_active = block < n_sched_blocks;
    const int my_first = block_active ? __ldg(sched_first + block) : 0; // the units of this group: a static, cost-balanced schedule
    const int my_units = block_active ? __ldg(sched_first + block + 1) - my_first : 0;
    KspaceWarpStage& st = sm.stage[warp];

    // this thread's move and x-index (the x entries of the tables carry the charges)
    const int m = tid >> 2;
    const int xi = tid & 3;
    const int mg = lane >> 2; // move inside the warp = fragment row
    const bool move_active = m < n;

    // staging plan of this lane: lane = t + 10 j takes table entry t (X0..3, Y0..1, Z0..3) of the warp's positions
    // j, j + 3, …, 15 (lanes 30, 31 only carry √A_k·Q_k and √A_k), so the table offset is one min() per unit
    const int st_t = lane % kUnitEntries;
    const int st_j = lane / kUnitEntries;
    const int st_kind = st_t < 4 ? 0 : (st_t < 6 ? 1 : 2);
    const int st_base = st_kind == 0 ? 0 : (st_kind == 1 ? geo.ncc + 1 : (geo.ncc + 1) + (2 * geo.ncc + 1));
    const int st_local = st_t - (st_kind == 0 ? 0 : (st_kind == 1 ? 4 : 6));
    const int st_limit = st_kind == 0 ? geo.ncc : 2 * geo.ncc;
    const int st_positions = max(0, min(16, 2 * (n - 8 * warp))); // positions of the warp that belong to moves < n
    const double2* __restrict__ table = cur.table + static_cast<size_t>(16 * warp) * geo.table_stride;
    const unsigned st_dst = static_cast<unsigned>(__cvta_generic_to_shared(&st.tab[0][0][st_t]));
    const unsigned st_kq = static_cast<unsigned>(__cvta_generic_to_shared(&st.kq[lane + (lane >> 3)]));
    const unsigned st_sa = static_cast<unsigned>(__cvta_generic_to_shared(&st.sa[lane + (lane >> 3)]));
    unsigned steps_next = 0; // slot columns (jj, l) of the unit being staged that hold a k-vector
    int u_next = my_units > 0 ? __ldg(sched_units + my_first) : 0; // read one unit ahead: off the staging chain
    auto issue = [&](int c) {
        const int u = u_next;
        if (c + 1 < my_units) {
            u_next = __ldg(sched_units + my_first + c + 1);
        }
        steps_next = __ldg(unit_steps + u);
        const int4 info = __ldg(unit_info + u);
        const int first = st_kind == 0 ? info.y : (st_kind == 1 ? info.z : info.w);
        const double2* src = table + st_base + min(first + st_local, st_limit); // beyond the table: slots without k-vector
        if (st_j < 3) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const int v = st_j + 3 * r; // position of the warp: 2·(move in warp) + (trial | old)
                if (v < st_positions) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(st_dst + v * (kUnitEntries * 16u)),
                                 "l"(src + v * geo.table_stride));
                }
            }
        }
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(st_kq), "l"(kq + static_cast<size_t>(u) * kUnitSlots + lane));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(st_sa), "l"(unit_sa + static_cast<size_t>(u) * kUnitSlots + lane));
        cpAsyncCommit();
    };

    // Gram tiles of this warp: (w, w) and the partner warps (w + d) mod 8, d = 1, 2, 3 (4 for w < 4), that hold moves;
    // the list is compacted so that the tile loop below runs without predicates (a predicated mma.sync costs a
    // WARPSYNC each: 17 % of the samples of the first version)
    int p0 = 0, p1 = 0, p2 = 0, p3 = 0; // (scalars: an array indexed by the running count would live in local memory)
    int n_partners = 0;
    if (warp_active) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int p = (warp + d + 1) & 7;
            if (p < n_active_warps && (d < 3 || warp < 4)) {
                p0 = n_partners == 0 ? p : p0;
                p1 = n_partners == 1 ? p : p1;
                p2 = n_partners == 2 ? p : p2;
                p3 = n_partners == 3 ? p : p3;
                ++n_partners;
            }
        }
    }
    const int partner[4] = {p0, p1, p2, p3};
    double g_diag[2] = {0.0, 0.0};
    double g_off[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    double racc = 0.0;

    if (my_units > 0 && warp_active) {
        issue(0);
    }
    for (int c = 0; c < my_units; ++c) {
        const int buf = c & 1;
        const unsigned steps = steps_next; // of unit c (issue(c + 1) below replaces steps_next)
        double dreg[16];
        if (warp_active) {
            cpAsyncWaitAll();
            __syncwarp(); // the warp's tables and the unit's √A_k·Q_k are in place
            // ---- δ of this thread's move for the 8 slots of its x-index, scaled by √A_k: the Gram fragments
            if (move_active) {
                const double2 xn = st.tab[mg][0][xi];
                const double2 xo = st.tab[mg][1][xi];
                double2 zn[4], zo[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    zn[l] = st.tab[mg][0][6 + l];
                    zo[l] = st.tab[mg][1][6 + l];
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const double2 xyn = cmul(xn, st.tab[mg][0][4 + jj]);
                    const double2 xyo = cmul(xo, st.tab[mg][1][4 + jj]);
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const int s = 4 * jj + l;
                        const int idx = unitSlotIndex(xi, s);
                        double2 d = cmul(xyn, zn[l]);
                        d.x = fma(-xyo.x, zo[l].x, fma(xyo.y, zo[l].y, d.x));
                        d.y = fma(-xyo.x, zo[l].y, fma(-xyo.y, zo[l].x, d.y));
                        const double sa = st.sa[idx];
                        const double2 q = st.kq[idx];
                        d.x *= sa;
                        d.y *= sa;
                        racc = fma(q.x, d.x, fma(q.y, d.y, racc));
                        dreg[2 * s] = d.x;
                        dreg[2 * s + 1] = d.y;
                    }
                }
            }
            else {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    dreg[s] = 0.0;
                }
            }
            __syncwarp(); // every lane is done with the staging area
            if (c + 1 < my_units) {
                issue(c + 1); // lands while the Gram update runs
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                sm.frag[buf][warp][s][lane] = make_double2(dreg[2 * s], dreg[2 * s + 1]);
            }
        }
        // the fragments of unit c of all warps of the GROUP are in place; every warp is through with the Gram update
        // of unit c − 1, so the other fragment buffer may be overwritten after this barrier (named: the other group
        // of the block runs its own units at its own pace)
        if constexpr (kKsGroups == 1) {
            __syncthreads();
        }
        else {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kKsGroupThreads) : "memory");
        }
        if (warp_active) { // ---- Gram update on the FP64 tensor path
            const double2* fr = &sm.frag[buf][0][0][lane];
#ifndef FB_KS_INTERLEAVE
#define FB_KS_INTERLEAVE 1
#endif
#if FB_KS_INTERLEAVE
// the re step of every tile, then the im step of every tile: two tensor instructions on the same accumulator are
// 1 + NP apart instead of back to back (DMMA: 16 cycles between issues of a sub-partition, ≈ 26 until the result)
#define FB_GRAM(NP, CHECK)                                                                                 \
    _Pragma("unroll") for (int s = 0; s < 8; ++s)                                                          \
    {                                                                                                      \
        if (CHECK && !((steps >> s) & 1u)) { /* block-uniform: no k-vector in this column, every δ is zero */ \
            continue;                                                                                      \
        }                                                                                                  \
        double2 f[NP > 0 ? NP : 1];                                                                        \
        _Pragma("unroll") for (int d = 0; d < NP; ++d) { f[d] = fr[(partner[d] * 8 + s) * 32]; }           \
        dmma884(g_diag[0], g_diag[1], dreg[2 * s], dreg[2 * s]);                                           \
        _Pragma("unroll") for (int d = 0; d < NP; ++d) { dmma884(g_off[d][0], g_off[d][1], dreg[2 * s], f[d].x); }         \
        dmma884(g_diag[0], g_diag[1], dreg[2 * s + 1], dreg[2 * s + 1]);                                   \
        _Pragma("unroll") for (int d = 0; d < NP; ++d) { dmma884(g_off[d][0], g_off[d][1], dreg[2 * s + 1], f[d].y); }     \
    }
#else
#define FB_GRAM(NP, CHECK)                                                                                 \
    _Pragma("unroll") for (int s = 0; s < 8; ++s)                                                          \
    {                                                                                                      \
        if (CHECK && !((steps >> s) & 1u)) { /* block-uniform: no k-vector in this column, every δ is zero */ \
            continue;                                                                                      \
        }                                                                                                  \
        double2 f[NP > 0 ? NP : 1];                                                                        \
        _Pragma("unroll") for (int d = 0; d < NP; ++d) { f[d] = fr[(partner[d] * 8 + s) * 32]; }           \
        dmma884(g_diag[0], g_diag[1], dreg[2 * s], dreg[2 * s]);                                           \
        dmma884(g_diag[0], g_diag[1], dreg[2 * s + 1], dreg[2 * s + 1]);                                   \
        _Pragma("unroll") for (int d = 0; d < NP; ++d)                                                     \
        {                                                                                                  \
            dmma884(g_off[d][0], g_off[d][1], dreg[2 * s], f[d].x);                                        \
            dmma884(g_off[d][0], g_off[d][1], dreg[2 * s + 1], f[d].y);                                    \
        }                                                                                                  \
    }
#endif
            // (switch FB_KS_SKIP = 2: a unit whose eight slot columns all hold k-vectors — 71 % of the units at S1 — takes
            // a copy of the loop without the per-step tests, which cost a WARPSYNC each: 47 instead of 12 in the SASS)
#ifndef FB_KS_SKIP
#define FB_KS_SKIP 1 // 0: no step is skipped; 1: every step tests its bit; 2: full units take the test-free loop
                     // measured at S1: 8.79e5 / 8.97e5 / 8.90e5 moves/s — the tests (a WARPSYNC each) cost less than the
                     // second copy of the loop does
#endif
            if (FB_KS_SKIP == 0 || (FB_KS_SKIP == 2 && steps == 0xffu)) {
                switch (n_partners) {
                case 4:
                    FB_GRAM(4, false)
                    break;
                case 3:
                    FB_GRAM(3, false)
                    break;
                case 2:
                    FB_GRAM(2, false)
                    break;
                case 1:
                    FB_GRAM(1, false)
                    break;
                default:
                    FB_GRAM(0, false)
                }
            }
            else {
                switch (n_partners) {
                case 4:
                    FB_GRAM(4, true)
                    break;
                case 3:
                    FB_GRAM(3, true)
                    break;
                case 2:
                    FB_GRAM(2, true)
                    break;
                case 1:
                    FB_GRAM(1, true)
                    break;
                default:
                    FB_GRAM(0, true)
                }
            }
#undef FB_GRAM
        }
    }

    // ---- partial sums of this block: group 1 hands its sums over through its (now idle) staging area, group 0 adds
    // them to its own — always own + other — and stores the row
    const int frag_g = lane >> 2;
    const int frag_t = lane & 3;
    const size_t row = static_cast<size_t>(blockIdx.x);
    // R[m]: the four x-indices of a move sit in adjacent lanes; the diagonal of G carries Σ A_k |δ|²
    double rs = racc;
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    static_assert(sizeof(KspaceWarpStage) >= 11 * 32 * sizeof(double), "hand-over area of a warp");
    double* hand = reinterpret_cast<double*>(&reinterpret_cast<KspaceSmem*>(ks_smem_raw)[kKsGroups - 1].stage[warp]);
    if (kKsGroups == 2 && group == 1 && warp_active) {
        hand[0 * 32 + lane] = rs;
        hand[1 * 32 + lane] = g_diag[0];
        hand[2 * 32 + lane] = g_diag[1];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            hand[(3 + 2 * d) * 32 + lane] = g_off[d][0];
            hand[(4 + 2 * d) * 32 + lane] = g_off[d][1];
        }
    }
    if constexpr (kKsGroups == 2) {
        __syncthreads();
    }
    if (group == 0 && warp_active) {
        if constexpr (kKsGroups == 2) {
            rs += hand[0 * 32 + lane];
            g_diag[0] += hand[1 * 32 + lane];
            g_diag[1] += hand[2 * 32 + lane];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                g_off[d][0] += hand[(3 + 2 * d) * 32 + lane];
                g_off[d][1] += hand[(4 + 2 * d) * 32 + lane];
            }
        }
        if ((frag_g >> 1) == frag_t && move_active) {
            r_partials[row * stride + m] = 2.0 * rs + g_diag[frag_g & 1];
        }
        // G: lane (g, t) of tile (w, p) holds G[8w + g][8p + 2t + {0, 1}]; stored as [a][m] with a < m
        double* G = g_partials + row * static_cast<size_t>(stride) * stride;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int a = 8 * warp + frag_g;
            const int b = 8 * warp + 2 * frag_t + e;
            if (a < b) {
                G[a * stride + b] = g_diag[e];
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d < n_partners) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a = 8 * warp + frag_g;
                    const int b = 8 * partner[d] + 2 * frag_t + e;
                    G[min(a, b) * stride + max(a, b)] = g_off[d][e];
                }
            }
        }
    }
}

} // namespace fbdev

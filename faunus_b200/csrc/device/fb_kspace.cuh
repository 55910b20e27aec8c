// k-space part of a window (sm_100a): R[m], G[a][m], Σ_k A_k |Q_k|² and the commit of the previous window's
// accepted moves into Q(k) — the quantities defined at the top of fb_batch.cuh — in ONE persistent kernel.
//
// Work unit = HALF of a 4×4×4 cell of integer triplets (the two y-rows ly ∈ {2h, 2h+1}): 32 k-slots that need only
// 10 phase-table entries per position (4 x, 2 y, 4 z). Block b walks the units b, b + grid, …
//
// Thread ↔ (move m = tid / 4, x-index i = tid % 4); warp w owns the moves 8w … 8w+7. A thread forms the 8 phases
// e^{ik·r} of its x-index from registers — xy_j = X_i·Y_j (once per y), then xy_j·Z_l — for the trial and the old
// position of ITS move: 14 FP64 instructions per (k, move) against 28 with lane ↔ k and six LDS.128 table reads.
// The 16 doubles √A_k·δ_m,k it ends up with ARE the m8n8k4 fragments of the Gram update: with lane = 4·(m % 8) + i
// the lane (g, t) of warp w holds D[8w + g][column (t, s)], which is both the A fragment A[g][t] of row block w and
// the B fragment B[t][g] of column block w for reduction step s (the order of the reduction columns is free). The
// diagonal tile of a warp needs no memory at all; for an off-diagonal tile the partner warp's fragments come from
// shared memory, stored lane by lane (conflict-free LDS.128, two steps per load). Tiles are dealt as a circulant:
// warp w takes (w, w) and (w, w+d mod 8) for d = 1, 2, 3 (and d = 4 for w < 4) — 4.5 DMMA per fragment load.
// Σ_k A_k |δ_m,k|² is the diagonal of G and is taken from there; R keeps only the 2 Re(conj(Q) δ) part.
//
// Commit of the previous window: thread ↔ (x-index, y-index, group of accepted moves); partial ΔQ per group through
// shared memory, summed in fixed order (results do not depend on scheduling), Q(k) updated in place by the unit's
// block, which is the only one that touches those k-vectors.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

constexpr int kUnitSlots = 32;    //!< k-slots of a unit: ks = 8·i + 4·jj + l
constexpr int kUnitEntries = 10;  //!< phase-table entries per position: X0..3, Y0..1, Z0..3
constexpr int kKsThreads = 256;
constexpr int kKsWarps = kKsThreads / 32;
constexpr int kCommitGroups = 32; //!< partial ΔQ sums per k-slot
constexpr int kDqStride = 38;     //!< double2 per row: 608 B ≡ 96 (mod 128): the fixed-order sum reads conflict-free

/** skewed slot index: the four x-indices of a warp land in different bank groups */
__device__ __forceinline__ int unitSlotIndex(int i, int s) { return 9 * i + s; }

struct KspaceUnitSmem
{
    double2 tab_new[kBatchMax][2][kUnitEntries]; //!< [move][trial | old]; 320 B per move ≡ 64 (mod 128)
    double2 tab_com[kBatchMax][2][kUnitEntries]; //!< accepted moves of the previous window
    union
    {
        double2 frag[kKsWarps][8][32];           //!< [warp][step pair][lane] = √A_k δ (re, im) of one k-slot
        double2 dq[kCommitGroups][kDqStride];    //!< partial ΔQ of the commit groups
    };
    double2 Q[kUnitSlots];   //!< staged Q(k) and {A_k, √A_k} of the unit's slots (absent slots: see map)
    double2 aks[kUnitSlots];
    double2 kq[36];          //!< √A_k · Q_k at the skewed index
    double sa[36];           //!< √A_k (0 for absent slots)
    int map[kUnitSlots];     //!< index of the slot's k-vector inside its cell's storage range, −1: none
    double cqn[kBatchMax], cqo[kBatchMax];
    int ctable[kBatchMax];   //!< first table entry (trial position) of the accepted move in the previous window's tables
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

/**
 * @param aks        [K] {A_k, √A_k}, storage order
 * @param unit_info  [n_units] {first k of the unit's cell, x table index of slot i = 0, y table index of jj = 0, z of l = 0}
 * @param unit_map   [n_units][32] index of each slot's k-vector inside the cell's storage range, 255: none
 * @param r_partials [grid][stride]   2 Σ_k A_k Re(conj(Q_k) δ_m,k) + Σ_k A_k |δ_m,k|² over the block's units
 * @param g_partials [grid][stride²]  entries [a][m], a < m: Σ_k A_k Re(conj(δ_a,k) δ_m,k)
 */
__global__ void __launch_bounds__(kKsThreads, 2)
    windowKspaceKernel(EwaldView E, const double2* __restrict__ aks, const int4* __restrict__ unit_info,
                       const unsigned char* __restrict__ unit_map, int n_units, BatchBuffers cur, BatchBuffers prev,
                       PhaseGeometry geo, int stride, double* __restrict__ r_partials, double* __restrict__ g_partials,
                       double* __restrict__ e_partials)
{
    extern __shared__ __align__(16) unsigned char ks_smem_raw[];
    KspaceUnitSmem& sm = *reinterpret_cast<KspaceUnitSmem*>(ks_smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int n = cur.in->n;
    const CommitList& commit = cur.in->commit;
    const int ncommit = min(commit.n, kBatchMax);
    const int n_active_warps = (n + 7) >> 3;

    if (tid < ncommit) {
        const int m = commit.index[tid];
        sm.cqn[tid] = prev.in->pnew[m].w;
        sm.cqo[tid] = prev.pold[m].w;
        sm.ctable[tid] = 2 * m * geo.table_stride;
    }
    __syncthreads();

    const int my_units = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int x_base = 0, y_base = geo.ncc + 1, z_base = (geo.ncc + 1) + (2 * geo.ncc + 1);

    // asynchronous staging of unit number `c` of this block: tables of the 2n positions of the window and of the
    // accepted moves of the previous one, Q(k) and {A_k, √A_k} of the unit's slots
    auto issue = [&](int c) {
        const int u = blockIdx.x + c * gridDim.x;
        const int4 info = __ldg(unit_info + u);
        auto offset_of = [&](int t) {
            if (t < 4) {
                return x_base + min(info.y + t, geo.ncc);
            }
            if (t < 6) {
                return y_base + min(info.z + (t - 4), 2 * geo.ncc);
            }
            return z_base + min(info.w + (t - 6), 2 * geo.ncc);
        };
        const int n_new = 2 * n * kUnitEntries;
        for (int e = tid; e < n_new; e += kKsThreads) {
            const int v = e / kUnitEntries;
            const int t = e - v * kUnitEntries;
            cpAsync16(&sm.tab_new[v >> 1][v & 1][t], cur.table + v * geo.table_stride + offset_of(t));
        }
        const int n_com = 2 * ncommit * kUnitEntries;
        for (int e = tid; e < n_com; e += kKsThreads) {
            const int v = e / kUnitEntries;
            const int t = e - v * kUnitEntries;
            cpAsync16(&sm.tab_com[v >> 1][v & 1][t], prev.table + sm.ctable[v >> 1] + (v & 1) * geo.table_stride + offset_of(t));
        }
        if (tid < kUnitSlots) {
            const int t = __ldg(unit_map + static_cast<size_t>(u) * kUnitSlots + tid);
            sm.map[tid] = t == 255 ? -1 : t;
            if (t != 255) {
                cpAsync16(&sm.Q[tid], E.Q + info.x + t);
                cpAsync16(&sm.aks[tid], aks + info.x + t);
            }
        }
        cpAsyncCommit();
    };

    // this thread's move and x-index; charges of the trial / old position
    const int m = tid >> 2;
    const int xi = tid & 3;
    const bool move_active = m < n;
    const double qn = move_active ? cur.in->pnew[m].w : 0.0;
    const double qo = move_active ? cur.pold[m].w : 0.0;
    const bool warp_active = warp < n_active_warps;

    // Gram tiles of this warp: (w, w) and the partners (w + d) mod 8
    int partner[4];
    bool partner_on[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        partner[d] = (warp + d + 1) & 7;
        partner_on[d] = warp_active && partner[d] < n_active_warps && (d < 3 || warp < 4);
    }
    double g_diag[2] = {0.0, 0.0};
    double g_off[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    double racc = 0.0;
    double eacc = 0.0;

    if (my_units > 0) {
        issue(0);
    }
    for (int c = 0; c < my_units; ++c) {
        const int u = blockIdx.x + c * gridDim.x;
        const int p0 = __ldg(unit_info + u).x;
        cpAsyncWaitAll();
        __syncthreads(); // unit c is staged for everybody; everybody is done with the fragments of unit c − 1

        // ---- ΔQ of the accepted moves of the previous window
        if (ncommit > 0) {
            const int ci = tid & 3;
            const int cj = (tid >> 2) & 1;
            const int cg = tid >> 3;
            double2 acc[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                acc[l] = make_double2(0.0, 0.0);
            }
            for (int a = cg; a < ncommit; a += kCommitGroups) {
#pragma unroll
                for (int pos = 0; pos < 2; ++pos) {
                    const double q = pos ? -sm.cqo[a] : sm.cqn[a];
                    double2 xy = cmul(sm.tab_com[a][pos][ci], sm.tab_com[a][pos][4 + cj]);
                    xy.x *= q;
                    xy.y *= q;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const double2 z = sm.tab_com[a][pos][6 + l];
                        acc[l].x = fma(xy.x, z.x, fma(-xy.y, z.y, acc[l].x));
                        acc[l].y = fma(xy.x, z.y, fma(xy.y, z.x, acc[l].y));
                    }
                }
            }
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                sm.dq[cg][unitSlotIndex(ci, 4 * cj + l)] = acc[l];
            }
            __syncthreads();
        }
        if (tid < 4 * kUnitSlots) { // four threads per slot: fixed-order sum of the 32 group shares
            const int ks = tid >> 2;
            const int part = tid & 3;
            const int idx = unitSlotIndex(ks >> 3, ks & 7);
            double2 dq = make_double2(0.0, 0.0);
            if (ncommit > 0) {
#pragma unroll
                for (int r = 0; r < kCommitGroups / 4; ++r) {
                    const double2 v = sm.dq[part + 4 * r][idx];
                    dq.x += v.x;
                    dq.y += v.y;
                }
                dq.x += __shfl_xor_sync(0xffffffffu, dq.x, 1);
                dq.y += __shfl_xor_sync(0xffffffffu, dq.y, 1);
                dq.x += __shfl_xor_sync(0xffffffffu, dq.x, 2);
                dq.y += __shfl_xor_sync(0xffffffffu, dq.y, 2);
            }
            if (part == 0) {
                const int t = sm.map[ks];
                double2 Q = make_double2(0.0, 0.0);
                double A = 0.0, sA = 0.0;
                if (t >= 0) {
                    Q = sm.Q[ks];
                    A = sm.aks[ks].x;
                    sA = sm.aks[ks].y;
                    if (ncommit > 0) {
                        Q.x += dq.x;
                        Q.y += dq.y;
                        E.Q[p0 + t] = Q; // only this block touches the unit's k-vectors
                    }
                }
                sm.kq[idx] = make_double2(sA * Q.x, sA * Q.y);
                sm.sa[idx] = sA;
                eacc += A * (Q.x * Q.x + Q.y * Q.y);
            }
        }
        __syncthreads(); // kq / sa are ready; dq (= frag) is free

        // ---- δ of this thread's move for the 8 slots of its x-index, scaled by √A_k: the Gram fragments
        double dreg[16];
        if (warp_active) {
            if (move_active) {
                double2 xn = sm.tab_new[m][0][xi];
                double2 xo = sm.tab_new[m][1][xi];
                xn.x *= qn;
                xn.y *= qn;
                xo.x *= qo;
                xo.y *= qo;
                double2 zn[4], zo[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    zn[l] = sm.tab_new[m][0][6 + l];
                    zo[l] = sm.tab_new[m][1][6 + l];
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const double2 xyn = cmul(xn, sm.tab_new[m][0][4 + jj]);
                    const double2 xyo = cmul(xo, sm.tab_new[m][1][4 + jj]);
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const int s = 4 * jj + l;
                        const int idx = unitSlotIndex(xi, s);
                        double2 d = cmul(xyn, zn[l]);
                        d.x = fma(-xyo.x, zo[l].x, fma(xyo.y, zo[l].y, d.x));
                        d.y = fma(-xyo.x, zo[l].y, fma(-xyo.y, zo[l].x, d.y));
                        const double sa = sm.sa[idx];
                        const double2 kq = sm.kq[idx];
                        d.x *= sa;
                        d.y *= sa;
                        racc = fma(kq.x, d.x, fma(kq.y, d.y, racc));
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
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                sm.frag[warp][s][lane] = make_double2(dreg[2 * s], dreg[2 * s + 1]);
            }
        }
        __syncthreads(); // fragments of all warps are in place; the tables are free
        if (c + 1 < my_units) {
            issue(c + 1); // lands while the Gram update runs
        }

        // ---- Gram update on the FP64 tensor path
        if (warp_active) {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                double2 f[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (partner_on[d]) {
                        f[d] = sm.frag[partner[d]][s][lane];
                    }
                }
                dmma884(g_diag[0], g_diag[1], dreg[2 * s], dreg[2 * s]);
                dmma884(g_diag[0], g_diag[1], dreg[2 * s + 1], dreg[2 * s + 1]);
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if (partner_on[d]) {
                        dmma884(g_off[d][0], g_off[d][1], dreg[2 * s], f[d].x);
                        dmma884(g_off[d][0], g_off[d][1], dreg[2 * s + 1], f[d].y);
                    }
                }
            }
        }
    }

    // ---- partial sums of this block
    {
        double es = warpSum(eacc); // the slot threads live in warps 0 … 3
        __syncthreads();
        double* scratch = reinterpret_cast<double*>(&sm.Q[0]);
        if (lane == 0 && warp < 4) {
            scratch[warp] = es;
        }
        __syncthreads();
        if (tid == 0) {
            e_partials[blockIdx.x] = ((scratch[0] + scratch[1]) + scratch[2]) + scratch[3];
        }
    }
    const int frag_g = lane >> 2;
    const int frag_t = lane & 3;
    const size_t row = static_cast<size_t>(blockIdx.x);
    if (warp_active) {
        // R[m]: the four x-indices of a move sit in adjacent lanes; the diagonal of G carries Σ A_k |δ|²
        double rs = racc;
        rs += __shfl_xor_sync(0xffffffffu, rs, 1);
        rs += __shfl_xor_sync(0xffffffffu, rs, 2);
        if ((frag_g >> 1) == frag_t && move_active) {
            r_partials[row * stride + m] = 2.0 * rs + g_diag[frag_g & 1];
        }
        // G: lane (g, t) of tile (w, p) holds G[8w + g][8p + 2t + {0, 1}]; stored as [a][m] with a < m
        double* G = g_partials + row * static_cast<size_t>(stride) * stride;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int a = 8 * warp + frag_g;
            const int b = 8 * warp + 2 * frag_t + e;
            if (a < b && b < stride) {
                G[a * stride + b] = g_diag[e];
            }
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (partner_on[d]) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a = 8 * warp + frag_g;
                    const int b = 8 * partner[d] + 2 * frag_t + e;
                    const int lo = min(a, b), hi = max(a, b);
                    if (hi < stride) {
                        G[lo * stride + hi] = g_off[d][e];
                    }
                }
            }
        }
    }
    else {
        // moves beyond the window: R = 0 (the sums kernel reads only m < n, but the row is defined)
        if (m < stride && xi == 0) {
            r_partials[row * stride + m] = 0.0;
        }
    }
}

} // namespace fbdev

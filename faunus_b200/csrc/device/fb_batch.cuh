// Speculative evaluation of a WINDOW of sequential single-atom trial moves in one pass (sm_100a).
//
// The reference evaluates one trial move at a time (src/montecarlo.cpp:139-187): u_new = energy of the
// moved atom at its trial position with every other particle, u_old the same at the old position
// (GroupPairingPolicy::group2all + groupInternal, src/energy.h:1182-1195, 885-914) and, with Ewald,
// the reciprocal energy before/after the partial Q(k) update (src/energy.cpp:219-247, 524-531).
// A window of B such moves on DISTINCT atoms is evaluated here against the window-start state S0:
//
//   pair     u_new[m], u_old[m]     = Σ_{j≠p_m} u(trial_m | old_m , r_j(S0))           batchPairKernel
//   cross    C_new[a][m], C_old[a][m] = u(x_m, new_a) − u(x_m, old_a)                   batchFinishKernel
//   k-space  R[m]  = Σ_k A_k (2 Re(conj(Q_k) δ_m,k) + |δ_m,k|²)                        batchEwaldKernel
//            G[a][m] = Σ_k A_k Re(conj(δ_a,k) δ_m,k),  δ_m,k = q (e^{ik·new_m} − e^{ik·old_m})
//
// so that the energies of move m in the state where the accepted moves a < m have been applied are
//   u_x[m] + Σ_{a accepted} C_x[a][m]          and     ΔU_rec = pref (R[m] + 2 Σ_{a accepted} G[a][m]),
// exact identities — only the summation order differs from the one-move-at-a-time evaluation. The
// caller walks the window in order, decides each move, and tells the next launch which were accepted
// (batchCommitKernel writes their positions into both mirrors and adds their δ to Q(k)).
//
// e^{ik·r} is factorised into per-axis phase tables e^{i 2π n x/L} (k = 2π n/L on an orthogonal box),
// built once per window by batchPhaseKernel: 2 complex products per (k, position) instead of a sincos.
#pragma once
#include "fb_kernels.cuh"

namespace fbdev {

constexpr int kBatchMax = 64;    //!< moves per window
constexpr int kBatchTile = 256;  //!< particles staged per block of the pair kernel
constexpr int kBatchDeltaElems = 2048; //!< double2 elements of the δ tile in shared memory (32 KB + padding)
static_assert(kBatchTile == kBlock, "the pair kernel stages one particle per thread");

/** Host → device description of one window (copied as one block) */
struct BatchInput
{
    int n;
    int with_ewald;
    int slot[kBatchMax];
    int id[kBatchMax];
    double4 pnew[kBatchMax];
};

/** Device-resident working set of one window */
struct BatchBuffers
{
    BatchInput* in;
    double4* pold;   //!< [kBatchMax] positions at window start (gathered from the mirror)
    int* idold;      //!< [kBatchMax]
    double2* table;  //!< [2·kBatchMax][table_stride] phase factors, variant 2m = new, 2m+1 = old
};

struct PhaseGeometry
{
    int ncc;          //!< ceil(n_cutoff): n_x ∈ [0, ncc], n_y, n_z ∈ [−ncc, ncc]
    int table_stride; //!< entries per position: (ncc+1) + 2(2ncc+1)
    double len[3];    //!< box lengths
};

struct CommitList
{
    int n;
    int index[kBatchMax]; //!< indices into the PREVIOUS window
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

/** e^{ik·r} for k = 2π(nx, ny, nz)/L from the phase table of one position */
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
    if (blockIdx.x == 0 && threadIdx.x < commit.n) {
        const int m = commit.index[threadIdx.x];
        const int s = prev.in->slot[m];
        const double4 p = prev.in->pnew[m];
        const int id = prev.in->id[m];
        M0.posq[s] = p;
        M0.atom_id[s] = id;
        M1.posq[s] = p;
        M1.atom_id[s] = id;
    }
    if (!with_ewald) {
        return;
    }
    double e = 0.0;
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < E.K; k += gridDim.x * kBlock) {
        const int4 n = __ldg(kn + k);
        double2 Q = E.Q[k];
        for (int a = 0; a < commit.n; ++a) {
            const int m = commit.index[a];
            const double2 en = tablePhase(prev.table + static_cast<size_t>(2 * m) * geo.table_stride, geo, n.x, n.y, n.z);
            const double2 eo =
                tablePhase(prev.table + static_cast<size_t>(2 * m + 1) * geo.table_stride, geo, n.x, n.y, n.z);
            const double qn = prev.in->pnew[m].w;
            const double qo = prev.pold[m].w;
            Q.x += qn * en.x - qo * eo.x;
            Q.y += qn * en.y - qo * eo.y;
        }
        if (commit.n > 0) {
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
__global__ void __launch_bounds__(kBlock) batchPhaseKernel(SlotView M0, BatchBuffers cur, PhaseGeometry geo)
{
    const int n = cur.in->n;
    const int tid = blockIdx.x * kBlock + threadIdx.x;
    if (tid < n) {
        const int s = cur.in->slot[tid];
        cur.pold[tid] = M0.posq[s];
        cur.idold[tid] = M0.atom_id[s];
    }
    if (!cur.in->with_ewald) {
        return;
    }
    const int total = 2 * n * geo.table_stride;
    for (int t = tid; t < total; t += gridDim.x * kBlock) {
        const int variant = t / geo.table_stride;
        const int e = t - variant * geo.table_stride;
        const int m = variant >> 1;
        const double4 p = (variant & 1) ? M0.posq[cur.in->slot[m]] : cur.in->pnew[m];
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
        cur.table[t] = make_double2(cs, sn);
    }
}

// ------------------------------------------------------------------------------------------------
// pair part: thread ↔ (move variant, j sub-range); blocks sweep tiles of particles staged in shared
// memory (broadcast reads). Inactive particles are staged with NaN coordinates, so r² compares false.
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchPairKernel(SlotView M0, PotParams P, BatchBuffers cur, double cut2, int stride,
                    double* __restrict__ partials /*[gridDim.x][2·stride]*/)
{
    __shared__ double4 s_pos[kBatchTile];
    __shared__ int s_id[kBatchTile];
    __shared__ double s_red[kBlock];

    const int n = cur.in->n;
    const int nv = 2 * n;                 // variants
    const int nsub = kBlock / nv;         // j sub-ranges per tile (≥ 2 for n ≤ 64)
    const int v = threadIdx.x % nv;
    const int sub = threadIdx.x / nv;
    const bool worker = sub < nsub;
    const int m = v >> 1;

    double4 me = make_double4(0, 0, 0, 0);
    int my_id = 0;
    int my_slot = -1;
    if (worker) {
        my_slot = cur.in->slot[m];
        if (v & 1) {
            me = cur.pold[m];
            my_id = cur.idold[m];
        }
        else {
            me = cur.in->pnew[m];
            my_id = cur.in->id[m];
        }
    }
    const double hx = M0.half[0], hy = M0.half[1], hz = M0.half[2];
    const double lx = M0.len_or_zero[0], ly = M0.len_or_zero[1], lz = M0.len_or_zero[2];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    double e0 = 0.0, e1 = 0.0;
    const int j0 = blockIdx.x * kBatchTile;
    {
        const int j = j0 + threadIdx.x;
        if (j < M0.n_slots) {
            double4 p = M0.posq[j];
            if (M0.gid[j] < 0) {
                p.x = nan;
            }
            s_pos[threadIdx.x] = p;
            s_id[threadIdx.x] = M0.atom_id[j];
        }
        else {
            s_pos[threadIdx.x] = make_double4(nan, 0, 0, 0);
            s_id[threadIdx.x] = 0;
        }
    }
    __syncthreads();
    if (worker) {
        const int per = (kBatchTile + nsub - 1) / nsub;
        const int jb = sub * per;
        const int je = min(kBatchTile, jb + per);
        auto one = [&](int jj, double& acc) {
            const double4 pj = s_pos[jj];
            double dx = fabs(me.x - pj.x);
            double dy = fabs(me.y - pj.y);
            double dz = fabs(me.z - pj.z);
            dx -= (dx > hx) ? lx : 0.0;
            dy -= (dy > hy) ? ly : 0.0;
            dz -= (dz > hz) ? lz : 0.0;
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < cut2 && (j0 + jj) != my_slot) {
                acc += pairEnergy<KIND>(P, my_id, s_id[jj], me.w, pj.w, r2);
            }
        };
        int jj = jb;
        for (; jj + 1 < je; jj += 2) { // two independent chains
            one(jj, e0);
            one(jj + 1, e1);
        }
        if (jj < je) {
            one(jj, e0);
        }
    }
    s_red[threadIdx.x] = worker ? (e0 + e1) : 0.0;
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < nv) {
        double s = 0.0;
        for (int t = 0; t < nsub; ++t) {
            s += s_red[t * nv + threadIdx.x];
        }
        partials[static_cast<size_t>(blockIdx.x) * (2 * stride) + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// k-space part. Per tile of KT k-vectors: (1) lane ↔ k, warp ↔ moves: δ_m,k → R[m] accumulators and
// s_delta[k][m] = sqrt(A_k) δ_m,k; (2) thread ↔ 4×4 tile of G (and a k sub-group): rank-KT update.
// BT = stride / 4 ∈ {4, 8, 16}.
// ------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(kBlock)
    batchEwaldKernel(EwaldView E, const int4* __restrict__ kn, BatchBuffers cur, PhaseGeometry geo, int tiles_per_block,
                     double* __restrict__ r_partials /*[gridDim.x][stride]*/,
                     double* __restrict__ g_partials /*[gridDim.x][stride²]*/)
{
    constexpr int STRIDE = BT * 4;
    constexpr int KT = kBatchDeltaElems / STRIDE;   // k-vectors per tile: 128, 64, 32
    constexpr int NTILE = BT * BT;                  // 4×4 output tiles
    constexpr int KG = kBlock / NTILE;              // k sub-groups: 16, 4, 1
    constexpr int MPW = STRIDE / (kBlock / 32);     // moves per warp: 2, 4, 8
    constexpr int KPL = KT / 32;                    // k-vectors per lane: 4, 2, 1

    constexpr int LD = KT + 1;                      // padded leading dimension of s_delta[m][k]
    __shared__ double2 s_delta[STRIDE * LD];

    const int n = cur.in->n;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int tile_id = threadIdx.x % NTILE;
    const int kg = threadIdx.x / NTILE;
    const int ta = tile_id / BT; // this thread owns G[ta + BT·i][tm + BT·j], i, j < 4 (bank-conflict-free reads)
    const int tm = tile_id % BT;

    double racc[MPW];
#pragma unroll
    for (int i = 0; i < MPW; ++i) {
        racc[i] = 0.0;
    }
    double gacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gacc[i][j] = 0.0;
        }
    }

    // per-warp move data
    double qn[MPW], qo[MPW];
#pragma unroll
    for (int i = 0; i < MPW; ++i) {
        const int m = warp * MPW + i;
        qn[i] = (m < n) ? cur.in->pnew[m].w : 0.0;
        qo[i] = (m < n) ? cur.pold[m].w : 0.0;
    }

    const int first_tile = blockIdx.x * tiles_per_block;
    for (int tile = first_tile; tile < first_tile + tiles_per_block; ++tile) {
        const int k0 = tile * KT;
        if (k0 >= E.K) {
            break;
        }
        __syncthreads(); // previous tile's phase 2 is done with s_delta
#pragma unroll
        for (int kk = 0; kk < KPL; ++kk) {
            const int kl = lane + 32 * kk;
            const int k = k0 + kl;
            const bool valid = k < E.K;
            int4 nn = make_int4(0, 0, 0, 0);
            double2 Q = make_double2(0, 0);
            double A = 0.0;
            if (valid) {
                nn = __ldg(kn + k);
                Q = E.Q[k];
                A = E.kA[k].w;
            }
            const double sqrtA = sqrt(A);
#pragma unroll
            for (int i = 0; i < MPW; ++i) {
                const int m = warp * MPW + i;
                double2 d = make_double2(0, 0);
                if (valid && m < n) {
                    const double2 en =
                        tablePhase(cur.table + static_cast<size_t>(2 * m) * geo.table_stride, geo, nn.x, nn.y, nn.z);
                    const double2 eo =
                        tablePhase(cur.table + static_cast<size_t>(2 * m + 1) * geo.table_stride, geo, nn.x, nn.y, nn.z);
                    d.x = qn[i] * en.x - qo[i] * eo.x;
                    d.y = qn[i] * en.y - qo[i] * eo.y;
                    racc[i] += A * (2.0 * (Q.x * d.x + Q.y * d.y) + (d.x * d.x + d.y * d.y));
                }
                s_delta[m * LD + kl] = make_double2(sqrtA * d.x, sqrtA * d.y);
            }
        }
        __syncthreads();
        if (kg < KG) {
#pragma unroll 4
            for (int kl = kg; kl < KT; kl += KG) {
                double2 da[4], dm[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    da[i] = s_delta[(ta + BT * i) * LD + kl];
                    dm[i] = s_delta[(tm + BT * i) * LD + kl];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        gacc[i][j] = fma(da[i].x, dm[j].x, fma(da[i].y, dm[j].y, gacc[i][j]));
                    }
                }
            }
        }
    }

    // R[m]: warp-level sums (each warp owns its moves)
#pragma unroll
    for (int i = 0; i < MPW; ++i) {
        const double s = warpSum(racc[i]);
        if (lane == 0) {
            r_partials[static_cast<size_t>(blockIdx.x) * STRIDE + warp * MPW + i] = s;
        }
    }
    // G: reduce the k sub-groups through shared memory (fixed order), then one store per element
    __syncthreads();
    double* s_g = reinterpret_cast<double*>(s_delta); // [KG][NTILE][16] = 4096 doubles = 32 KB
    if (kg < KG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s_g[(kg * NTILE + tile_id) * 16 + i * 4 + j] = gacc[i][j];
            }
        }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < NTILE * 16; o += kBlock) {
        double s = 0.0;
        for (int g = 0; g < KG; ++g) {
            s += s_g[g * NTILE * 16 + o];
        }
        const int t = o / 16;
        const int ij = o % 16;
        const int a = (t / BT) + BT * (ij / 4);
        const int m = (t % BT) + BT * (ij % 4);
        g_partials[static_cast<size_t>(blockIdx.x) * (STRIDE * STRIDE) + a * STRIDE + m] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// final ordered sums + the pair cross terms
// result layout (doubles), S = stride:
//   [0] Σ_k A_k|Q_k|² at window start   [1] n
//   [8 + 0·S ..) u_new   [8 + 1·S ..) u_old   [8 + 2·S ..) R
//   [8 + 3·S + 0·S² ..) C_new[a][m]   [+1·S²) C_old   [+2·S²) C_max   [+3·S²) G
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t batchResultDoubles(int stride)
{
    return 8 + 3 * static_cast<size_t>(stride) + 4 * static_cast<size_t>(stride) * stride;
}

template <int KIND>
__global__ void __launch_bounds__(kBlock)
    batchFinishKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, int n_pair_blocks,
                      const double* __restrict__ pair_partials, int n_ewald_blocks,
                      const double* __restrict__ r_partials, const double* __restrict__ g_partials,
                      int n_commit_blocks, const double* __restrict__ e_partials, double* __restrict__ result)
{
    const int n = cur.in->n;
    const int with_ewald = cur.in->with_ewald;
    const int S = stride;
    const int tid = blockIdx.x * kBlock + threadIdx.x;
    const int nthreads = gridDim.x * kBlock;
    double* u = result + 8;
    double* cross = result + 8 + 3 * S;
    if (tid == 0) {
        double e = 0.0;
        if (with_ewald) {
            for (int b = 0; b < n_commit_blocks; ++b) {
                e += e_partials[b];
            }
        }
        result[0] = e;
        result[1] = static_cast<double>(n);
    }
    // pair sums: 2S columns (variant order new/old interleaved → split)
    for (int t = tid; t < 2 * S; t += nthreads) {
        double s = 0.0;
        if (t < 2 * n) {
            for (int b = 0; b < n_pair_blocks; ++b) {
                s += pair_partials[static_cast<size_t>(b) * (2 * S) + t];
            }
        }
        u[(t & 1) * S + (t >> 1)] = s;
    }
    for (int t = tid; t < S; t += nthreads) {
        double s = 0.0;
        if (with_ewald && t < n) {
            for (int b = 0; b < n_ewald_blocks; ++b) {
                s += r_partials[static_cast<size_t>(b) * S + t];
            }
        }
        u[2 * S + t] = s;
    }
    for (int t = tid; t < S * S; t += nthreads) {
        const int a = t / S;
        const int m = t % S;
        double g = 0.0;
        double cn = 0.0, co = 0.0, cmax = 0.0;
        if (a < m && m < n) {
            if (with_ewald) {
                for (int b = 0; b < n_ewald_blocks; ++b) {
                    g += g_partials[static_cast<size_t>(b) * (S * S) + t];
                }
            }
            // how the energies of move m change when the earlier move a has been accepted
            const double4 na = cur.in->pnew[a];
            const double4 oa = cur.pold[a];
            const int ida_n = cur.in->id[a];
            const int ida_o = cur.idold[a];
            const double4 nm = cur.in->pnew[m];
            const double4 om = cur.pold[m];
            const int idm_n = cur.in->id[m];
            const int idm_o = cur.idold[m];
            const double t1 = pairEnergy<KIND>(P, idm_n, ida_n, nm.w, na.w, minImageR2(M0, nm.x, nm.y, nm.z, na.x, na.y, na.z));
            const double t2 = pairEnergy<KIND>(P, idm_n, ida_o, nm.w, oa.w, minImageR2(M0, nm.x, nm.y, nm.z, oa.x, oa.y, oa.z));
            const double t3 = pairEnergy<KIND>(P, idm_o, ida_n, om.w, na.w, minImageR2(M0, om.x, om.y, om.z, na.x, na.y, na.z));
            const double t4 = pairEnergy<KIND>(P, idm_o, ida_o, om.w, oa.w, minImageR2(M0, om.x, om.y, om.z, oa.x, oa.y, oa.z));
            cn = t1 - t2;
            co = t3 - t4;
            cmax = fmax(fmax(fabs(t1), fabs(t2)), fmax(fabs(t3), fabs(t4)));
            if (t1 != t1 || t2 != t2 || t3 != t3 || t4 != t4) {
                cmax = __longlong_as_double(0x7ff0000000000000LL);
            }
        }
        cross[t] = cn;
        cross[S * S + t] = co;
        cross[2 * S * S + t] = cmax;
        cross[3 * S * S + t] = g;
    }
}

} // namespace fbdev

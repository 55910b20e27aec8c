// Runs of windows decided ON THE DEVICE (the batched sequential-MC driver).
//
// A window (fb_batch.cuh) still needs the host between its evaluation and the next one: read the result
// block back, walk the moves in order (Metropolis), report the accepted ones. Everything the walk needs is
// independent of the energies — the proposals, the Metropolis uniforms (src/montecarlo.cpp:17-34: the uniform
// is ALWAYS drawn) and the energies of the caller's own Hamiltonian terms for a single-atom move — so the
// caller can ship a RUN of up to kRunMax proposals on distinct atoms at once. The device then repeats
//
//   runSetupKernel   window = the next ≤ stride undecided moves of the run → BatchInput (+ the commit list)
//   … the kernels of a window, unchanged (pair / cross terms, phase tables, k-space, sums) …
//   runDecideKernel  the walk of B200WindowEvaluator::energies + MetropolisMonteCarlo::decideWindow:
//                    corrected energies of move m given the accepted a < m, Hamiltonian sum in term order with
//                    the reference's early exit (src/energy.cpp:1227-1247), getEnergyChange
//                    (src/montecarlo.cpp:193-209), Metropolis; accepted moves → the next window's commit list
//
// without a host round trip, and returns per move {accepted, u_new, u_old} for the host to replay into its
// Space. A window that meets a cancellation (|pair energy| in a correction ≥ the limit) stops there; the
// next window starts at that move (the cursor lives on the device), the host launches extra steps if needed.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

constexpr int kRunMax = 1024; //!< moves per run
constexpr int kRunTerms = 6;  //!< Hamiltonian terms at most

enum RunTermKind : int
{
    RUN_TERM_HOST = 0,      //!< evaluated by the caller, travels with the move
    RUN_TERM_NONBONDED = 1, //!< pair energy of the moved atom
    RUN_TERM_EWALD = 2      //!< reciprocal-space energy of the whole system
};

struct RunMove
{
    double4 pnew;
    double4 pold;
    int slot, id, idold, pad;
    double uniform;              //!< the Metropolis uniform of this move
    double host_new[kRunTerms];  //!< caller-evaluated terms in the trial state, Hamiltonian order (others unused)
    double host_old[kRunTerms];  //!< … in the accepted state
};

struct RunHeader
{
    int n_moves;
    int with_ewald;
    int n_terms;
    int pad;
    int term_kind[kRunTerms];
    double max_energy;         //!< Hamiltonian::maximumAllowedEnergy: the sum stops after a term ≥ this (or NaN)
    double cancellation_limit; //!< |pair energy| in a correction from which on a move is evaluated afresh
    double rec_prefactor;      //!< 2π lB / V
};

struct RunState
{
    int cursor;       //!< first undecided move of the run
    int window_first; //!< first move and size of the window decided last
    int window_n;
    int steps;        //!< windows evaluated for this run
    CommitList commit; //!< its accepted moves (indices into that window): what the next window has to apply
};

struct RunOutput
{
    double u_new, u_old; //!< Hamiltonian energies of the move in the trial / accepted state at its turn
    int accepted;
    int step;            //!< which window of the run decided it
};

/** first launch of a run: where the device stands (the pending accepted moves of whatever came before) */
__global__ void __launch_bounds__(kBatchMax) runInitKernel(RunState* st, CommitList pending)
{
    if (threadIdx.x == 0) {
        st->cursor = 0;
        st->window_first = 0;
        st->window_n = 0;
        st->steps = 0;
        st->commit.n = pending.n;
    }
    if (static_cast<int>(threadIdx.x) < pending.n) {
        st->commit.index[threadIdx.x] = pending.index[threadIdx.x];
    }
}

/** the next window of the run: moves [cursor, cursor + stride) */
__global__ void __launch_bounds__(kBatchMax)
    runSetupKernel(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves, const RunState* __restrict__ st,
                   BatchInput* __restrict__ in, int stride)
{
    const int cursor = st->cursor;
    const int n = max(0, min(stride, hdr->n_moves - cursor));
    const int t = threadIdx.x;
    if (t == 0) {
        in->n = n;
        in->with_ewald = hdr->with_ewald;
        in->n_groups = 0;
        in->commit.n = st->commit.n;
        in->commit_moves.n = 0;
    }
    if (t < st->commit.n) {
        in->commit.index[t] = st->commit.index[t];
    }
    if (t < n) {
        const RunMove& mv = moves[cursor + t];
        in->slot[t] = mv.slot;
        in->id[t] = mv.id;
        in->idold[t] = mv.idold;
        in->pnew[t] = mv.pnew;
        in->pold[t] = mv.pold;
    }
}

constexpr int kDecideThreads = 256;

/** dynamic shared memory of runDecideKernel: four S × (S + 1) matrices, TRANSPOSED ([a][m], padded rows) */
inline size_t runDecideSmemBytes(int stride) { return sizeof(double) * 4 * static_cast<size_t>(stride) * (stride + 1); }

/**
 * One block. Thread m < n owns move m of the window: its running pair energies and reciprocal change, corrected
 * every time an earlier move a is accepted (additions in the order a = 0, 1, … — the order of the host walk).
 * Step a of the loop: thread a decides its move, everybody else folds the outcome in.
 */
__global__ void __launch_bounds__(kDecideThreads)
    runDecideKernel(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves, RunState* __restrict__ st,
                    BatchBuffers cur, int stride, int cell_list, const double* __restrict__ result,
                    RunOutput* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char run_smem[];
    __shared__ int s_accepted[kBatchMax];
    __shared__ double s_rec_change;
    __shared__ int s_flag; // 0 rejected, 1 accepted, 2 stop (cancellation)

    const int S = stride;
    const int LD = S + 1;
    double* s_cn = reinterpret_cast<double*>(run_smem); // [a][m]
    double* s_co = s_cn + S * LD;
    double* s_cmax = s_co + S * LD;
    double* s_g = s_cmax + S * LD;
    const int n = cur.in->n;
    const int cursor = st->cursor;
    const double* u = result + 8;
    const double* cross = result + 8 + 3 * S;
    // only a < m < n is ever read
    for (int t = threadIdx.x; t < n * S; t += kDecideThreads) {
        const int m = t / S;
        const int a = t - m * S;
        if (a < m) {
            s_cn[a * LD + m] = cross[t];
            s_co[a * LD + m] = cross[S * S + t];
            s_cmax[a * LD + m] = cross[2 * S * S + t];
            s_g[a * LD + m] = cross[3 * S * S + t];
        }
    }
    const int m = threadIdx.x;
    const bool mine = m < n;
    const bool with_ewald = hdr->with_ewald != 0;
    double nb_new = 0.0, nb_old = 0.0, rec = 0.0;
    RunMove mv{};
    if (mine) {
        nb_new = u[m];
        nb_old = u[S + m];
        rec = with_ewald ? u[2 * S + m] : 0.0;
        mv = moves[cursor + m];
    }
    double rec_running = result[0]; // Σ_k A_k |Q_k|² of the window-start state (0 without Ewald)
    bool cancelled = false;         // a correction of this move met a huge pair energy
    const bool overflow = cell_list && result[2] != 0.0; // a cell bucket ran full: nothing of this window counts
    const int n_terms = hdr->n_terms;
    const double limit = hdr->max_energy;
    const double cancellation_limit = hdr->cancellation_limit;
    const double pref = hdr->rec_prefactor;
    __syncthreads();

    int n_decided = 0;
    for (int a = 0; a < n; ++a) {
        if (m == a) {
            int flag = 2;
            if (!cancelled && !overflow) {
                // Hamiltonian::energy on the trial and on the accepted state (no FMA contraction: the host adds
                // the same numbers one by one)
                double total[2];
#pragma unroll
                for (int is_old = 0; is_old < 2; ++is_old) {
                    double sum = 0.0;
                    for (int i = 0; i < n_terms; ++i) {
                        double e;
                        const int kind = hdr->term_kind[i];
                        if (kind == RUN_TERM_NONBONDED) {
                            e = is_old ? nb_old : nb_new;
                        }
                        else if (kind == RUN_TERM_EWALD) {
                            e = __dmul_rn(pref, is_old ? rec_running : __dadd_rn(rec_running, rec));
                        }
                        else {
                            e = is_old ? mv.host_old[i] : mv.host_new[i];
                        }
                        sum = __dadd_rn(sum, e);
                        if (e >= limit || e != e) {
                            break;
                        }
                    }
                    total[is_old] = sum;
                }
                const double u_new = total[0], u_old = total[1];
                // getEnergyChange, src/montecarlo.cpp:193-209
                double du;
                if (u_old != u_old && u_new == u_new) {
                    du = -__longlong_as_double(0x7ff0000000000000LL);
                }
                else if (u_new != u_new) {
                    du = __longlong_as_double(0x7ff0000000000000LL);
                }
                else if (u_new > 0.0 && isinf(u_new)) {
                    du = __longlong_as_double(0x7ff0000000000000LL);
                }
                else {
                    du = __dsub_rn(u_new, u_old);
                    if (du != du) {
                        du = 0.0;
                    }
                }
                // metropolisCriterion, src/montecarlo.cpp:17-34
                bool accept;
                if (isinf(du) && du < 0.0) {
                    accept = true;
                }
                else if (-du > 709.782712893384) {
                    accept = true;
                }
                else {
                    accept = mv.uniform <= exp(-du);
                }
                flag = accept ? 1 : 0;
                RunOutput o;
                o.u_new = u_new;
                o.u_old = u_old;
                o.accepted = flag;
                o.step = st->steps;
                out[cursor + a] = o;
                s_rec_change = rec;
            }
            s_flag = flag;
        }
        __syncthreads();
        const int flag = s_flag;
        if (flag == 2) {
            break;
        }
        n_decided = a + 1;
        if (m == a) {
            s_accepted[a] = flag;
        }
        if (flag == 1) {
            if (mine && m > a) {
                const int t = a * LD + m;
                if (!(s_cmax[t] < cancellation_limit)) {
                    cancelled = true;
                }
                nb_new = __dadd_rn(nb_new, s_cn[t]);
                nb_old = __dadd_rn(nb_old, s_co[t]);
                if (with_ewald) {
                    rec = __dadd_rn(rec, __dmul_rn(2.0, s_g[t]));
                }
            }
            if (with_ewald) {
                rec_running = __dadd_rn(rec_running, s_rec_change);
            }
        }
        __syncthreads(); // s_flag / s_rec_change are rewritten in the next step
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        for (int a = 0; a < n_decided; ++a) {
            if (s_accepted[a]) {
                st->commit.index[k++] = a;
            }
        }
        st->commit.n = k;
        st->window_first = cursor;
        st->window_n = n;
        st->cursor = cursor + n_decided;
        st->steps += 1;
    }
}

} // namespace fbdev

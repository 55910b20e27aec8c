// Runs of windows decided ON THE DEVICE (the batched sequential-MC driver).
//
// A window (fb_batch.cuh) still needs the host between its evaluation and the next one: read the result
// block back, walk the moves in order (Metropolis), report the accepted ones. Everything the walk needs is
// independent of the energies — the proposals, the Metropolis uniforms (src/montecarlo.cpp:17-34: the uniform
// is ALWAYS drawn) and the energies of the caller's own Hamiltonian terms for a single-atom move — so the
// caller can ship a RUN of up to kRunMax proposals on distinct atoms at once. The device then repeats
//
//   … the kernels of a window, unchanged (pair / cross terms, phase tables, k-space, sums) …
//   runDecideKernel  the walk of B200WindowEvaluator::energies + MetropolisMonteCarlo::decideWindow:
//                    corrected energies of move m given the accepted a < m, Hamiltonian sum in term order with
//                    the reference's early exit (src/energy.cpp:1227-1247), getEnergyChange
//                    (src/montecarlo.cpp:193-209), Metropolis; accepted moves → the next window's commit list;
//                    then the next window = the next ≤ stride undecided moves of the run → BatchInput
//
// without a host round trip, and returns per move {accepted, u_new, u_old} for the host to replay into its
// Space. A window that meets a cancellation (|pair energy| in a correction ≥ the limit) stops there; the
// next window starts at that move (the cursor lives on the device), the host launches extra steps if needed.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

constexpr int kRunMax = 1024; //!< moves per run

/** closed: the in-order sum of the caller's terms ended early (a term ≥ max_energy or NaN, src/energy.cpp:1238-1244) */
enum RunMoveFlags : int
{
    RUN_HOST_NEW_CLOSED = 1,
    RUN_HOST_OLD_CLOSED = 2,
    RUN_ALT_HOST_NEW_CLOSED = 4,
    RUN_ALT_HOST_OLD_CLOSED = 8
};

constexpr int kRunDepPrevious = 1 << 30; //!< RunMove::dep refers to the run this one is queued behind

/**
 * One proposal. A proposal on an atom that an EARLIER, still undecided proposal `dep` already moves comes in two
 * variants — it starts where that move leaves the atom: (pnew, pold, host_*) if `dep` is accepted, the *_alt
 * fields if it is rejected. The window set-up picks the variant once `dep` is decided (it never shares a window
 * with it).
 */
struct RunMove
{
    double4 pnew;
    double4 pold;
    double4 pnew_alt;
    double4 pold_alt;
    int slot, id, idold, flags;
    double uniform;   //!< the Metropolis uniform of this move
    double host_new;  //!< in-order sum of the caller's Hamiltonian terms (they precede the device terms), trial state
    double host_old;  //!< … accepted state
    double host_new_alt, host_old_alt;
    int dep;          //!< −1, or the index of the move this one depends on (| kRunDepPrevious: in the previous run)
    int pad;
};

struct RunHeader
{
    int n_moves;
    int with_ewald;
    int prepair; //!< evaluate the pair sums of a window one window ahead (see runSetupWindow)
    int pad;
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
    int rounds;       //!< rounds of the fixed-point walk, summed over the windows
    int halted;       //!< queued behind a run that did not finish in its scheduled windows: does nothing, is launched again
    int pad;
    CommitList commit; //!< its accepted moves (indices into that window): what the next window has to apply
};

struct RunOutput
{
    double u_new, u_old; //!< Hamiltonian energies of the move in the trial / accepted state at its turn
    int accepted;
    int step;            //!< which window of the run decided it
};

__device__ __forceinline__ void barrier64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

/** was the move that proposal `mv` depends on accepted? (it is decided: earlier window or earlier run) */
__device__ __forceinline__ bool runDependencyAccepted(const RunMove& mv, const RunOutput* out, const RunOutput* prev_out)
{
    return (mv.dep & kRunDepPrevious) ? prev_out[mv.dep & (kRunDepPrevious - 1)].accepted != 0 : out[mv.dep].accepted != 0;
}

/**
 * The next window of the run: moves [cursor, cursor + n), n ≤ stride, cut before the first proposal that depends on
 * a move of this very window; with the commit list the window before it left. Called by 64 threads (t = 0 … 63).
 *
 * Pair sums one window ahead: `ahead` (may be null) receives the window that will MOST LIKELY follow this one —
 * moves [cursor + n, …) — so that its pair sums can be evaluated while this window's k-space kernel runs
 * (batchPairKernel on `ahead`, then batchPairFixKernel once this window is decided). `guess` (may be null) is what
 * was predicted for THIS window a step ago: if it is exactly this window, its sums are there (pair_ready).
 */
__device__ __forceinline__ void runSetupWindow(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves,
                                               int cursor, const CommitList& commit, BatchInput* __restrict__ in,
                                               int stride, int t, const RunOutput* out /* written by this very kernel */,
                                               const RunOutput* prev_out, const BatchInput* guess,
                                               BatchInput* __restrict__ ahead)
{
    __shared__ int s_window_n, s_ahead_n, s_ahead_ok;
    const int n_moves = hdr->n_moves;
    const int n_max = max(0, min(stride, n_moves - cursor));
    if (t == 0) {
        s_window_n = n_max;
    }
    barrier64();
    int dep = -1;
    if (t < n_max) {
        dep = moves[cursor + t].dep;
        if (dep >= cursor && !(dep & kRunDepPrevious)) {
            atomicMin(&s_window_n, t);
        }
    }
    barrier64();
    const int n = s_window_n;
    if (t == 0) {
        in->n = n;
        in->with_ewald = hdr->with_ewald;
        in->n_groups = 0;
        in->commit.n = commit.n;
        in->commit_moves.n = 0;
        in->first = cursor;
        in->pair_ready = (guess != nullptr && guess->pair_ready && guess->first == cursor && guess->n == n && n > 0) ? 1 : 0;
    }
    if (t < commit.n) {
        in->commit.index[t] = commit.index[t];
    }
    if (t < n) {
        const RunMove& mv = moves[cursor + t];
        const bool alt = dep >= 0 && !runDependencyAccepted(mv, out, prev_out);
        in->slot[t] = mv.slot;
        in->id[t] = mv.id;
        in->idold[t] = mv.idold;
        in->pnew[t] = alt ? mv.pnew_alt : mv.pnew;
        in->pold[t] = alt ? mv.pold_alt : mv.pold;
    }
    if (ahead == nullptr) {
        return;
    }
    // the window after this one, if this one decides all its moves; a proposal that depends on a move of THIS window
    // has no known start yet: no prediction then
    const int first = cursor + n;
    const int a_max = hdr->prepair ? max(0, min(stride, n_moves - first)) : 0;
    if (t == 0) {
        s_ahead_n = a_max;
        s_ahead_ok = 1;
    }
    barrier64();
    int adep = -1;
    if (t < a_max) {
        adep = moves[first + t].dep;
        if (adep >= 0 && !(adep & kRunDepPrevious)) {
            if (adep >= first) {
                atomicMin(&s_ahead_n, t);
            }
            else if (adep >= cursor) {
                s_ahead_ok = 0;
            }
        }
    }
    barrier64();
    const int an = s_ahead_ok ? s_ahead_n : 0;
    if (t == 0) {
        ahead->n = an;
        ahead->with_ewald = hdr->with_ewald;
        ahead->n_groups = 0;
        ahead->commit.n = 0;
        ahead->commit_moves.n = 0;
        ahead->first = first;
        ahead->pair_ready = an > 0 ? 1 : 0;
    }
    if (t < an) {
        const RunMove& mv = moves[first + t];
        const bool alt = adep >= 0 && !runDependencyAccepted(mv, out, prev_out); // decided: before this window
        ahead->slot[t] = mv.slot;
        ahead->id[t] = mv.id;
        ahead->idold[t] = mv.idold;
        ahead->pnew[t] = alt ? mv.pnew_alt : mv.pnew;
        ahead->pold[t] = alt ? mv.pold_alt : mv.pold;
    }
}

/**
 * First launch of a run: where the device stands (the pending accepted moves of whatever came before) and the
 * first window. Every later window is set up by the runDecideKernel of the window before it.
 */
__global__ void __launch_bounds__(kBatchMax)
    runInitKernel(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves, RunState* st, CommitList pending,
                  BatchInput* __restrict__ in, int stride, const RunOutput* __restrict__ out,
                  const RunOutput* __restrict__ prev_out, BatchInput* __restrict__ ahead)
{
    if (threadIdx.x == 0) {
        st->cursor = 0;
        st->window_first = 0;
        st->window_n = 0;
        st->steps = 0;
        st->rounds = 0;
        st->halted = 0;
        st->commit.n = pending.n;
    }
    if (static_cast<int>(threadIdx.x) < pending.n) {
        st->commit.index[threadIdx.x] = pending.index[threadIdx.x];
    }
    runSetupWindow(hdr, moves, 0, pending, in, stride, threadIdx.x, out, prev_out, nullptr, ahead);
}

/**
 * First launch of a run queued BEHIND a run that is still in flight: it starts from what that run left (the
 * accepted moves of its last window) — unless that run did not get through in the windows scheduled for it (a
 * cancellation or a full cell bucket cost it extra windows): then this run halts, its windows do nothing, and the
 * host launches it again once the earlier run is complete.
 */
__global__ void __launch_bounds__(kBatchMax)
    runChainKernel(const RunHeader* __restrict__ prev_hdr, const RunState* __restrict__ prev, const RunHeader* __restrict__ hdr,
                   const RunMove* __restrict__ moves, RunState* st, BatchInput* __restrict__ in, int stride,
                   const RunOutput* __restrict__ out, const RunOutput* __restrict__ prev_out,
                   BatchInput* __restrict__ ahead)
{
    const bool halted = prev->cursor < prev_hdr->n_moves || prev->halted != 0;
    if (threadIdx.x == 0) {
        st->cursor = 0;
        st->window_first = 0;
        st->window_n = 0;
        st->steps = 0;
        st->rounds = 0;
        st->halted = halted ? 1 : 0;
        st->commit.n = 0;
    }
    if (halted) {
        if (threadIdx.x == 0) {
            in->n = 0;
            in->with_ewald = hdr->with_ewald;
            in->n_groups = 0;
            in->commit.n = 0;
            in->commit_moves.n = 0;
            in->pair_ready = 0;
            ahead->n = 0;
            ahead->pair_ready = 0;
        }
        return;
    }
    runSetupWindow(hdr, moves, 0, prev->commit, in, stride, threadIdx.x, out, prev_out, nullptr, ahead);
}

/** a window of a run that is continued after the host looked at it: set up from the cursor on the device */
__global__ void __launch_bounds__(kBatchMax)
    runSetupKernel(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves, const RunState* __restrict__ st,
                   BatchInput* __restrict__ in, int stride, const RunOutput* __restrict__ out,
                   const RunOutput* __restrict__ prev_out, BatchInput* __restrict__ ahead)
{
    runSetupWindow(hdr, moves, st->cursor, st->commit, in, stride, threadIdx.x, out, prev_out, nullptr, ahead);
}

constexpr int kDecideThreads = 1024; //!< 16 threads per move of the window
static_assert(kDecideThreads == kFinishThreads, "the last block of the finish kernel walks the window");
static_assert(kDecideThreads == 16 * kBatchMax, "16 threads per move");

/** dynamic shared memory of the walk: three S × (S + 1) matrices, TRANSPOSED ([a][m], padded rows) */
inline size_t runDecideSmemBytes(int stride) { return sizeof(double) * 3 * static_cast<size_t>(stride) * (stride + 1); }

/** Σ over the 16 lanes of a move (half a warp), fixed tree */
__device__ __forceinline__ double sumOverMoveLanes(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 8, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 4, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 2, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 1, 16);
    return v;
}

/**
 * The in-order walk of a window as a fixed-point iteration, by ONE block of 1024 threads. All threads stage the
 * correction matrices in shared memory (transposed: the lanes of a move read column m of the rows a, conflict-free),
 * then 16 threads own each move m. Given a GUESS of which moves are accepted (a 64-bit mask, initially none), the
 * energies of every move are evaluated as the host walk would — corrections of the accepted a < m (each of the 16
 * lanes adds its a ≡ lane (mod 16) in ascending order, the lanes are added in a fixed tree), Hamiltonian sum in term
 * order with the early exit, getEnergyChange, Metropolis — all moves in parallel. The decision of move m only depends
 * on the bits below m, so everything up to and including the first move whose decision contradicts the guess is
 * FINAL; the guess is replaced by the new decisions and the iteration repeats from there. It ends after (1 + number
 * of decisions that the corrections overturned) rounds — a handful — instead of n dependent steps. (With one thread
 * per move the two in-order sums over the ≈ 57 accepted earlier moves of an S1 window were 2 × 57 dependent
 * shared-memory round trips per round: a third of the kernel.) A move whose correction meets a huge pair energy
 * (cancellation) ends the window: it is re-evaluated at the head of the next one. At the end the next window of the
 * run is set up (`next`).
 */
__device__ __forceinline__ void runDecideBlock(unsigned char* run_smem, const RunHeader* __restrict__ hdr,
                                               const RunMove* __restrict__ moves, RunState* __restrict__ st,
                                               BatchBuffers cur, BatchInput* __restrict__ next, int stride, int cell_list,
                                               const double* __restrict__ result, RunOutput* __restrict__ out,
                                               const RunOutput* __restrict__ prev_out, const BatchInput* predicted,
                                               BatchInput* __restrict__ ahead)
{
    __shared__ CommitList s_commit;
    __shared__ double s_rec[kBatchMax];
    __shared__ unsigned long long s_cancel[kBatchMax]; // bit a of row m: the correction (a, m) meets a huge pair energy
    __shared__ unsigned s_bits[3][kDecideThreads / 32]; // per warp (two moves): accepted, mismatch with the guess, stop
    __shared__ unsigned long long s_mask[3];

    const int S = stride;
    const int LD = S + 1;
    double* s_cn = reinterpret_cast<double*>(run_smem); // [a][m]
    double* s_co = s_cn + S * LD;
    double* s_g = s_co + S * LD;
    const int halted = st->halted;
    const int n = cur.in->n;
    const int cursor = st->cursor;
    const int step = st->steps;
    const double* u = result + 8;
    const double* cross = result + 8 + 3 * S;
    const int shift = S == 64 ? 6 : (S == 32 ? 5 : 4);
    const double cancellation_limit = hdr->cancellation_limit;
    if (threadIdx.x < kBatchMax) {
        s_cancel[threadIdx.x] = 0ull;
    }
    __syncthreads();
    // all S rows (rows ≥ n are zero padding): the staging does not wait for n, and it has no branch, so that the
    // loads of all rounds are in flight together
#pragma unroll 4
    for (int t = threadIdx.x; t < S * S; t += kDecideThreads) {
        const int m = t >> shift;
        const int a = t & (S - 1);
        const double v0 = __ldcg(cross + t), v1 = __ldcg(cross + S * S + t), v2 = __ldcg(cross + 2 * S * S + t),
                     v3 = __ldcg(cross + 3 * S * S + t);
        s_cn[a * LD + m] = v0;
        s_co[a * LD + m] = v1;
        s_g[a * LD + m] = v3;
        if (!(v2 < cancellation_limit)) {
            atomicOr(&s_cancel[m], 1ull << a);
        }
    }
    if (halted) { // behind an unfinished run: nothing was evaluated, nothing is decided, nothing follows
        if (threadIdx.x == 0) {
            next->n = 0;
            next->with_ewald = hdr->with_ewald;
            next->n_groups = 0;
            next->commit.n = 0;
            next->commit_moves.n = 0;
            next->pair_ready = 0;
            ahead->n = 0;
            ahead->pair_ready = 0;
        }
        return;
    }
    if (threadIdx.x < kBatchMax && cursor + n + static_cast<int>(threadIdx.x) < hdr->n_moves) {
        // the proposals the next window most likely starts with (this window decided completely)
        const char* ahead_moves = reinterpret_cast<const char*>(moves + cursor + n + threadIdx.x);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ahead_moves));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ahead_moves + 128));
    }
    const int m = threadIdx.x >> 4;  // the move of this thread
    const int part = threadIdx.x & 15;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool mine = m < n;
    const bool with_ewald = hdr->with_ewald != 0;
    double u_new0 = 0.0, u_old0 = 0.0, rec0 = 0.0, uniform = 0.0, host_new = 0.0, host_old = 0.0;
    int flags = 0;
    if (mine) {
        u_new0 = __ldcg(u + m);
        u_old0 = __ldcg(u + S + m);
        rec0 = with_ewald ? __ldcg(u + 2 * S + m) : 0.0;
        const RunMove& mv = moves[cursor + m];
        const bool alt = mv.dep >= 0 && !runDependencyAccepted(mv, out, prev_out); // as the window set-up chose
        uniform = mv.uniform;
        host_new = alt ? mv.host_new_alt : mv.host_new;
        host_old = alt ? mv.host_old_alt : mv.host_old;
        flags = alt ? (mv.flags >> 2) : mv.flags;
    }
    const double rec_start = __ldcg(result); // Σ_k A_k |Q_k|² of the window-start state (0 without Ewald)
    const bool overflow = cell_list && __ldcg(result + 2) != 0.0; // a cell bucket ran full: nothing of this window counts
    const double limit = hdr->max_energy;
    const double pref = hdr->rec_prefactor;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();

    unsigned long long guess = 0ull; // bit a: move a taken as accepted
    int fixed = 0;                   // decisions [0, fixed) are final
    int n_decided = n;
    int decision = 0;                // of this thread's move: 0 rejected, 1 accepted, 2 stop
    double total_new = 0.0, total_old = 0.0;
    int rounds = 0;
    if (overflow) {
        n_decided = 0;
        fixed = n;
    }
    while (fixed < n) { // uniform over the block
        rounds++;
        const bool active = mine && m >= fixed;
        const unsigned long long below = guess & ((1ull << m) - 1ull);
        // corrections for the accepted earlier moves: this lane's share a = part, part + 16, …
        double pn = 0.0, po = 0.0, pg = 0.0;
        if (active) {
#pragma unroll
            for (int j = 0; j < kBatchMax / 16; ++j) {
                const int a = part + 16 * j;
                if ((below >> a) & 1ull) {
                    const int t = a * LD + m;
                    pn += s_cn[t];
                    po += s_co[t];
                    pg += s_g[t];
                }
            }
        }
        pn = sumOverMoveLanes(pn);
        po = sumOverMoveLanes(po);
        pg = sumOverMoveLanes(pg);
        const bool cancelled = active && (s_cancel[m] & below) != 0ull;
        const double nb_new = u_new0 + pn, nb_old = u_old0 + po, rec = rec0 + 2.0 * pg;
        if (active && part == 0) {
            s_rec[m] = rec;
        }
        __syncthreads();
        double pr = 0.0;
        if (active && with_ewald) { // reciprocal sum of the state move m starts from: the accepted earlier moves' changes
#pragma unroll
            for (int j = 0; j < kBatchMax / 16; ++j) {
                const int a = part + 16 * j;
                if ((below >> a) & 1ull) {
                    pr += s_rec[a];
                }
            }
        }
        pr = sumOverMoveLanes(pr);
        if (active) {
            const double rec_running = rec_start + pr;
            // Hamiltonian::energy on the trial and on the accepted state: the caller's terms, the non-bonded term,
            // the Ewald term; the sum stops after a term ≥ limit or NaN (no FMA contraction: the host adds the same
            // numbers one by one)
            total_new = host_new;
            if (!(flags & RUN_HOST_NEW_CLOSED)) {
                total_new = __dadd_rn(total_new, nb_new);
                if (with_ewald && !(nb_new >= limit || nb_new != nb_new)) {
                    total_new = __dadd_rn(total_new, __dmul_rn(pref, __dadd_rn(rec_running, rec)));
                }
            }
            total_old = host_old;
            if (!(flags & RUN_HOST_OLD_CLOSED)) {
                total_old = __dadd_rn(total_old, nb_old);
                if (with_ewald && !(nb_old >= limit || nb_old != nb_old)) {
                    total_old = __dadd_rn(total_old, __dmul_rn(pref, rec_running));
                }
            }
            // getEnergyChange, src/montecarlo.cpp:193-209
            double du;
            if (total_old != total_old && total_new == total_new) {
                du = -inf;
            }
            else if (total_new != total_new) {
                du = inf;
            }
            else if (total_new > 0.0 && isinf(total_new)) {
                du = inf;
            }
            else {
                du = __dsub_rn(total_new, total_old);
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
                accept = uniform <= exp(-du);
            }
            decision = cancelled ? 2 : (accept ? 1 : 0);
        }
        // the decisions of the two moves of a warp sit in lanes 0 and 16
        const bool taken = ((guess >> m) & 1ull) != 0ull;
        const bool speaker = part == 0;
        const unsigned b_acc = __ballot_sync(0xffffffffu, speaker && mine && decision == 1);
        const unsigned b_mis = __ballot_sync(0xffffffffu, speaker && active && ((decision == 1) != taken));
        const unsigned b_stop = __ballot_sync(0xffffffffu, speaker && active && decision == 2);
        if (lane == 0) {
            s_bits[0][warp] = (b_acc & 1u) | ((b_acc >> 15) & 2u);
            s_bits[1][warp] = (b_mis & 1u) | ((b_mis >> 15) & 2u);
            s_bits[2][warp] = (b_stop & 1u) | ((b_stop >> 15) & 2u);
        }
        __syncthreads();
        if (warp < 3) { // warp k assembles mask k: warp w of the block holds the moves 2w, 2w + 1
            const unsigned v = s_bits[warp][lane];
            const unsigned lo = __reduce_or_sync(0xffffffffu, lane < 16 ? v << (2 * lane) : 0u);
            const unsigned hi = __reduce_or_sync(0xffffffffu, lane >= 16 ? v << (2 * (lane - 16)) : 0u);
            if (lane == 0) {
                s_mask[warp] = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
            }
        }
        __syncthreads();
        const unsigned long long accepted = s_mask[0];
        const unsigned long long mismatch = s_mask[1];
        const unsigned long long stops = s_mask[2];
        const int first_mismatch = mismatch ? __ffsll(static_cast<long long>(mismatch)) - 1 : n;
        const int first_stop = stops ? __ffsll(static_cast<long long>(stops)) - 1 : n;
        guess = accepted; // final below `fixed` (those threads kept their decision), the new guess above
        if (first_stop <= first_mismatch && first_stop < n) { // a final stop: the window ends there
            n_decided = first_stop;
            fixed = n;
        }
        else {
            fixed = min(n, first_mismatch + 1);
        }
    }
    const unsigned long long decided_mask = n_decided >= 64 ? ~0ull : ((1ull << n_decided) - 1ull);
    const unsigned long long accepted = guess & decided_mask;
    if (mine && part == 0 && m < n_decided) {
        RunOutput o;
        o.u_new = total_new;
        o.u_old = total_old;
        o.accepted = decision == 1 ? 1 : 0;
        o.step = step;
        out[cursor + m] = o;
        if (decision == 1) {
            const int k = __popcll(accepted & ((1ull << m) - 1ull));
            s_commit.index[k] = m;
            st->commit.index[k] = m;
        }
    }
    if (threadIdx.x == 0) {
        const int k = __popcll(accepted);
        s_commit.n = k;
        st->commit.n = k;
        st->window_first = cursor;
        st->window_n = n;
        st->cursor = cursor + n_decided;
        st->steps = step + 1;
        st->rounds += rounds;
    }
    __threadfence_block(); // the set-up below reads out[] entries written by other threads of this block
    __syncthreads();
    if (threadIdx.x >= kBatchMax) {
        return; // the set-up of the next window is the business of the first two warps (named barrier)
    }
    runSetupWindow(hdr, moves, cursor + n_decided, s_commit, next, stride, threadIdx.x, out, prev_out, predicted, ahead);
}

__global__ void __launch_bounds__(kDecideThreads)
    runDecideKernel(const RunHeader* __restrict__ hdr, const RunMove* __restrict__ moves, RunState* __restrict__ st,
                    BatchBuffers cur, BatchInput* __restrict__ next, int stride, int cell_list,
                    const double* __restrict__ result, RunOutput* __restrict__ out,
                    const RunOutput* __restrict__ prev_out, const BatchInput* predicted, BatchInput* __restrict__ ahead)
{
    extern __shared__ __align__(16) unsigned char run_smem[];
    runDecideBlock(run_smem, hdr, moves, st, cur, next, stride, cell_list, result, out, prev_out, predicted, ahead);
}

/**
 * The tail of a window of a run in ONE launch: the sums of windowFinishKernel and, by whichever block finishes last
 * (ticket), the walk and the set-up of the next window. The result block stays in L2 between the two.
 */
#ifndef FB_TAIL_BLOCKS_PER_SM
#define FB_TAIL_BLOCKS_PER_SM 1
#endif
template <int KIND>
__global__ void __launch_bounds__(kFinishThreads, FB_TAIL_BLOCKS_PER_SM)
    windowTailKernel(SlotView M0, PotParams P, BatchBuffers cur, int stride, int with_ewald, int n_rows, int n_e_rows,
                     const double* __restrict__ r_partials, const double* __restrict__ g_partials,
                     const double* __restrict__ e_partials, int n_pair_blocks, const double* __restrict__ pair_partials,
                     double* __restrict__ result, unsigned* __restrict__ ticket, const RunHeader* __restrict__ hdr,
                     const RunMove* __restrict__ moves, RunState* __restrict__ st, BatchInput* __restrict__ next,
                     RunOutput* __restrict__ out, const RunOutput* __restrict__ prev_out, const BatchInput* predicted,
                     BatchInput* __restrict__ ahead, int cross_done)
{
    extern __shared__ __align__(16) unsigned char run_smem[];
    __shared__ unsigned s_last;
    FB_GRID_DEPENDENCY_WAIT(); // the partial sums are the k-space kernel's
    FB_LAUNCH_DEPENDENTS();
#ifndef FB_TAIL_ONE_WAVE
#define FB_TAIL_ONE_WAVE 0
#endif
    const int k_blocks = kspaceFinishGrid(stride);
#if FB_TAIL_ONE_WAVE && FB_CROSS_PACKED
#error "FB_TAIL_ONE_WAVE walks the pair outputs with pairFinishWarp: build it with -DFB_CROSS_PACKED=0"
#endif
#if FB_TAIL_ONE_WAVE
    // One wave: a block of 1024 threads at 64 registers owns an SM, and the walk's shared memory is reserved for every
    // block, so a grid of (column blocks + pair blocks) = 263 ran as two waves on 148 SMs, each paying the full latency
    // of its loads. The grid is now at most one block per SM; a block takes the column blocks b, b + grid, … and
    // then the pair outputs of its warps.
    for (int kb = blockIdx.x; kb < k_blocks; kb += gridDim.x) {
        kspaceFinishBlock(cur, stride, with_ewald, n_rows, n_e_rows, r_partials, g_partials, e_partials, result, kb);
        __syncthreads(); // kspaceFinishBlock's staging area is free again
    }
    const int n_outputs = cross_done ? 2 * stride : 2 * stride + stride * stride;
    for (int w = static_cast<int>((blockIdx.x * kFinishThreads + threadIdx.x) >> 5); w < n_outputs;
         w += static_cast<int>(gridDim.x) * (kFinishThreads / 32)) {
        pairFinishWarp<KIND>(M0, P, cur, stride, n_pair_blocks, pair_partials, 0, nullptr, result, nullptr, nullptr, w);
    }
    // the writes of all threads of the block → (barrier) → thread 0 → (fence, cumulative) → device-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
#else
    if (static_cast<int>(blockIdx.x) < k_blocks) {
        kspaceFinishBlock(cur, stride, with_ewald, n_rows, n_e_rows, r_partials, g_partials, e_partials, result, blockIdx.x);
    }
    else {
        const int w = static_cast<int>(((blockIdx.x - k_blocks) * kFinishThreads + threadIdx.x) >> 5);
#if FB_CROSS_PACKED
        if (w < 2 * stride) { // the 2S pair sums, one warp each; then the S² cross terms, 16 per warp
            pairFinishWarp<KIND>(M0, P, cur, stride, n_pair_blocks, pair_partials, 0, nullptr, result, nullptr, nullptr, w);
        }
        else if (!cross_done) {
            pairCrossPacked<KIND>(M0, P, cur, stride, result, w - 2 * stride);
        }
#else
        if (!cross_done || w < 2 * stride) { // cross_done: windowCrossKernel took the cross terms at the start of the window
            pairFinishWarp<KIND>(M0, P, cur, stride, n_pair_blocks, pair_partials, 0, nullptr, result, nullptr, nullptr, w);
        }
#endif
    }
#ifndef FB_TAIL_LIGHT_FENCE
#define FB_TAIL_LIGHT_FENCE 1
#endif
#if FB_TAIL_LIGHT_FENCE
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence(); // cumulative: orders the writes of the whole block (seen through the barrier) before the ticket
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
#else
    __threadfence(); // this thread's part of the result block is visible device-wide before the ticket is drawn
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
#endif
    __syncthreads();
#endif
    if (!s_last) {
        return;
    }
    if (threadIdx.x == 0) {
        *ticket = 0u; // for the next window
    }
    __threadfence();
    runDecideBlock(run_smem, hdr, moves, st, cur, next, stride, 0, result, out, prev_out, predicted, ahead);
}

/** what the walk of a window of a run needs (launchWindow → windowTailKernel / runDecideKernel) */
struct RunTail
{
    const RunHeader* hdr;
    const RunMove* moves;
    RunState* st;
    BatchInput* next;
    RunOutput* out;
    const RunOutput* prev_out;
    const BatchInput* predicted;
    BatchInput* ahead;
};

} // namespace fbdev

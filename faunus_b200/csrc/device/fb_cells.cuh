// Device cell list for the windowed pair evaluation at large N (K9 of SURVEY §2.3). The reference's
// Nonbonded never uses a cell list (brute force, src/energy.h:1182-1195); its celllistimpl.h only fixes the
// geometry convention used here: cell edge = box / floor(box / cutoff) (src/celllistimpl.h:223-236).
//
//  * buckets of fixed capacity, one per cell; the entries of a bucket are kept SORTED BY PARTICLE SLOT, so
//    the list is a pure function of the positions (no dependence on build order or move history) and every
//    sum over it is reproducible;
//  * built on the device when windowed evaluation starts (atomic append + per-cell insertion sort),
//    updated incrementally (one thread per accepted move, ≤ 64 per window: tombstone, append, re-sort);
//  * only for all-atomic systems with a finite cutoff in a fully periodic orthogonal cell with ≥ 3 cells
//    per axis; everything else takes the brute-force kernels.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

struct CellGrid
{
    int n[3];          //!< cells per axis (≥ 3)
    double inv_edge[3];
    double half[3];
    int cap;           //!< bucket capacity
    int* count;        //!< [n_cells]
    int* bucket;       //!< [n_cells][cap] particle slots, ascending
    int* overflow;     //!< set when a bucket ran full
};

__device__ __forceinline__ int cellAxis(const CellGrid& g, int axis, double x)
{
    int c = static_cast<int>(floor((x + g.half[axis]) * g.inv_edge[axis]));
    c = c < 0 ? 0 : c;
    return c >= g.n[axis] ? g.n[axis] - 1 : c;
}

__device__ __forceinline__ int cellOf(const CellGrid& g, const double4& p)
{
    return (cellAxis(g, 0, p.x) * g.n[1] + cellAxis(g, 1, p.y)) * g.n[2] + cellAxis(g, 2, p.z);
}

__global__ void cellAppendKernel(SlotView V, CellGrid g)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V.n_slots || V.gid[j] < 0) {
        return;
    }
    const int c = cellOf(g, V.posq[j]);
    const int at = atomicAdd(g.count + c, 1);
    if (at < g.cap) {
        g.bucket[static_cast<size_t>(c) * g.cap + at] = j;
    }
    else {
        *g.overflow = 1;
    }
}

/** ascending slots inside every bucket (insertion sort, ≤ cap entries) */
__global__ void cellSortKernel(CellGrid g, int n_cells)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) {
        return;
    }
    int* b = g.bucket + static_cast<size_t>(c) * g.cap;
    const int n = min(g.count[c], g.cap);
    for (int i = 1; i < n; ++i) {
        const int key = b[i];
        int k = i - 1;
        while (k >= 0 && b[k] > key) {
            b[k + 1] = b[k];
            --k;
        }
        b[k + 1] = key;
    }
}

/**
 * After batchPrepKernel: the accepted moves of the previous window change cells. One block, one thread per
 * accepted move: tombstone the slot in its old bucket, append it to the new one, then every touched bucket is
 * re-sorted (tombstones sink to the end and are trimmed) by the first thread that lists it — the result is the
 * sorted set of slots again, whatever the order of the atomics.
 */
__global__ void __launch_bounds__(2 * kBatchMax) cellCommitKernel(CellGrid g, BatchBuffers cur, BatchBuffers prev)
{
    const CommitList& commit = cur.in->commit;
    __shared__ int s_cell[2 * kBatchMax]; // [a] old cell, [n + a] new cell; −1: nothing to do
    constexpr int tombstone = 0x7fffffff;
    const int n = commit.n;
    const int a = threadIdx.x;
    int slot = -1, c0 = -1, c1 = -1;
    if (a < n) {
        const int m = commit.index[a];
        slot = prev.in->slot[m];
        c0 = cellOf(g, prev.pold[m]);
        c1 = cellOf(g, prev.in->pnew[m]);
        if (c0 == c1) {
            c0 = c1 = -1;
        }
        s_cell[a] = c0;
        s_cell[n + a] = c1;
    }
    if (c0 >= 0) {
        int* b0 = g.bucket + static_cast<size_t>(c0) * g.cap;
        const int n0 = min(g.count[c0], g.cap);
        for (int i = 0; i < n0; ++i) {
            if (b0[i] == slot) {
                b0[i] = tombstone;
            }
        }
    }
    __syncthreads();
    if (c1 >= 0) {
        const int at = atomicAdd(g.count + c1, 1);
        if (at < g.cap) {
            g.bucket[static_cast<size_t>(c1) * g.cap + at] = slot;
        }
        else {
            *g.overflow = 1;
        }
    }
    __syncthreads();
    if (a < 2 * n) {
        const int c = s_cell[a];
        bool mine = c >= 0;
        for (int t = 0; t < a && mine; ++t) {
            mine = s_cell[t] != c;
        }
        if (mine) {
            int* b = g.bucket + static_cast<size_t>(c) * g.cap;
            const int cnt = min(g.count[c], g.cap);
            for (int i = 1; i < cnt; ++i) {
                const int key = b[i];
                int k = i - 1;
                while (k >= 0 && b[k] > key) {
                    b[k + 1] = b[k];
                    --k;
                }
                b[k + 1] = key;
            }
            int live = cnt;
            while (live > 0 && b[live - 1] == tombstone) {
                --live;
            }
            g.count[c] = live;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pair part of a window through the cell list: one block per move variant, one WARP per neighbour cell
// (27 warps), lanes the bucket entries. The two pair sums of each move go straight into the result block
// (no partials); the 27 cell sums are added in cell order.
// ------------------------------------------------------------------------------------------------
constexpr int kCellThreads = 27 * 32;

template <int KIND>
__global__ void __launch_bounds__(kCellThreads)
    batchPairCellKernel(SlotView M0, PotParams P, CellGrid g, BatchBuffers cur, double cut2, int stride,
                        double* __restrict__ result)
{
    __shared__ double s_sum[27];
    const int v = blockIdx.x;
    const int m = v >> 1;
    const int lane = threadIdx.x & 31;
    const int nb = threadIdx.x >> 5;
    const int my_slot = cur.in->slot[m];
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
    int ox = cellAxis(g, 0, a.x) + nb / 9 - 1, oy = cellAxis(g, 1, a.y) + (nb / 3) % 3 - 1,
        oz = cellAxis(g, 2, a.z) + nb % 3 - 1;
    ox = ox < 0 ? ox + g.n[0] : (ox >= g.n[0] ? ox - g.n[0] : ox);
    oy = oy < 0 ? oy + g.n[1] : (oy >= g.n[1] ? oy - g.n[1] : oy);
    oz = oz < 0 ? oz + g.n[2] : (oz >= g.n[2] ? oz - g.n[2] : oz);
    const int c = (ox * g.n[1] + oy) * g.n[2] + oz;
    const int cnt = min(g.count[c], g.cap);
    const int* b = g.bucket + static_cast<size_t>(c) * g.cap;
    double e = 0.0;
    for (int t = lane; t < cnt; t += 32) {
        const int j = b[t];
        if (j == my_slot) {
            continue;
        }
        const double4 pj = M0.posq[j];
        const double r2 = minImageR2(M0, a.x, a.y, a.z, pj.x, pj.y, pj.z);
        if (r2 < cut2) {
            e += pairEnergy<KIND>(P, id, M0.atom_id[j], a.w, pj.w, r2);
        }
    }
    e = warpSum(e);
    if (lane == 0) {
        s_sum[nb] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 27; ++w) {
            s += s_sum[w];
        }
        result[8 + (v & 1) * stride + m] = s;
    }
}

} // namespace fbdev

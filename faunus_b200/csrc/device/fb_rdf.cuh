// Atomic radial distribution function: the distance histogram of AtomRDF (src/analysis.cpp:1556-1600) on the
// device mirror (the dominant non-energy cost of examples/bulk, `atomrdf` every 10 steps).
//
//   id1 != id2 : all pairs (i of type id1, j of type id2)            AtomRDF::sampleDifferent
//   id1 == id2 : all pairs i < j of that type                        AtomRDF::sampleIdentical
//   distance   : minimum-image VECTOR a − b (fold by ±L when |d| > L/2, src/geometry.h:429-458), r = |d|
//   bin        : floor(r / dr) (Equidistant2DTable<double,double>, centerbin = false, xmin = 0,
//                src/aux/equidistant_table.h:32-40); with a slice direction the in-plane distance decides
//                whether the pair counts and the along-axis distance is binned (src/analysis.cpp:1559-1564)
//
// Integer work: the particles of the two types are compacted first (stable), then a block takes a 256 × 256 tile
// of pairs of the two lists, counts into a shared-memory histogram (32-bit, ≤ 65 536 increments per block) and
// adds it to the 64-bit global one — exact and order independent.
// The arithmetic of r is spelled out without FMA contraction so that a distance lands in the same bin as on
// the host.
#pragma once
#include "fb_kernels.cuh"

namespace fbdev {

constexpr int kRdfTile = 256;
constexpr int kRdfMaxBins = 12288; //!< 48 kB of shared counters

/**
 * Positions of the active particles of type id1 (list 0) and id2 (list 1; unused when the types are equal),
 * compacted IN STORAGE ORDER so that every thread of the histogram kernel has a particle and every inner iteration
 * a pair. Stable (one block walks the mirror 1024 slots at a time, ballot + scan): the ranks of a sharded sample
 * must cut the very same lists into tiles.
 */
constexpr int kRdfCompactThreads = 1024;

__global__ void __launch_bounds__(kRdfCompactThreads)
    atomRdfCompactKernel(SlotView V, int id1, int id2, double4* __restrict__ list0, double4* __restrict__ list1,
                         int* __restrict__ n_list /*[2]*/)
{
    __shared__ int s_warp[2][32];
    __shared__ int s_total[2];
    __shared__ int s_base[2];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    if (threadIdx.x < 2) {
        s_base[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int start = 0; start < V.n_slots; start += kRdfCompactThreads) {
        const int j = start + threadIdx.x;
        bool f0 = false, f1 = false;
        double4 p = make_double4(0, 0, 0, 0);
        if (j < V.n_slots && V.gid[j] >= 0) {
            const int id = V.atom_id[j];
            f0 = id == id1;
            f1 = id == id2 && id1 != id2;
            p = V.posq[j];
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, f0);
        const unsigned b1 = __ballot_sync(0xffffffffu, f1);
        if (lane == 0) {
            s_warp[0][warp] = __popc(b0);
            s_warp[1][warp] = __popc(b1);
        }
        __syncthreads();
        if (warp < 2) { // warp k scans the 32 warp counts of list k
            const int count = s_warp[warp][lane];
            int inclusive = count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, inclusive, o);
                if (lane >= o) {
                    inclusive += up;
                }
            }
            s_warp[warp][lane] = inclusive - count;
            if (lane == 31) {
                s_total[warp] = inclusive;
            }
        }
        __syncthreads();
        if (f0) {
            list0[s_base[0] + s_warp[0][warp] + __popc(b0 & below)] = p;
        }
        if (f1) {
            list1[s_base[1] + s_warp[1][warp] + __popc(b1 & below)] = p;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            s_base[threadIdx.x] += s_total[threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) {
        n_list[threadIdx.x] = s_base[threadIdx.x];
    }
}

/** the same for the mass centres of the active molecular groups of kind molid1 / molid2 ("molrdf") */
__global__ void __launch_bounds__(kRdfCompactThreads)
    moleculeRdfCompactKernel(SlotView V, int molid1, int molid2, double4* __restrict__ list0, double4* __restrict__ list1,
                             int* __restrict__ n_list /*[2]*/)
{
    __shared__ int s_warp[2][32];
    __shared__ int s_total[2];
    __shared__ int s_base[2];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    if (threadIdx.x < 2) {
        s_base[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int start = 0; start < V.n_groups; start += kRdfCompactThreads) {
        const int g = start + threadIdx.x;
        bool f0 = false, f1 = false;
        double4 p = make_double4(0, 0, 0, 0);
        if (g < V.n_groups && V.gsize[g] > 0) {
            const int molid = V.ginfo[g] >> 8;
            f0 = molid == molid1;
            f1 = molid == molid2 && molid1 != molid2;
            p = V.gcm[g];
        }
        const unsigned b0 = __ballot_sync(0xffffffffu, f0);
        const unsigned b1 = __ballot_sync(0xffffffffu, f1);
        if (lane == 0) {
            s_warp[0][warp] = __popc(b0);
            s_warp[1][warp] = __popc(b1);
        }
        __syncthreads();
        if (warp < 2) {
            const int count = s_warp[warp][lane];
            int inclusive = count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, inclusive, o);
                if (lane >= o) {
                    inclusive += up;
                }
            }
            s_warp[warp][lane] = inclusive - count;
            if (lane == 31) {
                s_total[warp] = inclusive;
            }
        }
        __syncthreads();
        if (f0) {
            list0[s_base[0] + s_warp[0][warp] + __popc(b0 & below)] = p;
        }
        if (f1) {
            list1[s_base[1] + s_warp[1][warp] + __popc(b1 & below)] = p;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            s_base[threadIdx.x] += s_total[threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x < 2) {
        n_list[threadIdx.x] = s_base[threadIdx.x];
    }
}

__global__ void __launch_bounds__(kRdfTile)
    atomRdfKernel(SlotView V, const double4* __restrict__ list0, const double4* __restrict__ list1,
                  const int* __restrict__ n_list, bool identical, bool fold_absolute, double dxinv, int sx, int sy, int sz,
                  double thickness,
                  int n_bins, int shard, int n_shards, unsigned long long* __restrict__ hist, int* __restrict__ out_of_range)
{
    extern __shared__ unsigned int s_hist[];
    __shared__ double s_x[kRdfTile], s_y[kRdfTile], s_z[kRdfTile];
    const int n_i = n_list[0];
    const int n_j = identical ? n_i : n_list[1];
    const double4* __restrict__ list_j = identical ? list0 : list1;
    const int ti = blockIdx.x * n_shards + shard, tj = blockIdx.y; // tile rows dealt round robin to the shards
    if (ti * kRdfTile >= n_i || tj * kRdfTile >= n_j || (identical && tj < ti)) {
        return; // beyond the lists; i < j: the upper triangle of tiles
    }
    for (int b = threadIdx.x; b < n_bins; b += kRdfTile) {
        s_hist[b] = 0u;
    }
    const int pj = tj * kRdfTile + threadIdx.x;
    if (pj < n_j) {
        const double4 p = list_j[pj];
        s_x[threadIdx.x] = p.x;
        s_y[threadIdx.x] = p.y;
        s_z[threadIdx.x] = p.z;
    }
    const int j_end = min(kRdfTile, n_j - tj * kRdfTile);
    const int pi = ti * kRdfTile + threadIdx.x;
    const bool oki = pi < n_i;
    double ax = 0.0, ay = 0.0, az = 0.0;
    if (oki) {
        const double4 p = list0[pi];
        ax = p.x;
        ay = p.y;
        az = p.z;
    }
    __syncthreads();
    const bool slice = sx + sy + sz > 0;
    if (oki) {
        const int j_begin = (identical && ti == tj) ? static_cast<int>(threadIdx.x) + 1 : 0;
        for (int j = j_begin; j < j_end; ++j) {
            double dx = __dsub_rn(ax, s_x[j]);
            double dy = __dsub_rn(ay, s_y[j]);
            double dz = __dsub_rn(az, s_z[j]);
            if (fold_absolute) { // Geometry::sqdist (src/geometry.h:460-470): |d| − L·[|d| > L/2], as MoleculeRDF uses it
                dx = fabs(dx);
                dy = fabs(dy);
                dz = fabs(dz);
                dx = dx > V.half[0] ? __dsub_rn(dx, V.len_or_zero[0]) : dx;
                dy = dy > V.half[1] ? __dsub_rn(dy, V.len_or_zero[1]) : dy;
                dz = dz > V.half[2] ? __dsub_rn(dz, V.len_or_zero[2]) : dz;
            }
            else if (V.len_or_zero[0] > 0.0) {
                dx = dx > V.half[0] ? __dsub_rn(dx, V.len_or_zero[0]) : (dx < -V.half[0] ? __dadd_rn(dx, V.len_or_zero[0]) : dx);
            }
            if (V.len_or_zero[1] > 0.0) {
                dy = dy > V.half[1] ? __dsub_rn(dy, V.len_or_zero[1]) : (dy < -V.half[1] ? __dadd_rn(dy, V.len_or_zero[1]) : dy);
            }
            if (V.len_or_zero[2] > 0.0) {
                dz = dz > V.half[2] ? __dsub_rn(dz, V.len_or_zero[2]) : (dz < -V.half[2] ? __dadd_rn(dz, V.len_or_zero[2]) : dz);
            }
            double r;
            if (slice) {
                // in-plane part |d ∘ (1 − s)| < thickness, then the part along the slice direction is binned
                const double px = sx ? 0.0 : dx, py = sy ? 0.0 : dy, pz = sz ? 0.0 : dz;
                const double in_plane = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz)));
                if (!(in_plane < thickness)) {
                    continue;
                }
                const double qx = sx ? dx : 0.0, qy = sy ? dy : 0.0, qz = sz ? dz : 0.0;
                r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz)));
            }
            else {
                r = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            }
            const int bin = static_cast<int>(floor(__dmul_rn(r, dxinv)));
            if (bin >= 0 && bin < n_bins) {
                atomicAdd(&s_hist[bin], 1u);
            }
            else {
                *out_of_range = 1;
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += kRdfTile) {
        const unsigned int count = s_hist[b];
        if (count != 0u) {
            atomicAdd(hist + b, static_cast<unsigned long long>(count));
        }
    }
}

} // namespace fbdev

// Full rebuild of Q(k) as a complex matrix product on the FP64 tensor path.
//
// PolicyIonIon::updateComplex (src/energy.cpp:191-206; the PBCEigen quirk :208-217) is
//     Q(nx, ny, nz) = Σ_j w_j · X_j(nx) · Y_j(ny) · Z_j(nz),     X_j(n) = e^{i 2π n x_j / Lx}, …
// i.e. C = A · B with A[(nx, ny), j] = w_j X_j(nx) Y_j(ny) and B[j, nz] = Z_j(nz), the particles as the inner dimension.
// ewaldFullCellKernel (fb_stream.cuh) evaluates the same sum with one block per 4×4×4 cell on the DFMA pipe from shared-memory
// operands and redoes the six sincos of every particle in every cell; here a block owns a TILE of 4 nx × 8 ny rows and up to
// 64 nz columns (8 column groups of 8, only the groups the cutoff sphere touches) for a range of particles:
//
//   phase A  per chunk of 64 particles: the per-axis tables of the tile (4 + 8 + 8·groups entries per particle) → shared memory,
//            every entry an integer power of the particle's three unit phases e^{i 2π x / L} (ewaldStepPhaseKernel: the only
//            sincos of the rebuild), by repeated squaring and running products; the x entries carry the weight (charge;
//            PBCEigen: the imaginary part is summed WITHOUT the charge, which needs a second copy of the x entries)
//   phase B  warp ↔ (nx of the tile, half of the column groups): a k-step is 4 particles; the lane forms its A element
//            X·Y (one complex product) in registers, the B elements are one LDS.128 per column group, and the complex
//            product of the fragments is three mma.sync.m8n8k4.f64 per group (Gauss's three-multiplication form, gemmStep)
//
// The particle ranges (blockIdx.y) leave their shares in `partials`, and ewaldFullGatherKernel adds them in range order into
// Q(k) — every sum has a fixed order.
#pragma once

#include "fb_kspace.cuh"

namespace fbdev {

constexpr int kGemmThreads = 256;
constexpr int kGemmChunk = 64;     //!< particles per chunk (16 k-steps)
constexpr int kGemmRows = 32;      //!< rows of a tile: 4 nx × 8 ny
constexpr int kGemmCols = 64;      //!< columns of a tile: up to 8 groups of 8 nz
constexpr int kGemmZStride = 66;   //!< z entries per particle, padded: the 4 particles of a k-step in different banks
constexpr int kGemmYStride = 68;   //!< particles per y row, padded: two y rows of a quarter warp in different banks
constexpr int kGemmShare = 2 * kGemmRows * kGemmCols; //!< doubles per (tile, particle range): re[32][64], im[32][64]

struct FullGemmSmem
{
    double2 x[4][kGemmChunk];       //!< w_re · X(bx + i)
    double2 xi[4][kGemmChunk];      //!< w_im · X(bx + i) (used by the PBCEigen quirk only)
    double2 y[8][kGemmYStride];     //!< Y(by + i)
    double2 z[kGemmChunk][kGemmZStride]; //!< Z(z0 − ncc + i)
};

/**
 * One k-step (4 particles) of a warp over NG column groups. A complex product costs THREE real ones (Gauss):
 *   P1 = (A_re + A_im)·B_re,  P2 = A_re·(B_im − B_re),  P3 = A_im·(B_re + B_im)   →   C_re = P1 − P3,  C_im = P1 + P2
 * so a group is 3 mma.sync.m8n8k4.f64 and two additions on the B element instead of 4 mma; acc[g] = {P1, P2, P3}. With the
 * PBCEigen quirk (MODE 1) the real and the imaginary part carry different weights, P1 cannot be shared: acc[g] = {C_re, C_im, –},
 * four products. IPBC (MODE 2; PolicyIonIonIPBC::updateComplex, src/energy.cpp:414-430) is the real product of the cosines — the
 * real parts of the same tables —, one mma per group.
 */
template <int NG, int MODE>
__device__ __forceinline__ void gemmStep(const FullGemmSmem& sm, int ix, int row, int g0, int j, double (&acc)[4][3][2])
{
    const double2 yv = sm.y[row][j];
    const double2* zrow = &sm.z[j][8 * g0 + row];
    if (MODE == 2) { // IPBC: q cos(kx x) cos(ky y) cos(kz z), a real product
        const double ar = sm.x[ix][j].x * yv.x;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            dmma884(acc[g][0][0], acc[g][0][1], ar, zrow[8 * g].x);
        }
        return;
    }
    const double2 a = cmul(sm.x[ix][j], yv);
    if (MODE == 1) {
        const double2 b = cmul(sm.xi[ix][j], yv);
        const double na = -a.y;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const double2 zv = zrow[8 * g];
            dmma884(acc[g][0][0], acc[g][0][1], a.x, zv.x);
            dmma884(acc[g][1][0], acc[g][1][1], b.x, zv.y);
            dmma884(acc[g][0][0], acc[g][0][1], na, zv.y);
            dmma884(acc[g][1][0], acc[g][1][1], b.y, zv.x);
        }
    }
    else {
        const double sa = a.x + a.y;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const double2 zv = zrow[8 * g];
            dmma884(acc[g][0][0], acc[g][0][1], sa, zv.x);
            dmma884(acc[g][1][0], acc[g][1][1], a.x, zv.y - zv.x);
            dmma884(acc[g][2][0], acc[g][2][1], a.y, zv.x + zv.y);
        }
    }
}

template <int NG, int MODE>
__device__ __forceinline__ void gemmChunk(const FullGemmSmem& sm, int ix, int g0, int lane, double (&acc)[4][3][2])
{
    const int row = lane >> 2;
#pragma unroll 2
    for (int step = 0; step < kGemmChunk / 4; ++step) {
        gemmStep<NG, MODE>(sm, ix, row, g0, 4 * step + (lane & 3), acc);
    }
}

/** b^n of a unit complex number by repeated squaring (n is the same for the whole block) */
__device__ __forceinline__ double2 cpowi(double2 b, int n)
{
    const bool negative = n < 0;
    n = negative ? -n : n;
    double2 r = make_double2(1.0, 0.0);
    while (n != 0) {
        if (n & 1) {
            r = cmul(r, b);
        }
        b = cmul(b, b);
        n >>= 1;
    }
    return negative ? make_double2(r.x, -r.y) : r;
}

/**
 * The three sincos of a particle, once per rebuild (every block of the product would otherwise redo them for its tile):
 * steps[3 j] = {w_re, w_im} (charge, or 0 for an inactive slot; PBCEigen: w_im = 1), steps[3 j + 1] = e^{i 2π x / Lx},
 * steps[3 j + 2] = e^{i 2π y / Ly}; zsteps[j] = e^{i 2π z / Lz}. Every table entry of the product is an integer power of these.
 */
__global__ void __launch_bounds__(256)
    ewaldStepPhaseKernel(SlotView V, PhaseGeometry geo, int quirk, double2* __restrict__ steps, double2* __restrict__ zsteps)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V.n_slots) {
        return;
    }
    const double two_pi = 2.0 * 3.141592653589793238462643383279502884;
    double4 p = make_double4(0.0, 0.0, 0.0, 0.0);
    const bool active = V.gid[j] >= 0;
    if (active) {
        p = V.posq[j];
    }
    double s, c;
    steps[3 * static_cast<size_t>(j)] = make_double2(active ? p.w : 0.0, active ? (quirk ? 1.0 : p.w) : 0.0);
    sincos(two_pi / geo.len[0] * p.x, &s, &c);
    steps[3 * static_cast<size_t>(j) + 1] = make_double2(c, s);
    sincos(two_pi / geo.len[1] * p.y, &s, &c);
    steps[3 * static_cast<size_t>(j) + 2] = make_double2(c, s);
    sincos(two_pi / geo.len[2] * p.z, &s, &c);
    zsteps[j] = make_double2(c, s);
}

/**
 * @param tiles    [n_tiles] {nx of the first row, y table index of the first row (ny + ncc), z table index of the first
 *                 column (nz + ncc; any integer), number of column groups} in storage order of the k-vectors
 * @param row_groups [n_tiles] per nx of the tile (8 bits each): first column group that holds one of its k-vectors | (last + 1) << 4
 * @param order    block → tile, heaviest tiles first (all tiles), or nullptr: block b takes tile tile_begin + b (a slab)
 * @param partials [tile − tile_begin][gridDim.y][2][32][64]
 */
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 2)
    ewaldFullGemmKernel(int n_slots, const double2* __restrict__ steps, const double2* __restrict__ zsteps,
                        const int4* __restrict__ tiles, const int* __restrict__ row_groups, const int* __restrict__ order,
                        int tile_begin, PhaseGeometry geo, int range_size, double* __restrict__ partials)
{
    extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
    FullGemmSmem& sm = *reinterpret_cast<FullGemmSmem*>(gemm_smem_raw);
    const int tile_index = order != nullptr ? __ldg(order + blockIdx.x) : tile_begin + static_cast<int>(blockIdx.x);
    const int4 tile = __ldg(tiles + tile_index);
    const int ng = tile.w;
    const int j_begin = static_cast<int>(blockIdx.y) * range_size;
    const int j_end = min(n_slots, j_begin + range_size);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // which nx of the tile a warp takes rotates with the block: the rows of a tile need different numbers of column groups
    // (the sphere), a warp's sub-partition is warp % 4, and the blocks that share an SM should not all have their short rows
    // on the same one
    const int ix = (warp + static_cast<int>(blockIdx.x + blockIdx.y)) & 3;
    const int half = warp >> 2;
    const int row_range = __ldg(row_groups + tile_index) >> (8 * ix); // groups [lo, hi) that hold a k-vector of this nx
    const int row_lo = row_range & 15;
    const int row_n = ((row_range >> 4) & 15) - row_lo;

    // the two warps of an nx share the column groups (the first ⌈ng / 2⌉ and the rest: both on the same SM sub-partition)
    const int g0 = row_lo + (half == 0 ? 0 : (row_n + 1) / 2);
    const int my_ng = half == 0 ? (row_n + 1) / 2 : row_n / 2;
    double acc[4][3][2];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            acc[g][t][0] = acc[g][t][1] = 0.0;
        }
    }

    const int pj = threadIdx.x & (kGemmChunk - 1);
    const int part = threadIdx.x / kGemmChunk; // warp-uniform: 0 = x and y entries, 1..3 = column groups part − 1, part + 2, …
    for (int c0 = j_begin; c0 < j_end; c0 += kGemmChunk) {
        __syncthreads(); // the previous chunk is consumed
        {
            const int j = c0 + pj;
            const bool in_range = j < j_end;
            if (part == 0) {
                double wre = 0.0, wim = 0.0;
                double2 sx = make_double2(1.0, 0.0), sy = sx;
                if (in_range) {
                    const double2 w = steps[3 * static_cast<size_t>(j)]; // (see ewaldStepPhaseKernel)
                    sx = steps[3 * static_cast<size_t>(j) + 1];
                    sy = steps[3 * static_cast<size_t>(j) + 2];
                    wre = w.x;
                    wim = w.y;
                }
                double2 e = cpowi(sx, tile.x);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sm.x[i][pj] = make_double2(wre * e.x, wre * e.y);
                    if (MODE == 1) {
                        sm.xi[i][pj] = make_double2(wim * e.x, wim * e.y);
                    }
                    e = cmul(e, sx);
                }
                e = cpowi(sy, tile.y - geo.ncc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sm.y[i][pj] = e;
                    e = cmul(e, sy);
                }
            }
            else if (part - 1 < ng) {
                const double2 sz = in_range ? zsteps[j] : make_double2(1.0, 0.0);
                double2 e = cpowi(sz, tile.z + 8 * (part - 1) - geo.ncc);
                double2 s24 = cmul(sz, sz);
                s24 = cmul(s24, s24);
                s24 = cmul(s24, s24);            // 8 steps
                s24 = cmul(cmul(s24, s24), s24); // 24 steps: this thread's next group
                for (int g = part - 1; g < ng; g += 3) {
                    double2 f = e;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        sm.z[pj][8 * g + i] = f;
                        f = cmul(f, sz);
                    }
                    e = cmul(e, s24);
                }
            }
        }
        __syncthreads();
        switch (my_ng) { // (a predicated mma.sync costs a WARPSYNC each: the group count is a template argument)
        case 1: gemmChunk<1, MODE>(sm, ix, g0, lane, acc); break;
        case 2: gemmChunk<2, MODE>(sm, ix, g0, lane, acc); break;
        case 3: gemmChunk<3, MODE>(sm, ix, g0, lane, acc); break;
        case 4: gemmChunk<4, MODE>(sm, ix, g0, lane, acc); break;
        default: break;
        }
    }
    // the block's share: row = 8·ix + lane/4, columns 8·g + 2·(lane%4) + {0, 1}
    double* share = partials + (static_cast<size_t>(tile_index - tile_begin) * gridDim.y + blockIdx.y) * kGemmShare;
    const int row = 8 * ix + (lane >> 2);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (g < my_ng) {
            const int col = 8 * (g0 + g) + 2 * (lane & 3);
            double2 re, im;
            if (MODE == 2) {
                re = make_double2(acc[g][0][0], acc[g][0][1]);
                im = make_double2(0.0, 0.0);
            }
            else if (MODE == 1) {
                re = make_double2(acc[g][0][0], acc[g][0][1]);
                im = make_double2(acc[g][1][0], acc[g][1][1]);
            }
            else {
                re = make_double2(acc[g][0][0] - acc[g][2][0], acc[g][0][1] - acc[g][2][1]);
                im = make_double2(acc[g][0][0] + acc[g][1][0], acc[g][0][1] + acc[g][1][1]);
            }
            *reinterpret_cast<double2*>(share + row * kGemmCols + col) = re;
            *reinterpret_cast<double2*>(share + kGemmRows * kGemmCols + row * kGemmCols + col) = im;
        }
    }
}

/** Σ over the particle ranges, in range order, of the tile shares of k-vector k */
__device__ __forceinline__ double2 gatherShares(int at, int tile_begin, int n_ranges, const double* __restrict__ partials)
{
    const int tile = at / (kGemmRows * kGemmCols) - tile_begin;
    const int inside = at % (kGemmRows * kGemmCols);
    const double* share = partials + static_cast<size_t>(tile) * n_ranges * kGemmShare + inside;
    double2 Q = make_double2(0.0, 0.0);
    for (int r = 0; r < n_ranges; ++r) {
        Q.x += share[static_cast<size_t>(r) * kGemmShare];
        Q.y += share[static_cast<size_t>(r) * kGemmShare + kGemmRows * kGemmCols];
    }
    return Q;
}

/**
 * Q(k) of all k-vectors from the shares.
 * @param index [K] (tile · 2048 + row · 64 + column) of every k-vector
 */
__global__ void __launch_bounds__(256)
    ewaldFullGatherKernel(EwaldView E, const int* __restrict__ index, int n_ranges, const double* __restrict__ partials)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < E.K) {
        E.Q[k] = gatherShares(__ldg(index + k), 0, n_ranges, partials);
    }
}

/**
 * Sharded reciprocal energy: Σ_k A_k |Q_k|² over the k-vectors [k_begin, k_end) of a slab of tiles, one partial sum per
 * block of 256 k-vectors (fixed shuffle tree, warps in order); Q(k) is not stored.
 */
__global__ void __launch_bounds__(256)
    ewaldFullGatherEnergyKernel(EwaldView E, const int* __restrict__ index, int tile_begin, int n_ranges,
                                const double* __restrict__ partials, int k_begin, int k_end, double* __restrict__ e_partials)
{
    __shared__ double s_e[8];
    const int k = k_begin + blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (k < k_end) {
        const double2 Q = gatherShares(__ldg(index + k), tile_begin, n_ranges, partials);
        e = E.kA[k].w * (Q.x * Q.x + Q.y * Q.y);
    }
    e = warpSum(e);
    if ((threadIdx.x & 31) == 0) {
        s_e[threadIdx.x >> 5] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            s += s_e[w];
        }
        e_partials[blockIdx.x] = s;
    }
}

} // namespace fbdev

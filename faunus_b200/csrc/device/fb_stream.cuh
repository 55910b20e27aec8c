// Streaming pair kernels for the two compute-bound passes (sm_100a, FP64 CUDA cores):
//   widomStreamKernel   B ghost insertions × all particles      (WidomInsertion::_sample, src/analysis.cpp:1255-1262)
//   fullStreamKernel    Σ_{i<j} over an all-atomic system       (GroupPairingPolicy::all, src/energy.h:1290-1326)
// Same inner loop as batchPairKernel (fb_batch.cuh): lane ↔ particle j held in registers (two per
// thread), the block walks "variants" (ghost atoms / particles i) staged in shared memory, two per
// iteration; the minimum-image fold of an axis is skipped for variants at least the cutoff away from
// both cell faces; the few pairs inside the cutoff are queued per warp in ballot order (deterministic)
// and evaluated afterwards with all lanes busy. With a potential that has no cutoff (DENSE) every pair
// is evaluated in place.
#pragma once
#include "fb_batch.cuh"

namespace fbdev {

constexpr int kStreamThreads = 128;
constexpr int kStreamChunk = 2 * kStreamThreads; //!< particles j per block iteration (two per thread)
constexpr int kStreamVariants = 64;              //!< variants per block
constexpr int kStreamQueue = 384;                //!< queued in-range candidates per warp

struct StreamFold
{
    double hx, hy, hz, lx, ly, lz;
};

__device__ __forceinline__ int foldFlags(const SlotView& V, const double4& a, double cut2)
{
    const double rc = sqrt(cut2); // +inf when some term has no cutoff
    int fold = 0;
    fold |= (V.len_or_zero[0] > 0.0 && !(fabs(a.x) + rc < V.half[0])) ? 1 : 0;
    fold |= (V.len_or_zero[1] > 0.0 && !(fabs(a.y) + rc < V.half[1])) ? 2 : 0;
    fold |= (V.len_or_zero[2] > 0.0 && !(fabs(a.z) + rc < V.half[2])) ? 4 : 0;
    return fold;
}

/** r² of variant `a` with the thread's two particles; fold only where the flags ask for it */
__device__ __forceinline__ void streamR2(const double4& a, int fold, const double4 (&p)[2], const StreamFold& f,
                                         double (&r2)[2])
{
    double dx[2], dy[2], dz[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        dx[t] = a.x - p[t].x;
        dy[t] = a.y - p[t].y;
        dz[t] = a.z - p[t].z;
    }
    if (fold & 1) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const double ad = fabs(dx[t]);
            dx[t] = (ad > f.hx) ? ad - f.lx : dx[t];
        }
    }
    if (fold & 2) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const double ad = fabs(dy[t]);
            dy[t] = (ad > f.hy) ? ad - f.ly : dy[t];
        }
    }
    if (fold & 4) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const double ad = fabs(dz[t]);
            dz[t] = (ad > f.hz) ? ad - f.lz : dz[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        r2[t] = dx[t] * dx[t] + dy[t] * dy[t] + dz[t] * dz[t];
    }
}

// ------------------------------------------------------------------------------------------------
// Widom: variant v = b · n_ghost_atoms + a (ghost atom a of insertion b). Block = 64 consecutive
// variants, loops over ALL particles. out[v] = energy of that ghost atom with every active particle
// outside the ghost group. Atomic ghost groups only (no mass-centre cutoff applies to them).
// ------------------------------------------------------------------------------------------------
template <int KIND, bool DENSE>
__global__ void __launch_bounds__(kStreamThreads)
    widomStreamKernel(SlotView V, PotParams P, int ghost_group, int n_ghost_atoms, int n_variants,
                      const double4* __restrict__ ghost_posq, const int* __restrict__ ghost_id, double cut2,
                      double* __restrict__ out /*[n_variants]*/)
{
    constexpr int NW = kStreamThreads / 32;
    __shared__ double4 s_var[kStreamVariants];
    __shared__ int s_vid[kStreamVariants];
    __shared__ int s_vfold[kStreamVariants];
    __shared__ double s_acc[NW][kStreamVariants];
    __shared__ double s_qr[DENSE ? 1 : NW][DENSE ? 1 : kStreamQueue];
    __shared__ unsigned s_qe[DENSE ? 1 : NW][DENSE ? 1 : kStreamQueue];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int v0 = blockIdx.x * kStreamVariants;
    const int nv = min(kStreamVariants, n_variants - v0);
    const StreamFold f{V.half[0], V.half[1], V.half[2], V.len_or_zero[0], V.len_or_zero[1], V.len_or_zero[2]};
    const double nan = __longlong_as_double(0x7ff8000000000000LL);

    for (int v = threadIdx.x; v < kStreamVariants; v += kStreamThreads) {
        double4 a = make_double4(nan, 0, 0, 0);
        int id = 0;
        if (v < nv) {
            a = ghost_posq[v0 + v];
            id = ghost_id[(v0 + v) % n_ghost_atoms];
        }
        s_var[v] = a;
        s_vid[v] = id;
        s_vfold[v] = v < nv ? foldFlags(V, a, cut2) : 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s_acc[w][v] = 0.0;
        }
    }
    __syncthreads();

    int queued = 0; // warp-uniform
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0xffu;
            const int j = ent >> 8;
            const double4 a = s_var[v];
            s_qr[warp][e] = pairEnergy<KIND>(P, s_vid[v], V.atom_id[j], a.w, V.posq[j].w, s_qr[warp][e]);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < kStreamVariants / 32; ++h) {
            const int myv = lane + 32 * h;
            double acc = s_acc[warp][myv]; // one running sum per variant: independent of when the queue is flushed
            for (int e = 0; e < queued; ++e) {
                if (static_cast<int>(s_qe[warp][e] & 0xffu) == myv) {
                    acc += s_qr[warp][e];
                }
            }
            s_acc[warp][myv] = acc;
        }
        __syncwarp();
        queued = 0;
    };

    const unsigned var_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_var[0]));
    const unsigned fold_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_vfold[0]));
    for (int base = 0; base < V.n_slots; base += kStreamChunk) {
        double4 p[2];
        int pid[2], pj[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = base + t * kStreamThreads + threadIdx.x;
            pj[t] = j;
            pid[t] = 0;
            p[t] = make_double4(nan, 0, 0, 0);
            if (j < V.n_slots) {
                const int g = V.gid[j];
                if (g >= 0 && g != ghost_group) {
                    p[t] = V.posq[j];
                    pid[t] = V.atom_id[j];
                }
            }
        }
        for (int vv = 0; vv < nv; vv += 2) {
            double4 a[2];
            int fold[2];
            double r2[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                loadVariant(var_addr, fold_addr, min(vv + u, nv - 1), a[u], fold[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                streamR2(a[u], fold[u], p, f, r2[u]);
            }
            if (!__any_sync(0xffffffffu, anyBelow(r2[0][0], r2[0][1], r2[1][0], r2[1][1], cut2))) {
                continue;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = vv + u;
                if (v >= nv) {
                    continue;
                }
                if (DENSE) {
                    double e = 0.0;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (r2[u][t] < cut2) {
                            e += pairEnergy<KIND>(P, s_vid[v], pid[t], s_var[v].w, p[t].w, r2[u][t]);
                        }
                    }
                    e = warpSum(e);
                    if (lane == 0) {
                        s_acc[warp][v] += e;
                    }
                }
                else {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const bool in = r2[u][t] < cut2;
                        const unsigned mask = __ballot_sync(0xffffffffu, in);
                        if (in) {
                            const int at = queued + __popc(mask & ((1u << lane) - 1u));
                            s_qe[warp][at] = static_cast<unsigned>(v) | (static_cast<unsigned>(pj[t]) << 8);
                            s_qr[warp][at] = r2[u][t];
                        }
                        queued += __popc(mask);
                    }
                }
            }
            if (!DENSE) {
                __syncwarp();
                if (queued > kStreamQueue - 128) {
                    flush();
                }
            }
        }
    }
    if (!DENSE) {
        flush();
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += kStreamThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s += s_acc[w][v];
        }
        out[v0 + v] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Widom with a finite cutoff: FP32 screening of the distance test, FP64 evaluation of the candidates (the scheme
// of batchPairScreenKernel, fb_batch.cuh — 99.9 % of the ghost–particle pairs of S1 lie beyond the cutoff and
// contribute exactly zero). Grid (blocks of 64 variants) × (splits of kWidomSplit particles): a slice of a few
// thousand insertions — the share of one of 8 GPUs — still fills the machine. The split size is a constant, so the
// partial sums out[split][variant], and their sum in split order, are the same numbers however the insertions are
// sliced over launches or ranks.
// ------------------------------------------------------------------------------------------------
constexpr int kWidomSplit = 8192;
constexpr int kWidomScreenPerThread = 4;
constexpr int kWidomScreenChunk = kStreamThreads * kWidomScreenPerThread;

template <int KIND>
__global__ void __launch_bounds__(kStreamThreads)
    widomScreenKernel(SlotView V, PotParams P, int ghost_group, int n_ghost_atoms, int n_variants,
                      const double4* __restrict__ ghost_posq, const int* __restrict__ ghost_id, double cut2,
                      float cut2_screen, double* __restrict__ out /*[gridDim.y][n_variants]*/)
{
    constexpr int NW = kStreamThreads / 32;
    __shared__ double4 s_var[kStreamVariants];
    __shared__ float4 s_varf[kStreamVariants];
    __shared__ int s_vid[kStreamVariants];
    __shared__ double s_acc[NW][kStreamVariants];
    __shared__ double s_qr[NW][kStreamQueue];
    __shared__ unsigned s_qe[NW][kStreamQueue]; //!< variant of the block | particle << 6

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int v0 = blockIdx.x * kStreamVariants;
    const int nv = min(kStreamVariants, n_variants - v0);
    const float hx = static_cast<float>(V.half[0]), hy = static_cast<float>(V.half[1]), hz = static_cast<float>(V.half[2]);
    const float lx = static_cast<float>(V.len_or_zero[0]), ly = static_cast<float>(V.len_or_zero[1]),
                lz = static_cast<float>(V.len_or_zero[2]);
    const float nanf_ = __int_as_float(0x7fc00000);

    for (int v = threadIdx.x; v < kStreamVariants; v += kStreamThreads) {
        double4 a = make_double4(0, 0, 0, 0);
        int id = 0;
        float4 af = make_float4(nanf_, 0.0f, 0.0f, 0.0f);
        if (v < nv) {
            a = ghost_posq[v0 + v];
            id = ghost_id[(v0 + v) % n_ghost_atoms];
            const float rc = sqrtf(cut2_screen) * 1.0001f;
            int fold = 0;
            fold |= (lx > 0.0f && !(fabsf(static_cast<float>(a.x)) + rc < hx)) ? 1 : 0;
            fold |= (ly > 0.0f && !(fabsf(static_cast<float>(a.y)) + rc < hy)) ? 2 : 0;
            fold |= (lz > 0.0f && !(fabsf(static_cast<float>(a.z)) + rc < hz)) ? 4 : 0;
            af = make_float4(static_cast<float>(a.x), static_cast<float>(a.y), static_cast<float>(a.z), __int_as_float(fold));
        }
        s_var[v] = a;
        s_varf[v] = af;
        s_vid[v] = id;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s_acc[w][v] = 0.0;
        }
    }
    __syncthreads();

    int queued = 0; // warp-uniform
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0x3fu;
            const int j = ent >> 6;
            const double4 a = s_var[v];
            const double4 b = V.posq[j]; // the candidates are few: double-precision positions from L2
            const double r2 = minImageR2(V, a.x, a.y, a.z, b.x, b.y, b.z);
            s_qr[warp][e] = r2 < cut2 ? pairEnergy<KIND>(P, s_vid[v], V.atom_id[j], a.w, b.w, r2) : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < kStreamVariants / 32; ++h) {
            const int myv = lane + 32 * h;
            double acc = s_acc[warp][myv]; // one running sum per variant: independent of when the queue is flushed
            for (int e = 0; e < queued; ++e) {
                if (static_cast<int>(s_qe[warp][e] & 0x3fu) == myv) {
                    acc += s_qr[warp][e];
                }
            }
            s_acc[warp][myv] = acc;
        }
        __syncwarp();
        queued = 0;
    };

    const int j_begin = blockIdx.y * kWidomSplit;
    const int j_end = min(V.n_slots, j_begin + kWidomSplit);
    for (int base = j_begin; base < j_end; base += kWidomScreenChunk) {
        float px[kWidomScreenPerThread], py[kWidomScreenPerThread], pz[kWidomScreenPerThread];
#pragma unroll
        for (int t = 0; t < kWidomScreenPerThread; ++t) {
            const int j = base + t * kStreamThreads + threadIdx.x;
            px[t] = nanf_; // inactive particles and the ghost group itself: never in range
            py[t] = pz[t] = 0.0f;
            if (j < j_end) {
                const int g = V.gid[j];
                if (g >= 0 && g != ghost_group) {
                    const double4 p = V.posq[j];
                    px[t] = static_cast<float>(p.x);
                    py[t] = static_cast<float>(p.y);
                    pz[t] = static_cast<float>(p.z);
                }
            }
        }
        for (int vv = 0; vv < nv; vv += 2) {
            float r2[2][kWidomScreenPerThread];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float4 a = s_varf[min(vv + u, nv - 1)];
                const int fold = __float_as_int(a.w);
                float dx[kWidomScreenPerThread], dy[kWidomScreenPerThread], dz[kWidomScreenPerThread];
#pragma unroll
                for (int t = 0; t < kWidomScreenPerThread; ++t) {
                    dx[t] = a.x - px[t];
                    dy[t] = a.y - py[t];
                    dz[t] = a.z - pz[t];
                }
                if (fold & 1) {
#pragma unroll
                    for (int t = 0; t < kWidomScreenPerThread; ++t) {
                        const float ad = fabsf(dx[t]);
                        dx[t] = (ad > hx) ? ad - lx : dx[t];
                    }
                }
                if (fold & 2) {
#pragma unroll
                    for (int t = 0; t < kWidomScreenPerThread; ++t) {
                        const float ad = fabsf(dy[t]);
                        dy[t] = (ad > hy) ? ad - ly : dy[t];
                    }
                }
                if (fold & 4) {
#pragma unroll
                    for (int t = 0; t < kWidomScreenPerThread; ++t) {
                        const float ad = fabsf(dz[t]);
                        dz[t] = (ad > hz) ? ad - lz : dz[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < kWidomScreenPerThread; ++t) {
                    r2[u][t] = fmaf(dz[t], dz[t], fmaf(dy[t], dy[t], dx[t] * dx[t]));
                }
            }
            bool any_in = false;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int t = 0; t < kWidomScreenPerThread; ++t) {
                    any_in = any_in || (r2[u][t] < cut2_screen);
                }
            }
            if (!__any_sync(0xffffffffu, any_in)) {
                continue;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int v = vv + u;
                if (v >= nv) {
                    continue;
                }
#pragma unroll
                for (int t = 0; t < kWidomScreenPerThread; ++t) {
                    const bool in = r2[u][t] < cut2_screen;
                    const unsigned mask = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int at = queued + __popc(mask & ((1u << lane) - 1u));
                        s_qe[warp][at] = static_cast<unsigned>(v) |
                                         (static_cast<unsigned>(base + t * kStreamThreads + threadIdx.x) << 6);
                    }
                    queued += __popc(mask);
                }
            }
            __syncwarp();
            if (queued > kStreamQueue - 2 * kWidomScreenPerThread * 32) {
                flush();
            }
        }
    }
    flush();
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += kStreamThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s += s_acc[w][v];
        }
        out[static_cast<size_t>(blockIdx.y) * n_variants + v0 + v] = s;
    }
}

/** du[b] = Σ_a out[b·n_g + a] + ghost-internal pairs (atomic ghosts with the internal flag) */
template <int KIND>
__global__ void widomStreamFinishKernel(SlotView V, PotParams P, int n_ghost_atoms, int n_insertions,
                                        const double4* __restrict__ ghost_posq, const int* __restrict__ ghost_id,
                                        int internal, const double* __restrict__ per_variant /*[n_split][variants]*/,
                                        int n_split, double* __restrict__ du)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_insertions) {
        return;
    }
    const size_t n_variants = static_cast<size_t>(n_insertions) * n_ghost_atoms;
    double e = 0.0;
    for (int a = 0; a < n_ghost_atoms; ++a) {
        double atom = 0.0;
        for (int split = 0; split < n_split; ++split) { // particle ranges in order
            atom += per_variant[split * n_variants + static_cast<size_t>(b) * n_ghost_atoms + a];
        }
        e += atom;
    }
    if (internal) {
        const double4* g = ghost_posq + static_cast<size_t>(b) * n_ghost_atoms;
        for (int i = 0; i < n_ghost_atoms - 1; ++i) {
            for (int j = i + 1; j < n_ghost_atoms; ++j) {
                const double r2 = minImageR2(V, g[i].x, g[i].y, g[i].z, g[j].x, g[j].y, g[j].z);
                e += pairEnergy<KIND>(P, ghost_id[i], ghost_id[j], g[i].w, g[j].w, r2);
            }
        }
    }
    du[b] = e;
}

// ------------------------------------------------------------------------------------------------
// Full energy of an all-atomic system: block (bi, bj ≥ bi·…) — variants are the particles i of tile bi
// (64 per block row), the block walks the particles j > i. grid.x = i-tiles of 64; every block loops over
// the j-chunks from its own tile on; shard/n_shards deal the i-tiles round robin (multi-GPU).
// partials[blockIdx.x] = Σ over the block's pairs.
// ------------------------------------------------------------------------------------------------
template <int KIND, bool DENSE>
__global__ void __launch_bounds__(kStreamThreads)
    fullStreamKernel(SlotView V, PotParams P, double cut2, int j_split, int shard, int n_shards,
                     double* __restrict__ partials /*[gridDim.x · gridDim.y]*/)
{
    constexpr int NW = kStreamThreads / 32;
    __shared__ double4 s_var[kStreamVariants];
    __shared__ int s_vid[kStreamVariants];
    __shared__ int s_vfold[kStreamVariants];
    __shared__ double s_qr[DENSE ? 1 : NW][DENSE ? 1 : kStreamQueue];
    __shared__ unsigned s_qe[DENSE ? 1 : NW][DENSE ? 1 : kStreamQueue];
    __shared__ double s_sum[NW];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * kStreamVariants;
    const StreamFold f{V.half[0], V.half[1], V.half[2], V.len_or_zero[0], V.len_or_zero[1], V.len_or_zero[2]};
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const size_t out_index = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    if (static_cast<int>(blockIdx.x) % n_shards != shard) {
        if (threadIdx.x == 0) {
            partials[out_index] = 0.0;
        }
        return;
    }
    for (int v = threadIdx.x; v < kStreamVariants; v += kStreamThreads) {
        const int i = i0 + v;
        double4 a = make_double4(nan, 0, 0, 0);
        int id = 0;
        if (i < V.n_slots && V.gid[i] >= 0) {
            a = V.posq[i];
            id = V.atom_id[i];
        }
        s_var[v] = a;
        s_vid[v] = id;
        s_vfold[v] = (a.x == a.x) ? foldFlags(V, a, cut2) : 0;
    }
    __syncthreads();

    double esum = 0.0;
    int queued = 0;
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0xffu;
            const int j = ent >> 8;
            const double4 a = s_var[v];
            esum += pairEnergy<KIND>(P, s_vid[v], V.atom_id[j], a.w, V.posq[j].w, s_qr[warp][e]);
        }
        __syncwarp();
        queued = 0;
    };

    const unsigned var_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_var[0]));
    const unsigned fold_addr = static_cast<unsigned>(__cvta_generic_to_shared(&s_vfold[0]));
    // j-chunks from the one containing i0 on; blockIdx.y takes every gridDim.y-th chunk
    const int first_chunk = i0 / kStreamChunk;
    const int n_chunks = (V.n_slots + kStreamChunk - 1) / kStreamChunk;
    (void)j_split;
    for (int chunk = first_chunk + blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
        const int base = chunk * kStreamChunk;
        double4 p[2];
        int pid[2], pj[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = base + t * kStreamThreads + threadIdx.x;
            pj[t] = j;
            pid[t] = 0;
            p[t] = make_double4(nan, 0, 0, 0);
            if (j < V.n_slots && V.gid[j] >= 0) {
                p[t] = V.posq[j];
                pid[t] = V.atom_id[j];
            }
        }
        const bool diagonal = base < i0 + kStreamVariants; // some j of this chunk are ≤ some i of the tile
        for (int vv = 0; vv < kStreamVariants; vv += 2) {
            double4 a[2];
            int fold[2];
            double r2[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                loadVariant(var_addr, fold_addr, vv + u, a[u], fold[u]);
            }
            bool in[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                streamR2(a[u], fold[u], p, f, r2[u]);
            }
            if (!__any_sync(0xffffffffu, anyBelow(r2[0][0], r2[0][1], r2[1][0], r2[1][1], cut2))) {
                continue;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    in[u][t] = r2[u][t] < cut2 && (!diagonal || pj[t] > i0 + vv + u);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (DENSE) {
                        if (in[u][t]) {
                            esum += pairEnergy<KIND>(P, s_vid[vv + u], pid[t], s_var[vv + u].w, p[t].w, r2[u][t]);
                        }
                    }
                    else {
                        const unsigned mask = __ballot_sync(0xffffffffu, in[u][t]);
                        if (in[u][t]) {
                            const int at = queued + __popc(mask & ((1u << lane) - 1u));
                            s_qe[warp][at] = static_cast<unsigned>(vv + u) | (static_cast<unsigned>(pj[t]) << 8);
                            s_qr[warp][at] = r2[u][t];
                        }
                        queued += __popc(mask);
                    }
                }
            }
            if (!DENSE) {
                __syncwarp();
                if (queued > kStreamQueue - 128) {
                    flush();
                }
            }
        }
    }
    if (!DENSE) {
        flush();
    }
    esum = warpSum(esum);
    if (lane == 0) {
        s_sum[warp] = esum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s += s_sum[w];
        }
        partials[out_index] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Full energy of an all-atomic system with a finite cutoff: the rows and chunks of fullStreamKernel with the FP32 screening of
// widomScreenKernel — lane ↔ 4 particles j in registers as FP32, the distance test on the FMA pipes against the cutoff widened
// by the bound of the FP32 rounding error (screeningCutoff, fb_api.cu), the candidates (0.1 % of the pairs at S1) queued per
// warp in ballot order and evaluated in FP64 from the FP64 positions with the reference's minimum-image arithmetic and the
// exact test r² < cut². partials[blockIdx.y · gridDim.x + blockIdx.x] = Σ over the block's pairs (lane sums in queue order,
// fixed shuffle tree, warps in order).
// ------------------------------------------------------------------------------------------------
constexpr int kFullScreenPerThread = 4;
constexpr int kFullScreenChunk = kStreamThreads * kFullScreenPerThread;

template <int KIND>
__global__ void __launch_bounds__(kStreamThreads)
    fullScreenKernel(SlotView V, PotParams P, double cut2, float cut2_screen, int shard, int n_shards,
                     double* __restrict__ partials /*[gridDim.x · gridDim.y]*/)
{
    constexpr int NW = kStreamThreads / 32;
    constexpr int PT = kFullScreenPerThread;
    __shared__ double4 s_var[kStreamVariants];
    __shared__ float4 s_varf[kStreamVariants];
    __shared__ int s_vid[kStreamVariants];
    __shared__ unsigned s_qe[NW][kStreamQueue]; //!< variant of the block | particle << 6
    __shared__ double s_sum[NW];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int i0 = blockIdx.x * kStreamVariants;
    const size_t out_index = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    if (static_cast<int>(blockIdx.x) % n_shards != shard) {
        if (threadIdx.x == 0) {
            partials[out_index] = 0.0;
        }
        return;
    }
    const float hx = static_cast<float>(V.half[0]), hy = static_cast<float>(V.half[1]), hz = static_cast<float>(V.half[2]);
    const float lx = static_cast<float>(V.len_or_zero[0]), ly = static_cast<float>(V.len_or_zero[1]),
                lz = static_cast<float>(V.len_or_zero[2]);
    const float nanf_ = __int_as_float(0x7fc00000);
    for (int v = threadIdx.x; v < kStreamVariants; v += kStreamThreads) {
        const int i = i0 + v;
        double4 a = make_double4(0, 0, 0, 0);
        int id = 0;
        float4 af = make_float4(nanf_, 0.0f, 0.0f, 0.0f); // inactive: never in range
        if (i < V.n_slots && V.gid[i] >= 0) {
            a = V.posq[i];
            id = V.atom_id[i];
            const float rc = sqrtf(cut2_screen) * 1.0001f;
            int fold = 0;
            fold |= (lx > 0.0f && !(fabsf(static_cast<float>(a.x)) + rc < hx)) ? 1 : 0;
            fold |= (ly > 0.0f && !(fabsf(static_cast<float>(a.y)) + rc < hy)) ? 2 : 0;
            fold |= (lz > 0.0f && !(fabsf(static_cast<float>(a.z)) + rc < hz)) ? 4 : 0;
            af = make_float4(static_cast<float>(a.x), static_cast<float>(a.y), static_cast<float>(a.z), __int_as_float(fold));
        }
        s_var[v] = a;
        s_varf[v] = af;
        s_vid[v] = id;
    }
    __syncthreads();

    double esum = 0.0;
    int queued = 0; // warp-uniform
    auto flush = [&]() {
        for (int e = lane; e < queued; e += 32) {
            const unsigned ent = s_qe[warp][e];
            const int v = ent & 0x3fu;
            const int j = ent >> 6;
            const double4 a = s_var[v];
            const double4 b = V.posq[j]; // the candidates are few: double-precision positions from L2
            const double r2 = minImageR2(V, a.x, a.y, a.z, b.x, b.y, b.z);
            if (r2 < cut2) {
                esum += pairEnergy<KIND>(P, s_vid[v], V.atom_id[j], a.w, b.w, r2);
            }
        }
        __syncwarp();
        queued = 0;
    };

    // j-chunks from the one containing i0 on; blockIdx.y takes every gridDim.y-th chunk
    const int first_chunk = i0 / kFullScreenChunk;
    const int n_chunks = (V.n_slots + kFullScreenChunk - 1) / kFullScreenChunk;
    for (int chunk = first_chunk + blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
        const int base = chunk * kFullScreenChunk;
        float px[PT], py[PT], pz[PT];
        int pj[PT];
#pragma unroll
        for (int t = 0; t < PT; ++t) {
            const int j = base + t * kStreamThreads + threadIdx.x;
            pj[t] = j;
            px[t] = nanf_;
            py[t] = pz[t] = 0.0f;
            if (j < V.n_slots && V.gid[j] >= 0) {
                const double4 p = V.posq[j];
                px[t] = static_cast<float>(p.x);
                py[t] = static_cast<float>(p.y);
                pz[t] = static_cast<float>(p.z);
            }
        }
        const bool diagonal = base < i0 + kStreamVariants; // some j of this chunk are ≤ some i of the tile
        for (int vv = 0; vv < kStreamVariants; vv += 2) {
            float r2[2][PT];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float4 a = s_varf[vv + u];
                const int fold = __float_as_int(a.w);
                float dx[PT], dy[PT], dz[PT];
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    dx[t] = a.x - px[t];
                    dy[t] = a.y - py[t];
                    dz[t] = a.z - pz[t];
                }
                if (fold & 1) {
#pragma unroll
                    for (int t = 0; t < PT; ++t) {
                        const float ad = fabsf(dx[t]);
                        dx[t] = (ad > hx) ? ad - lx : dx[t];
                    }
                }
                if (fold & 2) {
#pragma unroll
                    for (int t = 0; t < PT; ++t) {
                        const float ad = fabsf(dy[t]);
                        dy[t] = (ad > hy) ? ad - ly : dy[t];
                    }
                }
                if (fold & 4) {
#pragma unroll
                    for (int t = 0; t < PT; ++t) {
                        const float ad = fabsf(dz[t]);
                        dz[t] = (ad > hz) ? ad - lz : dz[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    r2[u][t] = fmaf(dz[t], dz[t], fmaf(dy[t], dy[t], dx[t] * dx[t]));
                }
            }
            bool any_in = false;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    any_in = any_in || (r2[u][t] < cut2_screen);
                }
            }
            if (!__any_sync(0xffffffffu, any_in)) {
                continue;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int t = 0; t < PT; ++t) {
                    const bool in = r2[u][t] < cut2_screen && (!diagonal || pj[t] > i0 + vv + u);
                    const unsigned mask = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int at = queued + __popc(mask & ((1u << lane) - 1u));
                        s_qe[warp][at] = static_cast<unsigned>(vv + u) | (static_cast<unsigned>(pj[t]) << 6);
                    }
                    queued += __popc(mask);
                }
            }
            __syncwarp();
            if (queued > kStreamQueue - 2 * PT * 32) {
                flush();
            }
        }
    }
    flush();
    esum = warpSum(esum);
    if (lane == 0) {
        s_sum[warp] = esum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            s += s_sum[w];
        }
        partials[out_index] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Full rebuild of Q(k) (PolicyIonIon::updateComplex, src/energy.cpp:191-206; PBCEigen quirk :208-217) with
// factorised phases: one block per 4×4×4 cell of k-vectors, particles in chunks of 256.
//   phase A  thread ↔ particle: 2 sincos per axis (cell base 2π b x/L and step 2π x/L), the other three
//            entries of each axis by complex products, then the 16 x·y products → shared memory
//   phase B  thread ↔ (k of the cell, quarter of the chunk): one complex product and two FMAs per
//            (k, particle) instead of a sincos
// Σ_k A_k |Q_k|² of the cell is optionally reduced to e_partials (sharded system energy: no store of Q).
// ------------------------------------------------------------------------------------------------
constexpr int kFullQChunk = 256;

struct FullQSmem
{
    double2 exy[kFullQChunk][16];
    double2 ez[kFullQChunk][4];
    double wre[kFullQChunk]; //!< weight of the real part (charge, 0 if inactive)
    double wim[kFullQChunk]; //!< weight of the imaginary part (charge; PBCEigen: 1 if active)
    double2 red[4][kTileK];
};

__global__ void __launch_bounds__(kBlock, 2)
    ewaldFullCellKernel(SlotView V, EwaldView E, const int4* __restrict__ kn, const int* __restrict__ cell_start,
                        int cell_begin, PhaseGeometry geo, int store_q, double* __restrict__ e_partials, int split_size = 0,
                        double2* __restrict__ q_partials = nullptr)
{
    // split_size > 0 (a slab of few cells on one of several GPUs): blockIdx.y takes the particles
    // [y·split_size, (y+1)·split_size) and leaves its share of Q(k) in q_partials[cell][y][64]; ewaldCellEnergyKernel adds
    // the shares in y order. The split size is a constant of the caller: the sums do not depend on the number of GPUs.
    const int j_begin = split_size > 0 ? static_cast<int>(blockIdx.y) * split_size : 0;
    const int j_end = split_size > 0 ? min(V.n_slots, j_begin + split_size) : V.n_slots;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FullQSmem& sm = *reinterpret_cast<FullQSmem*>(smem_raw);
    const int cell = cell_begin + blockIdx.x;
    const int p0 = cell_start[cell];
    const int len = cell_start[cell + 1] - p0;
    const CellBase base = cellBase(__ldg(kn + p0), geo.ncc);
    const double two_pi = 2.0 * 3.141592653589793238462643383279502884;
    const double kb[3] = {two_pi * static_cast<double>(base.nx) / geo.len[0], two_pi * static_cast<double>(base.ny) / geo.len[1],
                          two_pi * static_cast<double>(base.nz) / geo.len[2]};
    const double k1[3] = {two_pi / geo.len[0], two_pi / geo.len[1], two_pi / geo.len[2]};

    const int kl = threadIdx.x & (kTileK - 1);
    const int quarter = threadIdx.x / kTileK;
    int li = 0, lj = 0, ll = 0;
    const bool kvalid = kl < len;
    if (kvalid) {
        const int4 nn = __ldg(kn + p0 + kl);
        li = nn.x & 3;
        lj = (nn.y + geo.ncc) & 3;
        ll = (nn.z + geo.ncc) & 3;
    }
    const int ixy = li * 4 + lj;
    double qre = 0.0, qim = 0.0;

    for (int c0 = j_begin; c0 < j_end; c0 += kFullQChunk) {
        __syncthreads(); // previous chunk consumed
        {
            const int j = c0 + threadIdx.x;
            double4 p = make_double4(0, 0, 0, 0);
            bool active = false;
            if (j < j_end) {
                p = V.posq[j];
                active = V.gid[j] >= 0;
            }
            double2 e[3][4];
            const double x[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                double s, c;
                sincos(kb[ax] * x[ax], &s, &c);
                e[ax][0] = make_double2(c, s);
                sincos(k1[ax] * x[ax], &s, &c);
                const double2 step = make_double2(c, s);
#pragma unroll
                for (int i = 1; i < 4; ++i) {
                    e[ax][i] = cmul(e[ax][i - 1], step);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    sm.exy[threadIdx.x][i * 4 + jj] = cmul(e[0][i], e[1][jj]);
                }
                sm.ez[threadIdx.x][i] = e[2][i];
            }
            sm.wre[threadIdx.x] = active ? p.w : 0.0;
            sm.wim[threadIdx.x] = active ? (E.policy == 1 ? 1.0 : p.w) : 0.0;
        }
        __syncthreads();
        if (kvalid) {
            const int n = min(kFullQChunk, j_end - c0);
#pragma unroll 4
            for (int t = quarter; t < n; t += 4) {
                const double2 ph = cmul(sm.exy[t][ixy], sm.ez[t][ll]);
                qre = fma(sm.wre[t], ph.x, qre);
                qim = fma(sm.wim[t], ph.y, qim);
            }
        }
    }
    __syncthreads();
    sm.red[quarter][kl] = make_double2(qre, qim);
    __syncthreads();
    double e = 0.0;
    if (quarter == 0 && kvalid) {
        double2 Q = make_double2(0, 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            Q.x += sm.red[q][kl].x;
            Q.y += sm.red[q][kl].y;
        }
        if (store_q) {
            E.Q[p0 + kl] = Q;
        }
        if (q_partials != nullptr) {
            q_partials[(static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * kTileK + kl] = Q;
        }
        e = E.kA[p0 + kl].w * (Q.x * Q.x + Q.y * Q.y);
    }
    __syncthreads();
    if (e_partials != nullptr && q_partials == nullptr) {
        __shared__ double s_e[2];
        if (threadIdx.x < 64) {
            const double s = warpSum(e);
            if ((threadIdx.x & 31) == 0) {
                s_e[threadIdx.x >> 5] = s;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            e_partials[blockIdx.x] = s_e[0] + s_e[1];
        }
    }
}

/** Σ_k A_k |Σ_y Q_y(k)|² of one k-cell from the particle-range shares ewaldFullCellKernel left (y order) */
__global__ void __launch_bounds__(kTileK)
    ewaldCellEnergyKernel(EwaldView E, const int* __restrict__ cell_start, int cell_begin, int n_splits,
                          const double2* __restrict__ q_partials, double* __restrict__ e_partials)
{
    __shared__ double s_e[kTileK / 32];
    const int cell = cell_begin + blockIdx.x;
    const int p0 = cell_start[cell];
    const int len = cell_start[cell + 1] - p0;
    const int kl = threadIdx.x;
    double e = 0.0;
    if (kl < len) {
        double2 Q = make_double2(0.0, 0.0);
        for (int y = 0; y < n_splits; ++y) {
            const double2 q = q_partials[(static_cast<size_t>(blockIdx.x) * n_splits + y) * kTileK + kl];
            Q.x += q.x;
            Q.y += q.y;
        }
        e = E.kA[p0 + kl].w * (Q.x * Q.x + Q.y * Q.y);
    }
    const double s = warpSum(e);
    if ((threadIdx.x & 31) == 0) {
        s_e[threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        e_partials[blockIdx.x] = s_e[0] + s_e[1];
    }
}

} // namespace fbdev

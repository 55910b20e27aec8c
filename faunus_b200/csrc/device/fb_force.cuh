// Forces: the device side of `Energy::Nonbonded::force` (src/energy.h:1584-1597) and
// `Energy::Ewald::force` (src/energy.cpp:596-629), which `Hamiltonian::force` (src/energy.cpp:1162-1166) calls term
// by term for Langevin dynamics (src/forcemove.cpp:124-125).
//
// Restated as the reference has them, stubs included:
//   * Nonbonded::force runs over ALL particles of the particle vector, active or not, every pair i < j, no group
//     rules, no exclusions (`@todo A stub`); the pair force is `PairEnergy::force` (src/energy.h:441-447): minimum
//     image VECTOR b → a (`Chameleon::vdist`, src/geometry.h:429-458) handed to the pair potential's force.
//   * pair forces exist for Lennard-Jones (src/potentials.h:32-40), WCA (:173-184) and the CoulombGalore potential
//     (:600-606, `bjerrum_length · pot.ion_ion_force(qa, qb, r)`); plain `Coulomb`, `HardSphere`, `FunctorPotential`
//     and `SplinedPotential` inherit `PairPotential::force`, which throws (src/potentials.cpp:246-251) — so only
//     `nonbonded_coulomblj` and `nonbonded_coulombwca` have forces, here as there.
//   * Ewald::force ASSIGNS the surface term to the force of each particle, adds the k-space sum and scales the lot by
//     −4π lB / V — whatever the terms before it left in the vector is overwritten, and the surface term is there for
//     any `epss` (no tinfoil test). Reproduced.
//
// A thread owns particle i and adds the forces of the j ≠ i of a fixed range of particles (2048, at most 64 ranges) in index order; the
// shares of the ranges are added in range order: F_i = Σ_j f(i, j). The reference adds f(i, j) to i and subtracts it
// from j for i < j; vdist and the force laws are odd in the distance vector, so the two differ by the order of the
// sums only.
#pragma once
#include "fb_kernels.cuh"

namespace fbdev {

/** Andrea table of S'(q), q ∈ [0, 1] (fb_set_force_table) */
struct ForceTable
{
    int nk;
    const double* knots;
    const double* coef;
};

constexpr int kForceBlock = 128;

/** distance vector b → a, folded once per periodic axis (src/geometry.h:429-458) */
__device__ __forceinline__ double foldComponent(double d, double half, double len_or_zero)
{
    if (len_or_zero > 0.0) {
        if (d > half) {
            d -= len_or_zero;
        }
        else if (d < -half) {
            d += len_or_zero;
        }
    }
    return d;
}

/**
 * −dU/dr / r of the CoulombGalore energy u = lB qq / r · S(r/Rc) · e^{−κr} (the factor of the distance vector):
 * lB qq / r³ · [S(q)(1 + κr) − q S'(q)] e^{−κr} for r² < Rc² (CoulombGalore `ion_ion_force`: no +ε, cutoff tested on
 * r²), S and S' from their Andrea tables.
 */
__device__ __forceinline__ double coulombForceFactor(const PotParams& P, const ForceTable& T, double qq, double r2)
{
    if (!(r2 < P.Rc * P.Rc)) {
        return 0.0;
    }
    const double r = sqrt(r2);
    const double q = r * P.invRc;
    const double S = andreaEval(P.knots, P.coef, 0, P.nk, P.lut, P.nlut, static_cast<double>(P.nlut), q);
    const double dS = andreaEval(T.knots, T.coef, 0, T.nk, nullptr, 0, 0.0, q);
    double f = qq / (r2 * r) * (S * (1.0 + P.kappa * r) - q * dS);
    if (P.kappa > 0.0) {
        f *= exp(-P.kappa * r);
    }
    return P.lB * f;
}

/** src/potentials.h:32-40: 6 · 4ε σ⁶ (2σ⁶ − r⁶) / r¹⁴ */
__device__ __forceinline__ double ljForceFactor(const double* s2, const double* e4, int t, double r2)
{
    const double s = __ldg(s2 + t);
    const double s6 = s * s * s;
    const double r6 = r2 * r2 * r2;
    const double r14 = r6 * r6 * r2;
    return 6.0 * __ldg(e4 + t) * s6 * (2.0 * s6 - r6) / r14;
}

/** src/potentials.h:173-184 */
__device__ __forceinline__ double wcaForceFactor(const double* s2, const double* e4, int t, double r2)
{
    double x = __ldg(s2 + t);
    if (r2 > x * 1.2599210498948732) {
        return 0.0;
    }
    x = x / r2;
    x = x * x * x;
    return __ldg(e4 + t) * 6.0 * (2.0 * x * x - x) / r2;
}

/**
 * F_i = Σ_{j ≠ i} f(i, j) over every particle slot. grid = ⌈n/128⌉ × j-ranges, thread ↔ i; the j of the block's range
 * [y·j_chunk, (y+1)·j_chunk) run through shared memory in tiles of 128, in index order; the share of the range goes to
 * out[y][i][3] and forceSumKernel adds the shares in range order. The chunk is a constant of the caller, so the sums do not
 * depend on the grid. (One range only — 157 blocks of 4 warps at N = 20 000 — left 93 % of the warp slots empty.)
 */
template <int KIND>
__global__ void __launch_bounds__(kForceBlock)
    nonbondedForceKernel(SlotView V, PotParams P, ForceTable T, int j_chunk, double* __restrict__ out)
{
    static_assert(KIND == POT_COULOMB_LJ || KIND == POT_COULOMB_WCA, "the reference has forces for these two only");
    __shared__ double4 s_pos[kForceBlock];
    __shared__ int s_id[kForceBlock];
    const int i = blockIdx.x * kForceBlock + threadIdx.x;
    const bool mine = i < V.n_slots;
    const double4 a = mine ? V.posq[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    const int ida = mine ? V.atom_id[i] : 0;
    const int j_begin = static_cast<int>(blockIdx.y) * j_chunk;
    const int j_end = min(V.n_slots, j_begin + j_chunk);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int j0 = j_begin; j0 < j_end; j0 += kForceBlock) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < j_end) {
            s_pos[threadIdx.x] = V.posq[j];
            s_id[threadIdx.x] = V.atom_id[j];
        }
        __syncthreads();
        const int count = min(kForceBlock, j_end - j0);
        if (!mine) {
            continue;
        }
        for (int t = 0; t < count; ++t) {
            if (j0 + t == i) {
                continue;
            }
            const double4 b = s_pos[t];
            const double dx = foldComponent(a.x - b.x, V.half[0], V.len_or_zero[0]);
            const double dy = foldComponent(a.y - b.y, V.half[1], V.len_or_zero[1]);
            const double dz = foldComponent(a.z - b.z, V.half[2], V.len_or_zero[2]);
            const double r2 = dx * dx + dy * dy + dz * dz;
            const int tt = ida * P.n_types + s_id[t];
            double f = coulombForceFactor(P, T, a.w * b.w, r2);
            if constexpr (KIND == POT_COULOMB_LJ) {
                f += ljForceFactor(P.lj_s2, P.lj_e4, tt, r2);
            }
            else {
                f += wcaForceFactor(P.wca_s2, P.wca_e4, tt, r2);
            }
            fx += f * dx;
            fy += f * dy;
            fz += f * dz;
        }
    }
    if (mine) {
        double* share = out + 3 * (static_cast<size_t>(blockIdx.y) * V.n_slots + i);
        share[0] = fx;
        share[1] = fy;
        share[2] = fz;
    }
}

/** out[t] = Σ_y shares[y][t] in range order, t over the 3n force components */
__global__ void __launch_bounds__(kBlock) forceSumKernel(const double* __restrict__ shares, int n3, int n_ranges, double* __restrict__ out)
{
    const int t = blockIdx.x * kBlock + threadIdx.x;
    if (t < n3) {
        double s = 0.0;
        for (int y = 0; y < n_ranges; ++y) {
            s += shares[static_cast<size_t>(y) * n3 + t];
        }
        out[t] = s;
    }
}

/** Σ_j q_j r_j over EVERY particle slot (Ewald::force sums the particle vector, src/energy.cpp:602-607) → out[0..3) */
__global__ void __launch_bounds__(kBlock) dipoleAllKernel(SlotView V, double* partials, unsigned* ticket, double* out)
{
    __shared__ double scratch[kBlock / 32];
    double sx = 0, sy = 0, sz = 0;
    for (int j = blockIdx.x * kBlock + threadIdx.x; j < V.n_slots; j += gridDim.x * kBlock) {
        const double4 p = V.posq[j];
        sx += p.w * p.x;
        sy += p.w * p.y;
        sz += p.w * p.z;
    }
    double vals[3];
    vals[0] = blockSum<kBlock>(sx, scratch);
    vals[1] = blockSum<kBlock>(sy, scratch);
    vals[2] = blockSum<kBlock>(sz, scratch);
    finalReduce<kBlock>(vals, 3, partials, ticket, out, scratch);
}

constexpr int kEwaldForceParticles = 32; //!< particles per block
constexpr int kEwaldForceLanes = 4;      //!< k-lanes per particle: warp w takes the k-vectors k ≡ w (mod 4)
constexpr int kEwaldForceChunk = 256;    //!< k-vectors staged per pass

/**
 * src/energy.cpp:609-628. Per particle: F = μ_tot q / (2ε_s + 1); F += Σ_k Re(e^{ik·r} · (0 + iq) · conj(Q_k)) A_k k;
 * F *= −4π lB / V. Re(…) = q (cos(k·r) Im Q − sin(k·r) Re Q). The phase is cos/sin of k·r for every policy, as there.
 * Block = 32 particles × 4 k-lanes; the four lane sums of a particle are added in lane order.
 *
 * @param dipole   Σ q r over all slots (dipoleAllKernel)
 * @param surface  1 / (2ε_s + 1)
 * @param scale    −4π lB / V
 */
__global__ void __launch_bounds__(kEwaldForceParticles* kEwaldForceLanes)
    ewaldForceKernel(SlotView V, EwaldView E, const double* __restrict__ dipole, double surface, double scale,
                     double* __restrict__ out)
{
    __shared__ double4 s_k[kEwaldForceChunk];
    __shared__ double2 s_q[kEwaldForceChunk];
    __shared__ double s_f[kEwaldForceLanes][kEwaldForceParticles][3];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int i = blockIdx.x * kEwaldForceParticles + lane;
    const bool mine = i < V.n_slots;
    const double4 p = mine ? V.posq[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int k0 = 0; k0 < E.K; k0 += kEwaldForceChunk) {
        __syncthreads();
        for (int t = threadIdx.x; t < kEwaldForceChunk; t += blockDim.x) {
            if (k0 + t < E.K) {
                s_k[t] = E.kA[k0 + t];
                s_q[t] = E.Q[k0 + t];
            }
        }
        __syncthreads();
        const int count = min(kEwaldForceChunk, E.K - k0);
        for (int t = w; t < count; t += kEwaldForceLanes) {
            const double4 k = s_k[t];
            const double2 Q = s_q[t];
            double s, c;
            sincos(k.x * p.x + k.y * p.y + k.z * p.z, &s, &c);
            const double re = p.w * (c * Q.y - s * Q.x) * k.w;
            fx += re * k.x;
            fy += re * k.y;
            fz += re * k.z;
        }
    }
    s_f[w][lane][0] = fx;
    s_f[w][lane][1] = fy;
    s_f[w][lane][2] = fz;
    __syncthreads();
    if (w == 0 && mine) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double f = dipole[d] * p.w * surface;
            double ksum = s_f[0][lane][d];
#pragma unroll
            for (int l = 1; l < kEwaldForceLanes; ++l) {
                ksum += s_f[l][lane][d];
            }
            f += ksum;
            out[3 * i + d] = f * scale;
        }
    }
}

} // namespace fbdev

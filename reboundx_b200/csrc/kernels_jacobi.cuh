// kernels_jacobi.cuh -- centre-of-mass reference frames of rebx_com_force on the device
// (reference src/rebxtools.c:66-147), included by kernels.cu.
//
// BARYCENTRIC: every particle is evaluated against the total centre of mass; the reference
// then subtracts (m_i/M) f_i from EVERY particle, one i at a time -- O(N^2).  Here: one
// reduction for the mass moments, the ordinary sweep (FLAG_BARYCENTRIC) that also reduces
// sum_i (m_i/M) f_i, and one uniform add.  O(N).
//
// JACOBI: particle i is evaluated against the centre of mass of particles 0..i-1, which the
// reference obtains by peeling particles off the total COM backwards (rebxtools.c:6-30,97-99),
// and (m_i/(M_i+m_i)) f_i is subtracted from every j <= i -- O(N^2) again.  Here the recurrence
// is two scans:  exclusive PREFIX sums of (m, m r, m v) give COM_i, and an inclusive SUFFIX
// sum of t_i = (m_i/(M_i+m_i)) sum_e f_i^e gives what particle j must lose.  O(N), five
// kernels: tile sums -> scan of tile sums -> sweep (in-tile scan + forces) -> reverse scan of
// tile sums -> apply (in-tile reverse scan).
#pragma once

#include "jacobi_common.cuh"
#include "jacobi_fused.cuh"

namespace rebxb {

// ---- mass moments: sum m, sum m r, sum m v ------------------------------------------------------
// tile_sums[7][ntiles] (component-major, so the tile scan reads coalesced); one tile = kScanThreads consecutive particles.
// lead_mass: device pointer to the mass of GLOBAL particle 0 (defines the binade of the mass grid), or NULL:
// the shard's own first element when it starts at global index 0, no quantisation otherwise.
// grid_out / unsafe: written by block 0 / OR-ed by every block (scratch of the launcher).
__global__ void __launch_bounds__(kScanThreads)
moments_tile_kernel(rebx_b200_state st, double* __restrict__ tile_sums, const double* __restrict__ lead_mass,
                    MassGrid* __restrict__ grid_out, unsigned* __restrict__ unsafe, unsigned long long* __restrict__ first_massive) {
    __shared__ double s_warp[kScanThreads/32][7];
    const int64_t k = (int64_t)blockIdx.x*kScanThreads + threadIdx.x;
    const MassGrid g = lead_mass ? make_mass_grid(*lead_mass) : (st.index_base == 0 ? make_mass_grid(st.m[0]) : make_mass_grid(0.0));
    if (blockIdx.x == 0 && threadIdx.x == 0) *grid_out = g;
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    bool bad = false;
    if (k < st.n) {
        const double m = st.m[k];
        v[0] = quantise_mass(m, g, bad);
        v[1] = st.x[k]*m; v[2] = st.y[k]*m; v[3] = st.z[k]*m;
        v[4] = st.vx[k]*m; v[5] = st.vy[k]*m; v[6] = st.vz[k]*m;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(unsafe, 1u);
    // smallest local index >= (global index 1) whose mass is not zero on the grid: only the first tiles can
    // lower it, so nearly every CTA leaves after one read
    // (a candidate only while it could still lower the current minimum, so nearly every CTA leaves after one read;
    // __syncthreads_or makes the branch uniform although the threads may read different values of the global word)
    const bool massive = k < st.n && st.index_base + k >= 1 && v[0] != 0.0 && (unsigned long long)k < *first_massive;
    if (__syncthreads_or(massive)) {
        // one global atomic per CTA: the lowest massive lane of each warp competes in shared memory first (every
        // thread hitting the one global word cost 110 us at 10^6 particles)
        __shared__ unsigned s_first;
        if (threadIdx.x == 0) s_first = 0xffffffffu;
        __syncthreads();
        const unsigned ballot = __ballot_sync(0xffffffffu, massive);
        if (ballot != 0u && (threadIdx.x & 31) == (unsigned)(__ffs(ballot) - 1)) atomicMin(&s_first, threadIdx.x);
        __syncthreads();
        if (threadIdx.x == 0 && s_first != 0xffffffffu) atomicMin(first_massive, (unsigned long long)((int64_t)blockIdx.x*kScanThreads + s_first));
    }
    double total[7];
    block_exclusive_scan<7>(v, total, s_warp);
#pragma unroll
    for (int c = 0; c < 7; c++) if (threadIdx.x == c) tile_sums[(size_t)c*gridDim.x + blockIdx.x] = total[c];
}

// Status of the mass scan + the literal recurrences when the exactness argument does not hold.
//   status[0] (totals[7]) = 0: the quantised scan IS the reference's arithmetic; 1: it is not
//   (running sum left the binade, tie, negative mass, no positive leading mass).
// mref != NULL (single-shard evaluation): on status 1 the forward fold and, when `backward`, the backward peel
// are redone literally by one thread (lanes load 32 masses at a time, lane 0 adds them in order):
// totals[0] = T_{N-1}; mref[k] = B_k.  Cost when taken: 2 N dependent FP64 additions.
__global__ void mass_status_kernel(rebx_b200_state st, MassGrid* __restrict__ grid, const unsigned* __restrict__ unsafe,
                                   const unsigned long long* __restrict__ first_massive,
                                   double* __restrict__ totals, double* __restrict__ mref, int backward, double* __restrict__ totals_out) {
    const int lane = threadIdx.x;
    bool bad = (*unsafe != 0u) || !grid->ok || !(totals[0] < grid->top);
    if (!bad && grid->lead == 0.5*grid->top) {
        // leading mass an exact power of two: the particles in front of the first massive one sit on the finer grid
        // below 2^K once the peel has dipped; a non-zero mass among them (too small for the grid) is outside the
        // argument.  Normally there is none to look at (the first massive particle is particle 1).
        const unsigned long long r = *first_massive;
        const int64_t lo = st.index_base >= 1 ? 0 : 1 - st.index_base;
        const int64_t hi = r < (unsigned long long)st.n ? (int64_t)r : st.n;
        bool tiny = hi - lo > 65536;
        for (int64_t k = lo + lane; k < hi && !tiny; k += 32) if (st.m[k] != 0.0) tiny = true;
        bad = __any_sync(0xffffffffu, tiny);
    }
    __syncwarp();
    if (lane == 0) {
        totals[7] = bad ? 1.0 : 0.0;
        const unsigned long long r = *first_massive;
        if (grid->ok && r < (unsigned long long)st.n) {       // B_r = fl((m_0 + u q_r) - m_r), see the header comment
            const double mr = st.m[r];
            bool unused = false;
            const double mq = quantise_mass(mr, *grid, unused);
            grid->dip = (grid->lead + mq) - mr;
        }
    }
    if (!bad || mref == nullptr) {
        __syncwarp();
        if (totals_out && lane < 8) totals_out[lane] = totals[lane];
        return;
    }
    double T = 0.0;
    for (int64_t base = 0; base < st.n; base += 32) {
        const int64_t k = base + lane;
        const double mine = (k < st.n) ? st.m[k] : 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double mj = __shfl_sync(0xffffffffu, mine, j);
            if (base + j < st.n) T += mj;                    // c.m += m_i   (reb_simulation_com)
        }
    }
    if (lane == 0) totals[0] = T;
    __syncwarp();
    if (totals_out && lane < 8) totals_out[lane] = totals[lane];
    if (!backward) return;
    double B = T;
    for (int64_t top = st.n; top > 0; top -= 32) {
        const int64_t k = top - 1 - lane;                    // lane j holds particle top-1-j
        const double mine = (k >= 0) ? st.m[k] : 0.0;
        double out = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double mj = __shfl_sync(0xffffffffu, mine, j);
            if (top - 1 - j >= 0) B -= mj;                   // com.m -= p.m  (rebx_get_com_without_particle)
            if (lane == j) out = B;
        }
        if (k >= 0) mref[k] = out;                           // mass interior to particle k
    }
}

// In-place exclusive scan over the `ntiles` tile sums of NC components, stored component-major
// rows[c*ntiles + tile] (forward over the tiles, or backward when REV).  One CTA walks the tiles in chunks of
// kTileScanThreads, thread t <-> tile chunk*kTileScanThreads + t, so every access is coalesced (a first
// version gave each thread a contiguous span of [tile][NC] rows: 86 k uncoalesced line accesses from one SM,
// 43 us for 3907 tiles); the next chunk is loaded before the current one is scanned.  Fixed order =>
// deterministic.  totals[NC] gets the grand total.
// The components are independent: CTA c scans component c (grid = NC), so the CTAs run side by side.
constexpr int kTileScanThreads = 1024;
template <int NC, bool REV>
__global__ void __launch_bounds__(kTileScanThreads)
scan_tiles_kernel(double* __restrict__ rows, int64_t ntiles, double* __restrict__ totals) {
    __shared__ double s_warp[kTileScanThreads/32][1];
    const int t = threadIdx.x;
    double* __restrict__ comp = rows + (size_t)blockIdx.x*ntiles;
    auto at = [&](int64_t i) { return REV ? (ntiles - 1 - i) : i; };
    double carry = 0.0;
    double nx = (t < ntiles) ? comp[at(t)] : 0.0;
    for (int64_t base = 0; base < ntiles; base += kTileScanThreads) {
        const int64_t i = base + t, inext = i + kTileScanThreads;
        double x[1] = {nx}, total[1];
        nx = (inext < ntiles) ? comp[at(inext)] : 0.0;
        block_exclusive_scan<1, kTileScanThreads>(x, total, s_warp);       // x := sum of the chunk's tiles in front of tile i
        if (i < ntiles) comp[at(i)] = carry + x[0];
        carry += total[0];
    }
    if (totals && t == 0) totals[blockIdx.x] = carry;
}

// Body record of the barycentre from the grand totals (reb_simulation_com restated as a
// reduction: x = sum(m x)/sum(m); index -1 = "no particle").
__global__ void com_body_kernel(const double* __restrict__ totals, double* __restrict__ body, int nbodies) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double M = totals[0];
    for (int b = 0; b < nbodies; b++) {
        double* rec = body + (size_t)b*REBX_B200_BODY_DOUBLES;
        for (int j = 0; j < REBX_B200_BODY_DOUBLES; j++) rec[j] = 0.0;
        rec[REBX_B200_BODY_INDEX] = -1.0;
        for (int j = 0; j < 6; j++) rec[REBX_B200_BODY_X + j] = (M > 0.) ? totals[1 + j]/M : totals[1 + j];
        rec[REBX_B200_BODY_M] = M;
    }
}

// ---- Jacobi sweep --------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ Vec3 damping_force(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Particle& p, int64_t k,
                                              const Orbit* so) {
    if constexpr (K == REBX_B200_MODIFY_ORBITS) {
        return modify_orbits(fx, org, body, p, __ldcs(fx.tab[0] + k), __ldcs(fx.tab[1] + k), __ldcs(fx.tab[2] + k), so);
    } else if constexpr (K == REBX_B200_GAS_DAMPING) {
        return gas_damping(fx, org, body, p, __ldcs(fx.tab[0] + k), so);
    } else if constexpr (K == REBX_B200_TYPE_I) {
        return type_I_migration(fx, org, body, p, so);
    } else if constexpr (K == REBX_B200_EXP_MIGRATION) {
        return exponential_migration(fx, org, body, p, __ldcs(fx.tab[0] + k), __ldcs(fx.tab[1] + k), __ldcs(fx.tab[2] + k));
    } else {
        return Vec3{0.0, 0.0, 0.0};
    }
}

// tile_prefix[7][ntiles]: exclusive prefix of the mass moments at the tile's first particle.
// Writes a += f (particle's own force), tvec[3][n] = t_i, tile_tsums[3][ntiles] = sum of t in the tile.
template <int K0, int K1, int K2>
__global__ void __launch_bounds__(kScanThreads)
jacobi_sweep_kernel(const __grid_constant__ rebx_b200_pass P, const double* __restrict__ tile_prefix,
                    double* __restrict__ tvec, double* __restrict__ tile_tsums, const double* __restrict__ carry7,
                    const MassGrid* __restrict__ grid, const double* __restrict__ status, const double* __restrict__ mref) {
    __shared__ double s_warp[kScanThreads/32][7];
    const rebx_b200_state& st = P.st;
    const int64_t k = (int64_t)blockIdx.x*kScanThreads + threadIdx.x;
    const bool live = k < st.n;
    Particle p = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        p.x = st.x[k]; p.y = st.y[k]; p.z = st.z[k]; p.vx = st.vx[k]; p.vy = st.vy[k]; p.vz = st.vz[k]; p.m = st.m[k];
    }
    bool unused_flag = false;
    double v[7] = {live ? quantise_mass(p.m, *grid, unused_flag) : 0.0, p.x*p.m, p.y*p.m, p.z*p.m, p.vx*p.m, p.vy*p.m, p.vz*p.m};
    double total[7];
    block_exclusive_scan<7>(v, total, s_warp);
    double body[REBX_B200_BODY_DOUBLES];
    // carry7: mass moments of the shards in front of this one (index ranges below index_base), or NULL
    double M = tile_prefix[blockIdx.x] + v[0];                // mass interior to particle k: exact sums of multiples of u
    if (carry7) M = carry7[0] + M;
    if (grid->ok && M == grid->lead && st.index_base + k >= 1) M = grid->dip;   // bottom of the binade: the literal last peel
    if (mref != nullptr && status[0] != 0.0 && live) M = mref[k];     // the literal recurrences (mass_status_kernel)
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double S = tile_prefix[(size_t)(1 + j)*gridDim.x + blockIdx.x] + v[1 + j];
        if (carry7) S = carry7[1 + j] + S;
        body[REBX_B200_BODY_X + j] = (M > 0.) ? S/M : S;
    }
    body[REBX_B200_BODY_INDEX] = -1.0;
    body[REBX_B200_BODY_M] = M;

    double t[3] = {0.0, 0.0, 0.0};
    const long long gi = st.index_base + k;
    if (live && gi != 0) {                                              // particle 0 has no Jacobi coordinate (rebxtools.c:71-73)
        Vec3 a = {st.ax[k], st.ay[k], st.az[k]};
        const double mr = p.m/(M + p.m);
        Vec3 f;
        const Origin org = load_origin(body);
        // stacked damping effects evaluate the SAME two-body elements (same particle, same interior centre of
        // mass, same G): computed once, bit-identical to each effect's own call (as in pass_kernel)
        constexpr int NDAMP = (K0 >= 5 && K0 <= 7) + (K1 >= 5 && K1 <= 7) + (K2 >= 5 && K2 <= 7);
        Orbit oc;
        const Orbit* so = nullptr;
        if constexpr (NDAMP >= 2) { oc = orbit_from_particle<true, true, false>(P.fx[0].scal[0], p, org, body); so = &oc; }
        f = damping_force<K0>(P.fx[0], org, body, p, k, so);
        a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        if constexpr (K1 != 0) {
            f = damping_force<K1>(P.fx[1], org, body, p, k, so);
            a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        }
        if constexpr (K2 != 0) {
            f = damping_force<K2>(P.fx[2], org, body, p, k, so);
            a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        }
        st.ax[k] = a.x; st.ay[k] = a.y; st.az[k] = a.z;
    }
    if (live) { tvec[k] = t[0]; tvec[st.n + k] = t[1]; tvec[2*st.n + k] = t[2]; }
    // tile total of t (any order is fine; reuse the scan)
    double tt[3], tot3[3];
    tt[0] = t[0]; tt[1] = t[1]; tt[2] = t[2];
    block_exclusive_scan<3>(tt, tot3, reinterpret_cast<double (*)[3]>(&s_warp[0][0]));
#pragma unroll
    for (int c = 0; c < 3; c++) if (threadIdx.x == c) tile_tsums[(size_t)c*gridDim.x + blockIdx.x] = tot3[c];
}

// a_j -= sum_{i >= j} t_i : in-tile reverse inclusive scan + exclusive suffix of the tiles behind.
__global__ void __launch_bounds__(kScanThreads)
jacobi_apply_kernel(rebx_b200_state st, const double* __restrict__ tvec, const double* __restrict__ tile_suffix,
                    const double* __restrict__ carry3) {
    __shared__ double s_warp[kScanThreads/32][3];
    const int64_t tile_lo = (int64_t)blockIdx.x*kScanThreads;
    const int64_t tile_hi = (tile_lo + kScanThreads < st.n) ? tile_lo + kScanThreads : st.n;
    const int64_t k = tile_hi - 1 - threadIdx.x;                        // reversed order inside the tile
    const bool live = k >= tile_lo;
    double t[3] = {0.0, 0.0, 0.0};
    if (live) { t[0] = tvec[k]; t[1] = tvec[st.n + k]; t[2] = tvec[2*st.n + k]; }
    double own[3] = {t[0], t[1], t[2]}, tot[3];
    block_exclusive_scan<3>(t, tot, s_warp);
    if (live) {
        // carry3: sum of t over the shards behind this one (higher index ranges), or NULL
        double s0 = tile_suffix[(size_t)0*gridDim.x + blockIdx.x] + (t[0] + own[0]);
        double s1 = tile_suffix[(size_t)1*gridDim.x + blockIdx.x] + (t[1] + own[1]);
        double s2 = tile_suffix[(size_t)2*gridDim.x + blockIdx.x] + (t[2] + own[2]);
        if (carry3) { s0 = carry3[0] + s0; s1 = carry3[1] + s1; s2 = carry3[2] + s2; }
        st.ax[k] -= s0; st.ay[k] -= s1; st.az[k] -= s2;
    }
}

// ---- the whole pipeline in one cooperative launch (jacobi_fused.cuh), reference-exact arithmetic ----------------------
// The force of the stack on particle k against the centre of mass of the particles in front of it: the body of
// jacobi_sweep_kernel, operation for operation (IEEE divisions for COM_k = S/M, the effects' own functions).
template <int K0, int K1, int K2>
struct JacobiExactForce {
    struct Ops { int unused; };
    __device__ __forceinline__ void init(const rebx_b200_pass&) {}
    __device__ __forceinline__ Ops load(const rebx_b200_pass&, int64_t) { return Ops{0}; }
    // a += f_e and t += m/(M + m) f_e, effect by effect (the association order of jacobi_sweep_kernel)
    __device__ __forceinline__ void eval(const rebx_b200_pass& P, const Ops&, int64_t k, double px, double py, double pz, double pvx, double pvy, double pvz,
                                         double pm, double M, const double (&S)[6], Vec3& a, double (&t)[3]) {
        double body[REBX_B200_BODY_DOUBLES];
#pragma unroll
        for (int j = 0; j < 6; j++) body[REBX_B200_BODY_X + j] = (M > 0.) ? S[j]/M : S[j];
        body[REBX_B200_BODY_INDEX] = -1.0;
        body[REBX_B200_BODY_M] = M;
        const Particle p = {px, py, pz, pvx, pvy, pvz, pm, 0.0};
        const double mr = pm/(M + pm);
        const Origin org = load_origin(body);
        constexpr int NDAMP = (K0 >= 5 && K0 <= 7) + (K1 >= 5 && K1 <= 7) + (K2 >= 5 && K2 <= 7);
        Orbit oc;
        const Orbit* so = nullptr;
        if constexpr (NDAMP >= 2) { oc = orbit_from_particle<true, true, false>(P.fx[0].scal[0], p, org, body); so = &oc; }
        Vec3 f = damping_force<K0>(P.fx[0], org, body, p, k, so);
        a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        if constexpr (K1 != 0) {
            f = damping_force<K1>(P.fx[1], org, body, p, k, so);
            a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        }
        if constexpr (K2 != 0) {
            f = damping_force<K2>(P.fx[2], org, body, p, k, so);
            a.x += f.x; a.y += f.y; a.z += f.z;  t[0] += mr*f.x; t[1] += mr*f.y; t[2] += mr*f.z;
        }
    }
};

// Resident CTAs per SM of the exact fused kernel: 3 (80 registers, 156 bytes of spills) measured 0.153 ms at 10^6 against
// 0.160 ms with 2 (128 registers, no spills) -- the exact arithmetic is bound by dependent division chains, more warps win.
#ifndef REBXB_JF_EXACT_BLOCKS
#define REBXB_JF_EXACT_BLOCKS 3
#endif
template <int K0, int K1, int K2>
__global__ void __launch_bounds__(kScanThreads, REBXB_JF_EXACT_BLOCKS)
jacobi_fused_kernel(const __grid_constant__ rebx_b200_pass P, const FusedScratch F, const int wtiles) {
    jacobi_fused_body<JacobiExactForce<K0, K1, K2>>(P, F, wtiles);
}

typedef void (*jacobi_fused_fn)(const rebx_b200_pass, const FusedScratch, const int);
typedef void (*jacobi_fn)(const rebx_b200_pass, const double*, double*, double*, const double*, const MassGrid*, const double*, const double*);
struct JacobiCombo { int k[3]; jacobi_fn fn; jacobi_fused_fn fused; };
#define REBXB_JCOMBO(a, b, c) { {a, b, c}, jacobi_sweep_kernel<a, b, c>, jacobi_fused_kernel<a, b, c> },
static const JacobiCombo kJacobiCombos[] = {
    REBXB_JCOMBO(5,0,0) REBXB_JCOMBO(6,0,0) REBXB_JCOMBO(7,0,0) REBXB_JCOMBO(9,0,0)
    REBXB_JCOMBO(5,6,0) REBXB_JCOMBO(6,5,0) REBXB_JCOMBO(5,7,0) REBXB_JCOMBO(7,5,0) REBXB_JCOMBO(6,7,0) REBXB_JCOMBO(7,6,0)
    REBXB_JCOMBO(5,6,7) REBXB_JCOMBO(5,7,6) REBXB_JCOMBO(6,5,7) REBXB_JCOMBO(6,7,5) REBXB_JCOMBO(7,5,6) REBXB_JCOMBO(7,6,5)
};

}  // namespace rebxb

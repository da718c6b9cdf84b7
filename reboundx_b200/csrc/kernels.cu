// kernels.cu -- hand-written FP64 CUDA kernels (sm_100a) behind include/rebx_b200.h.
//
// pass_kernel<V, K0..K3>: one streaming sweep over a shard of SoA particle state that
// applies up to four stacked effects in list order.  It replaces the reference's
// dispatch loop (src/core.c:1031-1040) plus the per-particle loops of the effects
// (src/radiation_forces.c:73, src/gravitational_harmonics.c:184, src/gr_potential.c:63,
// src/yarkovsky_effect.c:224, src/rebxtools.c:92).
//
//  * V = 2: each thread owns two consecutive particles and moves them with 16-byte
//    double2 loads/stores (coalesced: a warp touches 512 contiguous bytes per array).
//  * Body records (source / primary / oblate body) are staged once per CTA in shared
//    memory and read back as broadcasts.
//  * Back-reaction onto the body is accumulated per thread, reduced with warp shuffles,
//    then per CTA into `partials`; finalize_kernel sums the CTA partials in a fixed
//    order, so results are deterministic for a given grid.
//  * HBM-bound (no reuse): grid = SM count x resident CTAs, grid-stride loop, streaming
//    cache hints.  No tensor cores: nothing here is a dense contraction.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (parity build; see
// DESIGN.md for why contraction is off).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>
#include "rebx_b200.h"
#include "pass_common.cuh"
#include "forces_math.cuh"

#ifndef REBXB_FASTMATH
#define REBXB_FASTMATH 0
#endif

namespace rebxb {

template <int K, int V, class M = IeeeMath>
__device__ __forceinline__ void apply(const rebx_b200_effect& fx, const Origin& org, const double* __restrict__ body, const Pre& pre,
                                      const Particle& p, const Params<K, V>& q, int v, Vec3& a, Vec3& red, unsigned& bad,
                                      const Orbit* so = nullptr) {
    if constexpr (K == REBX_B200_RADIATION) {
        radiation<M>(fx, org, body, pre, p, q.t[0][v], a, bad);
    } else if constexpr (K == REBX_B200_HARMONICS) {
        harmonics<M>(fx, org, body, pre, p, a, red, bad);
    } else if constexpr (K == REBX_B200_GR_POTENTIAL) {
        gr_potential<M>(fx, org, body, pre, p, a, red, bad);
    } else if constexpr (K == REBX_B200_YARKOVSKY) {
        YarkParams yp{q.t[0][v], q.t[1][v], q.t[2][v], q.t[3][v], q.t[4][v], q.t[5][v], q.t[6][v], q.t[7][v], q.t[8][v]};
        yarkovsky(fx, org, body, pre, p, yp, q.it[v], a);
    } else if constexpr (K == REBX_B200_MODIFY_ORBITS) {
        const Vec3 f = modify_orbits(fx, org, body, p, q.t[0][v], q.t[1][v], q.t[2][v], so);
        com_force_tail(fx, org, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_GAS_DAMPING) {
        const Vec3 f = gas_damping(fx, org, body, p, q.t[0][v], so);
        com_force_tail(fx, org, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_TYPE_I) {
        const Vec3 f = type_I_migration(fx, org, body, p, so);
        com_force_tail(fx, org, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_CENTRAL_FORCE) {
        central_force(fx, org, body, p, a, red);
    } else if constexpr (K == REBX_B200_EXP_MIGRATION) {
        const Vec3 f = exponential_migration(fx, org, body, p, q.t[0][v], q.t[1][v], q.t[2][v]);
        com_force_tail(fx, org, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_LENSE_THIRRING) {
        lense_thirring(fx, org, body, p, a, red);
    } else if constexpr (K == REBX_B200_GAS_DF) {
        gas_dynamical_friction(fx, org, body, p, a);
    }
}

// Resident CTAs per SM the register allocator is asked to make room for.  The light,
// HBM-bound sweeps (radiation, gr_potential) are latency-bound at 2 CTAs (round-1 ncu: 102
// registers, 16 warps/SM, 0.49 eligible warps/cycle, 80 % of HBM); capping them at 64
// registers (4 CTAs, 32 warps/SM) costs a few spilled doubles but lifts the radiation sweep
// to 93 % of the measured copy bandwidth.  The FP64-heavy sweeps keep their registers.
#ifndef REBXB_LIGHT_MINBLOCKS
#define REBXB_LIGHT_MINBLOCKS 4
#endif
template <int K0, int K1, int K2, int K3>
__host__ __device__ constexpr int min_blocks() {
    constexpr bool light = (K0 == REBX_B200_RADIATION || K0 == REBX_B200_GR_POTENTIAL) &&
                           (K1 == 0 || K1 == REBX_B200_RADIATION || K1 == REBX_B200_GR_POTENTIAL) && K2 == 0 && K3 == 0;
    return light ? REBXB_LIGHT_MINBLOCKS*(256/kThreads) : REBXB_HEAVY_MINBLOCKS;
}

// CTA geometry of the plain sweep per stack.  Occupancy moves in coarse steps with 256-thread CTAs (2 CTAs = 16
// warps at 128 registers, 3 CTAs need <= 85); 128-thread CTAs add the step in between: 5 CTAs = 20 warps at 96
// registers.  Measured on all four configurations (profiles/r01_variant_sweep.txt): only the J2/J4 (+) gr_potential
// pair fits 96 registers without further spills and gains (0.2385 -> 0.2278 ms); the others stay at 256 x 2 / 256 x 4.
#ifndef REBXB_PASS_PREFETCH_ALL
#define REBXB_PASS_PREFETCH_ALL 0
#endif
#ifndef REBXB_RING_128
#define REBXB_RING_128 1
#endif
template <int K0, int K1, int K2, int K3>
__host__ __device__ constexpr bool narrow_cta() {
    return REBXB_RING_128 && kThreads == 256 && K2 == 0 && K3 == 0 &&
           ((K0 == REBX_B200_HARMONICS && K1 == REBX_B200_GR_POTENTIAL) || (K0 == REBX_B200_GR_POTENTIAL && K1 == REBX_B200_HARMONICS));
}
template <int K0, int K1, int K2, int K3>
__host__ __device__ constexpr int pass_threads() { return narrow_cta<K0, K1, K2, K3>() ? 128 : kThreads; }
template <int K0, int K1, int K2, int K3>
__host__ __device__ constexpr int pass_blocks() { return narrow_cta<K0, K1, K2, K3>() ? 5 : min_blocks<K0, K1, K2, K3>(); }

template <int V, bool SHARED, int K0, int K1, int K2, int K3>
__global__ void __launch_bounds__(pass_threads<K0, K1, K2, K3>(), pass_blocks<K0, K1, K2, K3>())
pass_kernel(const __grid_constant__ rebx_b200_pass P, double* __restrict__ partials) {
    constexpr bool VEL  = Traits<K0>::vel  || Traits<K1>::vel  || Traits<K2>::vel  || Traits<K3>::vel;
    constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass || Traits<K3>::mass;
    constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad  || Traits<K3>::rad;
    constexpr int T = pass_threads<K0, K1, K2, K3>();      // threads per CTA of this stack

    __shared__ double s_body[REBX_B200_MAX_EFFECTS][REBX_B200_BODY_DOUBLES];
    for (int idx = threadIdx.x; idx < P.n_effects*REBX_B200_BODY_DOUBLES; idx += T) {
        const int e = idx/REBX_B200_BODY_DOUBLES, j = idx - e*REBX_B200_BODY_DOUBLES;
        s_body[e][j] = P.fx[e].body[j];
    }
    __syncthreads();
    const long long b0 = (long long)s_body[0][0];
    const long long b1 = (K1 != 0) ? (long long)s_body[1][0] : -1;
    const long long b2 = (K2 != 0) ? (long long)s_body[2][0] : -1;
    const long long b3 = (K3 != 0) ? (long long)s_body[3][0] : -1;

    // SHARED: every effect of the stack refers to the same central body, so position / velocity / mass are
    // read from record 0 for all of them and the compiler merges the common sub-expressions of the
    // stacked effects (relative position, r^2, the orbit elements of the damping trio).
    const Origin org0 = load_origin(s_body[0]);
    const Origin org1 = SHARED ? org0 : load_origin(s_body[1]);
    const Origin org2 = SHARED ? org0 : load_origin(s_body[2]);
    const Origin org3 = SHARED ? org0 : load_origin(s_body[3]);
    const Pre pre0 = prepare<K0>(P.fx[0], org0, s_body[0]);
    const Pre pre1 = prepare<K1>(P.fx[1], org1, s_body[1]);
    const Pre pre2 = prepare<K2>(P.fx[2], org2, s_body[2]);
    const Pre pre3 = prepare<K3>(P.fx[3], org3, s_body[3]);
    Vec3 red0 = {0., 0., 0.}, red1 = {0., 0., 0.}, red2 = {0., 0., 0.}, red3 = {0., 0., 0.};

    const int64_t n = P.st.n;
    const int64_t ntiles = (n + V - 1)/V;
    const int64_t stride = (int64_t)gridDim.x*T;
    for (int64_t t = (int64_t)blockIdx.x*T + threadIdx.x; t < ntiles; t += stride) {
        const int64_t k = t*V;
        const bool full = (k + V <= n);
        // the register-heavy sweeps (2 CTAs of 256 or 5 of 128 per SM) prefetch the next iteration's operands
        constexpr bool PREFETCH = REBXB_PASS_PREFETCH_ALL || min_blocks<K0, K1, K2, K3>() != 4;
        if constexpr (PREFETCH) {
            if (t + stride < ntiles) {
                const int64_t kn = (t + stride)*V;
                prefetch_next(P.st.x + kn); prefetch_next(P.st.y + kn); prefetch_next(P.st.z + kn);
                if constexpr (VEL)  { prefetch_next(P.st.vx + kn); prefetch_next(P.st.vy + kn); prefetch_next(P.st.vz + kn); }
                if constexpr (MASS) prefetch_next(P.st.m + kn);
                if constexpr (RAD)  prefetch_next(P.st.r + kn);
                prefetch_next(P.st.ax + kn); prefetch_next(P.st.ay + kn); prefetch_next(P.st.az + kn);
                prefetch_params<K0>(P.fx[0], kn);
                if constexpr (K1 != 0) prefetch_params<K1>(P.fx[1], kn);
                if constexpr (K2 != 0) prefetch_params<K2>(P.fx[2], kn);
                if constexpr (K3 != 0) prefetch_params<K3>(P.fx[3], kn);
            }
        }
        double x[V], y[V], z[V], vx[V], vy[V], vz[V], m[V], r[V], ax[V], ay[V], az[V];
        ld<V>(P.st.x, k, full, x); ld<V>(P.st.y, k, full, y); ld<V>(P.st.z, k, full, z);
        if constexpr (VEL)  { ld<V>(P.st.vx, k, full, vx); ld<V>(P.st.vy, k, full, vy); ld<V>(P.st.vz, k, full, vz); }
        if constexpr (MASS) { ld<V>(P.st.m, k, full, m); }
        if constexpr (RAD)  { ld<V>(P.st.r, k, full, r); }
        ld<V>(P.st.ax, k, full, ax); ld<V>(P.st.ay, k, full, ay); ld<V>(P.st.az, k, full, az);
        Params<K0, V> q0; Params<K1, V> q1; Params<K2, V> q2; Params<K3, V> q3;
        load_params<K0, V>(P.fx[0], k, full, q0);
        if constexpr (K1 != 0) load_params<K1, V>(P.fx[1], k, full, q1);
        if constexpr (K2 != 0) load_params<K2, V>(P.fx[2], k, full, q2);
        if constexpr (K3 != 0) load_params<K3, V>(P.fx[3], k, full, q3);
        // Light sweeps run at the 64-register cap: re-reading the body from shared memory per tile is
        // cheaper there than keeping its seven doubles live across the loop (0.173 vs 0.180 ms measured).
        constexpr bool LIGHT = min_blocks<K0, K1, K2, K3>() == 4;
        if constexpr (LIGHT) asm volatile("" ::: "memory");
        const Origin o0 = LIGHT ? load_origin(s_body[0]) : org0;
        const Origin o1 = LIGHT ? load_origin(s_body[SHARED ? 0 : 1]) : org1;
        auto particle = [&](int v) {
            Particle p;
            p.x = x[v]; p.y = y[v]; p.z = z[v];
            if constexpr (VEL)  { p.vx = vx[v]; p.vy = vy[v]; p.vz = vz[v]; } else { p.vx = p.vy = p.vz = 0.0; }
            if constexpr (MASS) { p.m = m[v]; } else { p.m = 0.0; }
            if constexpr (RAD)  { p.r = r[v]; } else { p.r = 0.0; }
            return p;
        };
        // reference semantics with the hardware operators: branches around the body itself
        auto exact_one = [&](int v) {
            const Particle p = particle(v);
            Vec3 a = {ax[v], ay[v], az[v]};
            unsigned unused = 0;
            const long long gi = P.st.index_base + k + v;
            if constexpr (SHARED) {
                // one central body: one guard around the whole stack, so the effects sit in one region and
                // their common sub-expressions (relative coordinates, orbit elements) are computed once
                if (gi != b0) {
                    // stacked damping effects evaluate the SAME two-body elements (same particle, same
                    // central body, same G): computed once here, bit-identical to each effect's own call
                    constexpr int NDAMP = (K0 >= 5 && K0 <= 7) + (K1 >= 5 && K1 <= 7) + (K2 >= 5 && K2 <= 7) + (K3 >= 5 && K3 <= 7);
                    Orbit oc;
                    const Orbit* so = nullptr;
                    if constexpr (NDAMP >= 2) { oc = orbit_from_particle<true, true, false>(P.fx[0].scal[0], p, o0, s_body[0]); so = &oc; }
                    apply<K0, V>(P.fx[0], o0, s_body[0], pre0, p, q0, v, a, red0, unused, so);
                    if constexpr (K1 != 0) apply<K1, V>(P.fx[1], o1, s_body[1], pre1, p, q1, v, a, red1, unused, so);
                    if constexpr (K2 != 0) apply<K2, V>(P.fx[2], org2, s_body[2], pre2, p, q2, v, a, red2, unused, so);
                    if constexpr (K3 != 0) apply<K3, V>(P.fx[3], org3, s_body[3], pre3, p, q3, v, a, red3, unused, so);
                }
            } else {
                if (gi != b0) apply<K0, V>(P.fx[0], o0, s_body[0], pre0, p, q0, v, a, red0, unused);
                if constexpr (K1 != 0) { if (gi != b1) apply<K1, V>(P.fx[1], o1, s_body[1], pre1, p, q1, v, a, red1, unused); }
                if constexpr (K2 != 0) { if (gi != b2) apply<K2, V>(P.fx[2], org2, s_body[2], pre2, p, q2, v, a, red2, unused); }
                if constexpr (K3 != 0) { if (gi != b3) apply<K3, V>(P.fx[3], org3, s_body[3], pre3, p, q3, v, a, red3, unused); }
            }
            ax[v] = a.x; ay[v] = a.y; az[v] = a.z;
        };
        // MEASURED (B200, round 1): the straight-line FastMath pass keeps every bit-exactness test green
        // but does not pay: J2/J4+gr_potential 0.2875 ms (hardware operators: 0.2857 ms), radiation
        // 0.197 ms vs 0.172 ms (the extra live values spill at the 64-register cap).  The FP64-heavy
        // sweeps are bound by registers -> warps x ILP, not by basic-block boundaries.  Compile with
        // -DREBXB_FASTMATH=1 to enable it; off by default.
        constexpr bool FAST = REBXB_FASTMATH && K0 <= REBX_B200_GR_POTENTIAL && K1 <= REBX_B200_GR_POTENTIAL &&
                              K2 <= REBX_B200_GR_POTENTIAL && K3 <= REBX_B200_GR_POTENTIAL;
        if constexpr (FAST) {
            // Straight-line pass over the thread's V particles (one basic block, so their dependent
            // FP64 chains interleave); operands outside FastMath's domain, the body itself and a
            // ragged tail are flagged and redone below with the hardware operators.
            unsigned bad = full ? 0u : 1u;
            const Vec3 s0 = red0, s1 = red1, s2 = red2, s3 = red3;
            Vec3 an[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
                const Particle p = particle(v);
                Vec3 a = {ax[v], ay[v], az[v]};
                const long long gi = P.st.index_base + k + v;
                bad |= (gi == b0 || gi == b1 || gi == b2 || gi == b3) ? 1u : 0u;
                apply<K0, V, FastMath>(P.fx[0], o0, s_body[0], pre0, p, q0, v, a, red0, bad);
                if constexpr (K1 != 0) apply<K1, V, FastMath>(P.fx[1], o1, s_body[1], pre1, p, q1, v, a, red1, bad);
                if constexpr (K2 != 0) apply<K2, V, FastMath>(P.fx[2], org2, s_body[2], pre2, p, q2, v, a, red2, bad);
                if constexpr (K3 != 0) apply<K3, V, FastMath>(P.fx[3], org3, s_body[3], pre3, p, q3, v, a, red3, bad);
                an[v] = a;
            }
            if (bad) {
                red0 = s0; red1 = s1; red2 = s2; red3 = s3;
#pragma unroll
                for (int v = 0; v < V; v++) { if (v > 0 && !full) break; exact_one(v); }
            } else {
#pragma unroll
                for (int v = 0; v < V; v++) { ax[v] = an[v].x; ay[v] = an[v].y; az[v] = an[v].z; }
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; v++) { if (v > 0 && !full) break; exact_one(v); }
        }
        st<V>(P.st.ax, k, full, ax); st<V>(P.st.ay, k, full, ay); st<V>(P.st.az, k, full, az);
    }
    finish_sums<T, K0, K1, K2, K3>(red0, red1, red2, red3, P, partials, ApplyTail{});
}

// ---------------------------------------------------------------------------------------------
// step_kernel<V, K0..K3>: one whole drift-kick-drift step in ONE sweep.
// Per particle:  x += dt/2 v;  a = pull of the central body + stacked effects;  v += dt a;
// x += dt/2 v.  The acceleration lives in registers only, so a step moves 96 B/particle (+ tables)
// instead of the ~370 B of the five separate sweeps (drift, gravity, force, kick, drift), with the
// identical per-particle arithmetic (bit-identical trajectories, tests/test_gpu_parity.py).
// Requires every effect of the stack to refer to the same central body; that body itself is skipped
// here and advanced by step_body_kernel once the back-reaction sums are final.
// ---------------------------------------------------------------------------------------------
struct StepArgs { double dt, hdt, G, soft2; };

static __device__ __noinline__ void step_body(const rebx_b200_pass& P, const StepArgs& S, unsigned red_mask) {
    double* rec0 = const_cast<double*>(P.fx[0].body);
    const long long bi = (long long)rec0[REBX_B200_BODY_INDEX];
    const long long k = bi - P.st.index_base;
    double pos[3] = {rec0[REBX_B200_BODY_X+0], rec0[REBX_B200_BODY_X+1], rec0[REBX_B200_BODY_X+2]};
    double vel[3] = {rec0[REBX_B200_BODY_VX+0], rec0[REBX_B200_BODY_VX+1], rec0[REBX_B200_BODY_VX+2]};
    double a[3] = {0.0, 0.0, 0.0};
    for (int e = 0; e < P.n_effects; e++)
        if (((red_mask >> e) & 1u) && P.fx[e].backreaction)
            for (int c = 0; c < 3; c++) a[c] += P.fx[e].backreaction[c];
    for (int c = 0; c < 3; c++) {
        pos[c] += S.hdt*vel[c];
        vel[c] += S.dt*a[c];
        pos[c] += S.hdt*vel[c];
    }
    for (int e = 0; e < P.n_effects; e++) {
        double* rec = const_cast<double*>(P.fx[e].body);
        for (int c = 0; c < 3; c++) { rec[REBX_B200_BODY_X + c] = pos[c]; rec[REBX_B200_BODY_VX + c] = vel[c]; }
    }
    if (k >= 0 && k < P.st.n) {
        const_cast<double*>(P.st.x)[k] = pos[0]; const_cast<double*>(P.st.y)[k] = pos[1]; const_cast<double*>(P.st.z)[k] = pos[2];
        const_cast<double*>(P.st.vx)[k] = vel[0]; const_cast<double*>(P.st.vy)[k] = vel[1]; const_cast<double*>(P.st.vz)[k] = vel[2];
    }
}
__global__ void step_body_kernel(const __grid_constant__ rebx_b200_pass P, const StepArgs S, unsigned red_mask) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    step_body(P, S, red_mask);
}
// step_kernel's tail in the last CTA of a single-shard step (REBX_B200_PASS_APPLY_BACKREACTION): every other CTA read
// the body record when it started and stored its particles before it took its ticket, so the body may move now.
struct StepBodyTail {
    StepArgs S; unsigned red_mask;
    __device__ __forceinline__ void operator()(const rebx_b200_pass& P) const {
        if (P.reserved & REBX_B200_PASS_APPLY_BACKREACTION) step_body(P, S, red_mask);
    }
};

template <int V, bool SHARED, int K0, int K1, int K2, int K3>
__global__ void __launch_bounds__(kThreads, min_blocks<K0, K1, K2, K3>())
step_kernel(const __grid_constant__ rebx_b200_pass P, double* __restrict__ partials, const StepArgs S) {
    constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass || Traits<K3>::mass;
    constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad  || Traits<K3>::rad;

    __shared__ double s_body[REBX_B200_MAX_EFFECTS][REBX_B200_BODY_DOUBLES];
    for (int idx = threadIdx.x; idx < P.n_effects*REBX_B200_BODY_DOUBLES; idx += kThreads) {
        const int e = idx/REBX_B200_BODY_DOUBLES, j = idx - e*REBX_B200_BODY_DOUBLES;
        s_body[e][j] = P.fx[e].body[j];
    }
    __syncthreads();
    if (threadIdx.x < P.n_effects*3) {          // the body's own first half drift, same arithmetic as a particle's
        const int e = threadIdx.x/3, c = threadIdx.x - e*3;
        s_body[e][REBX_B200_BODY_X + c] += S.hdt*s_body[e][REBX_B200_BODY_VX + c];
    }
    __syncthreads();
    const long long b0 = (long long)s_body[0][0];
    // SHARED: every effect of the stack refers to the same central body, so position / velocity / mass are
    // read from record 0 for all of them and the compiler merges the common sub-expressions of the
    // stacked effects (relative position, r^2, the orbit elements of the damping trio).
    const Origin org0 = load_origin(s_body[0]);
    const Origin org1 = SHARED ? org0 : load_origin(s_body[1]);
    const Origin org2 = SHARED ? org0 : load_origin(s_body[2]);
    const Origin org3 = SHARED ? org0 : load_origin(s_body[3]);
    const Pre pre0 = prepare<K0>(P.fx[0], org0, s_body[0]);
    const Pre pre1 = prepare<K1>(P.fx[1], org1, s_body[1]);
    const Pre pre2 = prepare<K2>(P.fx[2], org2, s_body[2]);
    const Pre pre3 = prepare<K3>(P.fx[3], org3, s_body[3]);
    Vec3 red0 = {0., 0., 0.}, red1 = {0., 0., 0.}, red2 = {0., 0., 0.}, red3 = {0., 0., 0.};
    const double bmass = s_body[0][REBX_B200_BODY_M];

    double* const X = const_cast<double*>(P.st.x);  double* const Y = const_cast<double*>(P.st.y);  double* const Z = const_cast<double*>(P.st.z);
    double* const VX = const_cast<double*>(P.st.vx); double* const VY = const_cast<double*>(P.st.vy); double* const VZ = const_cast<double*>(P.st.vz);
    const int64_t n = P.st.n;
    const int64_t ntiles = (n + V - 1)/V;
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t t = (int64_t)blockIdx.x*kThreads + threadIdx.x; t < ntiles; t += stride) {
        const int64_t k = t*V;
        const bool full = (k + V <= n);
        double x[V], y[V], z[V], vx[V], vy[V], vz[V], m[V], r[V];
        ld<V>(P.st.x, k, full, x); ld<V>(P.st.y, k, full, y); ld<V>(P.st.z, k, full, z);
        ld<V>(P.st.vx, k, full, vx); ld<V>(P.st.vy, k, full, vy); ld<V>(P.st.vz, k, full, vz);
        if constexpr (MASS) { ld<V>(P.st.m, k, full, m); }
        if constexpr (RAD)  { ld<V>(P.st.r, k, full, r); }
        Params<K0, V> q0; Params<K1, V> q1; Params<K2, V> q2; Params<K3, V> q3;
        load_params<K0, V>(P.fx[0], k, full, q0);
        if constexpr (K1 != 0) load_params<K1, V>(P.fx[1], k, full, q1);
        if constexpr (K2 != 0) load_params<K2, V>(P.fx[2], k, full, q2);
        if constexpr (K3 != 0) load_params<K3, V>(P.fx[3], k, full, q3);
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (v > 0 && !full) break;
            const long long gi = P.st.index_base + k + v;
            if (gi == b0) continue;                                 // the central body: step_body_kernel
            Particle p;
            p.x = x[v] + S.hdt*vx[v]; p.y = y[v] + S.hdt*vy[v]; p.z = z[v] + S.hdt*vz[v];
            p.vx = vx[v]; p.vy = vy[v]; p.vz = vz[v];
            if constexpr (MASS) { p.m = m[v]; } else { p.m = 0.0; }
            if constexpr (RAD)  { p.r = r[v]; } else { p.r = 0.0; }
            Vec3 a = {0.0, 0.0, 0.0};
            if (bmass != 0.0) {                                     // central gravity (gravity_kernel's arithmetic)
                const double dx = p.x - s_body[0][REBX_B200_BODY_X+0], dy = p.y - s_body[0][REBX_B200_BODY_X+1], dz = p.z - s_body[0][REBX_B200_BODY_X+2];
                const double r2 = dx*dx + dy*dy + dz*dz + S.soft2;
                const double pref = -S.G*bmass/(r2*sqrt(r2));
                a.x += pref*dx; a.y += pref*dy; a.z += pref*dz;
            }
            unsigned unused = 0;
            apply<K0, V>(P.fx[0], org0, s_body[0], pre0, p, q0, v, a, red0, unused);
            if constexpr (K1 != 0) apply<K1, V>(P.fx[1], org1, s_body[1], pre1, p, q1, v, a, red1, unused);
            if constexpr (K2 != 0) apply<K2, V>(P.fx[2], org2, s_body[2], pre2, p, q2, v, a, red2, unused);
            if constexpr (K3 != 0) apply<K3, V>(P.fx[3], org3, s_body[3], pre3, p, q3, v, a, red3, unused);
            vx[v] = p.vx + S.dt*a.x; vy[v] = p.vy + S.dt*a.y; vz[v] = p.vz + S.dt*a.z;
            x[v] = p.x + S.hdt*vx[v]; y[v] = p.y + S.hdt*vy[v]; z[v] = p.z + S.hdt*vz[v];
        }
        st<V>(X, k, full, x); st<V>(Y, k, full, y); st<V>(Z, k, full, z);
        st<V>(VX, k, full, vx); st<V>(VY, k, full, vy); st<V>(VZ, k, full, vz);
    }
    constexpr unsigned RED_MASK = (Traits<K0>::red ? 1u : 0u) | (Traits<K1>::red ? 2u : 0u) | (Traits<K2>::red ? 4u : 0u) | (Traits<K3>::red ? 8u : 0u);
    finish_sums<kThreads, K0, K1, K2, K3>(red0, red1, red2, red3, P, partials, StepBodyTail{S, RED_MASK}, (P.reserved & REBX_B200_PASS_APPLY_BACKREACTION) != 0);
}

// The central body's own step (one thread): half drift, kick by the back-reaction sums of the stack
// (in stack order, like apply_backreaction), half drift; then every body record is refreshed.
// ---------------------------------------------------------------------------------------------
// pass_kernel_tma<K0..K3>: same sweep, but the particle state and parameter tables are moved
// HBM -> shared memory by the TMA engine (cp.async.bulk, 1-D bulk copies completing on an
// mbarrier) through a STAGES-deep ring per CTA.  One producer warp issues the copies; kTile
// compute threads (one particle each per stage) pull their operands out of shared memory.
// The copies need no registers and no resident warps while in flight, so memory-level
// parallelism no longer depends on occupancy of the register-heavy FP64 division code
// (round-1 ncu: 102 regs -> 16 warps/SM, issue-active 38 %, eligible warps 0.49/cycle).
// Requires 16-byte aligned arrays; the ragged tail (n % kTile) is handled with plain loads.
// ---------------------------------------------------------------------------------------------
constexpr int kTile       = 256;
constexpr int kTmaThreads = kTile + 32;
constexpr int kTmaWarps   = kTmaThreads/32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int K0, int K1, int K2, int K3>
struct TmaCfg {
    static constexpr bool VEL  = Traits<K0>::vel  || Traits<K1>::vel  || Traits<K2>::vel  || Traits<K3>::vel;
    static constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass || Traits<K3>::mass;
    static constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad  || Traits<K3>::rad;
    static constexpr int oV = 3, oM = oV + (VEL ? 3 : 0), oR = oM + (MASS ? 1 : 0), oA = oR + (RAD ? 1 : 0);
    static constexpr int oT0 = oA + 3, oT1 = oT0 + Traits<K0>::ntab, oT2 = oT1 + Traits<K1>::ntab, oT3 = oT2 + Traits<K2>::ntab;
    static constexpr int ND = oT3 + Traits<K3>::ntab;                     // double arrays per stage
    static constexpr int oI0 = 0, oI1 = oI0 + (Traits<K0>::itab ? 1 : 0), oI2 = oI1 + (Traits<K1>::itab ? 1 : 0),
                         oI3 = oI2 + (Traits<K2>::itab ? 1 : 0);
    static constexpr int NI = oI3 + (Traits<K3>::itab ? 1 : 0);           // int arrays per stage
    static constexpr int stage_bytes = ND*kTile*8 + NI*kTile*4;
    static constexpr int budget = 72*1024;                                // ~3 CTAs per SM
    static constexpr int STAGES = (budget/stage_bytes) >= 4 ? 4 : ((budget/stage_bytes) >= 2 ? (budget/stage_bytes) : 2);
    static constexpr int smem_bytes = STAGES*stage_bytes + 2*STAGES*8 + 16;
};

template <int K, int OT, int OI>
__device__ __forceinline__ void smem_params(const unsigned char* stage, int nd, int tid, Params<K, 1>& q) {
#pragma unroll
    for (int j = 0; j < Traits<K>::ntab; j++) q.t[j][0] = reinterpret_cast<const double*>(stage)[(OT + j)*kTile + tid];
    if constexpr (Traits<K>::itab) q.it[0] = reinterpret_cast<const int32_t*>(stage + (size_t)nd*kTile*8)[OI*kTile + tid];
}

template <int K0, int K1, int K2, int K3>
__global__ void __launch_bounds__(kTmaThreads)
pass_kernel_tma(const __grid_constant__ rebx_b200_pass P, double* __restrict__ partials) {
    using C = TmaCfg<K0, K1, K2, K3>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double s_body[REBX_B200_MAX_EFFECTS][REBX_B200_BODY_DOUBLES];
    __shared__ double s_red[kTmaWarps][3];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES*C::stage_bytes);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    for (int idx = tid; idx < P.n_effects*REBX_B200_BODY_DOUBLES; idx += kTmaThreads) {
        const int e = idx/REBX_B200_BODY_DOUBLES, j = idx - e*REBX_B200_BODY_DOUBLES;
        s_body[e][j] = P.fx[e].body[j];
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kTile/32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t n = P.st.n;
    const int64_t nfull = n/kTile;                      // full tiles go through the TMA ring
    Vec3 red0 = {0., 0., 0.}, red1 = {0., 0., 0.}, red2 = {0., 0., 0.}, red3 = {0., 0., 0.};

    if (tid >= kTile) {
        // ---------------- producer warp: one elected lane feeds the ring --------------------
        if (tid == kTile) {
            const void* dsrc[C::ND];
            dsrc[0] = P.st.x; dsrc[1] = P.st.y; dsrc[2] = P.st.z;
            if constexpr (C::VEL)  { dsrc[C::oV] = P.st.vx; dsrc[C::oV + 1] = P.st.vy; dsrc[C::oV + 2] = P.st.vz; }
            if constexpr (C::MASS) { dsrc[C::oM] = P.st.m; }
            if constexpr (C::RAD)  { dsrc[C::oR] = P.st.r; }
            dsrc[C::oA] = P.st.ax; dsrc[C::oA + 1] = P.st.ay; dsrc[C::oA + 2] = P.st.az;
#pragma unroll
            for (int j = 0; j < Traits<K0>::ntab; j++) dsrc[C::oT0 + j] = P.fx[0].tab[j];
#pragma unroll
            for (int j = 0; j < Traits<K1>::ntab; j++) dsrc[C::oT1 + j] = P.fx[1].tab[j];
#pragma unroll
            for (int j = 0; j < Traits<K2>::ntab; j++) dsrc[C::oT2 + j] = P.fx[2].tab[j];
#pragma unroll
            for (int j = 0; j < Traits<K3>::ntab; j++) dsrc[C::oT3 + j] = P.fx[3].tab[j];
            const void* isrc[C::NI > 0 ? C::NI : 1];
            if constexpr (Traits<K0>::itab) isrc[C::oI0] = P.fx[0].itab;
            if constexpr (Traits<K1>::itab) isrc[C::oI1] = P.fx[1].itab;
            if constexpr (Traits<K2>::itab) isrc[C::oI2] = P.fx[2].itab;
            if constexpr (Traits<K3>::itab) isrc[C::oI3] = P.fx[3].itab;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < nfull; tile += gridDim.x, it++) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it/STAGES) - 1) & 1);
                unsigned char* stage = smem + (size_t)s*C::stage_bytes;
                mbar_expect_tx(&full_bar[s], C::stage_bytes);
                const int64_t k0 = tile*kTile;
#pragma unroll
                for (int j = 0; j < C::ND; j++)
                    bulk_g2s(stage + (size_t)j*kTile*8, static_cast<const double*>(dsrc[j]) + k0, kTile*8, &full_bar[s]);
#pragma unroll
                for (int j = 0; j < C::NI; j++)
                    bulk_g2s(stage + (size_t)C::ND*kTile*8 + (size_t)j*kTile*4, static_cast<const int32_t*>(isrc[j]) + k0, kTile*4, &full_bar[s]);
            }
        }
    } else {
        // ---------------- compute threads: one particle per thread per stage -------------------
        const long long b0 = (long long)s_body[0][0];
        const long long b1 = (K1 != 0) ? (long long)s_body[1][0] : -1;
        const long long b2 = (K2 != 0) ? (long long)s_body[2][0] : -1;
        const long long b3 = (K3 != 0) ? (long long)s_body[3][0] : -1;
        const Origin org0 = load_origin(s_body[0]), org1 = load_origin(s_body[1]), org2 = load_origin(s_body[2]), org3 = load_origin(s_body[3]);
        const Pre pre0 = prepare<K0>(P.fx[0], org0, s_body[0]);
        const Pre pre1 = prepare<K1>(P.fx[1], org1, s_body[1]);
        const Pre pre2 = prepare<K2>(P.fx[2], org2, s_body[2]);
        const Pre pre3 = prepare<K3>(P.fx[3], org3, s_body[3]);
        auto one = [&](const Particle& p, Vec3& a, long long gi, const Params<K0, 1>& q0, const Params<K1, 1>& q1,
                       const Params<K2, 1>& q2, const Params<K3, 1>& q3) {
            unsigned unused = 0;
            if (gi != b0) apply<K0, 1>(P.fx[0], org0, s_body[0], pre0, p, q0, 0, a, red0, unused);
            if constexpr (K1 != 0) { if (gi != b1) apply<K1, 1>(P.fx[1], org1, s_body[1], pre1, p, q1, 0, a, red1, unused); }
            if constexpr (K2 != 0) { if (gi != b2) apply<K2, 1>(P.fx[2], org2, s_body[2], pre2, p, q2, 0, a, red2, unused); }
            if constexpr (K3 != 0) { if (gi != b3) apply<K3, 1>(P.fx[3], org3, s_body[3], pre3, p, q3, 0, a, red3, unused); }
        };
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < nfull; tile += gridDim.x, it++) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it/STAGES) & 1);
            const unsigned char* stage = smem + (size_t)s*C::stage_bytes;
            const double* sd = reinterpret_cast<const double*>(stage);
            Particle p;
            p.x = sd[0*kTile + tid]; p.y = sd[1*kTile + tid]; p.z = sd[2*kTile + tid];
            if constexpr (C::VEL)  { p.vx = sd[(C::oV + 0)*kTile + tid]; p.vy = sd[(C::oV + 1)*kTile + tid]; p.vz = sd[(C::oV + 2)*kTile + tid]; }
            else { p.vx = p.vy = p.vz = 0.0; }
            if constexpr (C::MASS) { p.m = sd[C::oM*kTile + tid]; } else { p.m = 0.0; }
            if constexpr (C::RAD)  { p.r = sd[C::oR*kTile + tid]; } else { p.r = 0.0; }
            Vec3 a = {sd[(C::oA + 0)*kTile + tid], sd[(C::oA + 1)*kTile + tid], sd[(C::oA + 2)*kTile + tid]};
            Params<K0, 1> q0; Params<K1, 1> q1; Params<K2, 1> q2; Params<K3, 1> q3;
            smem_params<K0, C::oT0, C::oI0>(stage, C::ND, tid, q0);
            if constexpr (K1 != 0) smem_params<K1, C::oT1, C::oI1>(stage, C::ND, tid, q1);
            if constexpr (K2 != 0) smem_params<K2, C::oT2, C::oI2>(stage, C::ND, tid, q2);
            if constexpr (K3 != 0) smem_params<K3, C::oT3, C::oI3>(stage, C::ND, tid, q3);
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);       // operands are in registers: release the stage
            const int64_t k = tile*kTile + tid;
            one(p, a, P.st.index_base + k, q0, q1, q2, q3);
            __stcs(P.st.ax + k, a.x); __stcs(P.st.ay + k, a.y); __stcs(P.st.az + k, a.z);
        }
        // ragged tail: plain loads, spread over the CTAs
        const int64_t rem0 = nfull*kTile;
        for (int64_t k = rem0 + (int64_t)blockIdx.x*kTile + tid; k < n; k += (int64_t)gridDim.x*kTile) {
            Particle p;
            p.x = __ldcs(P.st.x + k); p.y = __ldcs(P.st.y + k); p.z = __ldcs(P.st.z + k);
            if constexpr (C::VEL)  { p.vx = __ldcs(P.st.vx + k); p.vy = __ldcs(P.st.vy + k); p.vz = __ldcs(P.st.vz + k); }
            else { p.vx = p.vy = p.vz = 0.0; }
            if constexpr (C::MASS) { p.m = __ldcs(P.st.m + k); } else { p.m = 0.0; }
            if constexpr (C::RAD)  { p.r = __ldcs(P.st.r + k); } else { p.r = 0.0; }
            Vec3 a = {__ldcs(P.st.ax + k), __ldcs(P.st.ay + k), __ldcs(P.st.az + k)};
            Params<K0, 1> q0; Params<K1, 1> q1; Params<K2, 1> q2; Params<K3, 1> q3;
            load_params<K0, 1>(P.fx[0], k, true, q0);
            if constexpr (K1 != 0) load_params<K1, 1>(P.fx[1], k, true, q1);
            if constexpr (K2 != 0) load_params<K2, 1>(P.fx[2], k, true, q2);
            if constexpr (K3 != 0) load_params<K3, 1>(P.fx[3], k, true, q3);
            one(p, a, P.st.index_base + k, q0, q1, q2, q3);
            __stcs(P.st.ax + k, a.x); __stcs(P.st.ay + k, a.y); __stcs(P.st.az + k, a.z);
        }
    }
    reduce_effect<K0, kTmaWarps>(red0, 0, s_red, partials);
    reduce_effect<K1, kTmaWarps>(red1, 1, s_red, partials);
    reduce_effect<K2, kTmaWarps>(red2, 2, s_red, partials);
    reduce_effect<K3, kTmaWarps>(red3, 3, s_red, partials);
}

// Sums the per-CTA partials in a fixed order and writes fx.backreaction (3 doubles each).
__global__ void __launch_bounds__(kThreads)
finalize_kernel(const __grid_constant__ rebx_b200_pass P, const double* __restrict__ partials, int nblocks, unsigned red_mask) {
    __shared__ double s[kThreads];
    for (int e = 0; e < P.n_effects; e++) {
        if (!((red_mask >> e) & 1u) || P.fx[e].backreaction == nullptr) continue;
        for (int c = 0; c < 3; c++) {
            double acc = 0.0;
            for (int b = threadIdx.x; b < nblocks; b += kThreads)
                acc += partials[((size_t)b*REBX_B200_MAX_EFFECTS + e)*3 + c];
            s[threadIdx.x] = acc;
            __syncthreads();
            for (int w = kThreads/2; w > 0; w >>= 1) {
                if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
                __syncthreads();
            }
            if (threadIdx.x == 0) P.fx[e].backreaction[c] = s[0];
            __syncthreads();
        }
    }
}

__global__ void apply_backreaction_kernel(const __grid_constant__ rebx_b200_pass P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int e = 0; e < P.n_effects; e++) {
        const rebx_b200_effect& fx = P.fx[e];
        if (fx.backreaction == nullptr) continue;
        const long long bi = (long long)fx.body[REBX_B200_BODY_INDEX];
        const long long k = bi - P.st.index_base;
        if (bi < 0 || k < 0 || k >= P.st.n) continue;
        P.st.ax[k] += fx.backreaction[0];
        P.st.ay[k] += fx.backreaction[1];
        P.st.az[k] += fx.backreaction[2];
    }
}

__global__ void __launch_bounds__(kThreads)
add_uniform_kernel(double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az, int64_t n, const double* __restrict__ delta) {
    const double dx = delta[0], dy = delta[1], dz = delta[2];
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        ax[k] += dx; ay[k] += dy; az[k] += dz;
    }
}

// ---- resident stepping (callers either side of the force path) ------------------------------------
__global__ void __launch_bounds__(kThreads)
drift_kernel(double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
             const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz, int64_t n, double dt) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        x[k] += dt*vx[k]; y[k] += dt*vy[k]; z[k] += dt*vz[k];
    }
}

__global__ void __launch_bounds__(kThreads)
gravity_kernel(rebx_b200_state st, const double* __restrict__ bodies, int nbodies, double G, double soft2) {
    __shared__ double s_b[8][REBX_B200_BODY_DOUBLES];
    for (int idx = threadIdx.x; idx < nbodies*REBX_B200_BODY_DOUBLES; idx += kThreads)
        s_b[idx/REBX_B200_BODY_DOUBLES][idx % REBX_B200_BODY_DOUBLES] = bodies[idx];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < st.n; k += stride) {
        const double x = st.x[k], y = st.y[k], z = st.z[k];
        double ax = 0.0, ay = 0.0, az = 0.0;
        const long long gi = st.index_base + k;
        for (int b = 0; b < nbodies; b++) {
            const double mb = s_b[b][REBX_B200_BODY_M];
            if ((long long)s_b[b][REBX_B200_BODY_INDEX] == gi || mb == 0.0) continue;
            const double dx = x - s_b[b][REBX_B200_BODY_X+0], dy = y - s_b[b][REBX_B200_BODY_X+1], dz = z - s_b[b][REBX_B200_BODY_X+2];
            const double r2 = dx*dx + dy*dy + dz*dz + soft2;
            const double pref = -G*mb/(r2*sqrt(r2));
            ax += pref*dx; ay += pref*dy; az += pref*dz;
        }
        st.ax[k] = ax; st.ay[k] = ay; st.az[k] = az;
    }
}

__global__ void gather_bodies_kernel(rebx_b200_state st, double* __restrict__ bodies, int nbodies) {
    const int b = threadIdx.x;
    if (b >= nbodies) return;
    double* rec = bodies + (size_t)b*REBX_B200_BODY_DOUBLES;
    const long long k = (long long)rec[REBX_B200_BODY_INDEX] - st.index_base;
    if (k < 0 || k >= st.n) return;
    rec[REBX_B200_BODY_X+0] = st.x[k]; rec[REBX_B200_BODY_X+1] = st.y[k]; rec[REBX_B200_BODY_X+2] = st.z[k];
    if (st.vx) { rec[REBX_B200_BODY_VX+0] = st.vx[k]; rec[REBX_B200_BODY_VX+1] = st.vy[k]; rec[REBX_B200_BODY_VX+2] = st.vz[k]; }
    if (st.m) rec[REBX_B200_BODY_M] = st.m[k];
}

// ---- diagnostic potentials: one reduction over the shard ---------------------------------------------
__global__ void __launch_bounds__(kThreads)
potential_kernel(rebx_b200_state st, const __grid_constant__ rebx_b200_effect fx, double* __restrict__ partials) {
    __shared__ double s_body[REBX_B200_BODY_DOUBLES];
    __shared__ double s_sum[kThreads];
    if (threadIdx.x < REBX_B200_BODY_DOUBLES) s_body[threadIdx.x] = fx.body[threadIdx.x];
    __syncthreads();
    const long long bi = (long long)s_body[REBX_B200_BODY_INDEX];
    const double bx = s_body[REBX_B200_BODY_X], by = s_body[REBX_B200_BODY_X+1], bz = s_body[REBX_B200_BODY_X+2], bm = s_body[REBX_B200_BODY_M];
    const double G = fx.scal[0];
    double H = 0.0;
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k == bi) continue;
        const double dx = st.x[k] - bx, dy = st.y[k] - by, dz = st.z[k] - bz;
        const double r2 = dx*dx + dy*dy + dz*dz;
        const double mj = st.m[k];
        if (fx.kind == REBX_B200_GR_POTENTIAL) {
            const double mu = G*bm;
            const double prefac = 3.*mu*mu/fx.scal[1];
            H -= prefac*mj/r2;
        } else if (fx.kind == REBX_B200_HARMONICS) {
            const double r = sqrt(r2);
            const double dw = s_body[REBX_B200_BODY_W]*dx + s_body[REBX_B200_BODY_W+1]*dy + s_body[REBX_B200_BODY_W+2]*dz;
            const double costheta = dw/r;
            const double c2 = costheta*costheta;
            const double J2 = s_body[REBX_B200_BODY_J2], J4 = s_body[REBX_B200_BODY_J4], R = s_body[REBX_B200_BODY_REQ];
            {
                const double f1 = G*bm*mj*J2*R*R/r2/r;
                const double f2 = 1.0/2.0*(3.0*c2 - 1.0);
                H += f1*f2;
            }
            if (J4 != 0.0) {
                const double f1 = G*bm*mj*J4*R*R*R*R/r2/r2/r;
                const double f2 = 1.0/8.0*(35.0*c2*c2 - 30.0*c2 + 3.0);
                H += f1*f2;
            }
        } else {    // CENTRAL_FORCE
            const double A = s_body[REBX_B200_BODY_J2], gamma = s_body[REBX_B200_BODY_J4];
            if (fabs(gamma + 1.) < 2.220446049250313e-16) H -= mj*A*log(sqrt(r2));
            else H -= mj*A*pow(r2, (gamma + 1.)/2.)/(gamma + 1.);
        }
    }
    s_sum[threadIdx.x] = H;
    __syncthreads();
    for (int w = kThreads/2; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_sum[threadIdx.x] += s_sum[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = s_sum[0];
}

__global__ void __launch_bounds__(kThreads)
potential_finalize_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
    __shared__ double s_sum[kThreads];
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += kThreads) acc += partials[b];
    s_sum[threadIdx.x] = acc;
    __syncthreads();
    for (int w = kThreads/2; w > 0; w >>= 1) {
        if (threadIdx.x < w) s_sum[threadIdx.x] += s_sum[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s_sum[0];
}

// ---- integrate_force schemes: element-wise stage updates --------------------------------------------
struct V3 { double *x, *y, *z; };
struct CV3 { const double *x, *y, *z; };

// out = v + s*a
__global__ void __launch_bounds__(kThreads) stage_axpy_kernel(V3 out, CV3 v, CV3 a, double s, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        out.x[k] = v.x[k] + s*a.x[k]; out.y[k] = v.y[k] + s*a.y[k]; out.z[k] = v.z[k] + s*a.z[k];
    }
}
// RK4 stage 4 set-up (integrator_rk4.c:73-80): v2 = v + dt*a3 ; a3 += a2 ; a2 = 0
__global__ void __launch_bounds__(kThreads) rk4_stage4_kernel(V3 v2, V3 a2, V3 a3, CV3 v, double dt, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        v2.x[k] = v.x[k] + dt*a3.x[k]; v2.y[k] = v.y[k] + dt*a3.y[k]; v2.z[k] = v.z[k] + dt*a3.z[k];
        a3.x[k] += a2.x[k]; a3.y[k] += a2.y[k]; a3.z[k] += a2.z[k];
        a2.x[k] = 0.; a2.y[k] = 0.; a2.z[k] = 0.;
    }
}
// v += w*(a1 + a4 + 2.*a23)   (integrator_rk4.c:84-89)
__global__ void __launch_bounds__(kThreads) rk4_final_kernel(V3 v, CV3 a1, CV3 a4, CV3 a23, double w, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        v.x[k] += w*(a1.x[k] + a4.x[k] + 2.*a23.x[k]);
        v.y[k] += w*(a1.y[k] + a4.y[k] + 2.*a23.y[k]);
        v.z[k] += w*(a1.z[k] + a4.z[k] + 2.*a23.z[k]);
    }
}
// v += b1*a1 + b2*a2   (integrator_rk2.c:60-65)
__global__ void __launch_bounds__(kThreads) rk2_final_kernel(V3 v, CV3 a1, CV3 a2, double b1, double b2, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        v.x[k] += b1*a1.x[k] + b2*a2.x[k]; v.y[k] += b1*a1.y[k] + b2*a2.y[k]; v.z[k] += b1*a1.z[k] + b2*a2.z[k];
    }
}
// implicit midpoint: v_avg = 0.5*(v_orig + v_final), a_avg = 0  (integrator_implicit_midpoint.c:33-49)
__global__ void __launch_bounds__(kThreads) im_average_kernel(V3 vavg, V3 aavg, CV3 vorig, CV3 vfinal, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        vavg.x[k] = 0.5*(vorig.x[k] + vfinal.x[k]); vavg.y[k] = 0.5*(vorig.y[k] + vfinal.y[k]); vavg.z[k] = 0.5*(vorig.z[k] + vfinal.z[k]);
        aavg.x[k] = 0.; aavg.y[k] = 0.; aavg.z[k] = 0.;
    }
}
// convergence sums of compare() (integrator_implicit_midpoint.c:51-66): sum |v1-v2|^2 and sum |v1|^2
__global__ void __launch_bounds__(kThreads) im_compare_kernel(CV3 v1, CV3 v2, int64_t n, double* __restrict__ sums) {
    __shared__ double s_a[kThreads], s_b[kThreads];
    double d2 = 0.0, t2 = 0.0;
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        const double dx = v1.x[k] - v2.x[k], dy = v1.y[k] - v2.y[k], dz = v1.z[k] - v2.z[k];
        d2 += dx*dx + dy*dy + dz*dz;
        t2 += v1.x[k]*v1.x[k] + v1.y[k]*v1.y[k] + v1.z[k]*v1.z[k];
    }
    s_a[threadIdx.x] = d2; s_b[threadIdx.x] = t2;
    __syncthreads();
    for (int w = kThreads/2; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s_a[threadIdx.x] += s_a[threadIdx.x + w]; s_b[threadIdx.x] += s_b[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(&sums[0], s_a[0]); atomicAdd(&sums[1], s_b[0]); }
}

// ---- AoS <-> SoA staging (host-facing path; sits behind PCIe, so kept simple) ----------------
__global__ void __launch_bounds__(kThreads)
unpack_aos_kernel(const unsigned char* __restrict__ aos, rebx_b200_aos_layout L,
                  double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                  double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                  double* __restrict__ m, double* __restrict__ r,
                  double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        const unsigned char* rec = aos + (size_t)k*L.stride;
        const double* px = reinterpret_cast<const double*>(rec + L.off_x);
        const double* pv = reinterpret_cast<const double*>(rec + L.off_v);
        const double* pa = reinterpret_cast<const double*>(rec + L.off_a);
        x[k] = px[0]; y[k] = px[1]; z[k] = px[2];
        if (vx) { vx[k] = pv[0]; vy[k] = pv[1]; vz[k] = pv[2]; }
        ax[k] = pa[0]; ay[k] = pa[1]; az[k] = pa[2];
        if (m) m[k] = *reinterpret_cast<const double*>(rec + L.off_m);
        if (r) r[k] = *reinterpret_cast<const double*>(rec + L.off_r);
    }
}

__global__ void __launch_bounds__(kThreads)
pack_acc_aos_kernel(unsigned char* __restrict__ aos, rebx_b200_aos_layout L,
                    const double* __restrict__ ax, const double* __restrict__ ay, const double* __restrict__ az, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < n; k += stride) {
        double* pa = reinterpret_cast<double*>(aos + (size_t)k*L.stride + L.off_a);
        pa[0] = ax[k]; pa[1] = ay[k]; pa[2] = az[k];
    }
}

}  // namespace rebxb
#include "kernels_jacobi.cuh"
#include "kernels_gr.cuh"
#include "kernels_kepler.cuh"
namespace rebxb {

// ---- launch plumbing ----------------------------------------------------------------------
typedef void (*pass_fn)(const rebx_b200_pass, double*);
typedef void (*step_fn)(const rebx_b200_pass, double*, const StepArgs);
struct Combo { int k[4]; pass_fn fn[2][2]; bool red[4]; pass_fn tma; int tma_smem; step_fn step[2][2]; int threads; };   // [shared body][V-1]

#define REBXB_COMBO(a, b, c, d) { {a, b, c, d}, { { pass_kernel<1, false, a, b, c, d>, pass_kernel<2, false, a, b, c, d> }, \
                                                   { pass_kernel<1, (b) != 0, a, b, c, d>, pass_kernel<2, (b) != 0, a, b, c, d> } }, \
                                  { Traits<a>::red, Traits<b>::red, Traits<c>::red, Traits<d>::red }, \
                                  pass_kernel_tma<a, b, c, d>, TmaCfg<a, b, c, d>::smem_bytes, \
                                  { { step_kernel<1, false, a, b, c, d>, step_kernel<2, false, a, b, c, d> }, \
                                    { step_kernel<1, (b) != 0, a, b, c, d>, step_kernel<2, (b) != 0, a, b, c, d> } }, \
                                  pass_threads<a, b, c, d>() },
static Combo kCombos[] = {
    // every effect alone (any other stack falls back to consecutive single sweeps, still in list order)
    REBXB_COMBO(1,0,0,0) REBXB_COMBO(2,0,0,0) REBXB_COMBO(3,0,0,0) REBXB_COMBO(4,0,0,0)
    REBXB_COMBO(5,0,0,0) REBXB_COMBO(6,0,0,0) REBXB_COMBO(7,0,0,0) REBXB_COMBO(8,0,0,0) REBXB_COMBO(9,0,0,0) REBXB_COMBO(10,0,0,0) REBXB_COMBO(11,0,0,0)
    // config 3: J2/J4 (+) gr_potential, either LIFO order
    REBXB_COMBO(2,3,0,0) REBXB_COMBO(3,2,0,0)
    // config 4: yarkovsky (+) gr_potential
    REBXB_COMBO(4,3,0,0) REBXB_COMBO(3,4,0,0)
    // radiation (+) gr_potential
    REBXB_COMBO(1,3,0,0) REBXB_COMBO(3,1,0,0)
    // config 5: damping trio in the two natural LIFO orders
    REBXB_COMBO(7,6,5,0) REBXB_COMBO(5,6,7,0)
};
constexpr int kNumCombos = sizeof(kCombos)/sizeof(kCombos[0]);

static const Combo* find_combo(const rebx_b200_effect* fx, int count) {
    for (int i = 0; i < kNumCombos; i++) {
        bool ok = true;
        for (int j = 0; j < 4; j++) {
            const int want = j < count ? fx[j].kind : 0;
            if (kCombos[i].k[j] != want) { ok = false; break; }
        }
        if (ok) return &kCombos[i];
    }
    return nullptr;
}

// kernels_tol.cu
bool tol_launch(const rebx_b200_pass& sub, int take, double* workspace, cudaStream_t s, int* grid_out);
bool tol_jacobi_launch(const rebx_b200_pass& P, unsigned ntiles, const double* tile_prefix, double* tvec, double* tile_tsums, const double* carry7,
                       const void* grid, const double* status, const double* mref, cudaStream_t s);
const void* tol_jacobi_fused_kernel(const rebx_b200_pass& P);
bool tol_wh_kick_drift_launch(const rebx_b200_pass& P, double G, double dt_kick, double dt_drift, cudaStream_t s);
bool tol_leapfrog_sweep_launch(const rebx_b200_pass& P, double dt, double G, double soft2, cudaStream_t s);

static bool use_tma() {
    // opt-in (REBX_B200_KERNEL=tma): measured slower than the occupancy-tuned plain sweep on B200
    // (radiation 10^7: 0.179 ms vs 0.171 ms; FP64-heavy sweeps 5-30 % slower), see DESIGN.md
    static const int mode = [] { const char* e = getenv("REBX_B200_KERNEL"); return (e && e[0] == 't') ? 1 : 0; }();
    return mode != 0;
}

static int check_pass(const rebx_b200_pass& P) {
    if (P.n_effects < 1 || P.n_effects > REBX_B200_MAX_EFFECTS) return cudaErrorInvalidValue;
    if (P.st.n < 0) return cudaErrorInvalidValue;
    if (!P.st.x || !P.st.y || !P.st.z || !P.st.ax || !P.st.ay || !P.st.az) return cudaErrorInvalidValue;
    for (int e = 0; e < P.n_effects; e++) {
        const rebx_b200_effect& fx = P.fx[e];
        if (fx.kind <= 0 || fx.kind >= REBX_B200_KIND_COUNT || !fx.body) return cudaErrorInvalidValue;
        bool vel = false, mass = false, rad = false; int ntab = 0; bool itab = false;
        switch (fx.kind) {
#define REBXB_CASE(K) case K: vel = Traits<K>::vel; mass = Traits<K>::mass; rad = Traits<K>::rad; ntab = Traits<K>::ntab; itab = Traits<K>::itab; break;
            REBXB_CASE(1) REBXB_CASE(2) REBXB_CASE(3) REBXB_CASE(4) REBXB_CASE(5) REBXB_CASE(6) REBXB_CASE(7) REBXB_CASE(8) REBXB_CASE(9) REBXB_CASE(10) REBXB_CASE(11)
#undef REBXB_CASE
        }
        if (vel && (!P.st.vx || !P.st.vy || !P.st.vz)) return cudaErrorInvalidValue;
        if (mass && !P.st.m) return cudaErrorInvalidValue;
        if (rad && !P.st.r) return cudaErrorInvalidValue;
        for (int j = 0; j < ntab; j++) if (!fx.tab[j]) return cudaErrorInvalidValue;
        if (itab && !fx.itab) return cudaErrorInvalidValue;
    }
    return 0;
}

}  // namespace rebxb

using namespace rebxb;

extern "C" {

int rebx_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* rebx_b200_error_string(int code) {
    return cudaGetErrorString(static_cast<cudaError_t>(code));
}

size_t rebx_b200_workspace_bytes(void) {
    return (kTicketOffset + 8)*sizeof(double);     // CTA partials + a few scalars + the CTA ticket counter (pass_common.cuh)
}

int rebx_b200_sm_count(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); return 0; }
    return sms;
}

int rebx_b200_launch_pass(const rebx_b200_pass* pass, void* workspace, void* stream, int* launches) {
    if (!pass || !workspace) return cudaErrorInvalidValue;
    const int bad = check_pass(*pass);
    if (bad) return bad;
    if (pass->st.n == 0) {
        // nothing to stream, but back-reaction outputs must still be defined (zero)
        for (int e = 0; e < pass->n_effects; e++)
            if (pass->fx[e].backreaction) {
                cudaError_t err = cudaMemsetAsync(pass->fx[e].backreaction, 0, 3*sizeof(double), static_cast<cudaStream_t>(stream));
                if (err != cudaSuccess) return err;
            }
        return 0;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool aligned = pass_is_vectorizable(*pass);
    int done = 0;
    while (done < pass->n_effects) {
        int take = pass->n_effects - done;
        const Combo* combo = nullptr;
        for (; take >= 1; take--) {
            combo = find_combo(pass->fx + done, take);
            if (combo) break;
        }
        if (!combo) return cudaErrorInvalidValue;
        rebx_b200_pass sub;
        sub.st = pass->st;
        sub.n_effects = take;
        sub.reserved = pass->reserved;
        for (int j = 0; j < REBX_B200_MAX_EFFECTS; j++) {
            if (j < take) sub.fx[j] = pass->fx[done + j];
            else { sub.fx[j] = rebx_b200_effect(); sub.fx[j].kind = 0; }
        }
        int grid;
        cudaError_t err;
        // Two particles per thread (double2 access) pay off for the light and the J2/J4 sweeps; the
        // register-heavy ones (yarkovsky, damping trio) run 5-20 % faster with one particle per
        // thread (measured: asteroid 0.098 vs 0.104 ms, planetesimal 0.135 vs 0.171 ms).
        bool heavy = false;
        for (int j = 0; j < take; j++) if (sub.fx[j].kind >= REBX_B200_YARKOVSKY) heavy = true;
        const int vsel = (aligned && !heavy) ? 1 : 0;
        const bool want_tol = (pass->reserved & REBX_B200_PASS_TOLERANCE) && (pass->reserved & REBX_B200_PASS_SHARED_BODY);
        bool separate_finalize = false;      // the plain and the tolerance sweeps finish their own sums (finalize_by_last_cta)
        if (want_tol && tol_launch(sub, take, static_cast<double*>(workspace), s, &grid)) {
            // tolerance-mode sweep launched (kernels_tol.cu)
        } else if (aligned && use_tma() && sub.st.n >= 4*kTile) {
            separate_finalize = true;
            // TMA-fed ring (needs 16-byte aligned arrays)
            pass_fn fn = combo->tma;
            err = cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, combo->tma_smem);
            if (err != cudaSuccess) return err;
            grid = grid_for(reinterpret_cast<const void*>(fn), sub.st.n, kTmaThreads, combo->tma_smem, kTile);
            fn<<<grid, kTmaThreads, combo->tma_smem, s>>>(sub, static_cast<double*>(workspace));
        } else {
            const int V = vsel ? 2 : 1;
            pass_fn fn = combo->fn[(pass->reserved & REBX_B200_PASS_SHARED_BODY) ? 1 : 0][vsel];
            const int64_t tiles = (sub.st.n + V - 1)/V;
            grid = grid_for(reinterpret_cast<const void*>(fn), tiles, combo->threads, 0, combo->threads);
            fn<<<grid, combo->threads, 0, s>>>(sub, static_cast<double*>(workspace));
        }
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        if (launches) (*launches)++;
        unsigned red_mask = 0;
        bool any = false;
        for (int j = 0; j < take; j++) if (combo->red[j]) { red_mask |= 1u << j; if (sub.fx[j].backreaction) any = true; }
        if (any && separate_finalize) {
            finalize_kernel<<<1, kThreads, 0, s>>>(sub, static_cast<const double*>(workspace), grid, red_mask);
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
            if (launches) (*launches)++;
            if (pass->reserved & REBX_B200_PASS_APPLY_BACKREACTION) {
                apply_backreaction_kernel<<<1, 32, 0, s>>>(sub);
                if (launches) (*launches)++;
            }
        }
        done += take;
    }
    return 0;
}

int rebx_b200_apply_backreaction(const rebx_b200_pass* pass, void* stream, int* launches) {
    if (!pass) return cudaErrorInvalidValue;
    if (pass->reserved & REBX_B200_PASS_APPLY_BACKREACTION) return 0;       // the sweep applied the sums itself
    bool any = false;
    for (int e = 0; e < pass->n_effects; e++) if (pass->fx[e].backreaction) any = true;
    if (!any) return 0;
    apply_backreaction_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*pass);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

// Scratch layout shared by the centre-of-mass stages (all sized by the shard's own n).
struct JacobiScratch {
    double *rows7, *rows3, *totals, *tvec, *mref;
    MassGrid* grid;
    unsigned* unsafe;
    unsigned long long* first;
    int64_t ntiles;
    JacobiScratch(void* scratch, int64_t n) {
        ntiles = (n + kScanThreads - 1)/kScanThreads;
        rows7 = static_cast<double*>(scratch);
        rows3 = rows7 + (size_t)(ntiles + 1)*7;
        totals = rows3 + (size_t)(ntiles + 1)*3;         // [0..7] totals8, [8..10] tsum3, [16..22] MassGrid, [24] unsafe flag, [25] first massive index
        grid = reinterpret_cast<MassGrid*>(totals + 16);
        unsafe = reinterpret_cast<unsigned*>(totals + 24);
        first = reinterpret_cast<unsigned long long*>(totals + 25);
        tvec = totals + 32;
        mref = tvec + (size_t)3*(n + 2);
    }
};
static_assert(sizeof(MassGrid) <= 8*sizeof(double), "MassGrid must fit its scratch slot");

size_t rebx_b200_scan_scratch_bytes(int64_t n) {
    const int64_t ntiles = (n + kScanThreads - 1)/kScanThreads + 1;
    return (size_t)(ntiles*10 + 32 + 4*(n + 2) + 2*ntiles)*sizeof(double);     // the last 2*ntiles: per-warp flags of the fused pipeline
}

// Stage 1 of both centre-of-mass frames: quantised mass + moment tile sums, their scan, the status / literal
// fallback kernel.  `backward`: also provide the interior masses (JACOBI).
static int com_moments_stage(const rebx_b200_state* st, void* scratch, double* totals8, const double* lead_mass, int flags,
                             bool backward, cudaStream_t s, int* launches) {
    JacobiScratch J(scratch, st->n);
    cudaError_t err = cudaMemsetAsync(J.unsafe, 0, sizeof(double), s);                 // flag = 0
    if (err == cudaSuccess) err = cudaMemsetAsync(J.first, 0xff, sizeof(unsigned long long), s);   // first massive = none
    if (err != cudaSuccess) return err;
    moments_tile_kernel<<<(unsigned)J.ntiles, kScanThreads, 0, s>>>(*st, J.rows7, lead_mass, J.grid, J.unsafe, J.first);
    scan_tiles_kernel<7, false><<<7, kTileScanThreads, 0, s>>>(J.rows7, J.ntiles, J.totals);
    mass_status_kernel<<<1, 32, 0, s>>>(*st, J.grid, J.unsafe, J.first, J.totals, (flags & REBX_B200_COM_WHOLE) ? J.mref : nullptr, backward ? 1 : 0, totals8);
    err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += 3;
    return err;
}

int rebx_b200_mass_moments(const rebx_b200_state* st, double* totals8, void* scratch, const double* lead_mass, int flags, void* stream, int* launches) {
    if (!st || !totals8 || !scratch || !st->m || !st->vx) return cudaErrorInvalidValue;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (st->n == 0) return cudaMemsetAsync(totals8, 0, 8*sizeof(double), s);
    return com_moments_stage(st, scratch, totals8, lead_mass, flags, false, s, launches);
}

int rebx_b200_com_bodies(const double* totals7, double* bodies, int nbodies, void* stream, int* launches) {
    if (!totals7 || !bodies || nbodies < 1) return cudaErrorInvalidValue;
    com_body_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(totals7, bodies, nbodies);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

static const JacobiCombo* find_jacobi_combo(const rebx_b200_pass& P) {
    for (const JacobiCombo& c : kJacobiCombos) {
        bool ok = true;
        for (int j = 0; j < 3; j++) if (c.k[j] != (j < P.n_effects ? P.fx[j].kind : 0)) ok = false;
        if (ok) return &c;
    }
    return nullptr;
}

int rebx_b200_jacobi_moments(const rebx_b200_state* st, void* scratch, double* totals8, const double* lead_mass, int flags, void* stream, int* launches) {
    if (!st || !scratch || !st->x || !st->vx || !st->m) return cudaErrorInvalidValue;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (st->n == 0) return totals8 ? cudaMemsetAsync(totals8, 0, 8*sizeof(double), s) : cudaSuccess;
    return com_moments_stage(st, scratch, totals8, lead_mass, flags, true, s, launches);
}

int rebx_b200_jacobi_sweep(const rebx_b200_pass* pass, void* scratch, const double* carry7, double* tsum3, int flags, void* stream, int* launches) {
    if (!pass || !scratch) return cudaErrorInvalidValue;
    const rebx_b200_pass& P = *pass;
    if (P.n_effects < 1 || P.n_effects > 3) return cudaErrorInvalidValue;
    if (!P.st.x || !P.st.vx || !P.st.m || !P.st.ax) return cudaErrorInvalidValue;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P.st.n == 0) return tsum3 ? cudaMemsetAsync(tsum3, 0, 3*sizeof(double), s) : cudaSuccess;
    const JacobiCombo* combo = find_jacobi_combo(P);
    if (!combo) return cudaErrorNotSupported;
    JacobiScratch J(scratch, P.st.n);
    const double* mref = (flags & REBX_B200_COM_WHOLE) ? J.mref : nullptr;
    // REBX_B200_PASS_TOLERANCE: the sweep in tolerance arithmetic where kernels_tol.cu has the stack (the scans are the same)
    if (!((P.reserved & REBX_B200_PASS_TOLERANCE) &&
          tol_jacobi_launch(P, (unsigned)J.ntiles, J.rows7, J.tvec, J.rows3, carry7, J.grid, J.totals + 7, mref, s)))
        combo->fn<<<(unsigned)J.ntiles, kScanThreads, 0, s>>>(P, J.rows7, J.tvec, J.rows3, carry7, J.grid, J.totals + 7, mref);
    scan_tiles_kernel<3, true><<<3, kTileScanThreads, 0, s>>>(J.rows3, J.ntiles, tsum3 ? tsum3 : J.totals + 8);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += 2;
    return err;
}

int rebx_b200_jacobi_apply(const rebx_b200_state* st, void* scratch, const double* carry3, void* stream, int* launches) {
    if (!st || !scratch || !st->ax) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    JacobiScratch J(scratch, st->n);
    jacobi_apply_kernel<<<(unsigned)J.ntiles, kScanThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, J.tvec, J.rows3, carry3);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

// The fused pipeline: cudaErrorNotSupported when the device cannot launch cooperatively (the caller falls back).
static int launch_jacobi_fused(const rebx_b200_pass& P, const JacobiCombo* combo, void* scratch, cudaStream_t s) {
    static thread_local std::vector<std::pair<int, int>> coop;            // (device, cooperative launch supported)
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return cudaErrorNotSupported;
    int ok = -1;
    for (const auto& c : coop) if (c.first == dev) ok = c.second;
    if (ok < 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) { cudaGetLastError(); v = 0; }
        coop.push_back({dev, v});
        ok = v;
    }
    if (!ok || !combo) return cudaErrorNotSupported;
    const void* fn = nullptr;
    if (P.reserved & REBX_B200_PASS_TOLERANCE) fn = tol_jacobi_fused_kernel(P);
    if (!fn) fn = reinterpret_cast<const void*>(combo->fused);
    JacobiScratch J(scratch, P.st.n);
    // resident grid: every CTA must be co-resident (cooperative launch); the per-CTA aggregates live in the tile rows of
    // the staged pipeline's scratch, so not more CTAs than 256-particle tiles
    int grid = grid_for(fn, J.ntiles*kScanThreads, kScanThreads, 0, kScanThreads);
    FusedScratch F;
    F.agg7 = J.rows7; F.tagg3 = J.rows3; F.tvec = J.tvec; F.mref = J.mref;
    double* extra = J.mref + (P.st.n + 2);
    F.first = reinterpret_cast<unsigned long long*>(extra);
    F.unsafe = reinterpret_cast<unsigned*>(extra + (J.ntiles + 1));
    int wtiles = (int)((P.st.n + 31)/32);                                  // 32-particle tiles, a contiguous range per warp
    void* args[] = { const_cast<rebx_b200_pass*>(&P), &F, &wtiles };
    const cudaError_t err = cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kScanThreads), args, 0, s);
    if (err == cudaErrorCooperativeLaunchTooLarge || err == cudaErrorNotSupported) { cudaGetLastError(); return cudaErrorNotSupported; }
    return err;
}

int rebx_b200_launch_jacobi(const rebx_b200_pass* pass, void* scratch, void* stream, int* launches) {
    if (!pass || !scratch) return cudaErrorInvalidValue;
    const rebx_b200_pass& P = *pass;
    if (P.n_effects < 1 || P.n_effects > 3 || P.st.index_base != 0) return cudaErrorInvalidValue;
    if (!P.st.x || !P.st.vx || !P.st.m || !P.st.ax) return cudaErrorInvalidValue;
    if (P.st.n == 0) return 0;
    if (!find_jacobi_combo(P)) {
        // no fused instantiation for this stack: one complete Jacobi pipeline per effect (equivalent,
        // the effects read positions and velocities only)
        if (P.n_effects == 1) return cudaErrorInvalidValue;
        for (int e = 0; e < P.n_effects; e++) {
            rebx_b200_pass one = P;
            one.n_effects = 1;
            one.fx[0] = P.fx[e];
            const int rc = rebx_b200_launch_jacobi(&one, scratch, stream, launches);
            if (rc) return rc;
        }
        return 0;
    }
    // One cooperative launch for the whole pipeline (jacobi_fused.cuh) where the device supports it; REBX_B200_JACOBI=staged
    // keeps the six-launch pipeline below (which is also what a sharded caller drives stage by stage).
    static const bool staged = [] { const char* e = getenv("REBX_B200_JACOBI"); return e && e[0] == 's'; }();
    if (!staged) {
        const int frc = launch_jacobi_fused(P, find_jacobi_combo(P), scratch, static_cast<cudaStream_t>(stream));
        if (frc == 0) { if (launches) (*launches)++; return 0; }
        if (frc != cudaErrorNotSupported) return frc;
    }
    int rc = rebx_b200_jacobi_moments(&P.st, scratch, nullptr, nullptr, REBX_B200_COM_WHOLE, stream, launches);
    if (rc == 0) rc = rebx_b200_jacobi_sweep(&P, scratch, nullptr, nullptr, REBX_B200_COM_WHOLE, stream, launches);
    if (rc == 0) rc = rebx_b200_jacobi_apply(&P.st, scratch, nullptr, stream, launches);
    return rc;
}

int rebx_b200_gr_pull(const rebx_b200_state* st, const rebx_b200_gr_args* args, double* pull, void* workspace, void* stream, int* launches) {
    if (!st || !args || !pull || !workspace || !st->m || !args->active) return cudaErrorInvalidValue;
    if (args->n_active < 1 || args->n_active > REBX_B200_GR_MAX_ACTIVE) return cudaErrorInvalidValue;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (st->n == 0) return cudaMemsetAsync(pull, 0, (size_t)args->n_active*3*sizeof(double), s);
    int grid = grid_for(reinterpret_cast<const void*>(gr_pull_kernel), st->n);
    const int cap = kMaxBlocks*REBX_B200_MAX_EFFECTS/kGrChunk;      // workspace rows available
    if (grid > cap) grid = cap;
    int nl = 0;
    for (int first = 0; first < args->n_active; first += kGrChunk) {        // kGrChunk actives per sweep over the test particles
        const int count = args->n_active - first < kGrChunk ? args->n_active - first : kGrChunk;
        gr_pull_kernel<<<grid, kThreads, 0, s>>>(*st, *args, first, static_cast<double*>(workspace));
        gr_pull_finalize_kernel<<<1, kThreads, 0, s>>>(static_cast<const double*>(workspace), grid, count, pull + (size_t)first*3);
        nl += 2;
    }
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += nl;
    return err;
}

int rebx_b200_gr_sweep(const rebx_b200_state* st, const rebx_b200_gr_args* args, int* not_converged, void* stream, int* launches) {
    if (!st || !args || !not_converged || !st->vx || !st->ax || !args->active) return cudaErrorInvalidValue;
    if (args->n_active < 1 || args->n_active > REBX_B200_GR_MAX_ACTIVE) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(gr_sweep_kernel), st->n);
    gr_sweep_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, *args, not_converged);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_potential(const rebx_b200_state* st, const rebx_b200_effect* fx, double* potential, void* workspace, void* stream, int* launches) {
    if (!st || !fx || !potential || !workspace || !fx->body || !st->x || !st->m) return cudaErrorInvalidValue;
    if (fx->kind != REBX_B200_GR_POTENTIAL && fx->kind != REBX_B200_HARMONICS && fx->kind != REBX_B200_CENTRAL_FORCE) return cudaErrorInvalidValue;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (st->n == 0) return cudaMemsetAsync(potential, 0, sizeof(double), s);
    int grid = grid_for(reinterpret_cast<const void*>(potential_kernel), st->n);
    potential_kernel<<<grid, kThreads, 0, s>>>(*st, *fx, static_cast<double*>(workspace));
    potential_finalize_kernel<<<1, kThreads, 0, s>>>(static_cast<const double*>(workspace), grid, potential);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += 2;
    return err;
}

int rebx_b200_add_uniform(const rebx_b200_state* st, const double* delta, void* stream, int* launches) {
    if (!st || !delta) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(add_uniform_kernel), st->n);
    add_uniform_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(st->ax, st->ay, st->az, st->n, delta);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

// The fused step in two halves, so that a SHARDED caller can all-reduce the back-reaction sums between
// them: (1) the sweep over the shard's particles + the reduction of this shard's back-reaction;
// (2) the central body's own step from the (summed) back-reaction, run identically by every rank on
// its replica of the body record.  A stack without back-reaction needs no exchange at all.
static int leapfrog_combo(const rebx_b200_pass* pass, const Combo** combo_out, unsigned* red_mask, bool* any) {
    if (!pass) return cudaErrorInvalidValue;
    const int bad = check_pass(*pass);
    if (bad) return bad;
    const rebx_b200_pass& P = *pass;
    if (!P.st.vx || !P.st.vy || !P.st.vz) return cudaErrorInvalidValue;
    const Combo* combo = find_combo(P.fx, P.n_effects);
    if (!combo) return cudaErrorNotSupported;            // stack not instantiated: use the separate sweeps
    for (int e = 1; e < P.n_effects; e++) if (P.fx[e].body == nullptr) return cudaErrorInvalidValue;
    *combo_out = combo;
    *red_mask = 0; *any = false;
    for (int j = 0; j < P.n_effects; j++) if (combo->red[j]) { *red_mask |= 1u << j; if (P.fx[j].backreaction) *any = true; }
    return 0;
}

int rebx_b200_leapfrog_sweep(const rebx_b200_pass* pass, double dt, double G, double softening, void* workspace, void* stream, int* launches) {
    if (!workspace) return cudaErrorInvalidValue;
    const Combo* combo = nullptr; unsigned red_mask = 0; bool any = false;
    const int rc = leapfrog_combo(pass, &combo, &red_mask, &any);
    if (rc) return rc;
    const rebx_b200_pass& P = *pass;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (P.st.n == 0) {      // an empty shard contributes zero back-reaction
        for (int j = 0; j < P.n_effects; j++)
            if (((red_mask >> j) & 1u) && P.fx[j].backreaction) cudaMemsetAsync(P.fx[j].backreaction, 0, 3*sizeof(double), s);
        return cudaGetLastError();
    }
    if ((P.reserved & REBX_B200_PASS_TOLERANCE) && red_mask == 0 && tol_leapfrog_sweep_launch(P, dt, G, softening*softening, s)) {
        // tolerance arithmetic (kernels_tol.cu: step_tol_kernel), one effect without back-reaction
        cudaError_t terr = cudaGetLastError();
        if (terr == cudaSuccess && launches) *launches += 1;
        return terr;
    }
    bool heavy = false;
    for (int j = 0; j < P.n_effects; j++) if (P.fx[j].kind >= REBX_B200_YARKOVSKY) heavy = true;
    const int vsel = (pass_is_vectorizable(P) && !heavy) ? 1 : 0;
    const int V = vsel ? 2 : 1;
    const StepArgs S = {dt, 0.5*dt, G, softening*softening};
    step_fn fn = combo->step[1][vsel];        // the fused step requires one central body for the whole stack
    const int grid = grid_for(reinterpret_cast<const void*>(fn), (P.st.n + V - 1)/V);
    fn<<<grid, kThreads, 0, s>>>(P, static_cast<double*>(workspace), S);       // its last CTA finishes the sums (and the body, see _step)
    (void)any;
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

int rebx_b200_leapfrog_body(const rebx_b200_pass* pass, double dt, void* stream, int* launches) {
    const Combo* combo = nullptr; unsigned red_mask = 0; bool any = false;
    const int rc = leapfrog_combo(pass, &combo, &red_mask, &any);
    if (rc) return rc;
    const StepArgs S = {dt, 0.5*dt, 0.0, 0.0};
    step_body_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*pass, S, any ? red_mask : 0u);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_leapfrog_step(const rebx_b200_pass* pass, double dt, double G, double softening, void* workspace, void* stream, int* launches) {
    if (pass && pass->st.n == 0) return check_pass(*pass);
    if (!pass) return cudaErrorInvalidValue;
    const Combo* combo = nullptr; unsigned red_mask = 0; bool any = false;
    int rc = leapfrog_combo(pass, &combo, &red_mask, &any);
    if (rc) return rc;
    if (red_mask != 0) {
        // a stack with back-reaction sums, single shard: the sweep's last CTA finishes the sums AND advances the body
        rebx_b200_pass Q = *pass;
        Q.reserved |= REBX_B200_PASS_APPLY_BACKREACTION;
        return rebx_b200_leapfrog_sweep(&Q, dt, G, softening, workspace, stream, launches);
    }
    // no sums (radiation_forces, yarkovsky_effect): those sweeps take no ticket (pass_common.cuh), the body follows in its own launch
    rebx_b200_pass Q = *pass;
    Q.reserved &= ~REBX_B200_PASS_APPLY_BACKREACTION;
    rc = rebx_b200_leapfrog_sweep(&Q, dt, G, softening, workspace, stream, launches);
    if (rc == 0) rc = rebx_b200_leapfrog_body(&Q, dt, stream, launches);
    return rc;
}

// One force evaluation of the stack on stage arrays (v, a): what the reference's schemes do with
// force->update_accelerations(sim, force, scratch_copy, N).  The body record follows the stage's
// velocity of the body, as `particles[source]` does there.
static int stage_eval(const rebx_b200_pass& base, CV3 v, V3 a, void* workspace, cudaStream_t s, int* launches) {
    rebx_b200_pass P = base;
    P.st.vx = v.x; P.st.vy = v.y; P.st.vz = v.z;
    P.st.ax = a.x; P.st.ay = a.y; P.st.az = a.z;
    int rc = rebx_b200_gather_bodies(&P.st, const_cast<double*>(P.fx[0].body), P.n_effects, s, launches);
    if (rc) return rc;
    rc = rebx_b200_launch_pass(&P, workspace, s, launches);
    if (rc) return rc;
    return rebx_b200_apply_backreaction(&P, s, launches);
}

int rebx_b200_integrate_force(const rebx_b200_pass* pass, double dt, int scheme, void* scratch, void* workspace,
                              void* stream, int* launches, int* iterations) {
    if (!pass || !scratch || !workspace) return cudaErrorInvalidValue;
    const int bad = check_pass(*pass);
    if (bad) return bad;
    const rebx_b200_pass& P = *pass;
    if (!P.st.vx) return cudaErrorInvalidValue;
    // the body records of a stack must be contiguous (as device.Shard and the dispatcher lay them out)
    for (int e = 1; e < P.n_effects; e++) if (P.fx[e].body != P.fx[0].body + (size_t)e*REBX_B200_BODY_DOUBLES) return cudaErrorInvalidValue;
    const int64_t n = P.st.n;
    if (n == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* w = static_cast<double*>(scratch);
    auto slot = [&](int i) { return V3{w + (size_t)(3*i)*n, w + (size_t)(3*i + 1)*n, w + (size_t)(3*i + 2)*n}; };
    auto c = [](V3 q) { return CV3{q.x, q.y, q.z}; };
    const V3 v = {const_cast<double*>(P.st.vx), const_cast<double*>(P.st.vy), const_cast<double*>(P.st.vz)};
    const V3 a1 = {P.st.ax, P.st.ay, P.st.az};
    const int grid = grid_for(reinterpret_cast<const void*>(stage_axpy_kernel), n);
    int nl = 0, rc = 0;
    cudaError_t err;
    auto zero = [&](V3 q) {
        cudaMemsetAsync(q.x, 0, n*sizeof(double), s); cudaMemsetAsync(q.y, 0, n*sizeof(double), s); cudaMemsetAsync(q.z, 0, n*sizeof(double), s);
    };
    zero(a1);                                        // rebx_reset_accelerations (integrate_force.c:46)
    if (scheme == 2) {                               // Euler
        if ((rc = stage_eval(P, c(v), a1, workspace, s, &nl))) return rc;
        stage_axpy_kernel<<<grid, kThreads, 0, s>>>(v, c(v), c(a1), dt, n); nl++;
    } else if (scheme == 3) {                        // RK2
        const V3 v2 = slot(0), a2 = slot(1);
        zero(a2);
        if ((rc = stage_eval(P, c(v), a1, workspace, s, &nl))) return rc;
        stage_axpy_kernel<<<grid, kThreads, 0, s>>>(v2, c(v), c(a1), 2.*dt/3., n); nl++;
        if ((rc = stage_eval(P, c(v2), a2, workspace, s, &nl))) return rc;
        rk2_final_kernel<<<grid, kThreads, 0, s>>>(v, c(a1), c(a2), dt/4., 3.*dt/4., n); nl++;
    } else if (scheme == 1) {                        // RK4
        const V3 v2 = slot(0), a2 = slot(1), v3 = slot(2), a3 = slot(3);
        const double dt2 = dt/2.;
        zero(a2); zero(a3);
        if ((rc = stage_eval(P, c(v), a1, workspace, s, &nl))) return rc;
        stage_axpy_kernel<<<grid, kThreads, 0, s>>>(v2, c(v), c(a1), dt2, n); nl++;
        if ((rc = stage_eval(P, c(v2), a2, workspace, s, &nl))) return rc;
        stage_axpy_kernel<<<grid, kThreads, 0, s>>>(v3, c(v), c(a2), dt2, n); nl++;
        if ((rc = stage_eval(P, c(v3), a3, workspace, s, &nl))) return rc;
        rk4_stage4_kernel<<<grid, kThreads, 0, s>>>(v2, a2, a3, c(v), dt, n); nl++;
        if ((rc = stage_eval(P, c(v2), a2, workspace, s, &nl))) return rc;
        rk4_final_kernel<<<grid, kThreads, 0, s>>>(v, c(a1), c(a2), c(a3), dt/6., n); nl++;
    } else if (scheme == 0) {                        // implicit midpoint
        const V3 vfinal = slot(0), vprev = slot(1), vavg = slot(2), aavg = slot(3);
        double* sums = static_cast<double*>(workspace) + (size_t)kMaxBlocks*REBX_B200_MAX_EFFECTS*3;        // scalar tail of the workspace
        for (int q = 0; q < 3; q++) {
            double* src[3] = {v.x, v.y, v.z};
            double* d1[3] = {vfinal.x, vfinal.y, vfinal.z};
            double* d2[3] = {vavg.x, vavg.y, vavg.z};
            cudaMemcpyAsync(d1[q], src[q], n*sizeof(double), cudaMemcpyDeviceToDevice, s);
            cudaMemcpyAsync(d2[q], src[q], n*sizeof(double), cudaMemcpyDeviceToDevice, s);
        }
        zero(aavg);
        int it = 0;
        for (it = 0; it < 10; it++) {
            for (int q = 0; q < 3; q++) {
                double* from[3] = {vfinal.x, vfinal.y, vfinal.z};
                double* to[3] = {vprev.x, vprev.y, vprev.z};
                cudaMemcpyAsync(to[q], from[q], n*sizeof(double), cudaMemcpyDeviceToDevice, s);
            }
            if ((rc = stage_eval(P, c(vavg), aavg, workspace, s, &nl))) return rc;
            stage_axpy_kernel<<<grid, kThreads, 0, s>>>(vfinal, c(v), c(aavg), dt, n); nl++;
            cudaMemsetAsync(sums, 0, 2*sizeof(double), s);
            im_compare_kernel<<<grid, kThreads, 0, s>>>(c(vfinal), c(vprev), n, sums); nl++;
            double h[2];
            if ((err = cudaMemcpyAsync(h, sums, sizeof h, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return err;
            if ((err = cudaStreamSynchronize(s)) != cudaSuccess) return err;
            if (h[0]/h[1] < 2.220446049250313e-16*2.220446049250313e-16) break;
            im_average_kernel<<<grid, kThreads, 0, s>>>(vavg, aavg, c(v), c(vfinal), n); nl++;
        }
        if (iterations) *iterations = it;
        for (int q = 0; q < 3; q++) {
            double* from[3] = {vfinal.x, vfinal.y, vfinal.z};
            double* to[3] = {v.x, v.y, v.z};
            cudaMemcpyAsync(to[q], from[q], n*sizeof(double), cudaMemcpyDeviceToDevice, s);
        }
    } else {
        return cudaErrorInvalidValue;
    }
    // leave the body records consistent with the updated velocities
    if ((rc = rebx_b200_gather_bodies(&P.st, const_cast<double*>(P.fx[0].body), P.n_effects, s, &nl))) return rc;
    err = cudaGetLastError();
    if (err == cudaSuccess && launches) *launches += nl;
    return err;
}

int rebx_b200_drift(const rebx_b200_state* st, double dt, void* stream, int* launches) {
    if (!st || !st->vx) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(drift_kernel), st->n);
    drift_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        const_cast<double*>(st->x), const_cast<double*>(st->y), const_cast<double*>(st->z), st->vx, st->vy, st->vz, st->n, dt);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_kick(const rebx_b200_state* st, double dt, void* stream, int* launches) {
    if (!st || !st->vx || !st->ax) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(drift_kernel), st->n);
    drift_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(          // same update, other arrays
        const_cast<double*>(st->vx), const_cast<double*>(st->vy), const_cast<double*>(st->vz), st->ax, st->ay, st->az, st->n, dt);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_kepler_drift(const rebx_b200_state* st, const double* body, double G, double dt, void* stream, int* launches) {
    if (!st || !body || !st->x || !st->y || !st->z || !st->vx || !st->vy || !st->vz) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(kepler_drift_kernel), st->n);
    kepler_drift_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, body, G, dt);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_wh_kick_drift(const rebx_b200_pass* pass, double G, double dt_kick, double dt_drift, void* stream, int* launches) {
    if (!pass) return cudaErrorInvalidValue;
    const int bad = check_pass(*pass);
    if (bad) return bad;
    const rebx_b200_pass& P = *pass;
    if (!P.st.vx || !P.st.vy || !P.st.vz) return cudaErrorInvalidValue;
    if (P.n_effects != 1) return cudaErrorNotSupported;
    if (P.st.n == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if ((P.reserved & REBX_B200_PASS_TOLERANCE) && (P.fx[0].kind == REBX_B200_RADIATION || P.fx[0].kind == REBX_B200_YARKOVSKY) &&
        tol_wh_kick_drift_launch(P, G, dt_kick, dt_drift, s)) {
        // tolerance arithmetic (kernels_tol.cu): FMA, Newton reciprocals, the force's own reciprocal distance reused by the drift
    } else if (P.fx[0].kind == REBX_B200_RADIATION) {
        const int grid = grid_for(reinterpret_cast<const void*>(wh_kick_drift_kernel<REBX_B200_RADIATION>), P.st.n);
        wh_kick_drift_kernel<REBX_B200_RADIATION><<<grid, kThreads, 0, s>>>(P, G, dt_kick, dt_drift);
    } else if (P.fx[0].kind == REBX_B200_YARKOVSKY) {
        const int grid = grid_for(reinterpret_cast<const void*>(wh_kick_drift_kernel<REBX_B200_YARKOVSKY>), P.st.n);
        wh_kick_drift_kernel<REBX_B200_YARKOVSKY><<<grid, kThreads, 0, s>>>(P, G, dt_kick, dt_drift);
    } else {
        return cudaErrorNotSupported;          // an effect with back-reaction: use the separate kernels
    }
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_zero_acc(const rebx_b200_state* st, void* stream, int* launches) {
    if (!st || !st->ax || !st->ay || !st->az) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t err = cudaMemsetAsync(st->ax, 0, (size_t)st->n*sizeof(double), s);
    if (err == cudaSuccess) err = cudaMemsetAsync(st->ay, 0, (size_t)st->n*sizeof(double), s);
    if (err == cudaSuccess) err = cudaMemsetAsync(st->az, 0, (size_t)st->n*sizeof(double), s);
    (void)launches;
    return err;
}

int rebx_b200_point_mass_gravity(const rebx_b200_state* st, const double* bodies, int nbodies, double G, double softening, void* stream, int* launches) {
    if (!st || !bodies || nbodies < 0 || nbodies > 8 || !st->ax) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(gravity_kernel), st->n);
    gravity_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(*st, bodies, nbodies, G, softening*softening);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_gather_bodies(const rebx_b200_state* st, double* bodies, int nbodies, void* stream, int* launches) {
    if (!st || !bodies || nbodies < 1 || nbodies > 32) return cudaErrorInvalidValue;
    gather_bodies_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*st, bodies, nbodies);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_unpack_aos(const void* aos, const rebx_b200_aos_layout* L, const rebx_b200_state* st, void* stream, int* launches) {
    if (!aos || !L || !st) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(unpack_aos_kernel), st->n);
    unpack_aos_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned char*>(aos), *L,
        const_cast<double*>(st->x), const_cast<double*>(st->y), const_cast<double*>(st->z),
        const_cast<double*>(st->vx), const_cast<double*>(st->vy), const_cast<double*>(st->vz),
        const_cast<double*>(st->m), const_cast<double*>(st->r), st->ax, st->ay, st->az, st->n);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

int rebx_b200_pack_acc_aos(void* aos, const rebx_b200_aos_layout* L, const rebx_b200_state* st, void* stream, int* launches) {
    if (!aos || !L || !st) return cudaErrorInvalidValue;
    if (st->n == 0) return 0;
    const int grid = grid_for(reinterpret_cast<const void*>(pack_acc_aos_kernel), st->n);
    pack_acc_aos_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<unsigned char*>(aos), *L, st->ax, st->ay, st->az, st->n);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
    return err;
}

}  // extern "C"

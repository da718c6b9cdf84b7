// jacobi_fused.cuh -- REBX_COORDINATES_JACOBI of rebx_com_force (reference src/rebxtools.c:66-128) for a WHOLE ensemble
// in ONE cooperative launch.  The staged pipeline (kernels_jacobi.cuh: tile sums -> scan of tile sums -> status -> sweep ->
// reverse scan -> apply, six launches and two memsets) costs 0.15-0.18 ms at 10^6 planetesimals of which less than half
// is the sweep: every stage is a 9-25 us launch that cannot hide its own latency.  Here a resident grid (co-residency
// guaranteed by the cooperative launch) gives every WARP a contiguous range of 32-particle tiles and runs the three phases
// with two grid-wide barriers between them; inside a phase a warp never meets a CTA barrier (warp scans by shuffles, the
// running carry in a per-warp shared-memory slot), so the warps of an SM drift apart and hide each other's latencies:
//   A  mass moments (sum rn_u(m), sum m r, sum m v) of the warp's range (kept in shared memory), summed over the CTA
//                                                                                       -> agg7[7][G], unsafe[G], first[G]
//      ---- grid barrier ----   every CTA: sum over the warps in front of it, grand totals, status of the mass-grid
//                               argument (jacobi_common.cuh), the literal last peel `dip`; every warp: its own prefix
//   B  tiles in ascending order, running carry: COM_k from the prefix, the stacked damping forces (Force functor:
//      reference-exact arithmetic in kernels.cu, tolerance arithmetic in kernels_tol.cu),
//      a += f, t_k = m_k/(M_k + m_k) f -> tvec, sum of t over the range, summed over the CTA   -> tagg3[3][G]
//      ---- grid barrier ----   every warp: the sum over the warps behind it
//   C  tiles in descending order: a_j -= sum_{i >= j} t_i
// (a first version walked 256-particle tiles per CTA with CTA-wide scans, five barriers per tile: 0.128 ms at 10^6 in
// tolerance arithmetic against 0.152 staged, and no gain at all in exact arithmetic.)
// Phases B and C re-read what A and B touched from L2 (the whole working set of 10^6 particles is ~100 MB).
// All sums run in a fixed order for a given grid: deterministic.  The sharded pipeline (an index range per rank with
// carries exchanged between the stages) stays the staged one; so does any caller that sets REBX_B200_JACOBI=staged.
#pragma once
#include <cooperative_groups.h>
#include "jacobi_common.cuh"

namespace rebxb {

struct FusedScratch {
    double* agg7;                   // [7][G]  mass moments of each CTA's tile range (component-major), G = CTAs of the grid
    double* tagg3;                  // [3][G]  sum of t over each CTA's tile range
    double* tvec;                   // [3][n]  t_k
    double* mref;                   // [n]     interior masses by the literal recurrences (status 1 only)
    unsigned long long* first;      // [G]     smallest index >= 1 in the range whose mass is not zero on the grid (~0: none)
    unsigned* unsafe;               // [G]     the range holds a mass outside the mass-grid argument
};
constexpr int kFusedWarps = kScanThreads/32;

// Phase B's state of the NEXT tile travels global -> shared by cp.async while the current tile's forces are evaluated
// (every thread copies its own ten operands and reads only those back after cp.async.wait_group: no barrier, no
// registers); without it the top of every iteration waits a full L2 / HBM round trip (ncu: 6 % of the kernel).
#ifndef REBXB_JF_STAGED
#define REBXB_JF_STAGED 1
#endif
__device__ __forceinline__ void jf_cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void jf_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void jf_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Sum over the entries i < lo (forward) or i >= hi (REV) and over all `W` entries of NC component-major rows, by the
// whole CTA in a fixed order.  part[c], all[c] valid in every thread.
template <int NC, bool REV>
__device__ __forceinline__ void cta_sum_of_aggregates(const double* __restrict__ rows, int W, int lo, int hi, double (&part)[NC], double (&all)[NC],
                                                      double (*s_warp)[NC]) {
    double vb[NC], va[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) { vb[c] = 0.0; va[c] = 0.0; }
    for (int i = threadIdx.x; i < W; i += kScanThreads) {
        const bool in = REV ? (i >= hi) : (i < lo);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const double v = rows[(size_t)c*W + i];
            va[c] += v;
            if (in) vb[c] += v;
        }
    }
    block_exclusive_scan<NC>(vb, part, s_warp);
    __syncthreads();
    block_exclusive_scan<NC>(va, all, s_warp);
    __syncthreads();
}

template <int NC>
__device__ __forceinline__ void warp_sum(double (&v)[NC]) {            // butterfly: every lane ends with the same sum
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], off);
    }
}
// v := exclusive prefix over the lanes, total[c] := sum over the warp (in every lane)
template <int NC>
__device__ __forceinline__ void warp_exclusive_scan(double (&v)[NC], double (&total)[NC]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double x = v[c];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        total[c] = __shfl_sync(0xffffffffu, x, 31);
        const double up = __shfl_up_sync(0xffffffffu, x, 1);
        v[c] = lane ? up : 0.0;
    }
}

// The literal forward fold and backward peel of the reference (reb_simulation_com, rebx_get_com_without_particle), one
// warp, O(n) dependent additions: the fallback when the mass-grid argument does not hold (as in mass_status_kernel).
__device__ __forceinline__ void literal_interior_masses(const rebx_b200_state& st, double* __restrict__ mref) {
    const int lane = threadIdx.x & 31;
    double T = 0.0;
    for (int64_t base = 0; base < st.n; base += 32) {
        const int64_t k = base + lane;
        const double mine = (k < st.n) ? st.m[k] : 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double mj = __shfl_sync(0xffffffffu, mine, j);
            if (base + j < st.n) T += mj;
        }
    }
    double B = T;
    for (int64_t top = st.n; top > 0; top -= 32) {
        const int64_t k = top - 1 - lane;
        const double mine = (k >= 0) ? st.m[k] : 0.0;
        double out = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const double mj = __shfl_sync(0xffffffffu, mine, j);
            if (top - 1 - j >= 0) B -= mj;
            if (lane == j) out = B;
        }
        if (k >= 0) mref[k] = out;
    }
}

// Force: per-CTA set-up `init(P)` (followed by a barrier here), `load(P, k)` -> Force::Ops (the particle's parameter
// loads, issued next to the state loads) and `eval(P, ops, k, p, M, S, a, t)`: a += the forces of the stack on particle k
// evaluated against the centre of mass S/M of the particles in front of it, t += m_k/(M + m_k) times those forces.
// `wtiles` = number of 32-particle tiles.
template <class Force>
__device__ __forceinline__ void jacobi_fused_body(const rebx_b200_pass& P, const FusedScratch F, const int wtiles) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_warp[kFusedWarps][7];
    __shared__ double s_carry[kFusedWarps][7];
    __shared__ double s_agg[kFusedWarps][7];                            // the warps' mass moments (phase A), later their sums of t (phase B)
    __shared__ unsigned long long s_wfirst[kFusedWarps];
    __shared__ unsigned s_wunsafe[kFusedWarps];
    __shared__ unsigned long long s_first;
    const rebx_b200_state& st = P.st;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int G = (int)gridDim.x, b = (int)blockIdx.x;
    const int W = G*kFusedWarps, gw = b*kFusedWarps + w;
    const int t0 = (int)((int64_t)wtiles*gw/W), t1 = (int)((int64_t)wtiles*(gw + 1)/W);     // this warp's tiles [t0, t1)
    const MassGrid mg0 = make_mass_grid(st.m[0]);

    // ---- phase A: mass moments of the warp's range ---------------------------------------------------------------------
    {
        double acc[7] = {0, 0, 0, 0, 0, 0, 0};
        bool bad = false;
        unsigned long long first = ~0ull;
#pragma unroll 2
        for (int tile = t0; tile < t1; tile++) {
            const int64_t k = (int64_t)tile*32 + lane;
            if (k < st.n) {
                const double m = st.m[k];
                const double q = quantise_mass(m, mg0, bad);
                acc[0] += q;                                            // multiples of u below the top of the binade: exact
                acc[1] += st.x[k]*m; acc[2] += st.y[k]*m; acc[3] += st.z[k]*m;
                acc[4] += st.vx[k]*m; acc[5] += st.vy[k]*m; acc[6] += st.vz[k]*m;
                if (k >= 1 && q != 0.0 && first == ~0ull) first = (unsigned long long)k;
            }
        }
        warp_sum<7>(acc);
        const bool any_bad = __any_sync(0xffffffffu, bad);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, first, off);
            first = o < first ? o : first;
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 7; c++) s_agg[w][c] = acc[c];
            s_wunsafe[w] = any_bad ? 1u : 0u;
            s_wfirst[w] = first;
        }
        __syncthreads();
        if (threadIdx.x < 7) {                                          // the CTA's moments: its warps in order
            double sum = 0.0;
            for (int j = 0; j < kFusedWarps; j++) sum += s_agg[j][threadIdx.x];
            F.agg7[(size_t)threadIdx.x*G + b] = sum;
        } else if (threadIdx.x == 32) {
            unsigned u = 0u;
            unsigned long long f = ~0ull;
            for (int j = 0; j < kFusedWarps; j++) { u |= s_wunsafe[j]; f = s_wfirst[j] < f ? s_wfirst[j] : f; }
            F.unsafe[b] = u;
            F.first[b] = f;
        }
    }
    __threadfence();
    grid.sync();

    // ---- every CTA: the sum in front of it, the totals, the status of the mass-grid argument ---------------------------
    double before[7], all[7];
    cta_sum_of_aggregates<7, false>(F.agg7, G, b, 0, before, all, s_warp);
    MassGrid mg = mg0;
    bool literal;
    {
        bool u = false;
        unsigned long long fm = ~0ull;
        for (int i = threadIdx.x; i < G; i += kScanThreads) {
            u = u || F.unsafe[i] != 0u;
            const unsigned long long f = F.first[i];
            fm = f < fm ? f : fm;
        }
        if (threadIdx.x == 0) s_first = ~0ull;
        const int any_unsafe = __syncthreads_or(u);
        if (fm != ~0ull) atomicMin(&s_first, fm);
        __syncthreads();
        const unsigned long long r = s_first;
        bool bad = any_unsafe || !mg.ok || !(all[0] < mg.top);
        if (!bad && mg.lead == 0.5*mg.top) {
            // leading mass an exact power of two: a non-zero mass too small for the grid in front of the first massive
            // particle is outside the argument (mass_status_kernel)
            const int64_t hi = r < (unsigned long long)st.n ? (int64_t)r : st.n;
            bool tiny = hi - 1 > 65536;
            for (int64_t k = 1 + threadIdx.x; k < hi && !tiny; k += kScanThreads) if (st.m[k] != 0.0) tiny = true;
            bad = __syncthreads_or(tiny) != 0;
        }
        if (mg.ok && r < (unsigned long long)st.n) {                    // B_r = fl((m_0 + u q_r) - m_r): the literal last peel
            const double mr = st.m[r];
            bool unused = false;
            const double mq = quantise_mass(mr, mg, unused);
            mg.dip = (mg.lead + mq) - mr;
        }
        literal = bad;
    }
    if (literal) {                                                      // uniform over the grid: every CTA derived it from the same data
        if (blockIdx.x == 0 && threadIdx.x < 32) literal_interior_masses(st, F.mref);
        __threadfence();
        grid.sync();
    }

    // ---- phase B: forces, tiles in ascending order --------------------------------------------------------------------
    Force force;
    force.init(P);
    if (lane == 0) {
        // this warp's prefix: the CTA's + the warps of this CTA in front of it, in order
#pragma unroll
        for (int c = 0; c < 7; c++) {
            double sum = before[c];
            for (int j = 0; j < w; j++) sum += s_agg[j][c];
            s_carry[w][c] = sum;
        }
    }
    __syncthreads();
    double tsum[3] = {0.0, 0.0, 0.0};
#if REBXB_JF_STAGED
    __shared__ double s_stage[2][10][kScanThreads];
    auto stage = [&](int buf, int64_t k) {
        if (k < st.n) {
            double* base = &s_stage[buf][0][threadIdx.x];
            jf_cp_async8(base + 0*kScanThreads, st.x + k);  jf_cp_async8(base + 1*kScanThreads, st.y + k);  jf_cp_async8(base + 2*kScanThreads, st.z + k);
            jf_cp_async8(base + 3*kScanThreads, st.vx + k); jf_cp_async8(base + 4*kScanThreads, st.vy + k); jf_cp_async8(base + 5*kScanThreads, st.vz + k);
            jf_cp_async8(base + 6*kScanThreads, st.m + k);
            jf_cp_async8(base + 7*kScanThreads, st.ax + k); jf_cp_async8(base + 8*kScanThreads, st.ay + k); jf_cp_async8(base + 9*kScanThreads, st.az + k);
        }
        jf_cp_async_commit();
    };
    int buf = 0;
    if (t0 < t1) stage(0, (int64_t)t0*32 + lane);
#endif
    for (int tile = t0; tile < t1; tile++) {
        const int64_t k = (int64_t)tile*32 + lane;
        const bool live = k < st.n;
        double px = 0, py = 0, pz = 0, pvx = 0, pvy = 0, pvz = 0, pm = 0, a0 = 0, a1 = 0, a2 = 0;
        typename Force::Ops ops = {};
#if REBXB_JF_STAGED
        if (live) ops = force.load(P, k);
        if (tile + 1 < t1) { stage(buf ^ 1, k + 32); jf_cp_async_wait<1>(); } else jf_cp_async_wait<0>();
        if (live) {
            const double* base = &s_stage[buf][0][threadIdx.x];
            px = base[0*kScanThreads]; py = base[1*kScanThreads]; pz = base[2*kScanThreads];
            pvx = base[3*kScanThreads]; pvy = base[4*kScanThreads]; pvz = base[5*kScanThreads]; pm = base[6*kScanThreads];
            a0 = base[7*kScanThreads]; a1 = base[8*kScanThreads]; a2 = base[9*kScanThreads];
        }
        buf ^= 1;
#else
        if (live) {
            px = st.x[k]; py = st.y[k]; pz = st.z[k]; pvx = st.vx[k]; pvy = st.vy[k]; pvz = st.vz[k]; pm = st.m[k];
            ops = force.load(P, k);
            a0 = st.ax[k]; a1 = st.ay[k]; a2 = st.az[k];
        }
#endif
        bool unused_flag = false;
        double v[7] = {live ? quantise_mass(pm, mg, unused_flag) : 0.0, px*pm, py*pm, pz*pm, pvx*pm, pvy*pm, pvz*pm};
        double total[7];
        warp_exclusive_scan<7>(v, total);
        double M, S[6];
        {
            const double c0 = s_carry[w][0];
            M = c0 + v[0];                                              // mass interior to particle k: exact sums of multiples of u
#pragma unroll
            for (int j = 0; j < 6; j++) S[j] = s_carry[w][1 + j] + v[1 + j];
            __syncwarp();
            if (lane == 0) {
                s_carry[w][0] = c0 + total[0];
#pragma unroll
                for (int j = 1; j < 7; j++) s_carry[w][j] += total[j];
            }
            __syncwarp();
        }
        if (mg.ok && M == mg.lead && k >= 1) M = mg.dip;                // bottom of the binade: the literal last peel
        if (literal && live) M = F.mref[k];
        if (live) {
            double t[3] = {0.0, 0.0, 0.0};
            if (k != 0) {                                               // particle 0 has no Jacobi coordinate (rebxtools.c:71-73)
                Vec3 a = {a0, a1, a2};
                force.eval(P, ops, k, px, py, pz, pvx, pvy, pvz, pm, M, S, a, t);
                __stcg(st.ax + k, a.x); __stcg(st.ay + k, a.y); __stcg(st.az + k, a.z);
                tsum[0] += t[0]; tsum[1] += t[1]; tsum[2] += t[2];
            }
            __stcg(F.tvec + k, t[0]); __stcg(F.tvec + st.n + k, t[1]); __stcg(F.tvec + 2*st.n + k, t[2]);
        }
    }
    warp_sum<3>(tsum);
    __syncthreads();                                                    // every warp has taken its prefix from s_agg
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) s_agg[w][c] = tsum[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double sum = 0.0;
        for (int j = 0; j < kFusedWarps; j++) sum += s_agg[j][threadIdx.x];
        F.tagg3[(size_t)threadIdx.x*G + b] = sum;
    }
    __threadfence();
    grid.sync();

    // ---- phase C: a_j -= sum_{i >= j} t_i, tiles in descending order -----------------------------------------------------
    double behind[3], all3[3];
    cta_sum_of_aggregates<3, true>(F.tagg3, G, 0, b + 1, behind, all3, reinterpret_cast<double (*)[3]>(&s_warp[0][0]));
    double carry[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double sum = behind[c];
        for (int j = kFusedWarps - 1; j > w; j--) sum += s_agg[j][c];     // the warps of this CTA behind this one
        carry[c] = sum;
    }
    for (int tile = t1 - 1; tile >= t0; tile--) {
        const int64_t tile_lo = (int64_t)tile*32;
        const int64_t tile_hi = (tile_lo + 32 < st.n) ? tile_lo + 32 : st.n;
        const int64_t k = tile_hi - 1 - lane;                           // reversed order inside the tile
        const bool live = k >= tile_lo;
        double t[3] = {0.0, 0.0, 0.0}, a[3] = {0.0, 0.0, 0.0};
        if (live) {
            t[0] = __ldcg(F.tvec + k); t[1] = __ldcg(F.tvec + st.n + k); t[2] = __ldcg(F.tvec + 2*st.n + k);
            a[0] = __ldcg(st.ax + k); a[1] = __ldcg(st.ay + k); a[2] = __ldcg(st.az + k);
        }
        const double own[3] = {t[0], t[1], t[2]};
        double tot[3];
        warp_exclusive_scan<3>(t, tot);
        if (live) {
            __stcs(st.ax + k, a[0] - (carry[0] + (t[0] + own[0])));
            __stcs(st.ay + k, a[1] - (carry[1] + (t[1] + own[1])));
            __stcs(st.az + k, a[2] - (carry[2] + (t[2] + own[2])));
        }
        carry[0] += tot[0]; carry[1] += tot[1]; carry[2] += tot[2];
    }
}

}  // namespace rebxb

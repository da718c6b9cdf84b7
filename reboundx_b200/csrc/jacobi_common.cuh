// jacobi_common.cuh -- the pieces of the centre-of-mass pipeline (kernels_jacobi.cuh) that the tolerance-mode Jacobi
// sweep in kernels_tol.cu needs as well: the CTA-wide exclusive scan and the mass grid on which the reference's
// sequentially rounded interior masses are reproduced.  Templates and inline functions only (two translation units).
#pragma once
#include <cuda_runtime.h>
#include <cmath>

namespace rebxb {

constexpr int kScanThreads = 256;

// Exclusive scan of NC doubles per thread across a CTA of kScanThreads; returns the CTA total
// in `total` (valid in every thread).  Fixed evaluation order => deterministic.
template <int NC, int THREADS = kScanThreads>
__device__ __forceinline__ void block_exclusive_scan(double (&v)[NC], double (&total)[NC], double (*s_warp)[NC]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl[NC], excl[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double x = v[c];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        incl[c] = x;
        const double up = __shfl_up_sync(0xffffffffu, x, 1);
        excl[c] = lane ? up : 0.0;
    }
    __syncthreads();
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < NC; c++) s_warp[warp][c] = incl[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; c++) {
        double before = 0.0, all = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS/32; w++) {
            const double s = s_warp[w][c];
            if (w < warp) before += s;
            all += s;
        }
        total[c] = all;
        v[c] = before + excl[c];
    }
}

// ---- the reference's sequentially rounded interior mass, reproduced in parallel -----------------------
// The reference obtains the mass interior to particle i (com.m in rebx_com_force) from two SEQUENTIAL
// double-precision recurrences: the forward fold of reb_simulation_com (T_k = fl(T_{k-1} + m_k)) and the
// backward peel of rebx_get_com_without_particle (B_i = fl(B_{i+1} - m_i), src/rebxtools.c:6-30,92-99).
// Measured (oracle_com_force_linear, com_mode 1-3): replacing that mass by the EXACT prefix sum changes the
// damping forces by 2e-12 norm-wise at N = 2e4 (mu = G (m + M_i) enters v^2 - mu/d, a difference of nearly
// equal numbers on a near-circular orbit), while the position / velocity moments tolerate any summation
// order (2e-15).  So the mass has to carry the reference's rounding, not a better one.
// It can, exactly: while the running sum stays inside one binade [2^K, 2^(K+1)) -- a central body at index 0
// that outweighs everything behind it up to the next power of two -- every T_k is a multiple of
// u = 2^(K-52), hence fl(T + m) = T + rn_u(m) with rn_u(m) = m rounded to the nearest multiple of u, and
// likewise fl(B - m) = B - rn_u(m): both recurrences are EXACT integer arithmetic on q_j = rn_u(m_j)/u, so
// B_i = T_{i-1} = u * sum_{j<i} q_j, and sums of multiples of u below 2^(K+1) are exact in any order.  The
// scans below therefore run on rn_u(m_j) and reproduce the reference's interior masses bit for bit.
// One more literal step is needed where the peel reaches the BOTTOM of the binade: for the first particle r >= 1
// with q_r != 0 the reference computes B_r = fl(T_r - m_r) = fl((m_0 + u q_r) - m_r), which is m_0 unless m_0 is an
// exact power of two (1.0 solar mass ...): then the difference may round on the finer grid below 2^K (to
// m_0 - u/2), and the massless particles in front of r inherit that value.  `dip` below is this B_r, evaluated
// literally; it stands in wherever the prefix equals the leading mass.
// Not covered by the argument (flagged per launch, see mass_status_kernel): a running sum that leaves the
// binade, a negative mass, or a tie (m_j/u an exact half-integer: the reference's result then depends on the
// parity of the running sum).  A single-shard evaluation then recomputes the two recurrences literally, one
// thread, O(N) dependent additions (mass_sequential path); a sharded one keeps the quantised sums, which are
// as accurate as the reference's (each within u/2 per term) but no longer bit-identical to them.
struct MassGrid { double u, inv_u, top, lead, dip; int ok; int pad; };

__device__ __forceinline__ MassGrid make_mass_grid(double m0) {
    MassGrid g; g.u = 0.0; g.inv_u = 0.0; g.top = 0.0; g.lead = m0; g.dip = m0; g.ok = 0; g.pad = 0;
    if (!(m0 > 0.0) || isinf(m0)) return g;
    int e;
    frexp(m0, &e);                                   // m0 = f 2^e, f in [0.5, 1): binade [2^(e-1), 2^e)
    if (e - 53 < -1000 || e > 1000) return g;        // keep u, 1/u and the scaled masses far from the exponent limits
    g.u = ldexp(1.0, e - 53); g.inv_u = ldexp(1.0, 53 - e); g.top = ldexp(1.0, e); g.ok = 1;
    return g;
}

// rn_u(m) and whether the reference's rounding of T + m could differ from T + rn_u(m)
__device__ __forceinline__ double quantise_mass(double m, const MassGrid& g, bool& unsafe) {
    if (!g.ok) return m;
    if (!(m >= 0.0) || !(m < g.top)) { unsafe = true; return m; }
    const double s = m*g.inv_u;                      // exact (power-of-two scaling, no overflow: m < top)
    const double q = rint(s);
    if (fabs(s - q) == 0.5) unsafe = true;           // tie: the outcome depends on the parity of the running sum
    return q*g.u;
}

}  // namespace rebxb

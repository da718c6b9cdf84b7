// kernels_kepler.cuh -- Kepler drift of a shard about its central body, included by kernels.cu.
//
// The caller either side of the force path under a Wisdom-Holman splitting (reference src/steppers.c:79-87,
// rebx_kepler_step, which hands the work to REBOUND's WHFast): every particle follows its two-body orbit
// about body record 0 (mu = G (m_body + m_i)) for dt while the body itself moves inertially.  With it a
// test-particle ensemble is integrated with Kepler drifts + kicks by the stacked effects without leaving
// HBM (device.Shard.wh_step).  Arithmetic: kepler_core.h, shared verbatim with the host harness.
#pragma once
#include "kepler_core.h"

namespace rebxb {

// Resident CTAs per SM asked of the register allocator for the Kepler kernels.  3 (80 registers, 24 warps per SM) is what
// the allocator chose by itself in round 1; since then it drifted to 94-96 registers (2 CTAs, 16 warps) and the fused
// radiation sweep went 0.390 -> 0.450 ms at 10^7, so the bound is stated.  4 CTAs (64 registers) spill and are slower
// (0.399 ms, measured round 1); the Yarkovsky variant needs 140 registers and keeps the allocator's choice.
#ifndef REBXB_KEPLER_MINBLOCKS
#define REBXB_KEPLER_MINBLOCKS 3
#endif
template <int K> constexpr int kepler_minblocks() { return K == REBX_B200_RADIATION ? REBXB_KEPLER_MINBLOCKS : 1; }

__global__ void __launch_bounds__(kThreads, REBXB_KEPLER_MINBLOCKS)
kepler_drift_kernel(rebx_b200_state st, const double* __restrict__ body, double G, double dt) {
    const long long bi = (long long)body[REBX_B200_BODY_INDEX];
    const double bx = body[REBX_B200_BODY_X+0], by = body[REBX_B200_BODY_X+1], bz = body[REBX_B200_BODY_X+2];
    const double bvx = body[REBX_B200_BODY_VX+0], bvy = body[REBX_B200_BODY_VX+1], bvz = body[REBX_B200_BODY_VX+2];
    const double bm = body[REBX_B200_BODY_M];
    const double nx = bx + dt*bvx, ny = by + dt*bvy, nz = bz + dt*bvz;      // where the body is after its inertial drift
    double* const X = const_cast<double*>(st.x);  double* const Y = const_cast<double*>(st.y);  double* const Z = const_cast<double*>(st.z);
    double* const VX = const_cast<double*>(st.vx); double* const VY = const_cast<double*>(st.vy); double* const VZ = const_cast<double*>(st.vz);
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k == bi) {          // the body's own element (on the shard that owns it)
            X[k] = nx; Y[k] = ny; Z[k] = nz;
            continue;
        }
        double x[3] = {__ldcs(X + k) - bx, __ldcs(Y + k) - by, __ldcs(Z + k) - bz};
        double v[3] = {__ldcs(VX + k) - bvx, __ldcs(VY + k) - bvy, __ldcs(VZ + k) - bvz};
        const double m = st.m ? __ldcs(st.m + k) : 0.0;
        rebxb_kepler_drift(G*(bm + m), dt, x, v);
        __stcs(X + k, nx + x[0]); __stcs(Y + k, ny + x[1]); __stcs(Z + k, nz + x[2]);
        __stcs(VX + k, bvx + v[0]); __stcs(VY + k, bvy + v[1]); __stcs(VZ + k, bvz + v[2]);
    }
}

// Fused Wisdom-Holman "kick + drift" for ONE effect that puts no back-reaction on the central body
// (radiation_forces, yarkovsky_effect): a = effect at the current state (from zero, like the kick of the
// splitting sees it), v += dt_kick a, Kepler drift about the body for dt_drift -- one pass over the state,
// the acceleration never exists in HBM.  Same per-particle arithmetic as rebx_b200_zero_acc +
// rebx_b200_launch_pass + rebx_b200_kick + rebx_b200_kepler_drift.  Without back-reaction the body keeps its
// velocity through the kick, so every CTA can advance its own copy of the body record; the caller refreshes
// the record afterwards (rebx_b200_gather_bodies).
template <int K>
__global__ void __launch_bounds__(kThreads, kepler_minblocks<K>())
wh_kick_drift_kernel(const __grid_constant__ rebx_b200_pass P, double G, double dt_kick, double dt_drift) {
    static_assert(!Traits<K>::red, "the fused kick+drift needs an effect without back-reaction");
    __shared__ double s_body[REBX_B200_BODY_DOUBLES];
    if (threadIdx.x < REBX_B200_BODY_DOUBLES) s_body[threadIdx.x] = P.fx[0].body[threadIdx.x];
    __syncthreads();
    const rebx_b200_state& st = P.st;
    const long long bi = (long long)s_body[REBX_B200_BODY_INDEX];
    const Origin org = load_origin(s_body);
    const Pre pre = prepare<K>(P.fx[0], org, s_body);
    const double nx = org.x + dt_drift*org.vx, ny = org.y + dt_drift*org.vy, nz = org.z + dt_drift*org.vz;
    double* const X = const_cast<double*>(st.x);  double* const Y = const_cast<double*>(st.y);  double* const Z = const_cast<double*>(st.z);
    double* const VX = const_cast<double*>(st.vx); double* const VY = const_cast<double*>(st.vy); double* const VZ = const_cast<double*>(st.vz);
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k == bi) {          // the body: no force on itself, no back-reaction => inertial
            X[k] = nx; Y[k] = ny; Z[k] = nz;
            continue;
        }
        Particle p;
        p.x = __ldcs(X + k); p.y = __ldcs(Y + k); p.z = __ldcs(Z + k);
        p.vx = __ldcs(VX + k); p.vy = __ldcs(VY + k); p.vz = __ldcs(VZ + k);
        p.m = st.m ? __ldcs(st.m + k) : 0.0;
        if constexpr (Traits<K>::rad) { p.r = __ldcs(st.r + k); } else { p.r = 0.0; }
        Params<K, 1> q;
        load_params<K, 1>(P.fx[0], k, true, q);
        Vec3 a = {0.0, 0.0, 0.0}, red = {0.0, 0.0, 0.0};
        unsigned unused = 0;
        apply<K, 1>(P.fx[0], org, s_body, pre, p, q, 0, a, red, unused);
        double v[3] = {(p.vx + dt_kick*a.x) - org.vx, (p.vy + dt_kick*a.y) - org.vy, (p.vz + dt_kick*a.z) - org.vz};
        double x[3] = {p.x - org.x, p.y - org.y, p.z - org.z};
        rebxb_kepler_drift(G*(org.m + p.m), dt_drift, x, v);
        __stcs(X + k, nx + x[0]); __stcs(Y + k, ny + x[1]); __stcs(Z + k, nz + x[2]);
        __stcs(VX + k, org.vx + v[0]); __stcs(VY + k, org.vy + v[1]); __stcs(VZ + k, org.vz + v[2]);
    }
}

}  // namespace rebxb

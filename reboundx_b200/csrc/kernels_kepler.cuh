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

__global__ void __launch_bounds__(kThreads)
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

}  // namespace rebxb

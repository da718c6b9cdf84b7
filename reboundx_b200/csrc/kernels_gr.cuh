// kernels_gr.cuh -- device part of the 1PN `gr` effect (reference src/gr.c:65-164), included by
// kernels.cu.  The reference recomputes Newtonian accelerations from the N_active massive bodies,
// moves to Jacobi coordinates, solves a per-particle fixed point for the Jacobi velocity
// (:114-129), evaluates the post-Newtonian acceleration (:135-148) and transforms back.
//
// Split used here (N_active <= REBX_B200_GR_MAX_ACTIVE = 4096 active bodies -- the host part is O(N_active^2) -- plus any
// number of test particles).  The actives {x, y, z, m} live in a device array; the sweep stages them through shared
// memory in tiles, the pull reduction takes them kGrChunk at a time (one launch per chunk, per-thread accumulators):
//   gr_pull_kernel   sum over test particles j of G m_j (x_i - x_j)/r^3 for each active body i --
//                    what the test particles contribute to the actives' Newtonian acceleration
//                    (gr.c:79-96 with j >= N_active); exact zeros for massless test particles.
//   host             the O(N_active^2) part for the active bodies, in the reference's order, and
//                    the Jacobi origin (centre of mass of the actives: position, velocity,
//                    Newtonian acceleration) for the test particles.
//   gr_sweep_kernel  per test particle: Newtonian pull of the actives (ascending i like the
//                    reference), Jacobi coordinates, fixed-point iteration, 1PN acceleration; a += .
#pragma once

namespace rebxb {

constexpr int kGrChunk = 8;         // actives per launch of the pull reduction (per-thread accumulators in registers)

// partials[block][kGrChunk][3] for the actives first .. first + kGrChunk - 1
__global__ void __launch_bounds__(kThreads)
gr_pull_kernel(rebx_b200_state st, const __grid_constant__ rebx_b200_gr_args A, int first, double* __restrict__ partials) {
    __shared__ double s_red[kWarps][3];
    __shared__ double s_act[kGrChunk][4];
    const int count = (A.n_active - first < kGrChunk) ? A.n_active - first : kGrChunk;
    if (threadIdx.x < count*4) s_act[threadIdx.x/4][threadIdx.x%4] = A.active[(size_t)(first + threadIdx.x/4)*4 + threadIdx.x%4];
    __syncthreads();
    double acc[kGrChunk][3];
#pragma unroll
    for (int i = 0; i < kGrChunk; i++) acc[i][0] = acc[i][1] = acc[i][2] = 0.0;
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t k = (int64_t)blockIdx.x*kThreads + threadIdx.x; k < st.n; k += stride) {
        if (st.index_base + k < A.n_active) continue;
        const double mj = st.m[k];
        if (mj == 0.0) continue;                    // contributes exact zeros
        const double xj = st.x[k], yj = st.y[k], zj = st.z[k];
#pragma unroll
        for (int i = 0; i < kGrChunk; i++) {
            if (i < count) {
                const double dx = s_act[i][0] - xj, dy = s_act[i][1] - yj, dz = s_act[i][2] - zj;
                const double r2 = dx*dx + dy*dy + dz*dz;
                const double r = sqrt(r2);
                const double prefac = A.G/(r2*r);
                acc[i][0] += prefac*mj*dx; acc[i][1] += prefac*mj*dy; acc[i][2] += prefac*mj*dz;
            }
        }
    }
    for (int i = 0; i < count; i++) {
        double v[3] = {acc[i][0], acc[i][1], acc[i][2]};
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v[c] += __shfl_down_sync(0xffffffffu, v[c], off);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = v[0]; s_red[threadIdx.x >> 5][1] = v[1]; s_red[threadIdx.x >> 5][2] = v[2]; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) s += s_red[w][threadIdx.x];
            partials[((size_t)blockIdx.x*kGrChunk + i)*3 + threadIdx.x] = s;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
gr_pull_finalize_kernel(const double* __restrict__ partials, int nblocks, int count, double* __restrict__ pull) {
    __shared__ double s[kThreads];
    for (int q = 0; q < count*3; q++) {
        double acc = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += kThreads) acc += partials[(size_t)b*kGrChunk*3 + q];
        s[threadIdx.x] = acc;
        __syncthreads();
        for (int w = kThreads/2; w > 0; w >>= 1) {
            if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) pull[q] = s[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads)
gr_sweep_kernel(rebx_b200_state st, const __grid_constant__ rebx_b200_gr_args A, int* __restrict__ not_converged) {
    constexpr int kTileActives = 256;
    __shared__ double s_act[kTileActives][4];
    const double C2 = A.C2, mu = A.mu;
    // acceleration of the Jacobi origin: from the host part, or -- one active body -- from the device-resident pull with the
    // host part's own operations (act.a = 0 - pull; s_a = m_0 act.a; com.a = s_a (1/m_0): rebound_shim.c shim_to_jacobi)
    double ca0 = A.com[6], ca1 = A.com[7], ca2 = A.com[8];
    if (A.pull1 != nullptr) {
        const double m0 = A.active[3], mi = 1./m0;
        ca0 = (m0*(0. - A.pull1[0]))*mi; ca1 = (m0*(0. - A.pull1[1]))*mi; ca2 = (m0*(0. - A.pull1[2]))*mi;
    }
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    // every thread of the CTA runs the same number of rounds (the tile loads below are CTA-wide barriers)
    const int64_t rounds = (st.n + stride - 1)/stride;
    for (int64_t round = 0; round < rounds; round++) {
        const int64_t k = round*stride + (int64_t)blockIdx.x*kThreads + threadIdx.x;
        const bool live = k < st.n && st.index_base + k >= A.n_active;       // actives are finished on the host
        const double x = live ? st.x[k] : 0.0, y = live ? st.y[k] : 0.0, z = live ? st.z[k] : 0.0;
        // Newtonian acceleration from the active bodies (gr.c:79-96, role "j"), ascending i like the reference
        double nx = 0.0, ny = 0.0, nz = 0.0;
        for (int base = 0; base < A.n_active; base += kTileActives) {
            const int cnt = (A.n_active - base < kTileActives) ? A.n_active - base : kTileActives;
            if (A.n_active > kTileActives || round == 0) {        // a single tile stays valid for all rounds
                __syncthreads();
                for (int q = threadIdx.x; q < cnt*4; q += kThreads) s_act[q/4][q%4] = A.active[(size_t)base*4 + q];
                __syncthreads();
            }
            if (live) for (int i = 0; i < cnt; i++) {
                const double dx = s_act[i][0] - x, dy = s_act[i][1] - y, dz = s_act[i][2] - z;
                const double r2 = dx*dx + dy*dy + dz*dz;
                const double r = sqrt(r2);
                const double prefac = A.G/(r2*r);
                nx += prefac*s_act[i][3]*dx; ny += prefac*s_act[i][3]*dy; nz += prefac*s_act[i][3]*dz;
            }
        }
        if (!live) continue;
        // Jacobi coordinates of a test particle: relative to the centre of mass of the actives
        const double px = x - A.com[0], py = y - A.com[1], pz = z - A.com[2];
        const double pvx = st.vx[k] - A.com[3], pvy = st.vy[k] - A.com[4], pvz = st.vz[k] - A.com[5];
        const double pax = nx - ca0, pay = ny - ca1, paz = nz - ca2;
        // fixed point for the Jacobi velocity (gr.c:105-129)
        double vix = pvx, viy = pvy, viz = pvz;
        double vi2 = vix*vix + viy*viy + viz*viz;
        const double ri = sqrt(px*px + py*py + pz*pz);
        int q = 0;
        double Aa = (0.5*vi2 + 3.*mu/ri)/C2;
        for (q = 0; q < A.max_iterations; q++) {
            const double ox = vix, oy = viy, oz = viz;
            vix = pvx/(1. - Aa);
            viy = pvy/(1. - Aa);
            viz = pvz/(1. - Aa);
            vi2 = vix*vix + viy*viy + viz*viz;
            Aa = (0.5*vi2 + 3.*mu/ri)/C2;
            const double dvx = vix - ox, dvy = viy - oy, dvz = viz - oz;
            if ((dvx*dvx + dvy*dvy + dvz*dvz)/vi2 < 2.220446049250313e-16*2.220446049250313e-16) break;
        }
        if (q == 10) *not_converged = 1;                        // the reference compares with 10, not max_iterations (gr.c:130-133)
        const double B = (mu/ri - 1.5*vi2)*mu/(ri*ri*ri)/C2;
        const double rdotrdot = px*pvx + py*pvy + pz*pvz;
        const double vdx = pax + B*px, vdy = pay + B*py, vdz = paz + B*pz;
        const double vdotvdot = vix*vdx + viy*vdy + viz*vdz;
        const double D = (vdotvdot - 3.*mu/(ri*ri*ri)*rdotrdot)/C2;
        const double gx = B*(1. - Aa)*px - Aa*pax - D*vix;
        const double gy = B*(1. - Aa)*py - Aa*pay - D*viy;
        const double gz = B*(1. - Aa)*pz - Aa*paz - D*viz;
        // back to inertial: a test particle's acceleration is its Jacobi acceleration plus s_a/eta with
        // s_a = 0 (the Jacobi acceleration of the centre of mass is zeroed, gr.c:151-155)
        st.ax[k] += gx + 0.0;
        st.ay[k] += gy + 0.0;
        st.az[k] += gz + 0.0;
    }
}

}  // namespace rebxb

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
#include "rebx_b200.h"
#include "forces_math.cuh"

namespace rebxb {

constexpr int kThreads   = 256;
constexpr int kMaxBlocks = 4096;
constexpr int kWarps     = kThreads/32;

template <int K> struct Traits;
template <> struct Traits<REBX_B200_NONE>          { static constexpr bool vel=false, mass=false, rad=false, red=false; static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_RADIATION>     { static constexpr bool vel=true,  mass=false, rad=false, red=false; static constexpr int ntab=1; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_HARMONICS>     { static constexpr bool vel=false, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_GR_POTENTIAL>  { static constexpr bool vel=false, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_YARKOVSKY>     { static constexpr bool vel=true,  mass=true,  rad=true,  red=false; static constexpr int ntab=9; static constexpr bool itab=true;  };
template <> struct Traits<REBX_B200_MODIFY_ORBITS> { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=3; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_GAS_DAMPING>   { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=1; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_TYPE_I>        { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };

// ---- vector loads / stores with streaming hints ------------------------------------------
template <int V>
__device__ __forceinline__ void ld(const double* __restrict__ p, int64_t k, bool full, double (&out)[V]) {
    if constexpr (V == 2) {
        if (full) {
            const double2 t = __ldcs(reinterpret_cast<const double2*>(p + k));
            out[0] = t.x; out[1] = t.y;
        } else {
            out[0] = __ldcs(p + k); out[1] = 0.0;
        }
    } else {
        out[0] = __ldcs(p + k);
    }
}
template <int V>
__device__ __forceinline__ void ldi(const int32_t* __restrict__ p, int64_t k, bool full, int (&out)[V]) {
    if constexpr (V == 2) {
        if (full) {
            const int2 t = __ldcs(reinterpret_cast<const int2*>(p + k));
            out[0] = t.x; out[1] = t.y;
        } else {
            out[0] = __ldcs(p + k); out[1] = 2;
        }
    } else {
        out[0] = __ldcs(p + k);
    }
}
template <int V>
__device__ __forceinline__ void st(double* __restrict__ p, int64_t k, bool full, const double (&v)[V]) {
    if constexpr (V == 2) {
        if (full) __stcs(reinterpret_cast<double2*>(p + k), make_double2(v[0], v[1]));
        else      __stcs(p + k, v[0]);
    } else {
        __stcs(p + k, v[0]);
    }
}

template <int K, int V> struct Params {
    double t[Traits<K>::ntab > 0 ? Traits<K>::ntab : 1][V];
    int    it[V];
};

template <int K, int V>
__device__ __forceinline__ void load_params(const rebx_b200_effect& fx, int64_t k, bool full, Params<K, V>& q) {
#pragma unroll
    for (int j = 0; j < Traits<K>::ntab; j++) ld<V>(fx.tab[j], k, full, q.t[j]);
    if constexpr (Traits<K>::itab) ldi<V>(fx.itab, k, full, q.it);
}

template <int K, int V>
__device__ __forceinline__ void apply(const rebx_b200_effect& fx, const double* __restrict__ body,
                                      const Particle& p, const Params<K, V>& q, int v, Vec3& a, Vec3& red) {
    if constexpr (K == REBX_B200_RADIATION) {
        radiation(fx, body, p, q.t[0][v], a);
    } else if constexpr (K == REBX_B200_HARMONICS) {
        harmonics(fx, body, p, a, red);
    } else if constexpr (K == REBX_B200_GR_POTENTIAL) {
        gr_potential(fx, body, p, a, red);
    } else if constexpr (K == REBX_B200_YARKOVSKY) {
        YarkParams yp{q.t[0][v], q.t[1][v], q.t[2][v], q.t[3][v], q.t[4][v], q.t[5][v], q.t[6][v], q.t[7][v], q.t[8][v]};
        yarkovsky(fx, body, p, yp, q.it[v], a);
    } else if constexpr (K == REBX_B200_MODIFY_ORBITS) {
        const Vec3 f = modify_orbits(fx, body, p, q.t[0][v], q.t[1][v], q.t[2][v]);
        com_force_tail(fx, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_GAS_DAMPING) {
        const Vec3 f = gas_damping(fx, body, p, q.t[0][v]);
        com_force_tail(fx, body, p, f, a, red);
    } else if constexpr (K == REBX_B200_TYPE_I) {
        const Vec3 f = type_I_migration(fx, body, p);
        com_force_tail(fx, body, p, f, a, red);
    }
}

template <int K>
__device__ __forceinline__ void reduce_effect(Vec3 red, int e, double (*s_red)[3], double* __restrict__ partials) {
    if constexpr (Traits<K>::red) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            red.x += __shfl_down_sync(0xffffffffu, red.x, off);
            red.y += __shfl_down_sync(0xffffffffu, red.y, off);
            red.z += __shfl_down_sync(0xffffffffu, red.z, off);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0) { s_red[warp][0] = red.x; s_red[warp][1] = red.y; s_red[warp][2] = red.z; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) s += s_red[w][threadIdx.x];
            partials[((size_t)blockIdx.x*REBX_B200_MAX_EFFECTS + e)*3 + threadIdx.x] = s;
        }
    }
}

template <int V, int K0, int K1, int K2, int K3>
__global__ void __launch_bounds__(kThreads)
pass_kernel(const __grid_constant__ rebx_b200_pass P, double* __restrict__ partials) {
    constexpr bool VEL  = Traits<K0>::vel  || Traits<K1>::vel  || Traits<K2>::vel  || Traits<K3>::vel;
    constexpr bool MASS = Traits<K0>::mass || Traits<K1>::mass || Traits<K2>::mass || Traits<K3>::mass;
    constexpr bool RAD  = Traits<K0>::rad  || Traits<K1>::rad  || Traits<K2>::rad  || Traits<K3>::rad;

    __shared__ double s_body[REBX_B200_MAX_EFFECTS][REBX_B200_BODY_DOUBLES];
    __shared__ double s_red[kWarps][3];
    for (int idx = threadIdx.x; idx < P.n_effects*REBX_B200_BODY_DOUBLES; idx += kThreads) {
        const int e = idx/REBX_B200_BODY_DOUBLES, j = idx - e*REBX_B200_BODY_DOUBLES;
        s_body[e][j] = P.fx[e].body[j];
    }
    __syncthreads();
    const long long b0 = (long long)s_body[0][0];
    const long long b1 = (K1 != 0) ? (long long)s_body[1][0] : -1;
    const long long b2 = (K2 != 0) ? (long long)s_body[2][0] : -1;
    const long long b3 = (K3 != 0) ? (long long)s_body[3][0] : -1;

    Vec3 red0 = {0., 0., 0.}, red1 = {0., 0., 0.}, red2 = {0., 0., 0.}, red3 = {0., 0., 0.};

    const int64_t n = P.st.n;
    const int64_t ntiles = (n + V - 1)/V;
    const int64_t stride = (int64_t)gridDim.x*kThreads;
    for (int64_t t = (int64_t)blockIdx.x*kThreads + threadIdx.x; t < ntiles; t += stride) {
        const int64_t k = t*V;
        const bool full = (k + V <= n);
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
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (v > 0 && !full) break;
            Particle p;
            p.x = x[v]; p.y = y[v]; p.z = z[v];
            if constexpr (VEL)  { p.vx = vx[v]; p.vy = vy[v]; p.vz = vz[v]; } else { p.vx = p.vy = p.vz = 0.0; }
            if constexpr (MASS) { p.m = m[v]; } else { p.m = 0.0; }
            if constexpr (RAD)  { p.r = r[v]; } else { p.r = 0.0; }
            Vec3 a = {ax[v], ay[v], az[v]};
            const long long gi = P.st.index_base + k + v;
            if (gi != b0) apply<K0, V>(P.fx[0], s_body[0], p, q0, v, a, red0);
            if constexpr (K1 != 0) { if (gi != b1) apply<K1, V>(P.fx[1], s_body[1], p, q1, v, a, red1); }
            if constexpr (K2 != 0) { if (gi != b2) apply<K2, V>(P.fx[2], s_body[2], p, q2, v, a, red2); }
            if constexpr (K3 != 0) { if (gi != b3) apply<K3, V>(P.fx[3], s_body[3], p, q3, v, a, red3); }
            ax[v] = a.x; ay[v] = a.y; az[v] = a.z;
        }
        st<V>(P.st.ax, k, full, ax); st<V>(P.st.ay, k, full, ay); st<V>(P.st.az, k, full, az);
    }
    reduce_effect<K0>(red0, 0, s_red, partials);
    reduce_effect<K1>(red1, 1, s_red, partials);
    reduce_effect<K2>(red2, 2, s_red, partials);
    reduce_effect<K3>(red3, 3, s_red, partials);
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

// ---- launch plumbing ----------------------------------------------------------------------
typedef void (*pass_fn)(const rebx_b200_pass, double*);
struct Combo { int k[4]; pass_fn fn[2]; bool red[4]; };   // fn[0]: V=1, fn[1]: V=2

#define REBXB_COMBO(a, b, c, d) { {a, b, c, d}, { pass_kernel<1, a, b, c, d>, pass_kernel<2, a, b, c, d> }, \
                                  { Traits<a>::red, Traits<b>::red, Traits<c>::red, Traits<d>::red } },
static const Combo kCombos[] = {
    // singles
    REBXB_COMBO(1,0,0,0) REBXB_COMBO(2,0,0,0) REBXB_COMBO(3,0,0,0) REBXB_COMBO(4,0,0,0)
    REBXB_COMBO(5,0,0,0) REBXB_COMBO(6,0,0,0) REBXB_COMBO(7,0,0,0)
    // config 3: J2/J4 (+) gr_potential, either order; two oblate bodies
    REBXB_COMBO(2,3,0,0) REBXB_COMBO(3,2,0,0) REBXB_COMBO(2,2,0,0)
    // config 4: yarkovsky (+) gr_potential
    REBXB_COMBO(4,3,0,0) REBXB_COMBO(3,4,0,0)
    // radiation stacks
    REBXB_COMBO(1,3,0,0) REBXB_COMBO(3,1,0,0) REBXB_COMBO(1,1,0,0)
    REBXB_COMBO(1,3,2,0) REBXB_COMBO(2,3,1,0)
    // config 5: damping trio in the two natural LIFO orders, and its pairs
    REBXB_COMBO(7,6,5,0) REBXB_COMBO(5,6,7,0)
    REBXB_COMBO(5,6,0,0) REBXB_COMBO(6,5,0,0) REBXB_COMBO(6,7,0,0) REBXB_COMBO(7,6,0,0)
    REBXB_COMBO(5,7,0,0) REBXB_COMBO(7,5,0,0)
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

struct DevInfo { int sms; bool ok; };
static DevInfo device_info() {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info = {0, false};
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return {0, false};
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return {0, false};
        info = {sms, true};
        cached_dev = dev;
    }
    return info;
}

static int grid_for(const void* fn, int64_t work_items) {
    const DevInfo di = device_info();
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t blocks = (work_items + kThreads - 1)/kThreads;
    int64_t cap = (int64_t)(di.ok ? di.sms : 148)*per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks - (kMaxBlocks % (di.ok ? di.sms : 148));
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static bool pass_is_vectorizable(const rebx_b200_pass& P) {
    const rebx_b200_state& s = P.st;
    const void* ptrs[] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.m, s.r, s.ax, s.ay, s.az};
    for (const void* p : ptrs) if (p && !aligned16(p)) return false;
    for (int e = 0; e < P.n_effects; e++) {
        for (int j = 0; j < REBX_B200_MAX_TABLES; j++) if (P.fx[e].tab[j] && !aligned16(P.fx[e].tab[j])) return false;
        if (P.fx[e].itab && (reinterpret_cast<uintptr_t>(P.fx[e].itab) & 7u)) return false;
    }
    return true;
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
            REBXB_CASE(1) REBXB_CASE(2) REBXB_CASE(3) REBXB_CASE(4) REBXB_CASE(5) REBXB_CASE(6) REBXB_CASE(7)
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
    return (size_t)kMaxBlocks*REBX_B200_MAX_EFFECTS*3*sizeof(double);
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
    const int vsel = pass_is_vectorizable(*pass) ? 1 : 0;
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
        sub.reserved = 0;
        for (int j = 0; j < REBX_B200_MAX_EFFECTS; j++) {
            if (j < take) sub.fx[j] = pass->fx[done + j];
            else { sub.fx[j] = rebx_b200_effect(); sub.fx[j].kind = 0; }
        }
        const int V = vsel ? 2 : 1;
        pass_fn fn = combo->fn[vsel];
        const int64_t tiles = (sub.st.n + V - 1)/V;
        const int grid = grid_for(reinterpret_cast<const void*>(fn), tiles);
        fn<<<grid, kThreads, 0, s>>>(sub, static_cast<double*>(workspace));
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        if (launches) (*launches)++;
        unsigned red_mask = 0;
        bool any = false;
        for (int j = 0; j < take; j++) if (combo->red[j]) { red_mask |= 1u << j; if (sub.fx[j].backreaction) any = true; }
        if (any) {
            finalize_kernel<<<1, kThreads, 0, s>>>(sub, static_cast<const double*>(workspace), grid, red_mask);
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
            if (launches) (*launches)++;
        }
        done += take;
    }
    return 0;
}

int rebx_b200_apply_backreaction(const rebx_b200_pass* pass, void* stream, int* launches) {
    if (!pass) return cudaErrorInvalidValue;
    bool any = false;
    for (int e = 0; e < pass->n_effects; e++) if (pass->fx[e].backreaction) any = true;
    if (!any) return 0;
    apply_backreaction_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*pass);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && launches) (*launches)++;
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

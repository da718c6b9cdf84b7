// pass_common.cuh -- pieces shared by the sweep kernels of kernels.cu (reference-exact arithmetic, -fmad=false) and
// kernels_tol.cu (tolerance-mode arithmetic, FMA on): effect traits, vector loads / stores with streaming hints,
// per-particle parameter loads, the back-reaction reduction, grid sizing.  Header-only; every translation unit
// gets its own copy (no relocatable device code).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <utility>
#include <vector>
#include "rebx_b200.h"

namespace rebxb {

#ifndef REBXB_HEAVY_MINBLOCKS
#define REBXB_HEAVY_MINBLOCKS 2
#endif
#ifndef REBXB_THREADS
#define REBXB_THREADS 256
#endif
constexpr int kThreads   = REBXB_THREADS;    // threads per CTA of the sweeps (128 was measured too)
constexpr int kMaxBlocks = 4096;
constexpr int kWarps     = kThreads/32;

// ---- particle / body records held in registers ------------------------------------------------
struct Particle { double x, y, z, vx, vy, vz, m, r; };
// Position, velocity and mass of the body an effect refers to, held in REGISTERS: stacked effects that
// share one central body then share one Origin, and the compiler can merge their common sub-expressions
// (values re-read from shared memory after every libdevice / division slow-path call could not be).
struct Origin { double x, y, z, vx, vy, vz, m; };
__device__ __forceinline__ Origin load_origin(const double* __restrict__ rec) {
    return Origin{rec[REBX_B200_BODY_X+0], rec[REBX_B200_BODY_X+1], rec[REBX_B200_BODY_X+2],
                  rec[REBX_B200_BODY_VX+0], rec[REBX_B200_BODY_VX+1], rec[REBX_B200_BODY_VX+2], rec[REBX_B200_BODY_M]};
}
struct Vec3     { double x, y, z; };

__device__ __forceinline__ bool is_absent(double v) {
    return (unsigned long long)__double_as_longlong(v) == REBX_B200_ABSENT_BITS;
}

template <int K> struct Traits;
template <> struct Traits<REBX_B200_NONE>          { static constexpr bool vel=false, mass=false, rad=false, red=false; static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_RADIATION>     { static constexpr bool vel=true,  mass=false, rad=false, red=false; static constexpr int ntab=1; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_HARMONICS>     { static constexpr bool vel=false, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_GR_POTENTIAL>  { static constexpr bool vel=false, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_YARKOVSKY>     { static constexpr bool vel=true,  mass=true,  rad=true,  red=false; static constexpr int ntab=9; static constexpr bool itab=true;  };
template <> struct Traits<REBX_B200_MODIFY_ORBITS> { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=3; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_GAS_DAMPING>   { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=1; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_TYPE_I>        { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_CENTRAL_FORCE> { static constexpr bool vel=false, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_LENSE_THIRRING> { static constexpr bool vel=true, mass=true,  rad=false, red=true;  static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_GAS_DF>        { static constexpr bool vel=true,  mass=true,  rad=true,  red=false; static constexpr int ntab=0; static constexpr bool itab=false; };
template <> struct Traits<REBX_B200_EXP_MIGRATION> { static constexpr bool vel=true,  mass=true,  rad=false, red=true;  static constexpr int ntab=3; static constexpr bool itab=false; };

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

template <int K, int NWARPS = kWarps>
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
            for (int w = 0; w < NWARPS; w++) s += s_red[w][threadIdx.x];
            partials[((size_t)blockIdx.x*REBX_B200_MAX_EFFECTS + e)*3 + threadIdx.x] = s;
        }
    }
}

// ---- launch plumbing ----------------------------------------------------------------------
struct DevInfo { int sms; bool ok; };
static inline DevInfo device_info() {
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

static inline int grid_for(const void* fn, int64_t work_items, int threads = kThreads, int smem = 0, int items_per_block = kThreads) {
    const DevInfo di = device_info();
    // resident CTAs per SM of this kernel: queried once per (kernel, launch shape) and thread
    struct Key { const void* fn; int threads, smem; };
    static thread_local std::vector<std::pair<Key, int>> cache;
    int per_sm = 0;
    for (const auto& kv : cache) if (kv.first.fn == fn && kv.first.threads == threads && kv.first.smem == smem) { per_sm = kv.second; break; }
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache.push_back({Key{fn, threads, smem}, per_sm});
    }
    int64_t blocks = (work_items + items_per_block - 1)/items_per_block;
    int64_t cap = (int64_t)(di.ok ? di.sms : 148)*per_sm;
    if (cap > kMaxBlocks) cap = kMaxBlocks - (kMaxBlocks % (di.ok ? di.sms : 148));
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline bool pass_is_vectorizable(const rebx_b200_pass& P) {
    const rebx_b200_state& s = P.st;
    const void* ptrs[] = {s.x, s.y, s.z, s.vx, s.vy, s.vz, s.m, s.r, s.ax, s.ay, s.az};
    for (const void* p : ptrs) if (p && !aligned16(p)) return false;
    for (int e = 0; e < P.n_effects; e++) {
        for (int j = 0; j < REBX_B200_MAX_TABLES; j++) if (P.fx[e].tab[j] && !aligned16(P.fx[e].tab[j])) return false;
        if (P.fx[e].itab && (reinterpret_cast<uintptr_t>(P.fx[e].itab) & 7u)) return false;
    }
    return true;
}


}  // namespace rebxb

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

// ---- software prefetch of the NEXT grid-stride iteration's operands (an option that did NOT pay) ---------------------
// The FP64-heavy sweeps hold 4-5 warps per scheduler (96-128 registers), each a dependent chain of several hundred
// FP64 instructions per particle, and ncu puts 18-22 % of the trio sweep's stall samples on the first use of the ~14
// operands loaded at the top of an iteration.  Register double-buffering would cost ~28 registers the kernels do not
// have; a prefetch into L2 / L1 (CCTL.E.PF2 / PF1 in the SASS) costs none.  MEASURED on a B200
// (tools/prefetch_variants.sh + prefetch_sweep.sh, profiles/r02_prefetch_sweep.txt): no gain anywhere -- ring exact
// 0.2306 / 0.2292 / 0.2290 ms (off / L1 / L2), asteroids exact 0.0709 / 0.0709 / 0.0707, trio exact 0.1031 / 0.1033 /
// 0.1027, trio tolerance 0.0584 / 0.0598 / 0.0599 (slower: 14 more address computations per iteration).  The exposed
// load is one of many serial waits of a warp, and with 4 warps per scheduler the others fill the gap either way.
// REBXB_PREFETCH: 0 off (default), 1 = prefetch.global.L1, 2 = prefetch.global.L2.
#ifndef REBXB_PREFETCH
#define REBXB_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_next(const void* p) {
#if REBXB_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#elif REBXB_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;
#endif
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

template <int K>
__device__ __forceinline__ void prefetch_params(const rebx_b200_effect& fx, int64_t k) {
#pragma unroll
    for (int j = 0; j < Traits<K>::ntab; j++) prefetch_next(fx.tab[j] + k);
    if constexpr (Traits<K>::itab) prefetch_next(fx.itab + k);
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

// ---- back-reaction sums finished inside the sweep -----------------------------------------------------------------
// The sweeps used to be followed by a one-CTA finalize launch (sum of the CTA partials) and, for a resident caller, a
// one-thread launch adding the sums to the body: 2 x ~4 us of launch + drain behind sweeps that take 40-230 us.  Now
// every CTA takes a ticket when its partials are out (threadfence + atomicInc on a counter in the workspace); the CTA
// that draws the last ticket sums the partials -- in the ORDER finalize_kernel uses (256 virtual lanes striding over the
// CTAs, then a binary tree), so the sums keep their bits for a given grid whatever the CTA size -- writes
// fx[e].backreaction and, when the pass asks for it (REBX_B200_PASS_APPLY_BACKREACTION), adds the sums to the body's
// acceleration like apply_backreaction_kernel.  atomicInc wraps the counter back to 0 on the last ticket, so a
// workspace that was zeroed once after allocation stays ready for every later launch.
//
// The whole epilogue is ONE pass for all effects and components (ncu at one iteration per thread: the epilogue of
// three reduce_effect calls, four fences and nine serial 256-lane reductions was 40 % of a 24 us launch, and
// tools/size_scaling.py put the fixed cost of a trio sweep at 25 us of its 56): one shuffle tree + one barrier for the
// CTA's partials, one fence, and the last CTA reduces all (effect, component) pairs side by side -- every pair in
// finalize_kernel's own order, so the bits do not change.
constexpr int kFinalLanes = 256;
constexpr size_t kTicketOffset = (size_t)kMaxBlocks*REBX_B200_MAX_EFFECTS*3 + 8;      // in doubles, behind the scalar tail

// CTA-uniform: true in the CTA that finishes last.  Called by ALL threads of every CTA of the grid, after the CTA's
// last global store.
__device__ __forceinline__ bool last_cta_ticket(double* partials) {
    __shared__ unsigned s_last;
    __syncthreads();                                    // this CTA's stores (partials, accelerations) are issued
    if (threadIdx.x == 0) {
        unsigned* ticket = reinterpret_cast<unsigned*>(partials + kTicketOffset);
        __threadfence();                                // cumulative: covers what the CTA's other threads stored before the barrier
        s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    return true;
}

// apply_backreaction_kernel's loop, by ONE thread: effects in stack order, only bodies inside this shard.  The whole
// grid is waiting for this thread, so every load it needs is issued before the first one is used (body indices and
// sums side by side), and consecutive effects on the same body -- the usual case: one central body -- share one
// read-modify-write of its acceleration.  The additions happen in stack order, as in the kernel.
static __device__ __noinline__ void apply_backreaction_inline(const rebx_b200_pass& P) {
    long long kk[REBX_B200_MAX_EFFECTS];
    double br[REBX_B200_MAX_EFFECTS][3];
#pragma unroll
    for (int e = 0; e < REBX_B200_MAX_EFFECTS; e++) {
        kk[e] = -1;
        br[e][0] = br[e][1] = br[e][2] = 0.0;
        if (e < P.n_effects && P.fx[e].backreaction != nullptr) {
            const long long bi = (long long)__ldcg(P.fx[e].body + REBX_B200_BODY_INDEX);
            const long long k = bi - P.st.index_base;
            if (bi >= 0 && k >= 0 && k < P.st.n) kk[e] = k;
            br[e][0] = __ldcg(P.fx[e].backreaction + 0); br[e][1] = __ldcg(P.fx[e].backreaction + 1); br[e][2] = __ldcg(P.fx[e].backreaction + 2);
        }
    }
    long long cur = -1;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int e = 0; e < REBX_B200_MAX_EFFECTS; e++) {
        if (kk[e] < 0) continue;
        if (kk[e] != cur) {
            if (cur >= 0) { P.st.ax[cur] = a0; P.st.ay[cur] = a1; P.st.az[cur] = a2; }
            cur = kk[e];
            a0 = __ldcg(P.st.ax + cur); a1 = __ldcg(P.st.ay + cur); a2 = __ldcg(P.st.az + cur);
        }
        a0 += br[e][0]; a1 += br[e][1]; a2 += br[e][2];
    }
    if (cur >= 0) { P.st.ax[cur] = a0; P.st.ay[cur] = a1; P.st.az[cur] = a2; }
}

// All (effect, component) pairs of the stack at once, by the last CTA; out of line (one CTA per launch runs it, and it
// must not weigh on the register allocation of the sweep's loop).  s_fin: [12][kFinalLanes] doubles.
template <int T, int K0, int K1, int K2, int K3>
__device__ __noinline__ void finalize_sums(const rebx_b200_pass& P, const double* partials, double* s_fin) {
    constexpr bool R[4] = {Traits<K0>::red, Traits<K1>::red, Traits<K2>::red, Traits<K3>::red};
    const int nblocks = gridDim.x;
    bool want[4];
#pragma unroll
    for (int e = 0; e < 4; e++) want[e] = R[e] && P.fx[e].backreaction != nullptr;
    for (int lane = threadIdx.x; lane < kFinalLanes; lane += T) {
        double acc[12];
#pragma unroll
        for (int q = 0; q < 12; q++) acc[q] = 0.0;
        for (int b = lane; b < nblocks; b += kFinalLanes) {
            const double* row = partials + (size_t)b*REBX_B200_MAX_EFFECTS*3;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (!R[e]) continue;
                if (want[e]) { acc[e*3 + 0] += __ldcg(row + e*3 + 0); acc[e*3 + 1] += __ldcg(row + e*3 + 1); acc[e*3 + 2] += __ldcg(row + e*3 + 2); }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (!R[e]) continue;
#pragma unroll
            for (int c = 0; c < 3; c++) s_fin[(e*3 + c)*kFinalLanes + lane] = acc[e*3 + c];
        }
    }
    __syncthreads();
    for (int w = kFinalLanes/2; w > 0; w >>= 1) {
        for (int lane = threadIdx.x; lane < w; lane += T) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (!R[e]) continue;
#pragma unroll
                for (int c = 0; c < 3; c++) s_fin[(e*3 + c)*kFinalLanes + lane] += s_fin[(e*3 + c)*kFinalLanes + lane + w];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 12) {
        const int e = threadIdx.x/3, c = threadIdx.x - e*3;
        if (want[e]) P.fx[e].backreaction[c] = s_fin[threadIdx.x*kFinalLanes];
    }
    __syncthreads();
}

// The end of a sweep, called by ALL threads of the CTA with the thread's running sums: CTA partials, ticket, and in
// the last CTA the sums and `tail(P)` (thread 0).  `always` (uniform over the grid) = take the ticket even when no
// effect of the stack wants its sum (all fx[].backreaction NULL), because the tail has other work.  s_red: [W][12].
template <int T, int K0, int K1, int K2, int K3, class Tail>
__device__ __forceinline__ void finish_sums(const Vec3& red0, const Vec3& red1, const Vec3& red2, const Vec3& red3,
                                            const rebx_b200_pass& P, double* partials, Tail tail, bool always = false) {
    constexpr bool R[4] = {Traits<K0>::red, Traits<K1>::red, Traits<K2>::red, Traits<K3>::red};
    constexpr bool RED = R[0] || R[1] || R[2] || R[3];
    // A stack without sums (radiation_forces, yarkovsky_effect) takes no ticket at all: those sweeps run at the
    // 64-register cap and the tail's presence alone cost them 3.5x the spill traffic (ptxas -v), 0.195 -> 0.251 ms
    // on the fused leapfrog step.  Their callers keep the separate one-thread launch for the body.
    if constexpr (!RED) return;
    constexpr int W = T/32;
    __shared__ double s_red[W][12];
    __shared__ double s_fin[RED ? 12*kFinalLanes : 1];
    double v[12] = {red0.x, red0.y, red0.z, red1.x, red1.y, red1.z, red2.x, red2.y, red2.z, red3.x, red3.y, red3.z};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (!R[e]) continue;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double x = v[e*3 + c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0) s_red[warp][e*3 + c] = x;
        }
    }
    __syncthreads();
    if (threadIdx.x < 12 && R[threadIdx.x/3]) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < W; w++) s += s_red[w][threadIdx.x];
        partials[(size_t)blockIdx.x*REBX_B200_MAX_EFFECTS*3 + threadIdx.x] = s;
    }
    bool any = always;
#pragma unroll
    for (int e = 0; e < 4; e++) if (R[e]) any |= P.fx[e].backreaction != nullptr;
    if (!any) return;                                   // uniform over the grid: nobody takes a ticket
    if (!last_cta_ticket(partials)) return;
    finalize_sums<T, K0, K1, K2, K3>(P, partials, s_fin);
    if (threadIdx.x == 0) tail(P);
}
// the force sweeps' tail: add the sums to the body when the pass asks for it.  Every other CTA stored its
// accelerations before taking its ticket, the body's own element included (unchanged).
struct ApplyTail {
    __device__ __forceinline__ void operator()(const rebx_b200_pass& P) const {
        if (P.reserved & REBX_B200_PASS_APPLY_BACKREACTION) apply_backreaction_inline(P);
    }
};

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

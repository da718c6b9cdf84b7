// dispatch.cpp -- the force dispatcher and its host<->device staging.
//
// Replaces, behind the unchanged plug-in signature (reference src/reboundx.h:234), the
// reference's per-evaluation work:
//   * rebx_additional_forces (src/core.c:1026-1041): walk the LIFO force list, warn, call
//     each force.  Here maximal runs of device-backed forces become ONE fused sweep.
//   * the per-particle rebx_get_param list walks inside every effect (src/core.c:725-746):
//     replaced by SoA parameter tables in HBM, rebuilt only when the owning extras'
//     generation counter says a parameter may have changed.
//   * the per-particle arithmetic: CUDA kernels behind include/rebx_b200.h.
// There is NO host arithmetic fallback: without a CUDA device every built-in force raises
// a REBOUND error and leaves the accelerations untouched.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "internal.h"
#include "transformations.h"     // REBOUND's Jacobi transforms, as the reference's src/gr.c:61 includes them

namespace rebxh {

constexpr int KIND_GR = 100;     // host-orchestrated effect, not a rebx_b200_launch_pass kind

// ---- small RAII helpers ------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    // zero-filled when (re)allocated: what rebx_b200_launch_pass asks of its workspace (ticket counter)
    cudaError_t ensure_zeroed(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        cudaError_t e = ensure(bytes);
        return e == cudaSuccess ? cudaMemset(p, 0, cap) : e;
    }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes; else p = nullptr;
        return e;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
};

static inline bool is_absent(double v) {
    unsigned long long bits;
    std::memcpy(&bits, &v, sizeof bits);
    return bits == REBX_B200_ABSENT_BITS;
}

static double absent_value() {
    const unsigned long long bits = REBX_B200_ABSENT_BITS;
    double d;
    std::memcpy(&d, &bits, sizeof d);
    return d;
}

// ---- SoA parameter tables of one force ------------------------------------------------------
struct Oblate { double J2, J4, Req, u[3], v[3], w[3]; };

struct ForceTables {
    int kind = 0;
    std::uint64_t gen = 0, epoch = 0;
    size_t N = 0;
    bool valid = false;
    std::vector<double> h[REBX_B200_MAX_TABLES];
    std::vector<int32_t> hi;
    // device copies: one slice [lo, hi) of every table per device the segment is spread over (index = slot in
    // DeviceContext::devs; a single-device run holds the whole range in slot 0)
    struct Slice { DevBuf d[REBX_B200_MAX_TABLES]; DevBuf di; size_t lo = 0, hi = 0; bool uploaded = false; };
    std::vector<std::unique_ptr<Slice>> dev;
    std::vector<int64_t> bodies;     // RADIATION: sources; HARMONICS: oblate bodies; damping trio: first "primary"
    std::vector<Oblate> oblate;
    int64_t n_incomplete = 0;        // particles whose parameter set the reference would report as incomplete
};

struct Stats {
    std::uint64_t calls = 0, launches = 0, h2d_bytes = 0, d2h_bytes = 0, table_builds = 0;
};

// What one device needs to run its share of a segment: pipeline streams and staging buffers.
struct DevSlot {
    static constexpr int kBuffers = 3;      // chunks in flight: H2D of c+1 and D2H of c-1 overlap the sweep of c
    int device = 0;
    cudaStream_t stream[kBuffers] = {nullptr, nullptr, nullptr};
    cudaEvent_t bodies_ready = nullptr;
    DevBuf aos[kBuffers], soa[kBuffers], work[kBuffers], bodies, br;
    PinnedBuf br_h;
    cudaError_t init(int dev) {
        device = dev;
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) return e;
        for (auto& s : stream) if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return e;
        return cudaEventCreateWithFlags(&bodies_ready, cudaEventDisableTiming);
    }
    ~DevSlot() {
        cudaSetDevice(device);
        for (auto& s : stream) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
        if (bodies_ready) cudaEventDestroy(bodies_ready);
        cudaGetLastError();
    }
};

struct DeviceContext {
    static constexpr int kBuffers = DevSlot::kBuffers;
    int device = 0;                         // primary device (devs[0]): zero-copy path, centre-of-mass frames, gr, potentials
    bool ready = false;
    // devs[0] is the primary; REBX_B200_DEVICES=all|<count>|<list> adds the others: a large PARTICLE-frame segment is
    // then split by index range over all of them (one process, one host thread issuing asynchronous work per device)
    std::vector<std::unique_ptr<DevSlot>> devs;
    DevBuf scan, totals, grbuf;
    PinnedBuf bodies_h, gr_h;
    std::unordered_map<struct rebx_force*, std::unique_ptr<ForceTables>> tables;
    void* registered_ptr = nullptr;
    size_t registered_bytes = 0;
    void* registered_dev = nullptr;         // device alias of the registered array (zero-copy path of small ensembles)
    size_t zero_copy_max = 3072;            // particles (measured crossover with pinned DMA: ~3000); REBX_B200_ZEROCOPY_MAX
    size_t multi_min = size_t(1) << 20;     // smallest ensemble worth spreading over several devices; REBX_B200_MULTI_MIN
    Stats stats;
    size_t chunk = size_t(1) << 19;         // 64 MB of AoS records per pipeline stage
    bool chunk_from_env = false;
    bool paranoid = false;
    bool tolerance = false;                 // REBX_B200_MATH=tolerance: kernels_tol.cu for the stacks it covers
    int d2h_width = 0;                      // REBX_B200_D2H=acc|accm: device->host as a 2-D copy of 24 (ax ay az) or 32 (+ m) bytes per record; 0 = whole records

    void unregister_host() {
        if (registered_ptr) { cudaHostUnregister(registered_ptr); cudaGetLastError(); }
        registered_ptr = nullptr; registered_bytes = 0; registered_dev = nullptr;
    }
    ~DeviceContext() {
        unregister_host();
        if (ready) cudaSetDevice(device);
        tables.clear();
        devs.clear();
    }
};

static void delete_ctx(DeviceContext* c) { delete c; }

// ---- kinds ------------------------------------------------------------------------------------
static int kind_of(const struct rebx_force* f) {
    auto fn = f->update_accelerations;
    if (fn == rebx_radiation_forces) return REBX_B200_RADIATION;
    if (fn == rebx_gravitational_harmonics) return REBX_B200_HARMONICS;
    if (fn == rebx_gr_potential) return REBX_B200_GR_POTENTIAL;
    if (fn == rebx_yarkovsky_effect) return REBX_B200_YARKOVSKY;
    if (fn == rebx_modify_orbits_forces) return REBX_B200_MODIFY_ORBITS;
    if (fn == rebx_gas_damping_timescale) return REBX_B200_GAS_DAMPING;
    if (fn == rebx_modify_orbits_with_type_I_migration) return REBX_B200_TYPE_I;
    if (fn == rebx_gr) return KIND_GR;
    if (fn == rebx_central_force) return REBX_B200_CENTRAL_FORCE;
    if (fn == rebx_exponential_migration) return REBX_B200_EXP_MIGRATION;
    if (fn == rebx_lense_thirring) return REBX_B200_LENSE_THIRRING;
    if (fn == rebx_gas_dynamical_friction) return REBX_B200_GAS_DF;
    return REBX_B200_NONE;
}

// ---- table builder (host) ----------------------------------------------------------------------
struct Want {
    const char* name;        // interned
    double* dst = nullptr;   // double table, or
    int32_t* idst = nullptr; // int table (value), with ipresent marking presence
    uint8_t* present = nullptr;
};

static inline bool same_name(const char* a, const char* b) {
    return a == b || (a[0] == b[0] && std::strcmp(a, b) == 0);
}

// uvw(): reference src/gravitational_harmonics.c:104-134 (body-fixed frame from the spin vector)
static void spin_frame(const struct reb_vec3d Omega, Oblate& o) {
    const double omega2 = Omega.x*Omega.x + Omega.y*Omega.y + Omega.z*Omega.z;
    const double omega = std::sqrt(omega2);
    const double sx = Omega.x/omega, sy = Omega.y/omega, sz = Omega.z/omega;
    o.w[0] = sx; o.w[1] = sy; o.w[2] = sz;
    const double fac = std::sqrt(sx*sx + sy*sy);
    if (fac != 0.0) { o.u[0] = -sy/fac; o.u[1] = sx/fac; o.u[2] = 0.0; }
    else            { o.u[0] = 1.0;     o.u[1] = 0.0;    o.u[2] = 0.0; }
    o.v[0] = -(o.u[1]*o.w[2] - o.u[2]*o.w[1]);
    o.v[1] = -(o.u[2]*o.w[0] - o.u[0]*o.w[2]);
    o.v[2] = -(o.u[0]*o.w[1] - o.u[1]*o.w[0]);
}

static void build_tables(ForceTables& T, int kind, struct rebx_force* force, const struct reb_particle* particles, size_t N) {
    T.kind = kind;
    T.N = N;
    T.bodies.clear();
    T.oblate.clear();
    T.n_incomplete = 0;
    for (auto& sl : T.dev) if (sl) sl->uploaded = false;
    const double ABSENT = absent_value();

    std::vector<const char*> dnames;     // double tables in tab[] order
    const char* flag_name = nullptr;     // presence-only int parameter collected into T.bodies
    const char* ival_name = nullptr;     // int parameter whose value is tabulated
    switch (kind) {
        case REBX_B200_RADIATION:     dnames = {"beta"}; flag_name = "radiation_source"; break;
        case REBX_B200_YARKOVSKY:     dnames = {"ye_body_density", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo",
                                                "ye_emissivity", "ye_k", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z"};
                                      ival_name = "ye_flag"; break;
        case REBX_B200_MODIFY_ORBITS: dnames = {"tau_a", "tau_e", "tau_inc"}; flag_name = "primary"; break;
        case REBX_B200_GAS_DAMPING:   dnames = {"d_factor"}; flag_name = "primary"; break;
        case REBX_B200_TYPE_I:        flag_name = "primary"; break;
        case REBX_B200_EXP_MIGRATION: dnames = {"em_tau_a", "em_aini", "em_afin"}; flag_name = "primary"; break;
        default: break;
    }
    for (auto& n : dnames) n = intern_name(n);
    if (flag_name) flag_name = intern_name(flag_name);
    if (ival_name) ival_name = intern_name(ival_name);
    const char* nJ2 = intern_name("J2"); const char* nJ4 = intern_name("J4");
    const char* nReq = intern_name("R_eq"); const char* nOmega = intern_name("Omega");
    const char* nAc = intern_name("Acentral"); const char* nGc = intern_name("gammacentral");

    const size_t nd = dnames.size();
    for (size_t j = 0; j < REBX_B200_MAX_TABLES; j++) { if (j < nd) T.h[j].assign(N, ABSENT); else { T.h[j].clear(); T.h[j].shrink_to_fit(); } }
    std::vector<uint8_t> ipresent;
    if (ival_name) { T.hi.assign(N, 2); ipresent.assign(N, 0); } else { T.hi.clear(); }

    unsigned hw = std::thread::hardware_concurrency();
    std::vector<std::vector<int64_t>> flagged(std::max(1u, std::min(hw ? hw : 1u, 32u)) + 1);
    std::vector<std::vector<std::pair<int64_t, Oblate>>> obl(flagged.size());

    parallel_ranges(N, [&](int t, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            const struct rebx_node* node = static_cast<const struct rebx_node*>(particles[i].ap);
            if (!node) continue;
            unsigned found = 0;            // bit j: double table j ; bit 30: flag ; bit 29: ival
            const double* J2 = nullptr; const double* J4 = nullptr; const double* Req = nullptr; const struct reb_vec3d* Om = nullptr;
            for (; node; node = node->next) {
                const struct rebx_param* p = static_cast<const struct rebx_param*>(node->object);
                const char* nm = p->name;
                bool matched = false;
                for (size_t j = 0; j < nd && !matched; j++) {
                    if (same_name(nm, dnames[j])) {
                        matched = true;
                        if (!(found & (1u << j)) && p->value) { T.h[j][i] = *static_cast<const double*>(p->value); }
                        found |= 1u << j;            // newest entry wins, like the reference's first-match walk
                    }
                }
                if (matched) continue;
                if (flag_name && same_name(nm, flag_name)) {
                    if (!(found & (1u << 30)) && p->value) flagged[t].push_back((int64_t)i);
                    found |= 1u << 30;
                } else if (ival_name && same_name(nm, ival_name)) {
                    if (!(found & (1u << 29)) && p->value) { T.hi[i] = *static_cast<const int*>(p->value); ipresent[i] = 1; }
                    found |= 1u << 29;
                } else if (kind == REBX_B200_CENTRAL_FORCE) {
                    // Acentral / gammacentral ride in the J2 / J4 slots of the collected record
                    if (!J2 && same_name(nm, nAc)) J2 = static_cast<const double*>(p->value);
                    else if (!J4 && same_name(nm, nGc)) J4 = static_cast<const double*>(p->value);
                } else if (kind == REBX_B200_HARMONICS) {
                    if (!J2 && same_name(nm, nJ2)) J2 = static_cast<const double*>(p->value);
                    else if (!J4 && same_name(nm, nJ4)) J4 = static_cast<const double*>(p->value);
                    else if (!Req && same_name(nm, nReq)) Req = static_cast<const double*>(p->value);
                    else if (!Om && same_name(nm, nOmega)) Om = static_cast<const struct reb_vec3d*>(p->value);
                }
            }
            if (kind == REBX_B200_CENTRAL_FORCE && J2 && J4) {
                // src/central_force.c:88-95: a central body has BOTH Acentral and gammacentral
                Oblate o;
                std::memset(&o, 0, sizeof o);
                o.J2 = *J2; o.J4 = *J4;
                obl[t].push_back({(int64_t)i, o});
            }
            if (kind == REBX_B200_HARMONICS && J2 && *J2 != 0.0 && Req) {
                // src/gravitational_harmonics.c:141-159: J2 set and non-zero, R_eq set; J4 and Omega optional
                Oblate o;
                o.J2 = *J2; o.J4 = J4 ? *J4 : 0.0; o.Req = *Req;
                struct reb_vec3d Omega = {0.0, 0.0, 1.0};
                if (Om) Omega = *Om;
                spin_frame(Omega, o);
                obl[t].push_back({(int64_t)i, o});
            }
        }
    });
    for (auto& v : flagged) T.bodies.insert(T.bodies.end(), v.begin(), v.end());
    if (kind == REBX_B200_HARMONICS || kind == REBX_B200_CENTRAL_FORCE) {
        for (auto& v : obl) for (auto& pr : v) { T.bodies.push_back(pr.first); T.oblate.push_back(pr.second); }
    }

    if (kind == REBX_B200_YARKOVSKY) {
        // src/yarkovsky_effect.c:244 (eligibility) and :127-130 (full version needs every parameter)
        const bool have_sb = find_value(force->ap, "ye_stef_boltz") != nullptr;
        int64_t incomplete = 0;
        for (size_t i = 0; i < N; i++) {
            int code = 2;
            const bool eligible = ipresent[i] && !is_absent(T.h[0][i]) && !is_absent(T.h[3][i]);
            if (eligible) {
                const int f = T.hi[i];
                if (f == 1 || f == -1) code = f;
                else if (f == 0) {
                    bool all = have_sb;
                    for (int j : {1, 2, 4, 5, 6, 7, 8}) if (is_absent(T.h[j][i])) all = false;
                    if (all) code = 0;
                    else if (particles[i].r != 0 && i > 0) incomplete++;
                }
                // any other flag value is undefined behaviour in the reference (uninitialised
                // magnitude, yarkovsky_effect.c:104-124): treated as "skip" here.
            }
            T.hi[i] = code;
        }
        T.n_incomplete = incomplete;
    }
    if (kind == REBX_B200_GAS_DAMPING) {
        int64_t missing = 0;
        for (size_t i = 0; i < N; i++) if (is_absent(T.h[0][i])) missing++;
        T.n_incomplete = missing;      // refined at dispatch (the reference body itself is exempt)
    }
    T.valid = true;
}

// Uploads rows [lo, hi) of every table of T to the CURRENT device as slice `slot` (no-op when that slice is up to date).
static cudaError_t upload_tables(ForceTables& T, Stats& st, size_t slot, size_t lo, size_t hi) {
    if (T.dev.size() <= slot) T.dev.resize(slot + 1);
    if (!T.dev[slot]) T.dev[slot] = std::make_unique<ForceTables::Slice>();
    ForceTables::Slice& S = *T.dev[slot];
    if (S.uploaded && S.lo == lo && S.hi == hi) return cudaSuccess;
    const size_t n = hi - lo;
    for (size_t j = 0; j < REBX_B200_MAX_TABLES; j++) {
        if (T.h[j].empty() || n == 0) continue;
        cudaError_t e = S.d[j].ensure(n*sizeof(double));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(S.d[j].p, T.h[j].data() + lo, n*sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        st.h2d_bytes += n*sizeof(double);
    }
    if (!T.hi.empty() && n > 0) {
        cudaError_t e = S.di.ensure(n*sizeof(int32_t));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(S.di.p, T.hi.data() + lo, n*sizeof(int32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        st.h2d_bytes += n*sizeof(int32_t);
    }
    S.lo = lo; S.hi = hi; S.uploaded = true;
    return cudaSuccess;
}

// Host-only table cache for callers without a device (tests of the builder).
static ForceTables& tables_for(Side* side, DeviceContext* ctx, struct rebx_force* force, int kind,
                               const struct reb_particle* particles, size_t N, std::unordered_map<struct rebx_force*, std::unique_ptr<ForceTables>>& store) {
    auto& slot = store[force];
    if (!slot) slot = std::make_unique<ForceTables>();
    ForceTables& T = *slot;
    const bool stale = !T.valid || T.kind != kind || T.N != N || T.gen != side->generation || T.epoch != global_epoch() || (ctx && ctx->paranoid);
    if (stale) {
        build_tables(T, kind, force, particles, N);
        T.gen = side->generation;
        T.epoch = global_epoch();
        if (ctx) ctx->stats.table_builds++;
    }
    return T;
}

// ---- device context -------------------------------------------------------------------------------
static DeviceContext* ensure_ctx(Side* side, std::string& why) {
    if (side->dev && side->dev->ready) { cudaSetDevice(side->dev->device); return side->dev.get(); }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        why = std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + ")";
        return nullptr;
    }
    std::unique_ptr<DeviceContext, void (*)(DeviceContext*)> ctx(new DeviceContext(), delete_ctx);
    const char* env = std::getenv("REBX_B200_DEVICE");
    ctx->device = env ? std::atoi(env) : 0;
    if (ctx->device < 0 || ctx->device >= count) ctx->device = 0;
    // REBX_B200_DEVICES = all | <count> | <comma-separated ordinals>: the devices a large segment is spread over.
    // The primary device (REBX_B200_DEVICE, default 0) always comes first.
    std::vector<int> ordinals = {ctx->device};
    if (const char* list = std::getenv("REBX_B200_DEVICES")) {
        std::vector<int> want;
        if (std::strcmp(list, "all") == 0) { for (int d = 0; d < count; d++) want.push_back(d); }
        else if (std::strchr(list, ',')) { for (const char* q = list; *q;) { char* end; long v = std::strtol(q, &end, 10); if (end == q) break; want.push_back((int)v); q = (*end == ',') ? end + 1 : end; } }
        else { const long k = std::atol(list); for (int d = 0; d < k && d < count; d++) want.push_back((ctx->device + d) % count); }
        for (int d : want) if (d >= 0 && d < count && std::find(ordinals.begin(), ordinals.end(), d) == ordinals.end()) ordinals.push_back(d);
    }
    for (int d : ordinals) {
        auto slot = std::make_unique<DevSlot>();
        if ((e = slot->init(d)) != cudaSuccess) { why = std::string("device ") + std::to_string(d) + ": " + cudaGetErrorString(e); return nullptr; }
        ctx->devs.push_back(std::move(slot));
    }
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) { why = cudaGetErrorString(e); return nullptr; }
    if (const char* c = std::getenv("REBX_B200_CHUNK")) { long v = std::atol(c); if (v >= 1024) { ctx->chunk = (size_t)v & ~size_t(1); ctx->chunk_from_env = true; } }
    if (const char* p = std::getenv("REBX_B200_PARANOID")) ctx->paranoid = std::atoi(p) != 0;
    if (const char* m = std::getenv("REBX_B200_MATH")) ctx->tolerance = (m[0] == 't' || m[0] == 'T');
    if (const char* z = std::getenv("REBX_B200_ZEROCOPY_MAX")) { long v = std::atol(z); if (v >= 0) ctx->zero_copy_max = (size_t)v; }
    if (const char* z = std::getenv("REBX_B200_D2H")) {
        // option (off by default): only the accelerations travel back -- the copy engine moves rows of 24 or
        // 32 bytes out of the 128-byte records instead of the whole records, which takes bytes off the link the uploads contend for
        if (std::strcmp(z, "acc") == 0) ctx->d2h_width = 24;
        else if (std::strcmp(z, "accm") == 0 && offsetof(struct reb_particle, m) == offsetof(struct reb_particle, ax) + 24) ctx->d2h_width = 32;
    }
    if (const char* z = std::getenv("REBX_B200_MULTI_MIN")) { long v = std::atol(z); if (v >= 2) ctx->multi_min = (size_t)v; }
    ctx->ready = true;
    side->dev = std::move(ctx);
    return side->dev.get();
}

// ---- one fused segment ----------------------------------------------------------------------------
struct Entry {
    int kind = 0, flags = 0;
    double scal[REBX_B200_MAX_SCALARS] = {0};
    int64_t body_index = -1;
    ForceTables* T = nullptr;
    const Oblate* ob = nullptr;
    bool reduces = false;
    int frame = 0;          // 0: body is a particle; 1: barycentre (second uniform sweep); 2: Jacobi (scan path); 3: gr
    double extra[3] = {0, 0, 0};    // lense_thirring: spin angular momentum of the source (body record slots 8..10)
    bool has_extra = false;
    int n_active = 0;       // gr only
    std::vector<struct reb_particle> gr_active;     // gr only: host result for the active bodies
};

static const double* get_d(struct rebx_force* f, const char* name) { return static_cast<const double*>(find_value(f->ap, name)); }

// Builds the effect entries of one force; returns false (after reporting) when the force does nothing.
static bool entries_for_force(struct reb_simulation* sim, Side* side, DeviceContext* ctx, struct rebx_force* force, int kind,
                              const struct reb_particle* particles, size_t N, std::vector<Entry>& out,
                              std::unordered_map<struct rebx_force*, std::unique_ptr<ForceTables>>& store) {
    Entry e;
    e.kind = kind;
    e.scal[0] = sim->G;
    switch (kind) {
    case REBX_B200_RADIATION: {
        const double* c = get_d(force, "c");
        if (!c) { reb_simulation_error(sim, "Need to set speed of light in radiation_forces effect.  See examples in documentation.\n"); return false; }
        e.scal[1] = *c;
        ForceTables& T = tables_for(side, ctx, force, kind, particles, N, store);
        e.T = &T;
        if (T.bodies.empty()) { e.body_index = 0; out.push_back(e); }          // default source: index 0 (radiation_forces.c:115-117)
        else for (int64_t s : T.bodies) { e.body_index = s; out.push_back(e); }  // one sweep per flagged source, ascending
        return true;
    }
    case REBX_B200_CENTRAL_FORCE:
    case REBX_B200_HARMONICS: {
        ForceTables& T = tables_for(side, ctx, force, kind, particles, N, store);
        e.T = &T; e.reduces = true;
        for (size_t b = 0; b < T.bodies.size(); b++) { e.body_index = T.bodies[b]; e.ob = &T.oblate[b]; out.push_back(e); }
        return !T.bodies.empty();
    }
    case REBX_B200_GR_POTENTIAL: {
        const double* c = get_d(force, "c");
        if (!c) { reb_simulation_error(sim, "REBOUNDx Error: Need to set speed of light in gr effect.  See examples in documentation.\n"); return false; }
        if (N < 2) return false;
        e.scal[1] = (*c)*(*c);
        e.body_index = 0; e.reduces = true;
        out.push_back(e);
        return true;
    }
    case REBX_B200_YARKOVSKY: {
        const double* lstar = get_d(force, "ye_lstar");
        const double* c = get_d(force, "ye_c");
        if (!lstar || !c || N < 2) return false;                               // yarkovsky_effect.c:244: silently inactive
        const double* sb = get_d(force, "ye_stef_boltz");
        e.scal[1] = *lstar; e.scal[2] = *c; e.scal[3] = sb ? *sb : 0.0;
        ForceTables& T = tables_for(side, ctx, force, kind, particles, N, store);
        e.T = &T; e.body_index = 0;
        if (T.n_incomplete > 0)
            reb_simulation_error(sim, "REBOUNDx Error: One or more parameters missing for this version of the Yarkovsky effect in Rebx. Please make sure you've given values to all variables for this version before running simulations. If you'd rather use the simplified version of this effect (requires fewer parameters), then please set 'ye_flag' to -1 or 1.\n\n");
        out.push_back(e);
        return true;
    }
    case REBX_B200_MODIFY_ORBITS:
    case REBX_B200_GAS_DAMPING:
    case REBX_B200_EXP_MIGRATION:
    case REBX_B200_TYPE_I: {
        const int* cptr = static_cast<const int*>(find_value(force->ap, "coordinates"));
        const int coordinates = cptr ? *cptr : REBX_COORDINATES_JACOBI;
        ForceTables& T = tables_for(side, ctx, force, kind, particles, N, store);
        e.T = &T; e.reduces = true;
        if (coordinates == REBX_COORDINATES_PARTICLE) {
            if (T.bodies.empty()) {
                reb_simulation_error(sim, "Coordinates set to REBX_COORDINATES_PARTICLE, but primary param was not found in any particle.  Need to set parameter.\n");
                return false;   // the reference goes on to index particles[-1] here (rebxtools.c:139); we stop
            }
            e.body_index = T.bodies.front();
        } else if (coordinates == REBX_COORDINATES_BARYCENTRIC) {
            e.frame = 1; e.flags |= REBX_B200_FLAG_BARYCENTRIC; e.body_index = -1;
        } else if (coordinates == REBX_COORDINATES_JACOBI) {
            e.frame = 2; e.body_index = 0;           // particle 0 has no Jacobi coordinate (rebxtools.c:71-73)
            if (N < 2) return false;
        } else {
            reb_simulation_error(sim, "Coordinates not supported in REBOUNDx.\n");
            return false;
        }
        if (kind == REBX_B200_MODIFY_ORBITS) {
            const double* dedge = get_d(force, "ide_position");
            const double* hedge = get_d(force, "ide_width");
            if (dedge && hedge) { e.flags |= REBX_B200_FLAG_TRAP; e.scal[1] = *dedge; e.scal[2] = *hedge; }
        } else if (kind == REBX_B200_GAS_DAMPING) {
            const double* cs = get_d(force, "cs_coeff");
            const double* tc = get_d(force, "tau_coeff");
            int64_t missing = T.n_incomplete;
            // the reference body is never evaluated, so its own missing d_factor is no error
            if (e.body_index >= 0 && is_absent(T.h[0][(size_t)e.body_index])) missing--;
            if (!cs || !tc || missing > 0)
                rebx_error(static_cast<struct rebx_extras*>(sim->extras), "Need to set d_factor, cs_coeff, tau_coeff parameters.  See examples in documentation.\n");
            if (!cs || !tc) return false;                                      // every particle returns a zero force
            e.scal[1] = *cs; e.scal[2] = *tc;
        } else if (kind == REBX_B200_EXP_MIGRATION) {
            e.scal[1] = sim->t;                                                 // exp(-t/tau), exponential_migration.c:96
        } else {
            const double* beta = get_d(force, "tIm_flaring_index");
            const double* h0 = get_d(force, "tIm_scale_height_1");
            const double* sd0 = get_d(force, "tIm_surface_density_1");
            const double* s = get_d(force, "tIm_surface_density_exponent");
            const double* dedge = get_d(force, "ide_position");
            const double* hedge = get_d(force, "ide_width");
            e.scal[1] = beta ? *beta : 0.0;  e.scal[2] = h0 ? *h0 : 0.01;      // defaults: type_I_migration.c:121-126
            e.scal[3] = sd0 ? *sd0 : 0.0;    e.scal[4] = s ? *s : 0.0;
            e.scal[5] = dedge ? *dedge : 0.0; e.scal[6] = hedge ? *hedge : 0.0;
        }
        out.push_back(e);
        return true;
    }
    case REBX_B200_LENSE_THIRRING: {
        const double* c = get_d(force, "lt_c");
        if (!c) { reb_simulation_error(sim, "REBOUNDx Error: Need to set speed of light in LT effect.  See examples in documentation.\n"); return false; }
        if (N < 2) return false;
        // the source is particles[0]; it needs both its moment of inertia I and its spin vector Omega (lense_thirring.c:109-115)
        const double* I = static_cast<const double*>(find_value(static_cast<struct rebx_node*>(particles[0].ap), "I"));
        const struct reb_vec3d* Om = static_cast<const struct reb_vec3d*>(find_value(static_cast<struct rebx_node*>(particles[0].ap), "Omega"));
        if (!I || !Om) return false;
        e.scal[1] = (*c)*(*c);
        e.body_index = 0; e.reduces = true;
        e.extra[0] = (*I)*Om->x; e.extra[1] = (*I)*Om->y; e.extra[2] = (*I)*Om->z;
        e.has_extra = true;
        out.push_back(e);
        return true;
    }
    case REBX_B200_GAS_DF: {
        static const char* const names[7] = {"gas_df_rhog", "gas_df_alpha_rhog", "gas_df_cs", "gas_df_alpha_cs", "gas_df_xmin", "gas_df_hr", "gas_df_Qd"};
        static const char* const msgs[7] = {"Need to specify a gas density\n", "Need to specify a profile for gas density\n", "Need to set a sound speed.\n",
                                            "Need to specify a profile for the sound speed\n", "Need to set a cutoff.\n", "Need an aspect ratio.\n", "Need to specify Qd"};
        bool ok = true;
        for (int q = 0; q < 7; q++) {
            const double* v = get_d(force, names[q]);
            if (!v) { reb_simulation_error(sim, msgs[q]); ok = false; }       // the reference reports and then dereferences NULL (gas_dynamical_friction.c:141-173)
            else e.scal[1 + q] = *v;
        }
        if (!ok || N < 2) return false;
        e.body_index = 0;
        out.push_back(e);
        return true;
    }
    case KIND_GR: {
        const double* c = get_d(force, "c");
        if (!c) { reb_simulation_error(sim, "REBOUNDx Error: Need to set speed of light in gr effect.  See examples in documentation.\n"); return false; }
        const size_t A = (sim->N_active == SIZE_MAX || sim->N_active > N) ? N : sim->N_active;
        if (A < 1 || N < 2) return false;
        if (A > REBX_B200_GR_MAX_ACTIVE) {
            reb_simulation_error(sim, "REBOUNDx-B200 Error: 'gr' handles up to 4096 active (massive) bodies -- their mutual terms are O(N_active^2) on the host -- plus any number of test particles; set sim->N_active. (With N_active unset the reference recomputes O(N^2) Newtonian gravity among ALL particles.)\n");
            return false;
        }
        const int* it = static_cast<const int*>(find_value(force->ap, "max_iterations"));
        e.scal[1] = (*c)*(*c);
        e.scal[2] = it ? (double)*it : 10.0;
        e.frame = 3; e.n_active = (int)A; e.body_index = 0;
        out.push_back(e);
        return true;
    }
    default: return false;
    }
}

// Host part of `gr` for the active bodies (reference src/gr.c:79-160 restricted to i < N_active):
// Newtonian accelerations among the actives minus the pull of the test particles (reduced on the
// device), Jacobi transform, fixed-point + 1PN term for actives 1..A-1, transform back.  Fills
// args.com (Jacobi origin of the test particles) and `act[i].a*` (what to add to the actives).
static void gr_host_actives(const struct reb_particle* particles, int A, double G, double C2, int max_iterations,
                            const double* pull, rebx_b200_gr_args& args, std::vector<struct reb_particle>& act, bool& not_converged) {
    act.assign(particles, particles + A);
    std::vector<struct reb_particle> pj(A);
    for (int i = 0; i < A; i++) { act[i].ax = 0.; act[i].ay = 0.; act[i].az = 0.; }
    for (int i = 0; i < A; i++) {
        const struct reb_particle pi = act[i];
        for (int j = i + 1; j < A; j++) {
            const struct reb_particle pq = act[j];
            const double dx = pi.x - pq.x, dy = pi.y - pq.y, dz = pi.z - pq.z;
            const double r2 = dx*dx + dy*dy + dz*dz;
            const double r = std::sqrt(r2);
            const double prefac = G/(r2*r);
            act[i].ax -= prefac*pq.m*dx; act[i].ay -= prefac*pq.m*dy; act[i].az -= prefac*pq.m*dz;
            act[j].ax += prefac*pi.m*dx; act[j].ay += prefac*pi.m*dy; act[j].az += prefac*pi.m*dz;
        }
        act[i].ax -= pull[3*i + 0]; act[i].ay -= pull[3*i + 1]; act[i].az -= pull[3*i + 2];
    }
    const double mu = G*act[0].m;
    reb_transformations_inertial_to_jacobi_posvelacc(act.data(), pj.data(), act.data(), A, A);
    args.com[0] = pj[0].x;  args.com[1] = pj[0].y;  args.com[2] = pj[0].z;
    args.com[3] = pj[0].vx; args.com[4] = pj[0].vy; args.com[5] = pj[0].vz;
    args.com[6] = pj[0].ax; args.com[7] = pj[0].ay; args.com[8] = pj[0].az;
    for (int i = 1; i < A; i++) {
        const struct reb_particle p = pj[i];
        double vix = p.vx, viy = p.vy, viz = p.vz;
        double vi2 = vix*vix + viy*viy + viz*viz;
        const double ri = std::sqrt(p.x*p.x + p.y*p.y + p.z*p.z);
        int q = 0;
        double Aa = (0.5*vi2 + 3.*mu/ri)/C2;
        for (q = 0; q < max_iterations; q++) {
            const double ox = vix, oy = viy, oz = viz;
            vix = p.vx/(1. - Aa); viy = p.vy/(1. - Aa); viz = p.vz/(1. - Aa);
            vi2 = vix*vix + viy*viy + viz*viz;
            Aa = (0.5*vi2 + 3.*mu/ri)/C2;
            const double dvx = vix - ox, dvy = viy - oy, dvz = viz - oz;
            if ((dvx*dvx + dvy*dvy + dvz*dvz)/vi2 < 2.220446049250313e-16*2.220446049250313e-16) break;
        }
        if (q == 10) not_converged = true;
        const double B = (mu/ri - 1.5*vi2)*mu/(ri*ri*ri)/C2;
        const double rdotrdot = p.x*p.vx + p.y*p.vy + p.z*p.vz;
        const double vdx = p.ax + B*p.x, vdy = p.ay + B*p.y, vdz = p.az + B*p.z;
        const double vdotvdot = vix*vdx + viy*vdy + viz*vdz;
        const double D = (vdotvdot - 3.*mu/(ri*ri*ri)*rdotrdot)/C2;
        pj[i].ax = B*(1. - Aa)*p.x - Aa*p.ax - D*vix;
        pj[i].ay = B*(1. - Aa)*p.y - Aa*p.ay - D*viy;
        pj[i].az = B*(1. - Aa)*p.z - Aa*p.az - D*viz;
    }
    pj[0].ax = 0.; pj[0].ay = 0.; pj[0].az = 0.;
    reb_transformations_jacobi_to_inertial_acc(act.data(), pj.data(), act.data(), A, A);
}

static void fill_body(double* rec, const Entry& e, const struct reb_particle* particles) {
    for (int j = 0; j < REBX_B200_BODY_DOUBLES; j++) rec[j] = 0.0;
    rec[REBX_B200_BODY_INDEX] = -1.0;
    if (e.body_index < 0) return;                   // barycentre: the record is computed on the device
    const struct reb_particle& b = particles[e.body_index];
    rec[REBX_B200_BODY_INDEX] = (double)e.body_index;
    rec[REBX_B200_BODY_X+0] = b.x;  rec[REBX_B200_BODY_X+1] = b.y;  rec[REBX_B200_BODY_X+2] = b.z;
    rec[REBX_B200_BODY_VX+0] = b.vx; rec[REBX_B200_BODY_VX+1] = b.vy; rec[REBX_B200_BODY_VX+2] = b.vz;
    rec[REBX_B200_BODY_M] = b.m;
    if (e.has_extra) { rec[8] = e.extra[0]; rec[9] = e.extra[1]; rec[10] = e.extra[2]; }
    if (e.ob) {
        rec[REBX_B200_BODY_J2] = e.ob->J2; rec[REBX_B200_BODY_J4] = e.ob->J4; rec[REBX_B200_BODY_REQ] = e.ob->Req;
        for (int j = 0; j < 3; j++) { rec[REBX_B200_BODY_U+j] = e.ob->u[j]; rec[REBX_B200_BODY_V+j] = e.ob->v[j]; rec[REBX_B200_BODY_W+j] = e.ob->w[j]; }
    }
}

#define REBXB_CUDA(call, what) do { cudaError_t err__ = (call); if (err__ != cudaSuccess) { fail = what; fail_code = err__; goto failed; } } while (0)

// DMA between the caller's particle array and device staging.  REBOUND's array comes from realloc, i.e. it
// starts 16 bytes into a page, and the copy engines run ~12 % slower on a transfer whose HOST address is not
// page-aligned (measured, tools/pcie_probe.cu: 10^7 records both ways 30.5 ms vs 27.2 ms).  Splitting each
// transfer into the short head up to the next page boundary and the page-aligned remainder restores the
// full rate.
static cudaError_t copy_page_split(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
    const uintptr_t host = reinterpret_cast<uintptr_t>(kind == cudaMemcpyHostToDevice ? src : dst);
    const size_t head = (size_t)((0 - host) & 4095u);
    if (head == 0 || head >= bytes || bytes < (size_t(1) << 20)) return cudaMemcpyAsync(dst, src, bytes, kind, s);
    cudaError_t e = cudaMemcpyAsync(dst, src, head, kind, s);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(static_cast<char*>(dst) + head, static_cast<const char*>(src) + head, bytes - head, kind, s);
}

static void run_segment(struct reb_simulation* sim, struct rebx_force** forces, int nf, struct reb_particle* particles, const int Nint) {
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    if (rebx == nullptr || Nint <= 0) return;
    const size_t N = (size_t)Nint;
    Side* side = side_of(rebx);
    std::string why;
    DeviceContext* ctx = ensure_ctx(side, why);
    if (!ctx) {
        std::string msg = "REBOUNDx-B200 Error: " + why + ". The built-in forces of this library run on the GPU only; accelerations were NOT updated.\n";
        reb_simulation_error(sim, msg.c_str());
        return;
    }
    ctx->stats.calls++;
    refresh_escaped(side);          // values written through pointers the getters handed out (no API call announces them)

    std::vector<Entry> entries;
    for (int f = 0; f < nf; f++) entries_for_force(sim, side, ctx, forces[f], kind_of(forces[f]), particles, N, entries, ctx->tables);
    if (entries.empty()) return;

    const char* fail = nullptr;
    cudaError_t fail_code = cudaSuccess;
    const size_t ne = entries.size();
    // Centre-of-mass frames need the whole ensemble on the device (a reduction or a scan over
    // all particles precedes the sweep), so such a segment runs as one resident chunk.
    bool whole = false;
    for (const Entry& e : entries) if (e.frame != 0) whole = true;
    // Devices of this segment: every configured device for a large segment in PARTICLE frames, split by index range
    // [lo_d, hi_d) with even boundaries (each device streams its range through its own three-stage pipeline; tables
    // are sharded the same way; the per-device, per-chunk back-reaction sums are added on the host in index order);
    // the primary device alone otherwise.
    const size_t ndev = (!whole && ctx->devs.size() > 1 && N >= ctx->multi_min) ? ctx->devs.size() : 1;
    size_t per_dev = (N + ndev - 1)/ndev;
    per_dev += per_dev & 1;
    // Pipeline granularity: the configured chunk for large ensembles; mid-size ones are cut into ~6
    // pieces (>= 32768 particles) so that H2D, sweep and D2H of different pieces overlap.
    size_t CH = ctx->chunk;
    if (!ctx->chunk_from_env && per_dev < 6*ctx->chunk) CH = std::max<size_t>(32768, ((per_dev/6) + 1) & ~size_t(1));
    if (whole) CH = (N + 1) & ~size_t(1);
    const size_t chunks_per_dev = (std::min(per_dev, N) + CH - 1)/CH;
    const size_t nchunks = chunks_per_dev*ndev;                 // slots (a trailing device may use fewer)
    const size_t chunk_cap = std::min(N, CH);
    bool need_vel = false, need_m = false, need_r = false, any_reduces = false;
    for (const Entry& e : entries) if (e.reduces) any_reduces = true;
    for (const Entry& e : entries) {
        switch (e.kind) {
            case REBX_B200_RADIATION: need_vel = true; break;
            case REBX_B200_HARMONICS: case REBX_B200_GR_POTENTIAL: case REBX_B200_CENTRAL_FORCE: need_m = true; break;
            case REBX_B200_YARKOVSKY: case REBX_B200_GAS_DF: need_vel = need_m = need_r = true; break;
            default: need_vel = need_m = true; break;     // damping trio, gr, lense_thirring
        }
    }
    const rebx_b200_aos_layout L = {(int32_t)sizeof(struct reb_particle), (int32_t)offsetof(struct reb_particle, x),
                                    (int32_t)offsetof(struct reb_particle, vx), (int32_t)offsetof(struct reb_particle, ax),
                                    (int32_t)offsetof(struct reb_particle, m), (int32_t)offsetof(struct reb_particle, r)};
    const size_t soa_stride = (chunk_cap + 1) & ~size_t(1);      // keep every array 16-byte aligned
    int launches = 0;
    double* bodies_h = nullptr;
    DevSlot& D0 = *ctx->devs[0];
    // a small ensemble is read and written by the staging kernels straight through the host mapping (below)
    const bool small = nchunks == 1 && !whole && N <= ctx->zero_copy_max;
    unsigned char* zc = nullptr;

    REBXB_CUDA(ctx->bodies_h.ensure(ne*REBX_B200_BODY_DOUBLES*sizeof(double)), "allocating pinned body records");
    for (size_t d = 0; d < ndev; d++) {
        DevSlot& D = *ctx->devs[d];
        const size_t lo = std::min(N, d*per_dev), hi = std::min(N, lo + per_dev);
        REBXB_CUDA(cudaSetDevice(D.device), "selecting a device");
        // tables to HBM (no-op unless rebuilt or re-partitioned)
        for (Entry& e : entries) if (e.T) REBXB_CUDA(upload_tables(*e.T, ctx->stats, d, ndev == 1 ? 0 : lo, ndev == 1 ? N : hi), "uploading parameter tables");
        REBXB_CUDA(D.bodies.ensure(ne*REBX_B200_BODY_DOUBLES*sizeof(double)), "allocating body records");
        REBXB_CUDA(D.br_h.ensure(chunks_per_dev*ne*3*sizeof(double)), "allocating pinned back-reaction sums");
        REBXB_CUDA(D.br.ensure(chunks_per_dev*ne*3*sizeof(double)), "allocating back-reaction sums");
        for (int b = 0; b < DeviceContext::kBuffers && (size_t)b < chunks_per_dev; b++) {
            REBXB_CUDA(D.aos[b].ensure(chunk_cap*sizeof(struct reb_particle)), "allocating AoS staging");
            REBXB_CUDA(D.soa[b].ensure(11*soa_stride*sizeof(double)), "allocating SoA state");
            REBXB_CUDA(D.work[b].ensure_zeroed(rebx_b200_workspace_bytes()), "allocating reduction scratch");
        }
    }
    REBXB_CUDA(cudaSetDevice(D0.device), "selecting a device");
    if (whole) {
        REBXB_CUDA(ctx->scan.ensure(rebx_b200_scan_scratch_bytes((int64_t)N)), "allocating scan scratch");
        REBXB_CUDA(ctx->totals.ensure(16*sizeof(double)), "allocating mass moments");
    }
    // Pin the caller's particle array once so the DMA engines can stream it (only worth it for the
    // simulation's own long-lived array, not for an integrator's scratch copy).  A small ensemble is
    // pinned as well: its records are then read and its accelerations written by the staging kernels
    // straight through the mapping (zero copy) -- no DMA descriptors, one PCIe round trip less each way.
    if (particles == sim->particles && (small || N*sizeof(struct reb_particle) >= (size_t(256) << 10))) {
        // REBOUND reallocs or frees its array without telling us: a cached registration is trusted only while the
        // runtime still reports the first and the last byte of the range as registered host memory
        if (ctx->registered_ptr == particles && ctx->registered_bytes == N*sizeof(struct reb_particle)) {
            cudaPointerAttributes a0, a1;
            const char* last = reinterpret_cast<const char*>(particles) + N*sizeof(struct reb_particle) - 1;
            if (cudaPointerGetAttributes(&a0, particles) != cudaSuccess || cudaPointerGetAttributes(&a1, last) != cudaSuccess ||
                a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost) { cudaGetLastError(); ctx->unregister_host(); }
        }
        if (ctx->registered_ptr != particles || ctx->registered_bytes != N*sizeof(struct reb_particle)) {
            ctx->unregister_host();
            if (cudaHostRegister(particles, N*sizeof(struct reb_particle), cudaHostRegisterMapped | cudaHostRegisterPortable) == cudaSuccess) {
                ctx->registered_ptr = particles; ctx->registered_bytes = N*sizeof(struct reb_particle);
                if (cudaHostGetDevicePointer(&ctx->registered_dev, particles, 0) != cudaSuccess) { ctx->registered_dev = nullptr; cudaGetLastError(); }
            } else {
                cudaGetLastError();     // pageable copies still work, just slower
            }
        }
    }
    if (small && ctx->registered_ptr == particles) zc = static_cast<unsigned char*>(ctx->registered_dev);

    bodies_h = static_cast<double*>(ctx->bodies_h.p);
    for (size_t e = 0; e < ne; e++) fill_body(bodies_h + e*REBX_B200_BODY_DOUBLES, entries[e], particles);
    for (size_t d = 0; d < ndev; d++) {
        DevSlot& D = *ctx->devs[d];
        if (ndev > 1) REBXB_CUDA(cudaSetDevice(D.device), "selecting a device");
        // zero-copy path: the sweeps read the few body records straight from the pinned host buffer (one copy less)
        if (!zc) REBXB_CUDA(cudaMemcpyAsync(D.bodies.p, bodies_h, ne*REBX_B200_BODY_DOUBLES*sizeof(double), cudaMemcpyHostToDevice, D.stream[0]), "copying body records");
        if (any_reduces) REBXB_CUDA(cudaMemsetAsync(D.br.p, 0, chunks_per_dev*ne*3*sizeof(double), D.stream[0]), "clearing back-reaction sums");
        if (chunks_per_dev > 1) {      // the other pipeline streams must see the body records
            REBXB_CUDA(cudaEventRecord(D.bodies_ready, D.stream[0]), "recording event");
            for (int b = 1; b < DeviceContext::kBuffers; b++) REBXB_CUDA(cudaStreamWaitEvent(D.stream[b], D.bodies_ready, 0), "waiting on event");
        }
        ctx->stats.h2d_bytes += ne*REBX_B200_BODY_DOUBLES*sizeof(double);
    }

    // chunk ci of every device is issued before chunk ci + 1 of any, so that all devices start streaming at once
    for (size_t ci = 0; ci < chunks_per_dev; ci++) for (size_t d = 0; d < ndev; d++) {
        DevSlot& D = *ctx->devs[d];
        const size_t lo = std::min(N, d*per_dev), hi = std::min(N, lo + per_dev);
        const size_t c0 = lo + ci*CH;
        if (c0 >= hi) continue;
        if (ndev > 1) REBXB_CUDA(cudaSetDevice(D.device), "selecting a device");
        const int b = (int)(ci % DeviceContext::kBuffers);
        cudaStream_t s = D.stream[b];
        const size_t n = std::min(CH, hi - c0);
        const size_t bytes = n*sizeof(struct reb_particle);
        const size_t tab0 = c0 - (ndev == 1 ? 0 : lo);                              // offset inside this device's table slice
        void* const aos_in = zc ? static_cast<void*>(zc) : D.aos[b].p;      // zc: one chunk on one device, so c0 == 0
        if (!zc) REBXB_CUDA(copy_page_split(D.aos[b].p, particles + c0, bytes, cudaMemcpyHostToDevice, s), "copying particles to the device");
        ctx->stats.h2d_bytes += bytes;
        double* base = static_cast<double*>(D.soa[b].p);
        rebx_b200_state st;
        st.n = (int64_t)n; st.index_base = (int64_t)c0;
        st.x = base + 0*soa_stride; st.y = base + 1*soa_stride; st.z = base + 2*soa_stride;
        st.vx = need_vel ? base + 3*soa_stride : nullptr; st.vy = need_vel ? base + 4*soa_stride : nullptr; st.vz = need_vel ? base + 5*soa_stride : nullptr;
        st.m = need_m ? base + 6*soa_stride : nullptr;
        st.r = need_r ? base + 7*soa_stride : nullptr;
        st.ax = base + 8*soa_stride; st.ay = base + 9*soa_stride; st.az = base + 10*soa_stride;
        REBXB_CUDA((cudaError_t)rebx_b200_unpack_aos(aos_in, &L, &st, s, &launches), "unpacking particles");
        for (size_t e0 = 0; e0 < ne;) {
            // consecutive entries sharing a reference-frame class form one launch group
            const int frame = entries[e0].frame;
            const size_t cap = frame == 3 ? 1 : (frame == 2 ? 3 : REBX_B200_MAX_EFFECTS);
            size_t e1 = e0;
            while (e1 < ne && e1 - e0 < cap && entries[e1].frame == frame) e1++;
            if (frame == 3) {
                // `gr`: device reduction -> host part for the active bodies -> device sweep (see kernels_gr.cuh)
                Entry& en = entries[e0];
                rebx_b200_gr_args ga;
                std::memset(&ga, 0, sizeof ga);
                ga.n_active = en.n_active; ga.max_iterations = (int)en.scal[2];
                ga.G = sim->G; ga.C2 = en.scal[1]; ga.mu = sim->G*particles[0].m;
                // device scratch: pull [3A] | convergence flag | actives {x, y, z, m} [4A]; pinned mirror of the first two + staging of the third
                const size_t Aq = (size_t)en.n_active;
                REBXB_CUDA(ctx->grbuf.ensure((7*Aq + 16)*sizeof(double)), "allocating gr scratch");
                REBXB_CUDA(ctx->gr_h.ensure((7*Aq + 16)*sizeof(double)), "allocating pinned gr scratch");
                double* pull_d = static_cast<double*>(ctx->grbuf.p);
                int* flag_d = reinterpret_cast<int*>(pull_d + 3*Aq);
                double* act_d = pull_d + 3*Aq + 8;
                double* pull_h = static_cast<double*>(ctx->gr_h.p);
                double* act_h = pull_h + 3*Aq + 8;
                for (size_t i = 0; i < Aq; i++) { act_h[4*i] = particles[i].x; act_h[4*i + 1] = particles[i].y; act_h[4*i + 2] = particles[i].z; act_h[4*i + 3] = particles[i].m; }
                REBXB_CUDA(cudaMemcpyAsync(act_d, act_h, 4*Aq*sizeof(double), cudaMemcpyHostToDevice, s), "copying the active bodies");
                ga.active = act_d;
                REBXB_CUDA(cudaMemsetAsync(flag_d, 0, sizeof(int), s), "clearing the convergence flag");
                REBXB_CUDA((cudaError_t)rebx_b200_gr_pull(&st, &ga, pull_d, D.work[b].p, s, &launches), "reducing the test particles' pull");
                bool host_nc = false;
                if (en.n_active == 1) {
                    // ONE active body (a star and its test particles, the common case): the host part needs nothing from the
                    // device -- the body's own 1PN acceleration is zero, the Jacobi origin is the body, and its acceleration
                    // (minus the test particles' pull) is taken from the device-resident reduction by the sweep itself.  No
                    // round trip, no synchronisation inside the pipeline.
                    pull_h[0] = pull_h[1] = pull_h[2] = 0.0;
                    gr_host_actives(particles, 1, sim->G, ga.C2, ga.max_iterations, pull_h, ga, en.gr_active, host_nc);
                    ga.pull1 = pull_d;
                } else {
                    REBXB_CUDA(cudaMemcpyAsync(pull_h, pull_d, (size_t)en.n_active*3*sizeof(double), cudaMemcpyDeviceToHost, s), "copying the pull");
                    REBXB_CUDA(cudaStreamSynchronize(s), "waiting for the device");
                    gr_host_actives(particles, en.n_active, sim->G, ga.C2, ga.max_iterations, pull_h, ga, en.gr_active, host_nc);
                }
                REBXB_CUDA((cudaError_t)rebx_b200_gr_sweep(&st, &ga, flag_d, s, &launches), "launching the gr sweep");
                REBXB_CUDA(cudaMemcpyAsync(pull_h + 3*Aq, flag_d, sizeof(int), cudaMemcpyDeviceToHost, s), "copying the convergence flag");
                *reinterpret_cast<int*>(pull_h + 3*Aq + 4) = host_nc ? 1 : 0;
                e0 = e1;
                continue;
            }
            rebx_b200_pass P;
            std::memset(&P, 0, sizeof P);
            P.st = st;
            P.n_effects = (int32_t)(e1 - e0);
            {   // one central body for the whole group => the sweep may share its relative coordinates
                bool same = true;
                for (size_t j = e0 + 1; j < e1; j++) if (entries[j].body_index != entries[e0].body_index) same = false;
                if (same) P.reserved |= REBX_B200_PASS_SHARED_BODY;
                if (ctx->tolerance) P.reserved |= REBX_B200_PASS_TOLERANCE;
            }
            for (int j = 0; j < P.n_effects; j++) {
                const Entry& en = entries[e0 + j];
                rebx_b200_effect& fx = P.fx[j];
                fx.kind = en.kind; fx.flags = en.flags;
                fx.body = (zc ? bodies_h : static_cast<const double*>(D.bodies.p)) + (e0 + j)*REBX_B200_BODY_DOUBLES;
                for (int k = 0; k < REBX_B200_MAX_SCALARS; k++) fx.scal[k] = en.scal[k];
                if (en.T) {
                    const ForceTables::Slice& S = *en.T->dev[d];
                    for (int k = 0; k < REBX_B200_MAX_TABLES; k++) fx.tab[k] = S.d[k].p ? static_cast<const double*>(S.d[k].p) + tab0 : nullptr;
                    fx.itab = S.di.p ? static_cast<const int32_t*>(S.di.p) + tab0 : nullptr;
                }
                fx.backreaction = en.reduces ? static_cast<double*>(D.br.p) + (ci*ne + e0 + j)*3 : nullptr;
            }
            if (frame == 2) {
                REBXB_CUDA((cudaError_t)rebx_b200_launch_jacobi(&P, ctx->scan.p, s, &launches), "launching the Jacobi-frame sweep");
            } else {
                if (frame == 1) {   // barycentre record computed on the device from the resident state
                    REBXB_CUDA((cudaError_t)rebx_b200_mass_moments(&st, static_cast<double*>(ctx->totals.p), ctx->scan.p, nullptr, REBX_B200_COM_WHOLE, s, &launches), "reducing mass moments");
                    REBXB_CUDA((cudaError_t)rebx_b200_com_bodies(static_cast<const double*>(ctx->totals.p),
                                                                 static_cast<double*>(D.bodies.p) + e0*REBX_B200_BODY_DOUBLES, P.n_effects, s, &launches), "building the barycentre record");
                }
                REBXB_CUDA((cudaError_t)rebx_b200_launch_pass(&P, D.work[b].p, s, &launches), "launching the force sweep");
                if (frame == 1)     // every particle loses sum_i (m_i/M) f_i  (rebxtools.c:108-115)
                    for (int j = 0; j < P.n_effects; j++)
                        REBXB_CUDA((cudaError_t)rebx_b200_add_uniform(&st, P.fx[j].backreaction, s, &launches), "applying the barycentric back-reaction");
            }
            e0 = e1;
        }
        REBXB_CUDA((cudaError_t)rebx_b200_pack_acc_aos(aos_in, &L, &st, s, &launches), "packing accelerations");
        if (!zc && ctx->d2h_width > 0) {
            const size_t off = offsetof(struct reb_particle, ax);
            REBXB_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char*>(particles + c0) + off, sizeof(struct reb_particle),
                                         static_cast<const char*>(D.aos[b].p) + off, sizeof(struct reb_particle),
                                         (size_t)ctx->d2h_width, n, cudaMemcpyDeviceToHost, s), "copying accelerations to the host");
            ctx->stats.d2h_bytes += n*(size_t)ctx->d2h_width;
        } else {
            if (!zc) REBXB_CUDA(copy_page_split(particles + c0, D.aos[b].p, bytes, cudaMemcpyDeviceToHost, s), "copying accelerations to the host");
            ctx->stats.d2h_bytes += zc ? n*3*sizeof(double) : bytes;
        }
    }
    for (size_t d = 0; d < ndev; d++) {
        DevSlot& D = *ctx->devs[d];
        if (ndev > 1) REBXB_CUDA(cudaSetDevice(D.device), "selecting a device");
        if (chunks_per_dev > 1) for (int b = 1; b < DeviceContext::kBuffers; b++) REBXB_CUDA(cudaStreamSynchronize(D.stream[b]), "waiting for the device");
        if (any_reduces) {
            REBXB_CUDA(cudaMemcpyAsync(D.br_h.p, D.br.p, chunks_per_dev*ne*3*sizeof(double), cudaMemcpyDeviceToHost, D.stream[0]), "copying back-reaction sums");
            ctx->stats.d2h_bytes += chunks_per_dev*ne*3*sizeof(double);
        }
        REBXB_CUDA(cudaStreamSynchronize(D.stream[0]), "waiting for the device");
    }
    if (ndev > 1) REBXB_CUDA(cudaSetDevice(D0.device), "selecting a device");
    ctx->stats.launches += (std::uint64_t)launches;

    // Back-reaction onto the source / primary / oblate bodies (gr_potential.c:74-76,
    // gravitational_harmonics.c:221-223, rebxtools.c:139-141): chunk sums added in chunk order.
    for (size_t e = 0; e < ne; e++) {
        if (entries[e].frame == 3) {                // gr: the active bodies were finished on the host
            for (int i = 0; i < entries[e].n_active; i++) {
                particles[i].ax += entries[e].gr_active[(size_t)i].ax;
                particles[i].ay += entries[e].gr_active[(size_t)i].ay;
                particles[i].az += entries[e].gr_active[(size_t)i].az;
            }
            const double* gh = static_cast<const double*>(ctx->gr_h.p) + 3*(size_t)entries[e].n_active;
            if (*reinterpret_cast<const int*>(gh) || *reinterpret_cast<const int*>(gh + 4))
                reb_simulation_warning(sim, "REBOUNDx Warning: 10 iterations in gr.c failed to converge. This is typically because the perturbation is too strong for the current implementation.");
            continue;
        }
        if (!entries[e].reduces || entries[e].frame != 0) continue;    // COM frames were applied on the device
        double sx = 0.0, sy = 0.0, sz = 0.0;          // device by device, chunk by chunk: ascending particle index
        for (size_t d = 0; d < ndev; d++) {
            const double* br_h = static_cast<const double*>(ctx->devs[d]->br_h.p);
            for (size_t c = 0; c < chunks_per_dev; c++) { sx += br_h[(c*ne + e)*3 + 0]; sy += br_h[(c*ne + e)*3 + 1]; sz += br_h[(c*ne + e)*3 + 2]; }
        }
        struct reb_particle& b = particles[entries[e].body_index];
        b.ax += sx; b.ay += sy; b.az += sz;
    }
    return;

failed:
    {
        cudaGetLastError();
        char buf[400];
        snprintf(buf, sizeof buf, "REBOUNDx-B200 Error: CUDA failure while %s: %s. Accelerations are NOT valid for this call.\n", fail, cudaGetErrorString(fail_code));
        reb_simulation_error(sim, buf);
        for (auto& D : ctx->devs) { cudaSetDevice(D->device); for (auto& st : D->stream) cudaStreamSynchronize(st); }
        cudaSetDevice(ctx->device);
        cudaGetLastError();
    }
}

// Diagnostic potentials (reference src/gr_potential.c:91-120, src/gravitational_harmonics.c:254-316,
// src/central_force.c:98-135): the whole ensemble goes to the device once, one reduction per body.
static double run_potential(struct rebx_extras* rebx, int kind, const struct rebx_force* force) {
    if (rebx == nullptr || rebx->sim == nullptr) { rebx_error(rebx, ""); return 0; }
    struct reb_simulation* sim = rebx->sim;
    const size_t N = sim->N;
    struct reb_particle* particles = sim->particles;
    if (N < 2) return 0;
    Side* side = side_of(rebx);
    std::string why;
    DeviceContext* ctx = ensure_ctx(side, why);
    if (!ctx) {
        std::string msg = "REBOUNDx-B200 Error: " + why + ". The potentials of this library are reduced on the GPU only.\n";
        reb_simulation_error(sim, msg.c_str());
        return 0;
    }
    DevSlot& P0 = *ctx->devs[0];
    refresh_escaped(side);
    std::vector<Entry> entries;
    Entry e;
    e.kind = kind; e.scal[0] = sim->G;
    if (kind == REBX_B200_GR_POTENTIAL) {
        const double* c = force ? static_cast<const double*>(find_value(force->ap, "c")) : nullptr;
        if (!c) { rebx_error(rebx, "REBOUNDx Error: Need to set speed of light in gr effect.  See examples in documentation.\n"); return 0; }
        e.scal[1] = (*c)*(*c); e.body_index = 0;
        entries.push_back(e);
    } else {
        // keyed by the extras address: these potentials take no force argument in the reference API
        ForceTables& T = tables_for(side, ctx, reinterpret_cast<struct rebx_force*>(rebx), kind, particles, N, ctx->tables);
        e.T = &T;
        for (size_t b = 0; b < T.bodies.size(); b++) { e.body_index = T.bodies[b]; e.ob = &T.oblate[b]; entries.push_back(e); }
    }
    if (entries.empty()) return 0;
    const char* fail = nullptr;
    cudaError_t fail_code = cudaSuccess;
    const size_t ne = entries.size();
    const size_t stride = (N + 1) & ~size_t(1);
    const rebx_b200_aos_layout L = {(int32_t)sizeof(struct reb_particle), (int32_t)offsetof(struct reb_particle, x),
                                    (int32_t)offsetof(struct reb_particle, vx), (int32_t)offsetof(struct reb_particle, ax),
                                    (int32_t)offsetof(struct reb_particle, m), (int32_t)offsetof(struct reb_particle, r)};
    int launches = 0;
    double total = 0.0;
    cudaStream_t s = P0.stream[0];
    double* base = nullptr;
    double* out_h = nullptr;
    rebx_b200_state st;
    REBXB_CUDA(P0.aos[0].ensure(N*sizeof(struct reb_particle)), "allocating AoS staging");
    REBXB_CUDA(P0.soa[0].ensure(11*stride*sizeof(double)), "allocating SoA state");
    REBXB_CUDA(P0.work[0].ensure_zeroed(rebx_b200_workspace_bytes()), "allocating reduction scratch");
    REBXB_CUDA(ctx->bodies_h.ensure(ne*REBX_B200_BODY_DOUBLES*sizeof(double)), "allocating pinned body records");
    REBXB_CUDA(P0.bodies.ensure(ne*REBX_B200_BODY_DOUBLES*sizeof(double)), "allocating body records");
    REBXB_CUDA(P0.br_h.ensure(ne*sizeof(double)), "allocating pinned results");
    REBXB_CUDA(P0.br.ensure(ne*sizeof(double)), "allocating results");
    for (size_t k = 0; k < ne; k++) fill_body(static_cast<double*>(ctx->bodies_h.p) + k*REBX_B200_BODY_DOUBLES, entries[k], particles);
    REBXB_CUDA(cudaMemcpyAsync(P0.bodies.p, ctx->bodies_h.p, ne*REBX_B200_BODY_DOUBLES*sizeof(double), cudaMemcpyHostToDevice, s), "copying body records");
    REBXB_CUDA(copy_page_split(P0.aos[0].p, particles, N*sizeof(struct reb_particle), cudaMemcpyHostToDevice, s), "copying particles to the device");
    ctx->stats.h2d_bytes += N*sizeof(struct reb_particle);
    base = static_cast<double*>(P0.soa[0].p);
    std::memset(&st, 0, sizeof st);
    st.n = (int64_t)N; st.index_base = 0;
    st.x = base; st.y = base + stride; st.z = base + 2*stride; st.m = base + 6*stride;
    st.ax = base + 8*stride; st.ay = base + 9*stride; st.az = base + 10*stride;
    REBXB_CUDA((cudaError_t)rebx_b200_unpack_aos(P0.aos[0].p, &L, &st, s, &launches), "unpacking particles");
    for (size_t k = 0; k < ne; k++) {
        rebx_b200_effect fx;
        std::memset(&fx, 0, sizeof fx);
        fx.kind = kind;
        fx.body = static_cast<const double*>(P0.bodies.p) + k*REBX_B200_BODY_DOUBLES;
        for (int q = 0; q < REBX_B200_MAX_SCALARS; q++) fx.scal[q] = entries[k].scal[q];
        REBXB_CUDA((cudaError_t)rebx_b200_potential(&st, &fx, static_cast<double*>(P0.br.p) + k, P0.work[0].p, s, &launches), "reducing the potential");
    }
    out_h = static_cast<double*>(P0.br_h.p);
    REBXB_CUDA(cudaMemcpyAsync(out_h, P0.br.p, ne*sizeof(double), cudaMemcpyDeviceToHost, s), "copying the potential");
    REBXB_CUDA(cudaStreamSynchronize(s), "waiting for the device");
    ctx->stats.launches += (std::uint64_t)launches;
    for (size_t k = 0; k < ne; k++) total += out_h[k];
    return total;
failed:
    {
        cudaGetLastError();
        char buf[400];
        snprintf(buf, sizeof buf, "REBOUNDx-B200 Error: CUDA failure while %s: %s.\n", fail, cudaGetErrorString(fail_code));
        reb_simulation_error(sim, buf);
        cudaStreamSynchronize(P0.stream[0]); cudaGetLastError();
    }
    return 0;
}

}  // namespace rebxh

using namespace rebxh;

extern "C" {

double rebx_gr_potential_potential(struct rebx_extras* const rebx, const struct rebx_force* const gr_potential) {
    return run_potential(rebx, REBX_B200_GR_POTENTIAL, gr_potential);
}
double rebx_gravitational_harmonics_potential(struct rebx_extras* const rebx) {
    return run_potential(rebx, REBX_B200_HARMONICS, nullptr);
}
double rebx_central_force_potential(struct rebx_extras* const rebx) {
    return run_potential(rebx, REBX_B200_CENTRAL_FORCE, nullptr);
}

// ---- plug-in entry points: the reference's force signature, one fused sweep each ---------------
#define REBXB_ENTRY(fn) \
    void fn(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N) { \
        struct rebx_force* one = force; run_segment(sim, &one, 1, particles, N); }
REBXB_ENTRY(rebx_radiation_forces)
REBXB_ENTRY(rebx_gravitational_harmonics)
REBXB_ENTRY(rebx_gr_potential)
REBXB_ENTRY(rebx_gr)
REBXB_ENTRY(rebx_central_force)
REBXB_ENTRY(rebx_exponential_migration)
REBXB_ENTRY(rebx_lense_thirring)
REBXB_ENTRY(rebx_gas_dynamical_friction)
REBXB_ENTRY(rebx_yarkovsky_effect)
REBXB_ENTRY(rebx_modify_orbits_forces)
REBXB_ENTRY(rebx_gas_damping_timescale)
REBXB_ENTRY(rebx_modify_orbits_with_type_I_migration)

// The dispatcher installed as sim->additional_forces (reference src/core.c:1026-1041).
void rebx_additional_forces(struct reb_simulation* sim) {
    if (sim->N_var != 0) {
        reb_simulation_warning(sim, "REBOUNDx: Variational particles have been added to the simulation but are not implemented in REBOUNDx and will not be evolved self-consistently.");
    }
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    if (!rebx) return;
    std::vector<struct rebx_force*> order;
    for (struct rebx_node* n = rebx->additional_forces; n; n = n->next) order.push_back(static_cast<struct rebx_force*>(n->object));
    const bool whfast_vel = sim->force_is_velocity_dependent && sim->integrator.name && std::strcmp(sim->integrator.name, "whfast") == 0;
    const double Nd = (double)sim->N;          // the reference stores N in a double before narrowing it (core.c:1037)
    const int N = (int)Nd;
    size_t i = 0;
    while (i < order.size()) {
        size_t j = i;
        while (j < order.size() && kind_of(order[j]) != REBX_B200_NONE) j++;
        if (j > i) {
            if (whfast_vel) for (size_t k = i; k < j; k++)
                reb_simulation_warning(sim, "REBOUNDx: Passing a velocity-dependent force to WHFAST, will accumulate errors proportional to the force. If forces get big, consider using IAS15 or applying force as an operator.");
            run_segment(sim, order.data() + i, (int)(j - i), sim->particles, N);
            i = j;
        } else {
            if (whfast_vel)
                reb_simulation_warning(sim, "REBOUNDx: Passing a velocity-dependent force to WHFAST, will accumulate errors proportional to the force. If forces get big, consider using IAS15 or applying force as an operator.");
            order[i]->update_accelerations(sim, order[i], sim->particles, N);   // custom / foreign force: host callback
            i++;
        }
    }
}

// ---- introspection used by tests and bench (host side only) -------------------------------------
// Builds (without touching the device) the SoA table `index` of `force` and copies it to `out`.
// Returns the number of elements, or -1.  index -1 selects the int table; index -2 copies the
// body index list (as doubles).
int64_t rebx_b200_host_table(struct rebx_extras* rebx, struct rebx_force* force, int index, double* out, int64_t cap) {
    if (!rebx || !rebx->sim || !force) return -1;
    const int kind = kind_of(force);
    if (kind == REBX_B200_NONE) return -1;
    static std::unordered_map<struct rebx_force*, std::unique_ptr<ForceTables>> scratch;
    scratch.erase(force);
    ForceTables& T = tables_for(side_of(rebx), nullptr, force, kind, rebx->sim->particles, rebx->sim->N, scratch);
    int64_t n = -1;
    if (index >= 0 && index < REBX_B200_MAX_TABLES) {
        n = (int64_t)T.h[index].size();
        for (int64_t k = 0; k < n && k < cap; k++) out[k] = T.h[index][(size_t)k];
    } else if (index == -1) {
        n = (int64_t)T.hi.size();
        for (int64_t k = 0; k < n && k < cap; k++) out[k] = (double)T.hi[(size_t)k];
    } else if (index == -2) {
        n = (int64_t)T.bodies.size();
        for (int64_t k = 0; k < n && k < cap; k++) out[k] = (double)T.bodies[(size_t)k];
    }
    scratch.erase(force);
    return n;
}

// Counters of the host-facing path: {calls, kernel launches, H2D bytes, D2H bytes, table builds}.
int rebx_b200_get_stats(struct rebx_extras* rebx, uint64_t out[5]) {
    if (!rebx) return -1;
    Side* side = side_of(rebx);
    if (!side->dev) { for (int k = 0; k < 5; k++) out[k] = 0; return 0; }
    const Stats& s = side->dev->stats;
    out[0] = s.calls; out[1] = s.launches; out[2] = s.h2d_bytes; out[3] = s.d2h_bytes; out[4] = s.table_builds;
    return 0;
}

}  // extern "C"

#include "session.inl"

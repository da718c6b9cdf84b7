// api_core.cpp -- registry, parameter and force-object API of the drop-in library.
//
// From-scratch implementation of the REBOUNDx 5.1.0 set-up API that surrounds the force
// path (reference src/core.c:51-1012,1091-1186 and src/linkedlist.c).  Same observable
// behaviour -- list order, error strings' meaning, ownership rules -- plus what the device
// path needs: every mutation bumps the owning extras' "generation" so the SoA parameter
// tables in HBM are rebuilt only when something may have changed, and parameter names
// are interned so the table builder matches them by pointer.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <set>

#include "internal.h"

#define REBXB_STR2(s) #s
#define REBXB_STR(s) REBXB_STR2(s)
#ifndef REBXGITHASH
#define REBXGITHASH b200native0000000000000000000000000000001
#endif

extern "C" {
const char* rebx_build_str   = __DATE__ " " __TIME__;
const char* rebx_version_str = "5.1.0";                 // API level of the reference this build mirrors
const char* rebx_githash_str = REBXB_STR(REBXGITHASH);
}

namespace rebxh {

// ---- side table ------------------------------------------------------------------------
static std::mutex g_mu;
static std::unordered_map<struct rebx_extras*, std::unique_ptr<Side>>& sides() {
    static auto* m = new std::unordered_map<struct rebx_extras*, std::unique_ptr<Side>>();
    return *m;
}
static std::atomic<std::uint64_t> g_epoch{1};

static void forget_side(Side* side);

Side* side_of(struct rebx_extras* rebx) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& m = sides();
    auto it = m.find(rebx);
    if (it == m.end()) it = m.emplace(rebx, std::make_unique<Side>()).first;
    return it->second.get();
}

void side_drop(struct rebx_extras* rebx) {
    std::unique_ptr<Side> victim;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = sides().find(rebx);
        if (it != sides().end()) forget_side(it->second.get());
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto& m = sides();
        auto it = m.find(rebx);
        if (it == m.end()) return;
        victim = std::move(it->second);
        m.erase(it);
    }
    // device buffers are released by ~Side outside the lock
}

void mark_dirty(struct rebx_extras* rebx) {
    if (rebx) side_of(rebx)->generation++;
}

// ---- escaped value pointers -----------------------------------------------------------------
static std::mutex g_esc_mu;
static std::unordered_map<const struct rebx_param*, std::pair<Side*, size_t>>& escaped_index() {
    static auto* m = new std::unordered_map<const struct rebx_param*, std::pair<Side*, size_t>>();
    return *m;
}
static std::atomic<size_t> g_escaped_count{0};

static bool value_bits(const struct rebx_param* p, std::uint64_t& out) {
    if (!p || !p->value) return false;
    switch (p->type) {
        case REBX_TYPE_DOUBLE: std::memcpy(&out, p->value, 8); return true;
        case REBX_TYPE_INT: case REBX_TYPE_UINT32: { std::uint32_t v; std::memcpy(&v, p->value, 4); out = v; return true; }
        case REBX_TYPE_VEC3D: {
            std::uint64_t w[3];
            std::memcpy(w, p->value, 24);
            out = w[0]*0x9e3779b97f4a7c15ULL ^ (w[1] + 0x7f4a7c15ULL)*0xbf58476d1ce4e5b9ULL ^ (w[2] + 0x1ce4e5b9ULL)*0x94d049bb133111ebULL;
            return true;
        }
        default: return false;          // pointers, strings, forces: nothing a parameter table is built from
    }
}

static void note_escape(struct rebx_extras* rebx, struct rebx_param* p) {
    std::uint64_t bits;
    if (!rebx || !value_bits(p, bits)) return;
    Side* side = side_of(rebx);
    std::lock_guard<std::mutex> lk(g_esc_mu);
    auto& idx = escaped_index();
    if (idx.find(p) != idx.end()) return;
    idx.emplace(p, std::make_pair(side, side->escaped.size()));
    side->escaped.push_back(Escaped{p, bits});
    g_escaped_count.fetch_add(1, std::memory_order_relaxed);
}

static void forget_escape(const struct rebx_param* p) {          // the parameter is being freed
    if (g_escaped_count.load(std::memory_order_relaxed) == 0) return;
    std::lock_guard<std::mutex> lk(g_esc_mu);
    auto& idx = escaped_index();
    auto it = idx.find(p);
    if (it == idx.end()) return;
    it->second.first->escaped[it->second.second].param = nullptr;
    idx.erase(it);
    g_escaped_count.fetch_sub(1, std::memory_order_relaxed);
}

static void forget_side(Side* side) {
    if (side->escaped.empty()) return;
    std::lock_guard<std::mutex> lk(g_esc_mu);
    auto& idx = escaped_index();
    for (const Escaped& e : side->escaped) if (e.param && idx.erase(e.param)) g_escaped_count.fetch_sub(1, std::memory_order_relaxed);
    side->escaped.clear();
}

void refresh_escaped(Side* side) {
    if (side->escaped.empty()) return;
    bool changed = false;
    auto scan = [&](int, size_t lo, size_t hi) {
        bool any = false;
        for (size_t k = lo; k < hi; k++) {
            Escaped& e = side->escaped[k];
            std::uint64_t bits;
            if (e.param && value_bits(e.param, bits) && bits != e.bits) { e.bits = bits; any = true; }
        }
        if (any) changed = true;        // benign race: every writer stores `true`
    };
    parallel_ranges(side->escaped.size(), scan);
    if (changed) side->generation++;
}
void mark_all_dirty() { g_epoch.fetch_add(1, std::memory_order_relaxed); }
std::uint64_t global_epoch() { return g_epoch.load(std::memory_order_relaxed); }

// ---- interned names ----------------------------------------------------------------------
struct CStrLess { bool operator()(const char* a, const char* b) const { return std::strcmp(a, b) < 0; } };
const char* intern_name(const char* name) {
    static std::mutex mu;
    static auto* pool = new std::set<const char*, CStrLess>();
    std::lock_guard<std::mutex> lk(mu);
    auto it = pool->find(name);
    if (it != pool->end()) return *it;
    char* copy = strdup(name);
    pool->insert(copy);
    return copy;
}

struct rebx_param* find_param(struct rebx_node* ap, const char* name) {
    for (struct rebx_node* n = ap; n != nullptr; n = n->next) {
        struct rebx_param* p = static_cast<struct rebx_param*>(n->object);
        if (p->name == name || std::strcmp(p->name, name) == 0) return p;
    }
    return nullptr;
}

}  // namespace rebxh

using namespace rebxh;

// Default registry: the names and types the reference registers in src/core.c:51-159.
namespace {
struct RegEntry { const char* name; enum rebx_param_type type; };
const char* const kDoubleParams[] = {
    "c", "tau_mass", "Acentral", "gammacentral", "J2", "J4", "R_eq", "beta", "luminosity", "p", "I", "k2", "tau",
    "tau_a", "tau_e", "tau_inc", "tau_omega", "tau_Omega", "tau_coeff", "cs_coeff", "d_factor",
    "em_aini", "em_afin", "em_tau_a", "ide_position", "ide_width",
    "tIm_flaring_index", "tIm_scale_height_1", "tIm_surface_density_1", "tIm_surface_density_exponent",
    "ye_c", "ye_body_density", "ye_lstar", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo",
    "ye_emissivity", "ye_k", "ye_stef_boltz", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z",
    "tctl_k2", "tctl_tau", "R_tides", "OmegaMag", "min_distance",
    "kappa", "kappa_x", "kappa_y", "kappa_z", "tau_kappa", "tau_kappa_x", "tau_kappa_y", "tau_kappa_z",
    "stochastic_force_r", "stochastic_force_phi", "stochastic_force_x", "stochastic_force_y", "stochastic_force_z",
    "gas_df_rhog", "gas_df_alpha_rhog", "gas_df_cs", "gas_df_alpha_cs", "gas_df_xmin", "gas_df_hr", "gas_df_Qd",
    "lt_p_hatx", "lt_p_haty", "lt_p_hatz", "lt_Mom_I_fac", "lt_rot_rate", "lt_R_eq", "lt_c",
    "td_dP_hat", "td_dP_crit", "td_EB0", "td_E_max", "td_E_resid", "td_dE_last", "td_last_apoapsis",
    "td_drag_coef", "td_M_last", "td_c_real", "td_c_imag",
};
const char* const kIntParams[] = {
    "gr_source", "max_iterations", "coordinates", "primary", "radiation_source", "integrator",
    "ye_flag", "tides_primary", "td_disruption_flag", "td_num_apoapsis",
};
const char* const kPointerParams[] = { "particle", "im_ps_final", "im_ps_prev", "im_ps_avg", "rk2_k2", "rk4_k2", "rk4_k3" };
const RegEntry kOtherParams[] = {
    {"force", REBX_TYPE_FORCE}, {"ode", REBX_TYPE_ODE}, {"min_distance_orbit", REBX_TYPE_ORBIT},
    {"Omega", REBX_TYPE_VEC3D}, {"min_distance_from", REBX_TYPE_STRING},
};
}  // namespace

extern "C" {

// ---- linked list (reference src/linkedlist.c:33-106): push-front, so lists are LIFO -------
void rebx_add_node(struct rebx_node** head, struct rebx_node* node) {
    node->next = *head;
    *head = node;
}

int rebx_remove_node(struct rebx_node** head, void* object) {
    for (struct rebx_node** link = head; *link != nullptr; link = &(*link)->next) {
        if ((*link)->object == object) {
            struct rebx_node* dead = *link;
            *link = dead->next;
            free(dead);
            return 1;
        }
    }
    return 0;
}

int rebx_len(struct rebx_node* head) {
    int n = 0;
    for (; head != nullptr; head = head->next) n++;
    return n;
}

// ---- errors and memory (reference src/core.c:854-862,1179-1186) ----------------------------
void rebx_error(struct rebx_extras* rebx, const char* const msg) {
    if (rebx == nullptr || rebx->sim == nullptr) {
        fprintf(stderr, "REBOUNDx Error: A Simulation is no longer attached to this REBOUNDx extras instance. Most likely the Simulation has been freed.\n");
        return;
    }
    reb_simulation_error(rebx->sim, msg);
}

void* rebx_malloc(struct rebx_extras* const rebx, size_t memsize) {
    void* ptr = malloc(memsize);
    if (ptr == nullptr && memsize > 0) rebx_error(rebx, "REBOUNDx Error: Could not allocate memory.\n");
    return ptr;
}

struct rebx_node* rebx_create_node(struct rebx_extras* rebx) {
    struct rebx_node* node = static_cast<struct rebx_node*>(rebx_malloc(rebx, sizeof(*node)));
    if (node) { node->object = nullptr; node->next = nullptr; }
    return node;
}

// A param's name is an interned string shared by all params of that name; never freed.
struct rebx_param* rebx_create_param(struct rebx_extras* rebx, const char* name, enum rebx_param_type type) {
    struct rebx_param* param = static_cast<struct rebx_param*>(rebx_malloc(rebx, sizeof(*param)));
    if (!param) return nullptr;
    param->type = type;
    param->value = nullptr;
    param->name = const_cast<char*>(intern_name(name));
    return param;
}

int rebx_add_param(struct rebx_extras* const rebx, struct rebx_node** apptr, struct rebx_param* param) {
    struct rebx_node* node = rebx_create_node(rebx);
    if (!node) return 0;
    node->object = param;
    rebx_add_node(apptr, node);
    mark_dirty(rebx);       // the binary loader inserts params through here (reference src/input.c:127)
    return 1;
}

// reference src/core.c:864-875: only INT and DOUBLE payloads are owned by the list
void rebx_free_param(struct rebx_param* param) {
    if (!param) return;
    forget_escape(param);
    // param->name is interned (rebx_create_param) and shared: never freed here
    if ((param->type == REBX_TYPE_INT || param->type == REBX_TYPE_DOUBLE) && param->value) free(param->value);
    free(param);
}

void rebx_free_ap(struct rebx_node** ap) {
    struct rebx_node* cur = *ap;
    while (cur) {
        struct rebx_node* next = cur->next;
        rebx_free_param(static_cast<struct rebx_param*>(cur->object));
        free(cur);
        cur = next;
    }
    *ap = nullptr;
}

// Installed as sim->free_particle_ap (REBOUND calls it when a particle is removed or freed).
void rebx_free_particle_ap(struct reb_particle* p) {
    // Unconditionally: REBOUND calls this whenever a particle is removed and then shifts or swaps the particles
    // behind it, which moves every per-index table entry and cached body index even if the removed particle
    // itself carried no parameters.
    mark_all_dirty();
    rebx_free_ap(reinterpret_cast<struct rebx_node**>(&p->ap));
}

void rebx_free_force(struct rebx_extras* rebx, struct rebx_force* force) {
    if (force->free_memory) force->free_memory(rebx, force);
    free(force->name);
    rebx_free_ap(&force->ap);
    free(force);
}

void rebx_reset_accelerations(struct reb_particle* const ps, const int N) {
    for (int i = 0; i < N; i++) { ps[i].ax = 0.; ps[i].ay = 0.; ps[i].az = 0.; }
}

// ---- registry ------------------------------------------------------------------------------
enum rebx_param_type rebx_get_type(struct rebx_extras* rebx, const char* name) {
    Side* s = side_of(rebx);
    auto it = s->types.find(name);
    if (it != s->types.end()) return static_cast<enum rebx_param_type>(it->second);
    // registry entries added behind our back (e.g. by a loader that links this library)
    struct rebx_param* p = find_param(rebx->registered_params, name);
    return p ? p->type : REBX_TYPE_NONE;
}

void rebx_register_param(struct rebx_extras* const rebx, const char* name, enum rebx_param_type type) {
    if (rebx_get_type(rebx, name) != REBX_TYPE_NONE) {
        char buf[300];
        snprintf(buf, sizeof buf, "REBOUNDx Error: Parameter name '%s' already in registered list. Cannot add duplicates.\n", name);
        rebx_error(rebx, buf);
        return;
    }
    struct rebx_param* param = rebx_create_param(rebx, name, type);
    if (!param) return;
    if (!rebx_add_param(rebx, &rebx->registered_params, param)) { rebx_free_param(param); return; }
    side_of(rebx)->types[name] = static_cast<int>(type);
}

void rebx_register_default_params(struct rebx_extras* rebx) {
    for (const char* n : kDoubleParams)  rebx_register_param(rebx, n, REBX_TYPE_DOUBLE);
    for (const char* n : kIntParams)     rebx_register_param(rebx, n, REBX_TYPE_INT);
    for (const char* n : kPointerParams) rebx_register_param(rebx, n, REBX_TYPE_POINTER);
    for (const RegEntry& e : kOtherParams) rebx_register_param(rebx, e.name, e.type);
}

// ---- attach / detach / free (reference src/core.c:186-236,926-1012) ---------------------------
void rebx_extras_cleanup(struct reb_simulation* sim) {
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(sim->extras);
    if (rebx) rebx->sim = nullptr;      // REBOUND is freeing the simulation under us
}

void rebx_initialize(struct reb_simulation* sim, struct rebx_extras* rebx) {
    rebx->sim = sim;
    if (sim == nullptr) { rebx_error(rebx, ""); return; }
    sim->extras = rebx;
    rebx->additional_forces = nullptr;
    rebx->pre_timestep_modifications = nullptr;
    rebx->post_timestep_modifications = nullptr;
    rebx->allocated_forces = nullptr;
    rebx->allocated_operators = nullptr;
    rebx->registered_params = nullptr;
    side_drop(rebx);                    // a recycled address must not inherit stale tables
    side_of(rebx);
    sim->free_particle_ap = rebx_free_particle_ap;
    sim->extras_cleanup = rebx_extras_cleanup;
    if (sim->additional_forces || sim->pre_timestep_modifications || sim->post_timestep_modifications) {
        reb_simulation_warning(sim, "REBOUNDx overwrites sim->additional_forces, sim->pre_timestep_modifications and sim->post_timestep_modifications whenever forces or operators that use them get added.  If you want to use REBOUNDx together with your own custom functions that use these callbacks, you should add them through REBOUNDx.");
    }
}

struct rebx_extras* rebx_attach(struct reb_simulation* sim) {
    if (sim == nullptr) {
        fprintf(stderr, "REBOUNDx Error: Simulation pointer passed to rebx_attach was NULL.\n");
        return nullptr;
    }
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(malloc(sizeof(*rebx)));
    if (!rebx) return nullptr;
    rebx_initialize(sim, rebx);
    rebx_register_default_params(rebx);
    return rebx;
}

void rebx_detach(struct rebx_extras* rebx) {
    side_drop(rebx);
    struct reb_simulation* sim = rebx->sim;
    if (sim == nullptr) return;
    if (sim->extras == rebx) {
        if (sim->additional_forces == rebx_additional_forces) sim->additional_forces = nullptr;
        if (sim->pre_timestep_modifications == rebx_pre_timestep_modifications) sim->pre_timestep_modifications = nullptr;
        if (sim->post_timestep_modifications == rebx_post_timestep_modifications) sim->post_timestep_modifications = nullptr;
        if (sim->extras_cleanup == rebx_extras_cleanup) sim->extras_cleanup = nullptr;
        if (sim->free_particle_ap == rebx_free_particle_ap) sim->free_particle_ap = nullptr;
        sim->extras = nullptr;
    }
}

static void free_node_list(struct rebx_node* cur, void (*free_object)(struct rebx_extras*, void*), struct rebx_extras* rebx) {
    while (cur) {
        struct rebx_node* next = cur->next;
        if (free_object) free_object(rebx, cur->object);
        free(cur);
        cur = next;
    }
}

void rebx_free_pointers(struct rebx_extras* rebx) {
    if (rebx == nullptr) return;
    side_drop(rebx);                    // device tables, staging buffers, host registration
    if (rebx->sim != nullptr) {
        for (size_t i = 0; i < rebx->sim->N; i++) rebx_free_particle_ap(&rebx->sim->particles[i]);
    }
    free_node_list(rebx->allocated_forces,
                   [](struct rebx_extras* r, void* f) { rebx_free_force(r, static_cast<struct rebx_force*>(f)); }, rebx);
    free_node_list(rebx->additional_forces, nullptr, rebx);
    // operators (reference src/core.c:273-287): the steps own nothing but themselves
    free_node_list(rebx->allocated_operators,
                   [](struct rebx_extras* r, void* o) {
                       struct rebx_operator* op = static_cast<struct rebx_operator*>(o);
                       if (op->free_memory) { op->free_memory(r, op); }
                       rebx_free_operator(op);
                   }, rebx);
    free_node_list(rebx->pre_timestep_modifications, [](struct rebx_extras*, void* st) { free(st); }, rebx);
    free_node_list(rebx->post_timestep_modifications, [](struct rebx_extras*, void* st) { free(st); }, rebx);
    free_node_list(rebx->registered_params,
                   [](struct rebx_extras*, void* p) { free(p); }, rebx);   // registry entries carry no value
    rebx->allocated_forces = rebx->additional_forces = rebx->allocated_operators = nullptr;
    rebx->pre_timestep_modifications = rebx->post_timestep_modifications = rebx->registered_params = nullptr;
}

void rebx_free(struct rebx_extras* rebx) {
    rebx_free_pointers(rebx);
    rebx_detach(rebx);
    free(rebx);
}

// ---- force objects (reference src/core.c:238-360,455-498,748-786) ------------------------------
struct rebx_force* rebx_create_force(struct rebx_extras* const rebx, const char* name) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return nullptr; }
    struct rebx_force* force = static_cast<struct rebx_force*>(rebx_malloc(rebx, sizeof(*force)));
    if (!force) return nullptr;
    force->ap = nullptr;
    force->sim = rebx->sim;
    force->force_type = REBX_FORCE_NONE;
    force->update_accelerations = nullptr;
    force->free_memory = nullptr;
    force->name = nullptr;
    if (name != nullptr) {
        force->name = static_cast<char*>(rebx_malloc(rebx, strlen(name) + 1));
        if (!force->name) { free(force); return nullptr; }
        strcpy(force->name, name);
    }
    struct rebx_node* node = rebx_create_node(rebx);
    if (!node) { rebx_free_force(rebx, force); return nullptr; }
    node->object = force;
    rebx_add_node(&rebx->allocated_forces, node);
    return force;
}

struct rebx_force* rebx_load_force(struct rebx_extras* const rebx, const char* name) {
    struct BuiltIn {
        const char* name; enum rebx_force_type type;
        void (*fn)(struct reb_simulation* const, struct rebx_force* const, struct reb_particle* const, const int);
    };
    // The device-backed subset of the reference's name table (src/core.c:282-350).
    static const BuiltIn table[] = {
        {"radiation_forces",        REBX_FORCE_VEL, rebx_radiation_forces},
        {"gravitational_harmonics", REBX_FORCE_POS, rebx_gravitational_harmonics},
        {"gr_potential",            REBX_FORCE_POS, rebx_gr_potential},
        {"gr",                      REBX_FORCE_VEL, rebx_gr},
        {"yarkovsky_effect",        REBX_FORCE_VEL, rebx_yarkovsky_effect},
        {"modify_orbits_forces",    REBX_FORCE_VEL, rebx_modify_orbits_forces},
        {"gas_damping_timescale",   REBX_FORCE_VEL, rebx_gas_damping_timescale},
        {"type_I_migration",        REBX_FORCE_VEL, rebx_modify_orbits_with_type_I_migration},
        {"central_force",           REBX_FORCE_POS, rebx_central_force},
        {"lense_thirring",          REBX_FORCE_VEL, rebx_lense_thirring},
        {"gas_dynamical_friction",  REBX_FORCE_VEL, rebx_gas_dynamical_friction},
        {"exponential_migration",   REBX_FORCE_VEL, rebx_exponential_migration},
    };
    struct rebx_force* force = rebx_create_force(rebx, name);
    if (!force) return nullptr;
    for (const BuiltIn& b : table) {
        if (strcmp(name, b.name) == 0) {
            force->update_accelerations = b.fn;
            force->force_type = b.type;
            return force;
        }
    }
    char buf[300];
    snprintf(buf, sizeof buf, "REBOUNDx error: Force '%s' not found in REBOUNDx library.\n", name);
    rebx_error(rebx, buf);
    rebx_remove_force(rebx, force);
    return nullptr;
}

int rebx_add_force(struct rebx_extras* rebx, struct rebx_force* force) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return 0; }
    struct reb_simulation* const sim = rebx->sim;
    if (sim->integrator.name && strcmp(sim->integrator.name, "whfast512") == 0) {
        reb_simulation_error(sim, "REBOUNDx Error: WHFast512 has been optimized for speed with options stripped out. This integrator will never be compatible with REBOUNDx.\n");
        return 0;
    }
    if (force == nullptr) { rebx_error(rebx, "REBOUNDx error: Passed NULL pointer to rebx_add_force.\n"); return 0; }
    if (force->update_accelerations == nullptr) {
        rebx_error(rebx, "REBOUNDx error: Need to set update_accelerations function pointer on force before calling rebx_add_force. See custom effects example.\n");
        return 0;
    }
    if (force->force_type == REBX_FORCE_NONE) {
        rebx_error(rebx, "REBOUNDx error: Need to set force_type field on force before calling rebx_add_force. See custom effects example.\n");
        return 0;
    }
    if (force->force_type == REBX_FORCE_VEL) sim->force_is_velocity_dependent = 1;
    struct rebx_node* node = rebx_create_node(rebx);
    if (!node) return 0;
    node->object = force;
    rebx_add_node(&rebx->additional_forces, node);
    if (sim->additional_forces != nullptr && sim->additional_forces != rebx_additional_forces) {
        reb_simulation_warning(sim, "REBOUNDx Warning: additional_forces was set and is being overwritten by REBOUNDx. To incorporate both, you can add your own custom effects through REBOUNDx.\n");
    }
    sim->additional_forces = rebx_additional_forces;
    mark_dirty(rebx);
    return 1;
}

int rebx_remove_force(struct rebx_extras* rebx, struct rebx_force* force) {
    mark_dirty(rebx);
    if (rebx_remove_node(&rebx->allocated_forces, force)) rebx_free_force(rebx, force);
    return rebx_remove_node(&rebx->additional_forces, force);
}

struct rebx_force* rebx_get_force(struct rebx_extras* const rebx, const char* const name) {
    for (struct rebx_node* n = rebx->allocated_forces; n; n = n->next) {
        struct rebx_force* f = static_cast<struct rebx_force*>(n->object);
        if (f->name && strcmp(f->name, name) == 0) return f;
    }
    return nullptr;
}

// ---- parameters (reference src/core.c:604-746) ------------------------------------------------
// Getters hand out interior pointers the caller may write through, so they mark the tables
// dirty; the library's own effects use rebxh::find_param instead.
struct rebx_param* rebx_get_param_struct(struct rebx_extras* const rebx, struct rebx_node* ap, const char* const param_name) {
    struct rebx_param* p = find_param(ap, param_name);
    if (p && rebx) { mark_dirty(rebx); note_escape(rebx, p); }      // the caller may write now AND keep the pointer for later
    return p;
}

void* rebx_get_param(struct rebx_extras* const rebx, struct rebx_node* ap, const char* const param_name) {
    struct rebx_param* p = rebx_get_param_struct(rebx, ap, param_name);
    return p ? p->value : nullptr;
}

extern "C++" { static struct rebx_param* get_or_add_param_impl(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name); }

struct rebx_param* rebx_get_or_add_param(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name) {
    struct rebx_param* p = get_or_add_param_impl(rebx, apptr, param_name);
    if (p && rebx) note_escape(rebx, p);       // exported: a foreign caller holds the struct (and its value pointer) from here on
    return p;
}

extern "C++" {
static struct rebx_param* get_or_add_param_impl(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name) {
    if (apptr == nullptr) {
        rebx_error(rebx, "REBOUNDx Error: Passed NULL apptr to rebx_add_param. See examples.\n");
        return nullptr;
    }
    mark_dirty(rebx);
    struct rebx_param* param = find_param(*apptr, param_name);
    if (param) return param;
    const enum rebx_param_type type = rebx_get_type(rebx, param_name);
    if (type == REBX_TYPE_NONE) {
        char buf[300];
        snprintf(buf, sizeof buf, "REBOUNDx Error: Need to register parameter name '%s' before using it. See examples.\n", param_name);
        rebx_error(rebx, buf);
        return nullptr;
    }
    param = rebx_create_param(rebx, param_name, type);
    if (!param) return nullptr;
    if (!rebx_add_param(rebx, apptr, param)) { rebx_free_param(param); return nullptr; }
    return param;
}
}  // extern "C++"

extern "C++" {
template <typename T>
static void set_scalar(struct rebx_extras* rebx, struct rebx_node** apptr, const char* name, T val) {
    struct rebx_param* param = get_or_add_param_impl(rebx, apptr, name);
    if (!param) return;
    if (param->value == nullptr) param->value = rebx_malloc(rebx, sizeof(T));
    if (param->value) *static_cast<T*>(param->value) = val;
}
}  // extern "C++"

void rebx_set_param_double(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, double val) { set_scalar<double>(rebx, apptr, param_name, val); }
void rebx_set_param_int(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, int val) { set_scalar<int>(rebx, apptr, param_name, val); }
void rebx_set_param_uint32(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, uint32_t val) { set_scalar<uint32_t>(rebx, apptr, param_name, val); }
void rebx_set_param_vec3d(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, struct reb_vec3d val) { set_scalar<struct reb_vec3d>(rebx, apptr, param_name, val); }

void rebx_set_param_pointer(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, void* val) {
    struct rebx_param* param = get_or_add_param_impl(rebx, apptr, param_name);
    if (param) param->value = val;
}

void rebx_set_param_string(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, const char* const val) {
    struct rebx_param* param = get_or_add_param_impl(rebx, apptr, param_name);
    if (!param) return;
    if (param->value == nullptr) param->value = rebx_malloc(rebx, sizeof(char*));
    if (param->value) *static_cast<const char**>(param->value) = reb_simulation_register_name(rebx->sim, val);
}

// reference src/core.c:1141-1177 (used by the binary writer, src/output.c)
size_t rebx_sizeof(struct rebx_extras* rebx, enum rebx_param_type type) {
    switch (type) {
        case REBX_TYPE_DOUBLE:  return sizeof(double);
        case REBX_TYPE_INT:     return sizeof(int);
        case REBX_TYPE_UINT32:  return sizeof(uint32_t);
        case REBX_TYPE_FORCE:   return sizeof(struct rebx_force);
        case REBX_TYPE_VEC3D:   return sizeof(struct reb_vec3d);
        case REBX_TYPE_POINTER: case REBX_TYPE_STRING: case REBX_TYPE_ORBIT: case REBX_TYPE_ODE: return 0;
        default:
            rebx_error(rebx, "REBOUNDx Error: Parameter type passed to rebx_sizeof is not registered.\n");
            return 0;
    }
}

void rebx_mark_params_dirty(struct rebx_extras* rebx) { mark_dirty(rebx); }

// Bulk ingest (extension): same list structure as N calls of rebx_set_param_*, but the type
// lookup and name interning are done once.
extern "C++" {
template <typename T>
static void set_array(struct rebx_extras* rebx, const char* name, const int64_t* index, const T* val, size_t n, enum rebx_param_type want) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return; }
    const enum rebx_param_type type = rebx_get_type(rebx, name);
    if (type != want) {
        char buf[300];
        snprintf(buf, sizeof buf, "REBOUNDx Error: Parameter '%s' is not registered with the type of this array setter.\n", name);
        rebx_error(rebx, buf);
        return;
    }
    const char* iname = intern_name(name);
    struct reb_simulation* sim = rebx->sim;
    // Distinct particles are independent (each owns its list; malloc is thread-safe), so an index list that is
    // strictly increasing -- or absent -- is filled by several threads; anything else stays on one.
    bool increasing = true;
    for (size_t k = 1; index && k < n && increasing; k++) increasing = index[k] > index[k - 1];
    std::atomic<int> status{0};         // 1: index out of range, 2: out of memory
    auto fill = [&](int, size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
            const size_t i = index ? static_cast<size_t>(index[k]) : k;
            if (i >= sim->N) { status = 1; return; }
            struct rebx_node** apptr = reinterpret_cast<struct rebx_node**>(&sim->particles[i].ap);
            struct rebx_param* param = find_param(*apptr, iname);
            if (!param) {
                param = static_cast<struct rebx_param*>(malloc(sizeof *param));
                struct rebx_node* node = static_cast<struct rebx_node*>(malloc(sizeof *node));
                if (!param || !node) { free(param); free(node); status = 2; return; }
                param->name = const_cast<char*>(iname); param->type = type; param->value = nullptr;
                node->object = param; node->next = *apptr; *apptr = node;
            }
            if (param->value == nullptr) param->value = malloc(sizeof(T));
            if (!param->value) { status = 2; return; }
            *static_cast<T*>(param->value) = val[k];
        }
    };
    if (increasing) parallel_ranges(n, fill); else fill(0, 0, n);
    if (status == 1) rebx_error(rebx, "REBOUNDx Error: particle index out of range in array parameter setter.\n");
    if (status == 2) rebx_error(rebx, "REBOUNDx Error: Could not allocate memory.\n");
    mark_dirty(rebx);
}
}  // extern "C++"

void rebx_set_param_double_array(struct rebx_extras* const rebx, const char* const param_name, const int64_t* index, const double* val, size_t n) {
    set_array<double>(rebx, param_name, index, val, n, REBX_TYPE_DOUBLE);
}
void rebx_set_param_int_array(struct rebx_extras* const rebx, const char* const param_name, const int64_t* index, const int* val, size_t n) {
    set_array<int>(rebx, param_name, index, val, n, REBX_TYPE_INT);
}

// ---- host scalar helpers (reference src/radiation_forces.c:120-125, src/inner_disk_edge.c:61-77) --
double rebx_rad_calc_beta(const double G, const double c, const double source_mass, const double source_luminosity, const double radius, const double density, const double Q_pr) {
    return 3.*source_luminosity*Q_pr/(16.*M_PI*G*source_mass*c*density*radius);
}
double rebx_rad_calc_particle_radius(const double G, const double c, const double source_mass, const double source_luminosity, const double beta, const double density, const double Q_pr) {
    return 3.*source_luminosity*Q_pr/(16.*M_PI*G*source_mass*c*density*beta);
}
double rebx_calculate_planet_trap(const double r, const double dedge, const double hedge) {
    if (r > dedge*(1.0 + hedge)) return 1.0;
    if (dedge*(1.0 - hedge) < r) return 5.5*cos(((dedge*(1.0 + hedge) - r)*2*M_PI)/(4*hedge*dedge)) - 4.5;
    return -10.0;
}

}  // extern "C"

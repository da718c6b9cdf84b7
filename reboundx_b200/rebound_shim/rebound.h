/*
 * rebound.h -- stand-in for the REBOUND N-body host (hannorein/rebound, >= 5.0.0).
 *
 * REBOUND is an external, un-vendored dependency of REBOUNDx (reference
 * pyproject.toml:7-9, src/reboundx.h:35) and is not present in this image.  This header
 * declares the subset of REBOUND's public types and functions that the
 * `rebx_additional_forces` hot path touches, so that
 *   (1) the drop-in library in reboundx_b200/csrc/ and
 *   (2) the reference's own hot-path .c files (built into oracle/_ref/)
 * compile against the SAME host ABI.  In a real deployment this directory is replaced
 * by REBOUND's own header and library: nothing in the product depends
 * on field OFFSETS of these structs, only on field NAMES (the device unpack kernels take
 * offsetof()/sizeof() values computed from whichever rebound.h was compiled in).
 *
 * The arithmetic restated in rebound_shim.c (orbit elements, centre of mass, Jacobi
 * transforms) is written from the published REBOUND algorithms; it cannot be diffed
 * against REBOUND in this environment ("parity unpinned" at this boundary).
 */
#ifndef REBOUND_SHIM_H
#define REBOUND_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct reb_simulation;

/* 128 bytes: twelve doubles followed by four pointers. */
struct reb_particle {
    double x, y, z;
    double vx, vy, vz;
    double ax, ay, az;
    double m;
    double r;
    double last_collision;
    void* c;
    const char* name;
    void* ap;                      /* head of the REBOUNDx per-particle parameter list */
    struct reb_simulation* sim;
};

struct reb_vec3d { double x, y, z; };

struct reb_rotation { double ix, iy, iz, r; };

struct reb_orbit {
    double d, v, h, P, n, a, e, inc, Omega, omega, pomega, f, M, l, theta, T, rhill;
    double pal_h, pal_k, pal_ix, pal_iy;
    struct reb_vec3d hvec, evec;
};

struct reb_integrator_ias15_state { double epsilon; };

struct reb_integrator_shim {
    const char* name;
    void* state;
};

/* Subset of struct reb_simulation: the fields REBOUNDx reads or installs. */
struct reb_simulation {
    double t;
    double G;
    double softening;
    double dt;
    double dt_last_done;
    size_t N;
    size_t N_var;
    size_t N_active;               /* SIZE_MAX == unset */
    struct reb_particle* particles;
    int force_is_velocity_dependent;
    struct reb_integrator_shim integrator;
    void* extras;
    void (*additional_forces)(struct reb_simulation* r);
    void (*pre_timestep_modifications)(struct reb_simulation* r);
    void (*post_timestep_modifications)(struct reb_simulation* r);
    void (*free_particle_ap)(struct reb_particle* p);
    void (*extras_cleanup)(struct reb_simulation* r);

    /* shim-only bookkeeping (never read by REBOUNDx code) */
    size_t N_allocated;
    char** messages;               /* queued "E..."/"W..." strings */
    size_t messages_N;
    size_t messages_allocated;
    char** names;
    size_t names_N;
    struct reb_integrator_ias15_state ias15;
};

/* --- life cycle (REBOUND names) ------------------------------------------------------ */
struct reb_simulation* reb_simulation_create(void);
void reb_simulation_free(struct reb_simulation* r);
void reb_simulation_add(struct reb_simulation* r, struct reb_particle pt);
void reb_simulation_remove_particle(struct reb_simulation* r, size_t index, int keep_sorted);

/* --- messages ------------------------------------------------------------------------ */
void reb_simulation_error(struct reb_simulation* r, const char* msg);
void reb_simulation_warning(struct reb_simulation* r, const char* msg);
void reb_simulation_update_acceleration(struct reb_simulation* r);
const char* reb_simulation_register_name(struct reb_simulation* r, const char* name);

/* --- celestial mechanics helpers used on the hot path --------------------------------- */
struct reb_orbit reb_orbit_from_particle_err(double G, struct reb_particle p, struct reb_particle primary, int* err);
struct reb_orbit reb_orbit_from_particle(double G, struct reb_particle p, struct reb_particle primary);
struct reb_particle reb_simulation_com(struct reb_simulation* r);
struct reb_particle reb_particle_com_of_pair(struct reb_particle p1, struct reb_particle p2);
void reb_particle_isub(struct reb_particle* p1, struct reb_particle* p2);      /* p1 -= p2 (used by src/gas_dynamical_friction.c:113) */

void reb_transformations_inertial_to_jacobi_posvel(const struct reb_particle* const particles, struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active);
void reb_transformations_inertial_to_jacobi_posvelacc(const struct reb_particle* const particles, struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active);
void reb_transformations_jacobi_to_inertial_acc(struct reb_particle* const particles, const struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active);

void reb_simulation_irotate(struct reb_simulation* r, const struct reb_rotation q);
void reb_vec3d_irotate(struct reb_vec3d* v, const struct reb_rotation q);

/* --- shim-only helpers for tests and benches (not REBOUND API) ------------------------- */
/* Replace the particle array by N particles built from SoA arrays (any pointer may be NULL => 0). */
void reb_shim_set_particles(struct reb_simulation* r, size_t N,
                            const double* x, const double* y, const double* z,
                            const double* vx, const double* vy, const double* vz,
                            const double* m, const double* rad);
void reb_shim_set_acc(struct reb_simulation* r, const double* ax, const double* ay, const double* az);
void reb_shim_get_acc(const struct reb_simulation* r, double* ax, double* ay, double* az);
void reb_shim_get_posvel(const struct reb_simulation* r, double* x, double* y, double* z, double* vx, double* vy, double* vz);
void reb_shim_set_integrator(struct reb_simulation* r, const char* name);
size_t reb_shim_message_count(const struct reb_simulation* r);
const char* reb_shim_message(const struct reb_simulation* r, size_t i);
void reb_shim_clear_messages(struct reb_simulation* r);
/* Bulk parameter set-up: calls `setter(rebx, &particles[idx[k]].ap, name, val[k])` for k<n.
 * `setter` is the rebx_set_param_double / rebx_set_param_int of whichever REBOUNDx build is loaded. */
typedef void (*reb_shim_set_double_fn)(void* rebx, void** apptr, const char* name, double val);
typedef void (*reb_shim_set_int_fn)(void* rebx, void** apptr, const char* name, int val);
void reb_shim_bulk_set_double(struct reb_simulation* r, reb_shim_set_double_fn setter, void* rebx,
                              const char* name, const int64_t* idx, const double* val, size_t n);
void reb_shim_bulk_set_int(struct reb_simulation* r, reb_shim_set_int_fn setter, void* rebx,
                           const char* name, const int64_t* idx, const int* val, size_t n);
/* Address of particles[i].ap, for passing to rebx_set_param_* from ctypes. */
void** reb_shim_particle_ap(struct reb_simulation* r, size_t i);
/* Calls r->additional_forces(r) if installed; returns 1 if it was called. */
int reb_shim_call_additional_forces(struct reb_simulation* r);
/* Host integrator harness (NOT REBOUND's integrators, which are not available here): advances the
 * simulation `nsteps` fixed steps of r->dt.  Accelerations = Newtonian gravity of the first N_active
 * bodies (all, when unset) on every particle + r->additional_forces(r).
 *   scheme 0: drift-kick-drift leapfrog, one force evaluation per step  (stands in for WHFast)
 *   scheme 1: classical RK4, four force evaluations per step
 *   scheme 2: 15th-order Gauss-Radau predictor-corrector, fixed step   (the scheme behind IAS15, epsilon = 0:
 *             7 force evaluations per sweep over the nodes, sweeps repeated to convergence)
 *   scheme 3: Wisdom-Holman drift-kick-drift: Kepler drifts about particle 0 + kicks by additional_forces
 *             only (the splitting behind WHFast, one massive central body, N_active = 1)
 *   scheme 4: the same with the half drifts between consecutive kicks merged into one drift of dt
 *             (WHFast between outputs); the state is synchronised at the end of the call only
 * Used to compare trajectories between two force implementations with everything else equal. */
void reb_shim_integrate(struct reb_simulation* r, int nsteps, int scheme);
/* Wall-clock seconds for `reps` calls of r->additional_forces(r) (best single call). */
double reb_shim_time_additional_forces(struct reb_simulation* r, int reps);

#ifdef __cplusplus
}
#endif
#endif /* REBOUND_SHIM_H */

/*
 * reboundx.h -- public C API of the B200-native drop-in for REBOUNDx's force path.
 *
 * This header re-declares, from scratch, the part of REBOUNDx 5.1.0's C API that the
 * `rebx_additional_forces` hot path and its set-up calls need.  Struct layouts and enum
 * values are ABI: the reference's ctypes wrapper mirrors them field for field
 * (reference reboundx/extras.py:251-258,319-333,376-381; reboundx/tools.py:4), so user C
 * code and the unmodified Python wrapper bind to this library exactly as they bind to the
 * reference's libreboundx.  Each declaration cites the reference declaration it replaces.
 *
 * The callers either side of the force path are declared too: the operator API with the
 * built-in operators "integrate_force", "drift" and "kick", and the binary state format.
 *
 * Out of scope: every other built-in operator, the splitting integrators,
 * interpolation and the other built-in forces.  Their declarations are intentionally
 * absent; a custom force or operator (user function pointer) still runs on the host through the
 * same dispatcher.
 */
#ifndef REBOUNDX_B200_REBOUNDX_H
#define REBOUNDX_B200_REBOUNDX_H

#include <stdint.h>
#include <stdio.h>
#include "rebound.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/reboundx.h:49 */
#define AP_PTR(p) ((struct rebx_node**)&(p).ap)

/* reference src/reboundx.h:50-52 / src/core.c:41-43 (read by reboundx/__init__.py:19-26) */
extern const char* rebx_build_str;
extern const char* rebx_version_str;
extern const char* rebx_githash_str;

/* reference src/reboundx.h:61-72 */
enum rebx_param_type {
    REBX_TYPE_NONE    = 0,
    REBX_TYPE_DOUBLE  = 1,
    REBX_TYPE_INT     = 2,
    REBX_TYPE_POINTER = 3,
    REBX_TYPE_FORCE   = 4,
    REBX_TYPE_UINT32  = 5,
    REBX_TYPE_ORBIT   = 6,
    REBX_TYPE_ODE     = 7,
    REBX_TYPE_VEC3D   = 8,
    REBX_TYPE_STRING  = 9
};

/* reference src/reboundx.h:77-81 */
enum REBX_COORDINATES {
    REBX_COORDINATES_JACOBI      = 0,
    REBX_COORDINATES_BARYCENTRIC = 1,
    REBX_COORDINATES_PARTICLE    = 2
};

/* reference src/reboundx.h:94-98 */
enum rebx_force_type {
    REBX_FORCE_NONE = 0,
    REBX_FORCE_POS  = 1,   /* derivable from a position-only potential */
    REBX_FORCE_VEL  = 2    /* depends on velocities */
};

/* reference src/reboundx.h:194-197 */
struct rebx_node {
    void* object;
    struct rebx_node* next;
};

/* reference src/reboundx.h:203-207 */
struct rebx_param {
    char* name;
    enum rebx_param_type type;
    void* value;
};

/* reference src/reboundx.h:213-223.  Exactly seven pointers: Python allocates this struct
 * itself (extras.py:39-45,327-333), so no field may be added -- all device state lives in
 * a side table keyed by the address of this struct. */
struct rebx_extras {
    struct reb_simulation* sim;
    struct rebx_node* additional_forces;
    struct rebx_node* pre_timestep_modifications;
    struct rebx_node* post_timestep_modifications;
    struct rebx_node* registered_params;
    struct rebx_node* allocated_forces;
    struct rebx_node* allocated_operators;
};

/* reference src/reboundx.h:228-236 */
struct rebx_force {
    char* name;
    struct rebx_node* ap;
    struct reb_simulation* sim;
    enum rebx_force_type force_type;
    void (*update_accelerations)(struct reb_simulation* const sim, struct rebx_force* const force,
                                 struct reb_particle* const particles, const int N);
    void (*free_memory)(struct rebx_extras* rebx, struct rebx_force* const force);
};

/* ---- operators (the callers either side of the force path) -------------------------
 * reference src/reboundx.h:86-89,103-107,170-176,241-258 */
enum rebx_timing { REBX_TIMING_PRE = -1, REBX_TIMING_POST = 1 };
enum rebx_operator_type { REBX_OPERATOR_NONE = 0, REBX_OPERATOR_UPDATER = 1, REBX_OPERATOR_RECORDER = 2 };
enum rebx_integrator {
    REBX_INTEGRATOR_NONE = -1, REBX_INTEGRATOR_IMPLICIT_MIDPOINT = 0, REBX_INTEGRATOR_RK4 = 1,
    REBX_INTEGRATOR_EULER = 2, REBX_INTEGRATOR_RK2 = 3
};
struct rebx_operator {
    char* name;
    struct rebx_node* ap;
    struct reb_simulation* sim;
    enum rebx_operator_type operator_type;
    void (*step_function)(struct reb_simulation* sim, struct rebx_operator* operator_, const double dt);
    void (*free_memory)(struct rebx_extras* rebx, struct rebx_operator* const operator_);
};
struct rebx_step {
    struct rebx_operator* operator_;    /* `operator` in the reference (a C++ keyword; same layout) */
    double dt_fraction;
};

/* ---- attach / free: reference src/reboundx.h:294-304, src/core.h:37-38 ---------------- */
struct rebx_extras* rebx_attach(struct reb_simulation* sim);
void rebx_initialize(struct reb_simulation* sim, struct rebx_extras* rebx);
void rebx_register_default_params(struct rebx_extras* rebx);
void rebx_detach(struct rebx_extras* rebx);
void rebx_extras_cleanup(struct reb_simulation* sim);
void rebx_free(struct rebx_extras* rebx);
void rebx_free_pointers(struct rebx_extras* rebx);

/* ---- forces: reference src/reboundx.h:306,355-359,385 --------------------------------- */
struct rebx_force* rebx_create_force(struct rebx_extras* const rebx, const char* name);
struct rebx_force* rebx_load_force(struct rebx_extras* const rebx, const char* name);
int rebx_add_force(struct rebx_extras* rebx, struct rebx_force* force);
int rebx_remove_force(struct rebx_extras* rebx, struct rebx_force* force);
struct rebx_force* rebx_get_force(struct rebx_extras* const rebx, const char* const name);

/* ---- parameters: reference src/reboundx.h:427-435,631, src/core.h:100-106 -------------- */
void  rebx_register_param(struct rebx_extras* const rebx, const char* name, enum rebx_param_type type);
enum rebx_param_type rebx_get_type(struct rebx_extras* rebx, const char* name);
void* rebx_get_param(struct rebx_extras* const rebx, struct rebx_node* ap, const char* const param_name);
struct rebx_param* rebx_get_param_struct(struct rebx_extras* const rebx, struct rebx_node* ap, const char* const param_name);
struct rebx_param* rebx_get_or_add_param(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name);
void rebx_set_param_pointer(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, void* val);
void rebx_set_param_double(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, double val);
void rebx_set_param_int(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, int val);
void rebx_set_param_uint32(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, uint32_t val);
void rebx_set_param_vec3d(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, struct reb_vec3d val);
void rebx_set_param_string(struct rebx_extras* const rebx, struct rebx_node** apptr, const char* const param_name, const char* const val);
struct rebx_param* rebx_create_param(struct rebx_extras* rebx, const char* name, enum rebx_param_type type);
int  rebx_add_param(struct rebx_extras* const rebx, struct rebx_node** apptr, struct rebx_param* param);
struct rebx_node* rebx_create_node(struct rebx_extras* rebx);
size_t rebx_sizeof(struct rebx_extras* rebx, enum rebx_param_type type);   /* reference src/core.h:52 */

/* ---- lists and memory: reference src/linkedlist.h:34-57, src/core.h:90-98 --------------- */
void rebx_add_node(struct rebx_node** head, struct rebx_node* node);
int  rebx_remove_node(struct rebx_node** head, void* object);
int  rebx_len(struct rebx_node* head);
void* rebx_malloc(struct rebx_extras* const rebx, size_t memsize);
void rebx_free_ap(struct rebx_node** ap);
void rebx_free_particle_ap(struct reb_particle* p);
void rebx_free_force(struct rebx_extras* rebx, struct rebx_force* force);
void rebx_free_param(struct rebx_param* param);
void rebx_error(struct rebx_extras* rebx, const char* const msg);
void rebx_reset_accelerations(struct reb_particle* const ps, const int N);

/* ---- operators: reference src/reboundx.h:307,353-358,393,684,691, src/core.h:77-80; src/core.c:362-453,500-602,
 * 1041-1089.  Built-in names of this build: "integrate_force" (the force path's own operator: velocities advanced
 * across dt under ONE force, src/integrate_force.c:34-78, on the device when that force is device-backed), "drift" and
 * "kick" (src/steppers.c:109-130).  Every other built-in operator name reports "not found"; custom operators
 * (rebx_create_operator + step_function) run on the host unchanged. */
struct rebx_operator* rebx_create_operator(struct rebx_extras* const rebx, const char* name);
struct rebx_operator* rebx_load_operator(struct rebx_extras* const rebx, const char* name);
int rebx_add_operator(struct rebx_extras* rebx, struct rebx_operator* operator_);
int rebx_add_operator_step(struct rebx_extras* rebx, struct rebx_operator* operator_, const double dt_fraction, enum rebx_timing timing);
int rebx_remove_operator(struct rebx_extras* rebx, struct rebx_operator* operator_);
struct rebx_operator* rebx_get_operator(struct rebx_extras* const rebx, const char* const name);
void rebx_free_operator(struct rebx_operator* operator_);
void rebx_pre_timestep_modifications(struct reb_simulation* sim);
void rebx_post_timestep_modifications(struct reb_simulation* sim);
void rebx_integrate_force(struct reb_simulation* const sim, struct rebx_operator* const operator_, const double dt);   /* src/integrate_force.c:34 */
void rebx_drift_step(struct reb_simulation* const sim, struct rebx_operator* const operator_, const double dt);        /* src/steppers.c:109 */
void rebx_kick_step(struct reb_simulation* const sim, struct rebx_operator* const operator_, const double dt);         /* src/steppers.c:120 */

/* ---- the hot path --------------------------------------------------------------------- */
/* Dispatcher installed as sim->additional_forces: reference src/core.c:1026-1041. */
void rebx_additional_forces(struct reb_simulation* sim);

/* Plug-in entry points with the reference's force signature (src/reboundx.h:234); each one
 * replaces the host loop of the same name (prototypes: reference src/core.h:59-75). */
void rebx_radiation_forces(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);        /* src/radiation_forces.c:100 */
void rebx_gravitational_harmonics(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N); /* src/gravitational_harmonics.c:136 */
void rebx_gr_potential(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);            /* src/gr_potential.c:80 */
void rebx_gr(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);                      /* src/gr.c:166 */
void rebx_yarkovsky_effect(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);        /* src/yarkovsky_effect.c:218 */
void rebx_modify_orbits_forces(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);    /* src/modify_orbits_forces.c:130 */
void rebx_gas_damping_timescale(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);   /* src/gas_damping_timescale.c:132 */
void rebx_modify_orbits_with_type_I_migration(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N); /* src/type_I_migration.c:205 */

/* Widened per SURVEY.md section 8f rank 3 (same-shape neighbouring forces): */
void rebx_central_force(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);         /* src/central_force.c:87 */
void rebx_lense_thirring(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N);        /* src/lense_thirring.c:102 */
void rebx_gas_dynamical_friction(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N); /* src/gas_dynamical_friction.c:139 */
void rebx_exponential_migration(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N); /* src/exponential_migration.c:103 */

/* Diagnostic potentials of the device-backed effects: reference
 * src/reboundx.h:553,575,582; reduced on the device. */
double rebx_gr_potential_potential(struct rebx_extras* const rebx, const struct rebx_force* const gr_potential);
double rebx_gravitational_harmonics_potential(struct rebx_extras* const rebx);
double rebx_central_force_potential(struct rebx_extras* const rebx);

/* Host scalar helpers: reference src/reboundx.h:497,510; src/inner_disk_edge.c:61 */
double rebx_rad_calc_beta(const double G, const double c, const double source_mass, const double source_luminosity, const double radius, const double density, const double Q_pr);
double rebx_rad_calc_particle_radius(const double G, const double c, const double source_mass, const double source_luminosity, const double beta, const double density, const double Q_pr);
double rebx_calculate_planet_trap(const double r, const double dedge, const double hedge);

/* ---- binary state files: reference src/reboundx.h:112-165,263-266,
 * 720-769; written from scratch in reboundx_b200/csrc/binary_io.cpp, byte-compatible with the files of
 * the reference's src/output.c / src/input.c.  Operators are out of scope: their lists are written empty
 * and skipped (with the reference's OPERATOR_NOT_LOADED warning) on load. */
enum rebx_binary_field_type {
    REBX_BINARY_FIELD_TYPE_NONE = 0, REBX_BINARY_FIELD_TYPE_OPERATOR = 1, REBX_BINARY_FIELD_TYPE_PARTICLE = 2,
    REBX_BINARY_FIELD_TYPE_REBX_STRUCTURE = 3, REBX_BINARY_FIELD_TYPE_PARAM = 4, REBX_BINARY_FIELD_TYPE_NAME = 5,
    REBX_BINARY_FIELD_TYPE_PARAM_TYPE = 6, REBX_BINARY_FIELD_TYPE_PARAM_VALUE = 7, REBX_BINARY_FIELD_TYPE_END = 8,
    REBX_BINARY_FIELD_TYPE_PARTICLE_INDEX = 9, REBX_BINARY_FIELD_TYPE_REBX_INTEGRATOR = 10, REBX_BINARY_FIELD_TYPE_FORCE_TYPE = 11,
    REBX_BINARY_FIELD_TYPE_OPERATOR_TYPE = 12, REBX_BINARY_FIELD_TYPE_STEP = 13, REBX_BINARY_FIELD_TYPE_STEP_DT_FRACTION = 14,
    REBX_BINARY_FIELD_TYPE_REGISTERED_PARAM = 15, REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCE = 16, REBX_BINARY_FIELD_TYPE_PARAM_LIST = 17,
    REBX_BINARY_FIELD_TYPE_REGISTERED_PARAMETERS = 18, REBX_BINARY_FIELD_TYPE_ALLOCATED_FORCES = 19,
    REBX_BINARY_FIELD_TYPE_ALLOCATED_OPERATORS = 20, REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCES = 21,
    REBX_BINARY_FIELD_TYPE_PRE_TIMESTEP_MODIFICATIONS = 22, REBX_BINARY_FIELD_TYPE_POST_TIMESTEP_MODIFICATIONS = 23,
    REBX_BINARY_FIELD_TYPE_PARTICLES = 24, REBX_BINARY_FIELD_TYPE_FORCE = 25, REBX_BINARY_FIELD_TYPE_SNAPSHOT = 26
};
enum rebx_input_binary_messages {
    REBX_INPUT_BINARY_WARNING_NONE = 0, REBX_INPUT_BINARY_ERROR_NOFILE = 1, REBX_INPUT_BINARY_ERROR_CORRUPT = 2,
    REBX_INPUT_BINARY_ERROR_NO_MEMORY = 4, REBX_INPUT_BINARY_ERROR_REBX_NOT_LOADED = 8,
    REBX_INPUT_BINARY_ERROR_REGISTERED_PARAM_NOT_LOADED = 16, REBX_INPUT_BINARY_WARNING_PARAM_NOT_LOADED = 32,
    REBX_INPUT_BINARY_WARNING_PARTICLE_PARAMS_NOT_LOADED = 64, REBX_INPUT_BINARY_WARNING_FORCE_NOT_LOADED = 128,
    REBX_INPUT_BINARY_WARNING_OPERATOR_NOT_LOADED = 256, REBX_INPUT_BINARY_WARNING_STEP_NOT_LOADED = 512,
    REBX_INPUT_BINARY_WARNING_ADDITIONAL_FORCE_NOT_LOADED = 1024, REBX_INPUT_BINARY_WARNING_FIELD_UNKNOWN = 2048,
    REBX_INPUT_BINARY_WARNING_LIST_UNKNOWN = 4096, REBX_INPUT_BINARY_WARNING_PARAM_VALUE_NULL = 8192,
    REBX_INPUT_BINARY_WARNING_VERSION = 16384, REBX_INPUT_BINARY_WARNING_FORCE_PARAM_NOT_LOADED = 32768
};
struct rebx_binary_field {
    enum rebx_binary_field_type type;
    long size;                          /* bytes that follow this header (children + END for a container) */
};
void rebx_output_binary(struct rebx_extras* rebx, char* filename);                                                  /* src/output.c:272 */
struct rebx_extras* rebx_create_extras_from_binary(struct reb_simulation* sim, const char* const filename);        /* src/input.c:634 */
void rebx_init_extras_from_binary(struct rebx_extras* rebx, const char* const filename, enum rebx_input_binary_messages* warnings); /* src/input.c:617 */
FILE* rebx_input_inspect_binary(const char* const filename, enum rebx_input_binary_messages* warnings);            /* src/input.c:692 */
struct rebx_binary_field rebx_input_read_binary_field(FILE* inf);                                                  /* src/input.c:702 */
void rebx_input_skip_binary_field(FILE* inf, long field_size);                                                     /* src/input.c:56 */

/* ---- additions of this build (all keep the rebx_ prefix, reference .github/workflows/c.yml:34) */
/* Forces a rebuild of the SoA parameter tables at the next force evaluation.  NOT needed for the reference's C
 * idiom of keeping a pointer from rebx_get_param and writing through it later: parameters whose pointer was handed
 * out are tracked and compared with the tables' copy at every call (api_core.cpp / dispatch.cpp, "escaped" params).
 * Kept for callers that change values behind the API in other ways. */
void rebx_mark_params_dirty(struct rebx_extras* rebx);
/* Bulk ingest: sets `param_name` on particles[index[k]] (index==NULL => k) for k<n. */
void rebx_set_param_double_array(struct rebx_extras* const rebx, const char* const param_name, const int64_t* index, const double* val, size_t n);
void rebx_set_param_int_array(struct rebx_extras* const rebx, const char* const param_name, const int64_t* index, const int* val, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* REBOUNDX_B200_REBOUNDX_H */

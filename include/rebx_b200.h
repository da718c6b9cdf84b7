/*
 * rebx_b200.h -- thin C-ABI device layer under the REBOUNDx force dispatcher.
 *
 * Plain pointers and sizes only: every `const double*` below is a DEVICE pointer into HBM
 * unless it says "host".  The layer is what the host dispatcher in
 * reboundx_b200/csrc/dispatch.cpp calls per force evaluation, and what a resident-state
 * caller (bench.py, a device integrator) calls directly with its own buffers and stream.
 *
 * One "pass" streams the particle state of one shard once and applies up to
 * REBX_B200_MAX_EFFECTS stacked effects in list order, i.e. it replaces that many
 * consecutive iterations of the reference's dispatch loop (src/core.c:1031-1040) and the
 * per-particle rebx_get_param walks inside each effect (src/core.c:725-746) by SoA
 * parameter tables.
 */
#ifndef REBX_B200_H
#define REBX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REBX_B200_MAX_EFFECTS   4
#define REBX_B200_BODY_DOUBLES  24
#define REBX_B200_MAX_TABLES    9
#define REBX_B200_MAX_SCALARS   8

/* "parameter absent on this particle" marker inside a double table: a quiet NaN with a
 * private payload, compared bit-wise (a NaN produced by arithmetic never has it). */
#define REBX_B200_ABSENT_BITS   0x7ff8b200dead0001ULL

/* Effect kinds; the comment names the reference loop each one replaces. */
enum rebx_b200_kind {
    REBX_B200_NONE            = 0,
    REBX_B200_RADIATION       = 1,  /* src/radiation_forces.c:69-98   */
    REBX_B200_HARMONICS       = 2,  /* src/gravitational_harmonics.c:184-224 (one oblate body per entry) */
    REBX_B200_GR_POTENTIAL    = 3,  /* src/gr_potential.c:60-78       */
    REBX_B200_YARKOVSKY       = 4,  /* src/yarkovsky_effect.c:76-216  */
    REBX_B200_MODIFY_ORBITS   = 5,  /* src/modify_orbits_forces.c:77-128 under src/rebxtools.c:66-147 */
    REBX_B200_GAS_DAMPING     = 6,  /* src/gas_damping_timescale.c:67-130 under rebx_com_force */
    REBX_B200_TYPE_I          = 7,  /* src/type_I_migration.c:115-203 under rebx_com_force */
    REBX_B200_CENTRAL_FORCE   = 8,  /* src/central_force.c:63-85 (one central body per entry)  [SURVEY 8f rank 3] */
    REBX_B200_EXP_MIGRATION   = 9,  /* src/exponential_migration.c:61-100 under rebx_com_force     [SURVEY 8f rank 3] */
    REBX_B200_LENSE_THIRRING  = 10, /* src/lense_thirring.c:65-100                                 [SURVEY 8f rank 3] */
    REBX_B200_GAS_DF          = 11, /* src/gas_dynamical_friction.c:66-137                         [SURVEY 8f rank 3] */
    REBX_B200_KIND_COUNT      = 12
};

/* Body record (REBX_B200_BODY_DOUBLES doubles, device memory): the source / primary /
 * oblate body an effect entry refers to.  Slots:
 *   0      global particle index of the body (stored as a double), or -1 for a pure
 *          reference frame that is no particle (barycentre)
 *   1..3   x y z     4..6  vx vy vz     7  m
 *   HARMONICS only: 8 J2, 9 J4 (0 when unset), 10 R_eq, 11..13 u-hat, 14..16 v-hat, 17..19 w-hat
 *   CENTRAL_FORCE only: 8 Acentral, 9 gammacentral
 *   LENSE_THIRRING only: 8..10 spin angular momentum J = I*Omega of the source
 */
enum { REBX_B200_BODY_INDEX = 0, REBX_B200_BODY_X = 1, REBX_B200_BODY_VX = 4, REBX_B200_BODY_M = 7,
       REBX_B200_BODY_J2 = 8, REBX_B200_BODY_J4 = 9, REBX_B200_BODY_REQ = 10,
       REBX_B200_BODY_U = 11, REBX_B200_BODY_V = 14, REBX_B200_BODY_W = 17 };

/* effect.flags bits */
#define REBX_B200_FLAG_TRAP        1   /* MODIFY_ORBITS: ide_position and ide_width both set */
#define REBX_B200_FLAG_BARYCENTRIC 2   /* damping trio: body = barycentre, mass ratio m/M */

/* SoA view of one shard of particle state resident in HBM. */
typedef struct rebx_b200_state {
    int64_t n;                 /* particles in the shard */
    int64_t index_base;        /* global particle index of element 0 */
    const double *x, *y, *z;
    const double *vx, *vy, *vz;   /* may be NULL when no effect in the pass reads velocities */
    const double *m;              /* may be NULL likewise */
    const double *r;              /* particle radius (yarkovsky only) */
    double *ax, *ay, *az;         /* read-modify-write: a += force */
} rebx_b200_state;

/* One stacked effect.  Per-kind meaning of tab[] / scal[]:
 *   RADIATION     tab0 beta                                   scal0 G, scal1 c
 *   HARMONICS     -                                           scal0 G
 *   GR_POTENTIAL  -                                           scal0 G, scal1 c*c
 *   YARKOVSKY     tab0 ye_body_density tab1 ye_rotation_period tab2 ye_thermal_inertia
 *                 tab3 ye_albedo tab4 ye_emissivity tab5 ye_k tab6..8 ye_spin_axis_x/y/z
 *                 itab: ye_flag code (-1,0,1 valid; 2 = skip)  scal0 G, scal1 ye_lstar,
 *                 scal2 ye_c, scal3 ye_stef_boltz
 *   MODIFY_ORBITS tab0 tau_a tab1 tau_e tab2 tau_inc            scal0 G, scal1 ide_position, scal2 ide_width
 *   GAS_DAMPING   tab0 d_factor                                 scal0 G, scal1 cs_coeff, scal2 tau_coeff
 *   TYPE_I        -                                             scal0 G, scal1 tIm_flaring_index, scal2 tIm_scale_height_1,
 *                                                               scal3 tIm_surface_density_1, scal4 tIm_surface_density_exponent,
 *                                                               scal5 ide_position, scal6 ide_width
 *   CENTRAL_FORCE -                                             (Acentral, gammacentral live in the body record)
 *   GAS_DF        -                                             scal0 G, scal1 gas_df_rhog, scal2 gas_df_alpha_rhog, scal3 gas_df_cs,
 *                                                               scal4 gas_df_alpha_cs, scal5 gas_df_xmin, scal6 gas_df_hr, scal7 gas_df_Qd
 *   LENSE_THIRRING -                                            scal0 G, scal1 lt_c*lt_c
 *   EXP_MIGRATION tab0 em_tau_a tab1 em_aini tab2 em_afin        scal0 G, scal1 sim->t
 * Tables are indexed like the state arrays (element k <-> global index index_base+k);
 * an unset double parameter holds REBX_B200_ABSENT_BITS. */
typedef struct rebx_b200_effect {
    int32_t kind;
    int32_t flags;
    const double* body;                       /* REBX_B200_BODY_DOUBLES doubles */
    const double* tab[REBX_B200_MAX_TABLES];
    const int32_t* itab;
    double scal[REBX_B200_MAX_SCALARS];
    double* backreaction;   /* 3 doubles: signed sum of the back-reaction terms this shard puts on
                               `body` (add it to the body's acceleration), or NULL to discard */
} rebx_b200_effect;

/* pass.reserved bit: every effect of the pass refers to the SAME central body (identical position,
 * velocity and mass in all body records).  Lets the fused sweep share the relative coordinates and the
 * orbit elements between stacked effects; results are unchanged. */
#define REBX_B200_PASS_SHARED_BODY 1
/* pass.reserved bit: TOLERANCE-MODE arithmetic for this pass (reboundx_b200/csrc/kernels_tol.cu): FMA contraction, one
 * reciprocal square root per particle shared by every power of the distance, Newton reciprocals instead of IEEE
 * divisions -- a few ulp per operation, ~1e-15 on a force vector, against the default build's bit identity with the
 * reference.  Honoured for stacks of up to three effects that all refer to one central body in PARTICLE
 * coordinates (REBX_B200_PASS_SHARED_BODY must be set as well); any other pass runs the exact kernels. */
#define REBX_B200_PASS_TOLERANCE 2

/* pass.reserved bit: the sweep itself adds each effect's back-reaction sum to the acceleration of the body it acts on
 * (when that body lies inside the shard), as rebx_b200_apply_backreaction would afterwards -- for a single-shard
 * caller, whose sums are final when the sweep ends.  A sharded caller leaves it clear, all-reduces the sums and
 * calls rebx_b200_apply_backreaction. */
#define REBX_B200_PASS_APPLY_BACKREACTION 4

typedef struct rebx_b200_pass {
    rebx_b200_state st;
    int32_t n_effects;
    int32_t reserved;                             /* REBX_B200_PASS_* bits */
    rebx_b200_effect fx[REBX_B200_MAX_EFFECTS];   /* applied in array order */
} rebx_b200_pass;

/* Layout of the host AoS particle record (taken from whichever rebound.h was compiled in). */
typedef struct rebx_b200_aos_layout {
    int32_t stride;                                   /* sizeof(struct reb_particle) */
    int32_t off_x, off_v, off_a, off_m, off_r;        /* byte offsets of x, vx, ax, m, r */
} rebx_b200_aos_layout;

/* ---- device management --------------------------------------------------------------- */
int  rebx_b200_device_count(void);                       /* 0 when no usable CUDA device */
const char* rebx_b200_error_string(int code);            /* text for a non-zero return code */
size_t rebx_b200_workspace_bytes(void);                  /* scratch a pass launch needs; ZERO it once after allocating it
                                                            (cudaMemset): it holds the ticket counter by which the last CTA
                                                            of a sweep knows to sum the back-reaction partials, which every
                                                            launch leaves at zero again.  One workspace per stream. */
int  rebx_b200_sm_count(int device);

/* ---- kernels (all asynchronous on `stream`, a cudaStream_t passed as void*) ------------ */
/* Streams the shard once and applies pass->fx[0..n_effects) in order.  `workspace` is
 * device scratch of rebx_b200_workspace_bytes().  Returns 0 or a cudaError_t value.
 * `launches` (host, may be NULL) is incremented by the number of kernels launched.  fx[].backreaction is
 * final when the launch completes (summed by the sweep's last CTA, no separate reduction launch). */
int rebx_b200_launch_pass(const rebx_b200_pass* pass, void* workspace, void* stream, int* launches);

/* a[index - index_base] += backreaction for each effect whose body lies inside the shard. */
int rebx_b200_apply_backreaction(const rebx_b200_pass* pass, void* stream, int* launches);

/* ---- centre-of-mass reference frames of rebx_com_force (reference src/rebxtools.c:66-147) ------
 * scratch: device memory of rebx_b200_scan_scratch_bytes(n), reusable between calls.
 *
 * Interior / total MASS (com.m of rebx_com_force) carries the reference's own sequential rounding, exactly:
 * the masses are scanned on the grid u = ulp of the leading body's mass, on which the reference's forward
 * fold (reb_simulation_com) and backward peel (rebx_get_com_without_particle, src/rebxtools.c:6-30) are
 * exact integer arithmetic (kernels_jacobi.cuh has the argument).  `lead_mass`: device pointer to the mass
 * of GLOBAL particle 0, or NULL for a state that starts at global index 0.  totals8[7] reports whether the
 * argument held for this shard (0) or not (1: running sum left the binade of the leading mass, an exact
 * tie, a negative mass, no positive leading mass).  With REBX_B200_COM_WHOLE in `flags` (the state is the
 * WHOLE ensemble) a launch with status 1 redoes the two recurrences literally -- one thread, O(n)
 * dependent additions -- so the masses are the reference's bits in every case; a sharded launch keeps the
 * grid sums (as accurate as the reference's own, no longer bit-identical to them). */
#define REBX_B200_COM_WHOLE 1
size_t rebx_b200_scan_scratch_bytes(int64_t n);
/* totals8 (device) = { sum m (on the grid), sum m x, sum m y, sum m z, sum m vx, sum m vy, sum m vz, status } of
 * the shard.  A sharded caller all-reduces totals8 before the next call. */
int rebx_b200_mass_moments(const rebx_b200_state* st, double* totals8, void* scratch, const double* lead_mass, int flags, void* stream, int* launches);
/* Writes `nbodies` consecutive body records (device) describing the barycentre (index -1). */
int rebx_b200_com_bodies(const double* totals7, double* bodies, int nbodies, void* stream, int* launches);
/* REBX_COORDINATES_JACOBI for 1..3 stacked damping effects (kinds 5,6,7) over the WHOLE ensemble
 * (index_base must be 0): prefix scan of the mass moments, sweep, suffix scan of the
 * back-reaction, all on the device in O(n).  fx[].body and fx[].backreaction are ignored.
 * ONE cooperative launch (reboundx_b200/csrc/jacobi_fused.cuh: a resident grid, contiguous tile ranges per warp, two
 * grid-wide barriers) where the device supports it, else -- or with REBX_B200_JACOBI=staged -- the three stages below. */
int rebx_b200_launch_jacobi(const rebx_b200_pass* pass, void* scratch, void* stream, int* launches);
/* The same pipeline in three stages, for an ensemble SHARDED by index range:
 * the shards exchange 7 + 3 doubles between the stages, everything else stays local.
 *   1. jacobi_moments: totals8 (device, may be NULL) = this shard's {sum m, sum m r, sum m v, status};
 *      the in-shard prefix stays in `scratch`.
 *      -> exchange: carry7 of a shard = sum of totals over the shards in FRONT of it (lower indices).
 *   2. jacobi_sweep: a += f against COM_i = (carry7 + in-shard prefix)/mass; tsum3 (device, may be NULL)
 *      = this shard's sum of t_i = m_i/(M_i+m_i) sum_e f_i^e.  carry7 NULL = the first shard.
 *      cudaErrorNotSupported for a stack without a fused instantiation (call per effect then).
 *      -> exchange: carry3 of a shard = sum of tsum3 over the shards BEHIND it (higher indices).
 *   3. jacobi_apply: a_j -= carry3 + sum_{i >= j in shard} t_i.  carry3 NULL = the last shard.
 * `scratch` (rebx_b200_scan_scratch_bytes(st.n)) must be the same untouched buffer through 1-3; `flags` the
 * same in 1 and 2. */
int rebx_b200_jacobi_moments(const rebx_b200_state* st, void* scratch, double* totals8, const double* lead_mass, int flags, void* stream, int* launches);
int rebx_b200_jacobi_sweep(const rebx_b200_pass* pass, void* scratch, const double* carry7, double* tsum3, int flags, void* stream, int* launches);
int rebx_b200_jacobi_apply(const rebx_b200_state* st, void* scratch, const double* carry3, void* stream, int* launches);

/* ---- 1PN `gr` effect (reference src/gr.c:65-164) for N_active <= REBX_B200_GR_MAX_ACTIVE ---------
 * Passed BY VALUE from host memory.  active (DEVICE pointer) = n_active x {x, y, z, m} of the active bodies;
 * com[] = the Jacobi origin of a test particle: position, velocity and Newtonian acceleration of the actives'
 * centre of mass; mu = G*m_0.  Particles with global index < n_active are skipped by both kernels (the
 * O(n_active^2) part is the caller's). */
#define REBX_B200_GR_MAX_ACTIVE 4096
typedef struct rebx_b200_gr_args {
    int32_t n_active;
    int32_t max_iterations;
    double G, C2, mu;
    double com[9];
    const double* active;
    const double* pull1;        /* n_active == 1 only, may be NULL: DEVICE pointer to the pull on the one active body (what
                                 * rebx_b200_gr_pull wrote); the sweep then takes the Jacobi origin's acceleration com[6..8] from
                                 * it -- (m_0 * (0 - pull_c)) * (1/m_0), the host part's operations -- and the caller needs no
                                 * device-to-host round trip between the reduction and the sweep */
} rebx_b200_gr_args;
/* pull[i*3+c] (device) = sum over the shard's test particles j of G m_j (x_i - x_j)_c / r^3. */
int rebx_b200_gr_pull(const rebx_b200_state* st, const rebx_b200_gr_args* args, double* pull, void* workspace, void* stream, int* launches);
/* a += 1PN acceleration for every test particle; *not_converged (device int) is set to 1 when a
 * fixed-point iteration ran 10 times without converging (the reference's warning, gr.c:130-133). */
int rebx_b200_gr_sweep(const rebx_b200_state* st, const rebx_b200_gr_args* args, int* not_converged, void* stream, int* launches);

/* ---- diagnostic potentials -------------------------------------------
 * *potential (device double) = this shard's part of the interaction potential of ONE effect entry
 * (`fx->kind` in GR_POTENTIAL, HARMONICS, CENTRAL_FORCE; fields as for the force sweep):
 *   GR_POTENTIAL   -sum_i 3 (G m_0)^2/c^2 m_i / r_i^2                 reference src/gr_potential.c:91-120
 *   HARMONICS      sum_j J2 and J4 terms of body `fx->body`            reference src/gravitational_harmonics.c:228-316
 *   CENTRAL_FORCE  -sum_i m_i A r^(gamma+1)/(gamma+1)  (log for -1)    reference src/central_force.c:98-121
 * A sharded caller all-reduces the result. */
int rebx_b200_potential(const rebx_b200_state* st, const rebx_b200_effect* fx, double* potential, void* workspace, void* stream, int* launches);

/* Barycentric second sweep: a[k] += delta[0..2] for every particle in the shard. */
int rebx_b200_add_uniform(const rebx_b200_state* st, const double* delta, void* stream, int* launches);

/* ---- resident stepping around the path ----------------------------
 * The callers either side of the force path, kept on the device so that a test-particle ensemble
 * never crosses PCIe between force evaluations.  (The state struct's const pointers are written.) */
/* x += dt v : reference rebx_drift_step, src/steppers.c:109-118 */
int rebx_b200_drift(const rebx_b200_state* st, double dt, void* stream, int* launches);
/* v += dt a : the update loop of reference rebx_kick_step, src/steppers.c:120-130 */
int rebx_b200_kick(const rebx_b200_state* st, double dt, void* stream, int* launches);
/* a = sum over `nbodies` body records (device) of -G m_b (r - r_b)/|r - r_b|^3 (overwrites a; the body
 * whose index equals the particle's is skipped): what REBOUND's gravity gives a test particle. */
int rebx_b200_point_mass_gravity(const rebx_b200_state* st, const double* bodies, int nbodies, double G, double softening, void* stream, int* launches);
/* Refreshes pos/vel/mass of every body record whose particle index lies inside the shard. */
int rebx_b200_gather_bodies(const rebx_b200_state* st, double* bodies, int nbodies, void* stream, int* launches);

/* Kepler drift: every particle of the shard follows its two-body orbit about `body` (one body record,
 * device; mu = G (m_body + m_i), m_i = 0 when st->m is NULL) for dt, the body's own element (if the shard
 * holds it) moves inertially; `body` itself is read, not written -- refresh it with rebx_b200_gather_bodies
 * afterwards.  The splitting of reference rebx_kepler_step (src/steppers.c:79-87 -> REBOUND's WHFast) for
 * one massive central body; arithmetic in reboundx_b200/csrc/kepler_core.h (universal variables). */
int rebx_b200_kepler_drift(const rebx_b200_state* st, const double* body, double G, double dt, void* stream, int* launches);
/* Fused Wisdom-Holman kick + drift for a pass of ONE effect without back-reaction (RADIATION, YARKOVSKY):
 * a = effect at the current state, v += dt_kick a, Kepler drift about fx[0].body for dt_drift, in one sweep
 * (104 B/particle for radiation instead of ~390 through zero_acc + launch_pass + kick + kepler_drift; the same
 * arithmetic per particle).  The body record is read, not written (refresh it with rebx_b200_gather_bodies).
 * cudaErrorNotSupported for any other pass: use the separate kernels.  With REBX_B200_PASS_TOLERANCE in pass->reserved
 * the sweep runs in tolerance arithmetic (kernels_tol.cu: FMA, Newton reciprocals, the force's reciprocal distance reused
 * by the drift): 0.39 -> 0.23 ms per step of 10^7 grains, trajectories within 1e-10 of the bit-exact sweep after 10^3 steps. */
int rebx_b200_wh_kick_drift(const rebx_b200_pass* pass, double G, double dt_kick, double dt_drift, void* stream, int* launches);
/* a = 0 for the shard (the kick of a Wisdom-Holman step sees the additional forces only). */
int rebx_b200_zero_acc(const rebx_b200_state* st, void* stream, int* launches);

/* One whole drift-kick-drift step (x += dt/2 v; a = pull of the stack's central body + stacked effects;
 * v += dt a; x += dt/2 v) in ONE sweep: accelerations stay in registers, 96 B/particle + tables.
 * All effects of the pass must refer to the same central body, whose record(s) and state the call
 * also advances (single shard: the back-reaction sums are used as they come out of this shard).
 * Same per-particle arithmetic as drift + gather_bodies + point_mass_gravity + launch_pass +
 * apply_backreaction + kick + drift.  Returns cudaErrorNotSupported for a stack without a fused
 * instantiation. */
int rebx_b200_leapfrog_step(const rebx_b200_pass* pass, double dt, double G, double softening, void* workspace, void* stream, int* launches);
/* The same step in two halves for an ensemble SHARDED by index range, every shard holding a replica of
 * the central body's record(s): `_sweep` advances the shard's particles and leaves this shard's
 * back-reaction sums in fx[].backreaction; the caller all-reduces those (nothing to exchange for a stack
 * without back-reaction, e.g. radiation_forces); `_body` then advances the central body -- its record(s),
 * and its state element on the shard that owns it -- identically on every rank.
 * rebx_b200_leapfrog_step == _sweep followed by _body. */
int rebx_b200_leapfrog_sweep(const rebx_b200_pass* pass, double dt, double G, double softening, void* workspace, void* stream, int* launches);
int rebx_b200_leapfrog_body(const rebx_b200_pass* pass, double dt, void* stream, int* launches);

/* Integrates the velocities across `dt` under the stacked effects of `pass` alone, positions fixed --
 * the device form of the reference's integrate_force operator and its schemes
 * (src/integrate_force.c:34-78; src/integrator_euler.c:33-41, _rk2.c:38-66, _rk4.c:40-90,
 * _implicit_midpoint.c:87-124).  `scheme` uses the reference's enum rebx_integrator values
 * (0 implicit midpoint, 1 RK4, 2 Euler, 3 RK2).  `scratch`: 12*n doubles of device memory.  Every
 * stage evaluates the sweep on scratch velocity/acceleration arrays, like the reference evaluates
 * update_accelerations on scratch particle copies.  *iterations (host, may be NULL) returns the
 * implicit-midpoint iteration count (10 = not converged, the reference warns).  Synchronises the
 * stream for the implicit-midpoint convergence test only. */
int rebx_b200_integrate_force(const rebx_b200_pass* pass, double dt, int scheme, void* scratch, void* workspace,
                              void* stream, int* launches, int* iterations);

/* AoS <-> SoA staging for the host-facing path (device-side transposes of one chunk). */
int rebx_b200_unpack_aos(const void* aos, const rebx_b200_aos_layout* layout, const rebx_b200_state* st_out_as_mutable,
                         void* stream, int* launches);
int rebx_b200_pack_acc_aos(void* aos, const rebx_b200_aos_layout* layout, const rebx_b200_state* st,
                           void* stream, int* launches);

/* ---- resident session: a simulation's particles kept in HBM between force evaluations, driven from C -------------
 * (the PCIe round trip of the host-facing call is the end-to-end bottleneck by ~100x.)
 * A session uploads sim->particles once, keeps SoA state + the parameter tables of every force on
 * rebx->additional_forces resident on the primary device, and advances / evaluates there; nothing crosses PCIe
 * until rebx_b200_session_sync_to_host.  Requirements: every added force is device-backed and refers to ONE central
 * body (particle 0) in PARTICLE coordinates, at most REBX_B200_MAX_EFFECTS effect entries; the steppers integrate
 * test-particle-like ensembles about particle 0 (N_active = 1 semantics, as the host harness reb_shim_integrate with
 * N_active = 1).  All calls return 0 or a cudaError_t value / -1 (a REBOUND error is queued on the simulation). */
struct reb_simulation;
typedef struct rebx_b200_session rebx_b200_session;
rebx_b200_session* rebx_b200_session_create(struct reb_simulation* sim, struct rebx_extras* rebx);
/* a += stacked effects at the resident state (what sim->additional_forces(sim) adds). */
int rebx_b200_session_evaluate(rebx_b200_session* s);
/* `nsteps` drift-kick-drift steps of dt: central gravity of particle 0 (sim->G, sim->softening) + the stacked effects;
 * one fused sweep per step where the stack has one (rebx_b200_leapfrog_step).  Same bits as reb_shim_integrate
 * scheme "leapfrog" with N_active = 1 calling this library's (or the reference's) force. */
int rebx_b200_session_step_leapfrog(rebx_b200_session* s, int nsteps, double dt);
/* `nsteps` Wisdom-Holman steps of dt: Kepler drift about particle 0 (universal variables), kick by the stacked effects
 * alone; the half drifts between consecutive kicks merged into one (scheme "wisdom_holman_merged" of the harness). */
int rebx_b200_session_step_wh(rebx_b200_session* s, int nsteps, double dt);
/* Velocities advanced across dt under the stacked effects alone, positions fixed: the integrate_force operator
 * (reference src/integrate_force.c:34-78); scheme = enum rebx_integrator value. */
int rebx_b200_session_integrate_force(rebx_b200_session* s, double dt, int scheme);
/* Writes positions, velocities and accelerations back into sim->particles and advances sim->t by the time stepped. */
int rebx_b200_session_sync_to_host(rebx_b200_session* s);
/* Kernel launches issued by the session so far. */
uint64_t rebx_b200_session_launches(const rebx_b200_session* s);
void rebx_b200_session_free(rebx_b200_session* s);

/* ---- the host dispatcher's knobs and counters (reboundx_b200/csrc/dispatch.cpp) ------------------------------------
 * Environment, read when an extras' device context is created (first force evaluation):
 *   REBX_B200_DEVICE=<ordinal>           primary device (default 0)
 *   REBX_B200_DEVICES=all|<count>|<list> devices a large PARTICLE-frame segment is split over by index range: one
 *                                        process, one host thread, per-device streams; parameter tables sharded the
 *                                        same way; back-reaction sums added on the host in index order
 *   REBX_B200_MULTI_MIN=<particles>      smallest ensemble worth splitting (default 2^20)
 *   REBX_B200_MATH=tolerance             REBX_B200_PASS_TOLERANCE for every pass
 *   REBX_B200_JACOBI=staged              rebx_b200_launch_jacobi keeps the staged pipeline (default: one cooperative launch)
 *   REBX_B200_D2H=acc|accm               device->host as a 2-D copy of the accelerations only (measured slower than whole records; off)
 *   REBX_B200_CHUNK, REBX_B200_ZEROCOPY_MAX, REBX_B200_PARANOID   pipeline chunk, zero-copy threshold, rebuild tables every call */
struct rebx_extras;
struct rebx_force;
/* Counters of the host-facing path: {calls, kernel launches, H2D bytes, D2H bytes, table builds}. */
int rebx_b200_get_stats(struct rebx_extras* rebx, uint64_t out[5]);
/* Builds (on the host, without touching a device) SoA table `index` of `force` into out[0..cap); index -1 = the int
 * table, -2 = the body index list.  Returns the number of elements, or -1. */
int64_t rebx_b200_host_table(struct rebx_extras* rebx, struct rebx_force* force, int index, double* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* REBX_B200_H */

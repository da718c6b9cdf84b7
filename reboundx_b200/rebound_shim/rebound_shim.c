/*
 * rebound_shim.c -- minimal stand-in for the REBOUND host library (see rebound.h here).
 *
 * Only what the REBOUNDx force path calls: simulation/particle storage, the error and
 * warning queue, reb_orbit_from_particle[_err], reb_simulation_com and the Jacobi
 * transforms.  The celestial-mechanics routines restate REBOUND's published algorithms
 * (REBOUND src/tools.c, src/transformations.c; module `rebound` >= 5.0.0, which is NOT
 * under /root/reference and not installable here).  Call sites in the reference:
 *   reb_orbit_from_particle      src/yarkovsky_effect.c:132, src/gas_damping_timescale.c:69
 *   reb_orbit_from_particle_err  src/modify_orbits_forces.c:102, src/type_I_migration.c:138
 *   reb_simulation_com           src/rebxtools.c:68
 *   reb_transformations_*        src/gr.c:101,155
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "rebound.h"
#include "../csrc/kepler_core.h"

#define SHIM_TINY 1.E-308

/* ------------------------------------------------------------------ life cycle ------- */

struct reb_simulation* reb_simulation_create(void){
    struct reb_simulation* r = calloc(1, sizeof(*r));
    if (!r) return NULL;
    r->G = 1.0;
    r->dt = 0.001;
    r->N_active = SIZE_MAX;
    r->integrator.name = "ias15";
    r->ias15.epsilon = 1e-9;
    r->integrator.state = &r->ias15;
    return r;
}

void reb_shim_clear_messages(struct reb_simulation* r){
    for (size_t i = 0; i < r->messages_N; i++) free(r->messages[i]);
    r->messages_N = 0;
}

static void shim_free_particles(struct reb_simulation* r){
    if (r->free_particle_ap){
        for (size_t i = 0; i < r->N; i++) r->free_particle_ap(&r->particles[i]);
    }
    free(r->particles);
    r->particles = NULL;
    r->N = 0;
    r->N_allocated = 0;
}

void reb_simulation_free(struct reb_simulation* r){
    if (!r) return;
    shim_free_particles(r);
    if (r->extras_cleanup) r->extras_cleanup(r);
    reb_shim_clear_messages(r);
    free(r->messages);
    for (size_t i = 0; i < r->names_N; i++) free(r->names[i]);
    free(r->names);
    free(r);
}

void reb_simulation_add(struct reb_simulation* r, struct reb_particle pt){
    if (r->N == r->N_allocated){
        size_t cap = r->N_allocated ? 2*r->N_allocated : 128;
        r->particles = realloc(r->particles, cap*sizeof(struct reb_particle));
        r->N_allocated = cap;
    }
    pt.sim = r;
    r->particles[r->N++] = pt;
}

void reb_simulation_remove_particle(struct reb_simulation* r, size_t index, int keep_sorted){
    if (index >= r->N) { reb_simulation_error(r, "Index out of range in reb_simulation_remove_particle."); return; }
    if (r->free_particle_ap) r->free_particle_ap(&r->particles[index]);
    if (keep_sorted){
        memmove(&r->particles[index], &r->particles[index+1], (r->N-1-index)*sizeof(struct reb_particle));
    } else {
        r->particles[index] = r->particles[r->N-1];
    }
    r->N--;
}

/* ------------------------------------------------------------------ messages --------- */

static void shim_push_message(struct reb_simulation* r, char kind, const char* msg){
    if (!r) { fprintf(stderr, "%s\n", msg); return; }
    if (r->messages_N >= 64) return;   /* bounded queue, like REBOUND's fixed-size message buffer */
    if (r->messages_N == r->messages_allocated){
        size_t cap = r->messages_allocated ? 2*r->messages_allocated : 8;
        r->messages = realloc(r->messages, cap*sizeof(char*));
        r->messages_allocated = cap;
    }
    size_t len = strlen(msg);
    char* s = malloc(len + 2);
    s[0] = kind;
    memcpy(s+1, msg, len+1);
    r->messages[r->messages_N++] = s;
}

void reb_simulation_error(struct reb_simulation* r, const char* msg){ shim_push_message(r, 'E', msg); }
void reb_simulation_warning(struct reb_simulation* r, const char* msg){ shim_push_message(r, 'W', msg); }
size_t reb_shim_message_count(const struct reb_simulation* r){ return r->messages_N; }
const char* reb_shim_message(const struct reb_simulation* r, size_t i){ return i < r->messages_N ? r->messages[i] : NULL; }

const char* reb_simulation_register_name(struct reb_simulation* r, const char* name){
    for (size_t i = 0; i < r->names_N; i++){
        if (strcmp(r->names[i], name) == 0) return r->names[i];
    }
    r->names = realloc(r->names, (r->names_N+1)*sizeof(char*));
    r->names[r->names_N] = strdup(name);
    return r->names[r->names_N++];
}

/* ------------------------------------------------------------------ orbit elements --- */

static double shim_acos2(double num, double denom, double disambiguator){
    double val;
    double cosine = num/denom;
    if (cosine > -1. && cosine < 1.){
        val = acos(cosine);
        if (disambiguator < 0.) val = -val;
    } else {
        val = (cosine <= -1.) ? M_PI : 0.;
    }
    return val;
}

static struct reb_orbit shim_orbit_nan(void){
    struct reb_orbit o;
    double* f = (double*)&o;
    for (size_t i = 0; i < sizeof(o)/sizeof(double); i++) f[i] = NAN;
    return o;
}

/* Restates REBOUND's reb_orbit_from_particle_err (heliocentric-style two-body elements
 * relative to `primary`, mu = G (m_p + m_primary)).  Only d, v, h, a, e, n, P, inc,
 * rhill, hvec, evec are filled; REBOUNDx's force path reads .a .e .inc .P only. */
struct reb_orbit reb_orbit_from_particle_err(double G, struct reb_particle p, struct reb_particle primary, int* err){
    struct reb_orbit o = shim_orbit_nan();
    if (primary.m <= SHIM_TINY){
        *err = 1;
        return o;
    }
    *err = 0;
    const double mu = G*(p.m + primary.m);
    const double dx = p.x - primary.x;
    const double dy = p.y - primary.y;
    const double dz = p.z - primary.z;
    const double dvx = p.vx - primary.vx;
    const double dvy = p.vy - primary.vy;
    const double dvz = p.vz - primary.vz;
    o.d = sqrt(dx*dx + dy*dy + dz*dz);

    const double vsquared = dvx*dvx + dvy*dvy + dvz*dvz;
    o.v = sqrt(vsquared);
    const double vcircsquared = mu/o.d;
    o.a = -mu/(vsquared - 2.*vcircsquared);
    o.rhill = o.a*cbrt(p.m/(3.*primary.m));

    const double hx = (dy*dvz - dz*dvy);
    const double hy = (dz*dvx - dx*dvz);
    const double hz = (dx*dvy - dy*dvx);
    o.h = sqrt(hx*hx + hy*hy + hz*hz);
    o.hvec.x = hx; o.hvec.y = hy; o.hvec.z = hz;

    const double vdiffsquared = vsquared - vcircsquared;
    if (o.d <= SHIM_TINY){
        *err = 2;
        return shim_orbit_nan();
    }
    const double vr = (dx*dvx + dy*dvy + dz*dvz)/o.d;
    const double rvr = o.d*vr;
    const double muinv = 1./mu;

    const double ex = muinv*(vdiffsquared*dx - rvr*dvx);
    const double ey = muinv*(vdiffsquared*dy - rvr*dvy);
    const double ez = muinv*(vdiffsquared*dz - rvr*dvz);
    o.evec.x = ex; o.evec.y = ey; o.evec.z = ez;
    o.e = sqrt(ex*ex + ey*ey + ez*ez);
    o.n = o.a/fabs(o.a)*sqrt(fabs(mu/(o.a*o.a*o.a)));
    o.P = 2*M_PI/o.n;
    o.inc = shim_acos2(hz, o.h, 1.);
    return o;
}

struct reb_orbit reb_orbit_from_particle(double G, struct reb_particle p, struct reb_particle primary){
    int err;
    return reb_orbit_from_particle_err(G, p, primary, &err);
}

/* ------------------------------------------------------------------ centre of mass --- */

struct reb_particle reb_particle_com_of_pair(struct reb_particle p1, struct reb_particle p2){
    p1.x  = p1.x*p1.m  + p2.x*p2.m;
    p1.y  = p1.y*p1.m  + p2.y*p2.m;
    p1.z  = p1.z*p1.m  + p2.z*p2.m;
    p1.vx = p1.vx*p1.m + p2.vx*p2.m;
    p1.vy = p1.vy*p1.m + p2.vy*p2.m;
    p1.vz = p1.vz*p1.m + p2.vz*p2.m;
    p1.ax = p1.ax*p1.m + p2.ax*p2.m;
    p1.ay = p1.ay*p1.m + p2.ay*p2.m;
    p1.az = p1.az*p1.m + p2.az*p2.m;
    p1.m += p2.m;
    if (p1.m > 0.){
        p1.x /= p1.m;  p1.y /= p1.m;  p1.z /= p1.m;
        p1.vx /= p1.m; p1.vy /= p1.m; p1.vz /= p1.m;
        p1.ax /= p1.m; p1.ay /= p1.m; p1.az /= p1.m;
    }
    return p1;
}

void reb_particle_isub(struct reb_particle* p1, struct reb_particle* p2){
    p1->x -= p2->x;   p1->y -= p2->y;   p1->z -= p2->z;
    p1->vx -= p2->vx; p1->vy -= p2->vy; p1->vz -= p2->vz;
    p1->ax -= p2->ax; p1->ay -= p2->ay; p1->az -= p2->az;
    p1->m -= p2->m;
}

struct reb_particle reb_simulation_com(struct reb_simulation* r){
    struct reb_particle com;
    memset(&com, 0, sizeof(com));
    const size_t N_real = r->N - r->N_var;
    for (size_t i = 0; i < N_real; i++) com = reb_particle_com_of_pair(com, r->particles[i]);
    return com;
}

/* ------------------------------------------------------------------ Jacobi transforms - */

static void shim_to_jacobi(const struct reb_particle* const particles, struct reb_particle* const p_j,
                           const struct reb_particle* const p_mass, const size_t N, const size_t N_active, int with_acc){
    double eta = p_mass[0].m;
    double s_x = eta*particles[0].x,  s_y = eta*particles[0].y,  s_z = eta*particles[0].z;
    double s_vx = eta*particles[0].vx, s_vy = eta*particles[0].vy, s_vz = eta*particles[0].vz;
    double s_ax = eta*particles[0].ax, s_ay = eta*particles[0].ay, s_az = eta*particles[0].az;
    for (size_t i = 1; i < N_active; i++){
        const double ei = 1./eta;
        const struct reb_particle pi = particles[i];
        eta += p_mass[i].m;
        const double pme = eta*ei;
        p_j[i].m = pi.m;
        p_j[i].x = pi.x - s_x*ei;   p_j[i].y = pi.y - s_y*ei;   p_j[i].z = pi.z - s_z*ei;
        p_j[i].vx = pi.vx - s_vx*ei; p_j[i].vy = pi.vy - s_vy*ei; p_j[i].vz = pi.vz - s_vz*ei;
        s_x = s_x*pme + p_mass[i].m*p_j[i].x;
        s_y = s_y*pme + p_mass[i].m*p_j[i].y;
        s_z = s_z*pme + p_mass[i].m*p_j[i].z;
        s_vx = s_vx*pme + p_mass[i].m*p_j[i].vx;
        s_vy = s_vy*pme + p_mass[i].m*p_j[i].vy;
        s_vz = s_vz*pme + p_mass[i].m*p_j[i].vz;
        if (with_acc){
            p_j[i].ax = pi.ax - s_ax*ei; p_j[i].ay = pi.ay - s_ay*ei; p_j[i].az = pi.az - s_az*ei;
            s_ax = s_ax*pme + p_mass[i].m*p_j[i].ax;
            s_ay = s_ay*pme + p_mass[i].m*p_j[i].ay;
            s_az = s_az*pme + p_mass[i].m*p_j[i].az;
        }
    }
    const double Mtotal = eta;
    const double Mtotali = 1./Mtotal;
    for (size_t i = N_active; i < N; i++){      /* test particles: relative to the COM of all active bodies */
        const struct reb_particle pi = particles[i];
        p_j[i].m = pi.m;
        p_j[i].x = pi.x - s_x*Mtotali;   p_j[i].y = pi.y - s_y*Mtotali;   p_j[i].z = pi.z - s_z*Mtotali;
        p_j[i].vx = pi.vx - s_vx*Mtotali; p_j[i].vy = pi.vy - s_vy*Mtotali; p_j[i].vz = pi.vz - s_vz*Mtotali;
        if (with_acc){
            p_j[i].ax = pi.ax - s_ax*Mtotali; p_j[i].ay = pi.ay - s_ay*Mtotali; p_j[i].az = pi.az - s_az*Mtotali;
        }
    }
    p_j[0].m = Mtotal;
    p_j[0].x = s_x*Mtotali;   p_j[0].y = s_y*Mtotali;   p_j[0].z = s_z*Mtotali;
    p_j[0].vx = s_vx*Mtotali; p_j[0].vy = s_vy*Mtotali; p_j[0].vz = s_vz*Mtotali;
    if (with_acc){
        p_j[0].ax = s_ax*Mtotali; p_j[0].ay = s_ay*Mtotali; p_j[0].az = s_az*Mtotali;
    }
}

void reb_transformations_inertial_to_jacobi_posvel(const struct reb_particle* const particles, struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active){
    shim_to_jacobi(particles, p_j, p_mass, N, N_active, 0);
}

void reb_transformations_inertial_to_jacobi_posvelacc(const struct reb_particle* const particles, struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active){
    shim_to_jacobi(particles, p_j, p_mass, N, N_active, 1);
}

void reb_transformations_jacobi_to_inertial_acc(struct reb_particle* const particles, const struct reb_particle* const p_j, const struct reb_particle* const p_mass, const size_t N, const size_t N_active){
    double eta = p_j[0].m;
    double s_ax = p_j[0].ax*eta, s_ay = p_j[0].ay*eta, s_az = p_j[0].az*eta;
    {
        const double ei = 1./eta;
        for (size_t i = N_active; i < N; i++){
            particles[i].ax = p_j[i].ax + s_ax*ei;
            particles[i].ay = p_j[i].ay + s_ay*ei;
            particles[i].az = p_j[i].az + s_az*ei;
        }
    }
    for (size_t i = N_active - 1; i > 0; i--){
        const struct reb_particle pji = p_j[i];
        const double ei = 1./eta;
        s_ax = (s_ax - p_mass[i].m*pji.ax)*ei;
        s_ay = (s_ay - p_mass[i].m*pji.ay)*ei;
        s_az = (s_az - p_mass[i].m*pji.az)*ei;
        particles[i].ax = pji.ax + s_ax;
        particles[i].ay = pji.ay + s_ay;
        particles[i].az = pji.az + s_az;
        eta -= p_mass[i].m;
        s_ax *= eta; s_ay *= eta; s_az *= eta;
    }
    const double mi = 1./eta;
    particles[0].ax = s_ax*mi;
    particles[0].ay = s_ay*mi;
    particles[0].az = s_az*mi;
}

/* ------------------------------------------------------------------ rotations (unused on the path) */

void reb_vec3d_irotate(struct reb_vec3d* v, const struct reb_rotation q){
    /* v' = q v q^-1 for a unit quaternion (ix,iy,iz,r) */
    const double tx = 2.*(q.iy*v->z - q.iz*v->y);
    const double ty = 2.*(q.iz*v->x - q.ix*v->z);
    const double tz = 2.*(q.ix*v->y - q.iy*v->x);
    const double nx = v->x + q.r*tx + (q.iy*tz - q.iz*ty);
    const double ny = v->y + q.r*ty + (q.iz*tx - q.ix*tz);
    const double nz = v->z + q.r*tz + (q.ix*ty - q.iy*tx);
    v->x = nx; v->y = ny; v->z = nz;
}

void reb_simulation_irotate(struct reb_simulation* r, const struct reb_rotation q){
    for (size_t i = 0; i < r->N; i++){
        struct reb_vec3d pos = {r->particles[i].x, r->particles[i].y, r->particles[i].z};
        struct reb_vec3d vel = {r->particles[i].vx, r->particles[i].vy, r->particles[i].vz};
        reb_vec3d_irotate(&pos, q);
        reb_vec3d_irotate(&vel, q);
        r->particles[i].x = pos.x; r->particles[i].y = pos.y; r->particles[i].z = pos.z;
        r->particles[i].vx = vel.x; r->particles[i].vy = vel.y; r->particles[i].vz = vel.z;
    }
}

/* ------------------------------------------------------------------ shim-only helpers - */

void reb_shim_set_particles(struct reb_simulation* r, size_t N,
                            const double* x, const double* y, const double* z,
                            const double* vx, const double* vy, const double* vz,
                            const double* m, const double* rad){
    shim_free_particles(r);
    r->particles = calloc(N ? N : 1, sizeof(struct reb_particle));
    r->N_allocated = N ? N : 1;
    r->N = N;
    for (size_t i = 0; i < N; i++){
        struct reb_particle* p = &r->particles[i];
        p->x = x ? x[i] : 0.;   p->y = y ? y[i] : 0.;   p->z = z ? z[i] : 0.;
        p->vx = vx ? vx[i] : 0.; p->vy = vy ? vy[i] : 0.; p->vz = vz ? vz[i] : 0.;
        p->m = m ? m[i] : 0.;
        p->r = rad ? rad[i] : 0.;
        p->sim = r;
    }
}

void reb_shim_set_acc(struct reb_simulation* r, const double* ax, const double* ay, const double* az){
    for (size_t i = 0; i < r->N; i++){
        r->particles[i].ax = ax ? ax[i] : 0.;
        r->particles[i].ay = ay ? ay[i] : 0.;
        r->particles[i].az = az ? az[i] : 0.;
    }
}

void reb_shim_get_acc(const struct reb_simulation* r, double* ax, double* ay, double* az){
    for (size_t i = 0; i < r->N; i++){
        ax[i] = r->particles[i].ax;
        ay[i] = r->particles[i].ay;
        az[i] = r->particles[i].az;
    }
}

void reb_shim_get_posvel(const struct reb_simulation* r, double* x, double* y, double* z, double* vx, double* vy, double* vz){
    for (size_t i = 0; i < r->N; i++){
        const struct reb_particle* p = &r->particles[i];
        x[i] = p->x; y[i] = p->y; z[i] = p->z;
        vx[i] = p->vx; vy[i] = p->vy; vz[i] = p->vz;
    }
}

void reb_shim_set_integrator(struct reb_simulation* r, const char* name){
    r->integrator.name = reb_simulation_register_name(r, name);
}

void reb_shim_bulk_set_double(struct reb_simulation* r, reb_shim_set_double_fn setter, void* rebx,
                              const char* name, const int64_t* idx, const double* val, size_t n){
    for (size_t k = 0; k < n; k++){
        const size_t i = idx ? (size_t)idx[k] : k;
        setter(rebx, &r->particles[i].ap, name, val[k]);
    }
}

void reb_shim_bulk_set_int(struct reb_simulation* r, reb_shim_set_int_fn setter, void* rebx,
                           const char* name, const int64_t* idx, const int* val, size_t n){
    for (size_t k = 0; k < n; k++){
        const size_t i = idx ? (size_t)idx[k] : k;
        setter(rebx, &r->particles[i].ap, name, val[k]);
    }
}

void** reb_shim_particle_ap(struct reb_simulation* r, size_t i){
    return &r->particles[i].ap;
}

int reb_shim_call_additional_forces(struct reb_simulation* r){
    if (!r->additional_forces) return 0;
    r->additional_forces(r);
    return 1;
}

double reb_shim_time_additional_forces(struct reb_simulation* r, int reps){
    double best = INFINITY;
    if (!r->additional_forces) return best;
    for (int k = 0; k < reps; k++){
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        r->additional_forces(r);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9*(double)(t1.tv_nsec - t0.tv_nsec);
        if (dt < best) best = dt;
    }
    return best;
}

/* ------------------------------------------------------------------ integrator harness */

static void shim_accelerations(struct reb_simulation* r){
    const size_t N = r->N;
    const size_t Nact = (r->N_active == SIZE_MAX || r->N_active > N) ? N : r->N_active;
    struct reb_particle* p = r->particles;
    const double soft2 = r->softening*r->softening;
    for (size_t i = 0; i < N; i++){ p[i].ax = 0.; p[i].ay = 0.; p[i].az = 0.; }
    for (size_t j = 0; j < Nact; j++){
        const double mj = p[j].m;
        if (mj == 0.) continue;
        for (size_t i = 0; i < N; i++){
            if (i == j) continue;
            const double dx = p[i].x - p[j].x, dy = p[i].y - p[j].y, dz = p[i].z - p[j].z;
            const double r2 = dx*dx + dy*dy + dz*dz + soft2;
            const double pref = -r->G*mj/(r2*sqrt(r2));
            p[i].ax += pref*dx; p[i].ay += pref*dy; p[i].az += pref*dz;
        }
    }
    if (r->additional_forces) r->additional_forces(r);
}

/* REBOUND's public name for "gravity + additional forces" (called by rebx_kick_step, reference src/steppers.c:121). */
void reb_simulation_update_acceleration(struct reb_simulation* r){ shim_accelerations(r); }

/* ---- scheme 2: 15th-order Gauss-Radau predictor-corrector with a FIXED step (the scheme behind
 * REBOUND's IAS15, Everhart 1985 / Rein & Spiegel 2015, without its step-size control, i.e.
 * epsilon = 0).  Written from the published algorithm: the acceleration along the step is the
 * polynomial a(s) = a0 + b0 s + ... + b6 s^7 (s in [0,1]) through the 8 Gauss-Radau nodes h_n; its
 * Newton form a0 + g0 s + g1 s(s-h1) + ... is updated node by node from divided differences, the b's
 * follow from the g's through the coefficients of prod_{i<=j}(s - h_i), and the sweep over the nodes
 * is repeated until b6 stops changing.  The coefficient tables are generated from the nodes here. */
static const double gr_h[8] = {0.0, 0.0562625605369221464656521910318, 0.180240691736892364987579942780,
                               0.352624717113169637373907769648, 0.547153626330555383001448554766,
                               0.734210177215410531523210605558, 0.885320946839095768090359771030,
                               0.977520613561287501891174488626};
static double gr_rr[28];        /* h_j - h_k, j = 1..7, k = 0..j-1 */
static double gr_c[7][7];       /* gr_c[j][k] = coefficient of s^k in prod_{i=1..j}(s - h_i) */
static int gr_ready = 0;
static void gr_setup(void){
    if (gr_ready) return;
    int idx = 0;
    for (int j = 1; j < 8; j++) for (int k = 0; k < j; k++) gr_rr[idx++] = gr_h[j] - gr_h[k];
    for (int j = 0; j < 7; j++) for (int k = 0; k < 7; k++) gr_c[j][k] = 0.;
    gr_c[0][0] = 1.;
    for (int j = 1; j < 7; j++)
        for (int k = 0; k <= j; k++)
            gr_c[j][k] = (k > 0 ? gr_c[j-1][k-1] : 0.) - gr_h[j]*(k < j ? gr_c[j-1][k] : 0.);
    gr_ready = 1;
}

static void shim_gauss_radau(struct reb_simulation* r, int nsteps){
    gr_setup();
    const size_t N = r->N, M = 3*N;
    struct reb_particle* p = r->particles;
    const double dt = r->dt;
    double* x0 = malloc(M*sizeof(double)); double* v0 = malloc(M*sizeof(double)); double* a0 = malloc(M*sizeof(double));
    double* b = calloc(7*M, sizeof(double)); double* g = calloc(7*M, sizeof(double));
    for (int s = 0; s < nsteps; s++){
        const double t0 = r->t;
        shim_accelerations(r);
        for (size_t i = 0; i < N; i++){
            x0[3*i] = p[i].x; x0[3*i+1] = p[i].y; x0[3*i+2] = p[i].z;
            v0[3*i] = p[i].vx; v0[3*i+1] = p[i].vy; v0[3*i+2] = p[i].vz;
            a0[3*i] = p[i].ax; a0[3*i+1] = p[i].ay; a0[3*i+2] = p[i].az;
        }
        /* g (and so b) of the previous step are the first guess of this one */
        for (int it = 0; it < 20; it++){
            double max_db6 = 0., max_a = 0.;
            for (int n = 1; n < 8; n++){
                const double h = gr_h[n];
                for (size_t q = 0; q < M; q++){
                    const double* bq = b + q;
                    const double xs = x0[q] + dt*h*(v0[q] + dt*h*(a0[q]/2. + h*(bq[0]/6. + h*(bq[M]/12. + h*(bq[2*M]/20. + h*(bq[3*M]/30.
                                      + h*(bq[4*M]/42. + h*(bq[5*M]/56. + h*bq[6*M]/72.))))))));
                    const double vs = v0[q] + dt*h*(a0[q] + h*(bq[0]/2. + h*(bq[M]/3. + h*(bq[2*M]/4. + h*(bq[3*M]/5.
                                      + h*(bq[4*M]/6. + h*(bq[5*M]/7. + h*bq[6*M]/8.)))))));
                    struct reb_particle* pi = &p[q/3];
                    switch (q % 3){ case 0: pi->x = xs; pi->vx = vs; break; case 1: pi->y = xs; pi->vy = vs; break; default: pi->z = xs; pi->vz = vs; }
                }
                r->t = t0 + h*dt;
                shim_accelerations(r);
                const double* rr = gr_rr + (n-1)*n/2;
                for (size_t q = 0; q < M; q++){
                    const struct reb_particle* pi = &p[q/3];
                    const double a = (q % 3 == 0) ? pi->ax : ((q % 3 == 1) ? pi->ay : pi->az);
                    double gk = (a - a0[q])/rr[0];
                    for (int m = 0; m < n-1; m++) gk = (gk - g[(size_t)m*M + q])/rr[m+1];
                    const double dg = gk - g[(size_t)(n-1)*M + q];
                    g[(size_t)(n-1)*M + q] = gk;
                    for (int k = 0; k < n; k++) b[(size_t)k*M + q] += dg*gr_c[n-1][k];
                    if (n == 7){
                        if (fabs(dg) > max_db6) max_db6 = fabs(dg);
                        if (fabs(a) > max_a) max_a = fabs(a);
                    }
                }
            }
            if (max_a == 0. || max_db6/max_a < 1e-16) break;
        }
        for (size_t q = 0; q < M; q++){
            const double* bq = b + q;
            const double xs = x0[q] + dt*v0[q] + dt*dt*(a0[q]/2. + bq[0]/6. + bq[M]/12. + bq[2*M]/20. + bq[3*M]/30. + bq[4*M]/42. + bq[5*M]/56. + bq[6*M]/72.);
            const double vs = v0[q] + dt*(a0[q] + bq[0]/2. + bq[M]/3. + bq[2*M]/4. + bq[3*M]/5. + bq[4*M]/6. + bq[5*M]/7. + bq[6*M]/8.);
            struct reb_particle* pi = &p[q/3];
            switch (q % 3){ case 0: pi->x = xs; pi->vx = vs; break; case 1: pi->y = xs; pi->vy = vs; break; default: pi->z = xs; pi->vz = vs; }
        }
        r->t = t0 + dt;
        r->dt_last_done = dt;
    }
    free(x0); free(v0); free(a0); free(b); free(g);
}

/* ---- scheme 3: Wisdom-Holman drift-kick-drift for ONE massive central body (particle 0) and any
 * number of light bodies -- the splitting behind REBOUND's WHFast, in heliocentric form: every body
 * i >= 1 follows its Kepler orbit about particle 0 (mu = G (m_0 + m_i)) for dt/2, receives the kick
 * dt * (a_i - a_0) of everything that is NOT that two-body attraction -- here only
 * r->additional_forces -- and drifts another dt/2.  Particle 0 moves inertially and takes its own
 * kick.  The Kepler drift (universal variables, Stumpff functions, Newton iteration) lives in
 * ../csrc/kepler_core.h, shared verbatim with the device kernel so that both execute the same IEEE
 * operations in the same order. */
static void shim_wh_half_drift(struct reb_simulation* r, double hdt){
    struct reb_particle* p = r->particles;
    const size_t N = r->N;
    const double x0[3] = {p[0].x, p[0].y, p[0].z}, v0[3] = {p[0].vx, p[0].vy, p[0].vz};
    const double x1[3] = {x0[0] + hdt*v0[0], x0[1] + hdt*v0[1], x0[2] + hdt*v0[2]};
    for (size_t i = 1; i < N; i++){
        double x[3] = {p[i].x - x0[0], p[i].y - x0[1], p[i].z - x0[2]};
        double v[3] = {p[i].vx - v0[0], p[i].vy - v0[1], p[i].vz - v0[2]};
        rebxb_kepler_drift(r->G*(p[0].m + p[i].m), hdt, x, v);
        p[i].x = x1[0] + x[0]; p[i].y = x1[1] + x[1]; p[i].z = x1[2] + x[2];
        p[i].vx = v0[0] + v[0]; p[i].vy = v0[1] + v[1]; p[i].vz = v0[2] + v[2];
    }
    p[0].x = x1[0]; p[0].y = x1[1]; p[0].z = x1[2];
}

static void shim_wh_kick(struct reb_simulation* r, double dt){
    struct reb_particle* p = r->particles;
    const size_t N = r->N;
    for (size_t i = 0; i < N; i++){ p[i].ax = 0.; p[i].ay = 0.; p[i].az = 0.; }
    if (r->additional_forces) r->additional_forces(r);
    for (size_t i = 0; i < N; i++){ p[i].vx += dt*p[i].ax; p[i].vy += dt*p[i].ay; p[i].vz += dt*p[i].az; }
}

/* merged = 0: every step is drift(dt/2) kick(dt) drift(dt/2), the state is synchronised after each step.
 * merged = 1: the two half drifts between consecutive kicks are taken as ONE drift of dt (what WHFast does
 * between outputs): drift(dt/2) [kick(dt) drift(dt)]^(n-1) kick(dt) drift(dt/2). */
static void shim_wisdom_holman(struct reb_simulation* r, int nsteps, int merged){
    const double dt = r->dt;
    if (nsteps <= 0) return;
    if (!merged){
        for (int s = 0; s < nsteps; s++){
            shim_wh_half_drift(r, 0.5*dt);
            r->t += 0.5*dt;
            shim_wh_kick(r, dt);
            shim_wh_half_drift(r, 0.5*dt);
            r->t += 0.5*dt;
            r->dt_last_done = dt;
        }
        return;
    }
    shim_wh_half_drift(r, 0.5*dt);
    r->t += 0.5*dt;
    for (int s = 0; s < nsteps; s++){
        shim_wh_kick(r, dt);
        if (s + 1 < nsteps){ shim_wh_half_drift(r, dt); r->t += dt; }
    }
    shim_wh_half_drift(r, 0.5*dt);
    r->t += 0.5*dt;
    r->dt_last_done = dt;
}

void reb_shim_integrate(struct reb_simulation* r, int nsteps, int scheme){
    if (scheme == 2){ shim_gauss_radau(r, nsteps); return; }
    if (scheme == 3){ shim_wisdom_holman(r, nsteps, 0); return; }
    if (scheme == 4){ shim_wisdom_holman(r, nsteps, 1); return; }
    const size_t N = r->N;
    struct reb_particle* p = r->particles;
    const double dt = r->dt;
    if (scheme == 0){
        for (int s = 0; s < nsteps; s++){
            if (r->pre_timestep_modifications) r->pre_timestep_modifications(r);     /* REBOUNDx operators, as REBOUND's step does */
            for (size_t i = 0; i < N; i++){ p[i].x += 0.5*dt*p[i].vx; p[i].y += 0.5*dt*p[i].vy; p[i].z += 0.5*dt*p[i].vz; }
            r->t += 0.5*dt;
            shim_accelerations(r);
            for (size_t i = 0; i < N; i++){ p[i].vx += dt*p[i].ax; p[i].vy += dt*p[i].ay; p[i].vz += dt*p[i].az; }
            for (size_t i = 0; i < N; i++){ p[i].x += 0.5*dt*p[i].vx; p[i].y += 0.5*dt*p[i].vy; p[i].z += 0.5*dt*p[i].vz; }
            r->t += 0.5*dt;
            r->dt_last_done = dt;
            if (r->post_timestep_modifications) r->post_timestep_modifications(r);
        }
        return;
    }
    /* RK4 on (x, v) */
    double* y0 = malloc(6*N*sizeof(double));
    double* k = malloc(4*6*N*sizeof(double));
    for (int s = 0; s < nsteps; s++){
        const double t0 = r->t;
        for (size_t i = 0; i < N; i++){
            y0[6*i+0] = p[i].x; y0[6*i+1] = p[i].y; y0[6*i+2] = p[i].z;
            y0[6*i+3] = p[i].vx; y0[6*i+4] = p[i].vy; y0[6*i+5] = p[i].vz;
        }
        const double c[4] = {0., 0.5, 0.5, 1.};
        for (int st = 0; st < 4; st++){
            if (st > 0){
                const double* kp = k + (size_t)(st-1)*6*N;
                for (size_t i = 0; i < N; i++){
                    p[i].x = y0[6*i+0] + c[st]*dt*kp[6*i+0]; p[i].y = y0[6*i+1] + c[st]*dt*kp[6*i+1]; p[i].z = y0[6*i+2] + c[st]*dt*kp[6*i+2];
                    p[i].vx = y0[6*i+3] + c[st]*dt*kp[6*i+3]; p[i].vy = y0[6*i+4] + c[st]*dt*kp[6*i+4]; p[i].vz = y0[6*i+5] + c[st]*dt*kp[6*i+5];
                }
            }
            r->t = t0 + c[st]*dt;
            shim_accelerations(r);
            double* kc = k + (size_t)st*6*N;
            for (size_t i = 0; i < N; i++){
                kc[6*i+0] = p[i].vx; kc[6*i+1] = p[i].vy; kc[6*i+2] = p[i].vz;
                kc[6*i+3] = p[i].ax; kc[6*i+4] = p[i].ay; kc[6*i+5] = p[i].az;
            }
        }
        for (size_t i = 0; i < N; i++){
            double y[6];
            for (int q = 0; q < 6; q++)
                y[q] = y0[6*i+q] + dt/6.*(k[6*i+q] + 2.*k[6*N+6*i+q] + 2.*k[12*N+6*i+q] + k[18*N+6*i+q]);
            p[i].x = y[0]; p[i].y = y[1]; p[i].z = y[2]; p[i].vx = y[3]; p[i].vy = y[4]; p[i].vz = y[5];
        }
        r->t = t0 + dt;
        r->dt_last_done = dt;
    }
    free(y0); free(k);
}

/* C user program for the resident session API (include/rebx_b200.h; GPU test, compiled and run by
 * tests/test_gpu_parity.py): BASELINE config 1a -- the reference's examples/rad_forces_debris_disk set-up, Sun + 1000 dust
 * grains with radiation_forces + Poynting-Robertson drag -- is integrated for 10^3 steps
 *   (a) by the host harness (reb_shim_integrate), which crosses PCIe twice per step through sim->additional_forces, and
 *   (b) by a resident session (rebx_b200_session_step_wh / _step_leapfrog): one upload, 10^3 steps in HBM, one download,
 * and the two trajectories are compared bit for bit.  Also: the session's force evaluation against the dispatcher's, and
 * the reference's operator entry (rebx_load_operator("integrate_force") + rebx_add_operator) against the session's
 * integrate_force.  No Python, no torch. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "rebound.h"
#include "reboundx.h"
#include "rebx_b200.h"

static double now(void){ struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9*(double)t.tv_nsec; }
static unsigned long long rs = 3;
static double uniform(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (double)(rs >> 11)/9007199254740992.0; }

static struct reb_simulation* make_sim(size_t n, struct rebx_extras** rebx_out, struct rebx_force** force_out){
    struct reb_simulation* sim = reb_simulation_create();
    const double AU = 1.5e11;
    sim->G = 6.674e-11; sim->dt = 1.e8; sim->N_active = 1;
    reb_shim_set_integrator(sim, "leapfrog");
    struct reb_particle sun = {0};
    sun.m = 1.99e30;
    reb_simulation_add(sim, sun);
    rs = 3;
    for (size_t i = 0; i < n; i++){
        const double a = AU*(100. + 20.*uniform()), e = 0.1, th = 6.283185307179586*uniform(), inc = 0.1*uniform();
        const double r = a*(1. - e*e)/(1. + e*cos(th)), v = sqrt(sim->G*sun.m*(2./r - 1./a));
        struct reb_particle p = {0};
        p.x = r*cos(th); p.y = r*sin(th)*cos(inc); p.z = r*sin(th)*sin(inc);
        p.vx = -v*sin(th); p.vy = v*cos(th)*cos(inc); p.vz = v*cos(th)*sin(inc);
        reb_simulation_add(sim, p);
    }
    struct rebx_extras* rebx = rebx_attach(sim);
    struct rebx_force* rad = rebx_load_force(rebx, "radiation_forces");
    rebx_set_param_double(rebx, &rad->ap, "c", 3.e8);
    for (size_t i = 1; i <= n; i++) rebx_set_param_double(rebx, (struct rebx_node**)&sim->particles[i].ap, "beta", 0.1);
    rebx_add_force(rebx, rad);
    *rebx_out = rebx;
    if (force_out) *force_out = rad;
    return sim;
}

static size_t differing(const struct reb_simulation* a, const struct reb_simulation* b, int with_acc){
    size_t d = 0;
    for (size_t i = 0; i < a->N; i++){
        const struct reb_particle *p = &a->particles[i], *q = &b->particles[i];
        d += (p->x != q->x) + (p->y != q->y) + (p->z != q->z) + (p->vx != q->vx) + (p->vy != q->vy) + (p->vz != q->vz);
        if (with_acc) d += (p->ax != q->ax) + (p->ay != q->ay) + (p->az != q->az);
    }
    return d;
}

int main(int argc, char** argv){
    const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : 1000;
    const int nsteps = argc > 2 ? atoi(argv[2]) : 1000;
    if (rebx_b200_device_count() < 1){ printf("{\"skipped\": \"no CUDA device\"}\n"); return 0; }
    int rc = 0;
    printf("{\"particles\": %zu, \"steps\": %d", n, nsteps);
    const char* names[2] = {"wisdom_holman_merged", "leapfrog"};
    for (int which = 0; which < 2; which++){
        struct rebx_extras *ra, *rb;
        struct reb_simulation* A = make_sim(n, &ra, NULL);
        struct reb_simulation* B = make_sim(n, &rb, NULL);
        double t0 = now();
        reb_shim_integrate(A, nsteps, which == 0 ? 4 : 0);
        const double t_host = now() - t0;
        rebx_b200_session* s = rebx_b200_session_create(B, rb);
        if (!s){ fprintf(stderr, "no session: %s\n", reb_shim_message_count(B) ? reb_shim_message(B, 0) : "?"); return 3; }
        t0 = now();
        int e = which == 0 ? rebx_b200_session_step_wh(s, nsteps, B->dt) : rebx_b200_session_step_leapfrog(s, nsteps, B->dt);
        e |= rebx_b200_session_sync_to_host(s);
        const double t_res = now() - t0;
        if (e){ fprintf(stderr, "session error %d: %s\n", e, reb_shim_message_count(B) ? reb_shim_message(B, 0) : "?"); return 4; }
        const size_t d = differing(A, B, 0);
        if (d != 0 || fabs(A->t - B->t) > 1e-6*fabs(A->t)) rc = 1;
        printf(", \"%s\": {\"differing_coordinates\": %zu, \"host_harness_ms_per_step\": %.4f, \"resident_ms_per_step\": %.4f, \"launches\": %llu}",
               names[which], d, 1e3*t_host/nsteps, 1e3*t_res/nsteps, (unsigned long long)rebx_b200_session_launches(s));
        rebx_b200_session_free(s);
        rebx_free(ra); reb_simulation_free(A);
        rebx_free(rb); reb_simulation_free(B);
    }
    {   /* force evaluation: session vs dispatcher, bit for bit */
        struct rebx_extras *ra, *rb;
        struct reb_simulation* A = make_sim(n, &ra, NULL);
        struct reb_simulation* B = make_sim(n, &rb, NULL);
        A->additional_forces(A);
        rebx_b200_session* s = rebx_b200_session_create(B, rb);
        int e = rebx_b200_session_evaluate(s) | rebx_b200_session_sync_to_host(s);
        const size_t d = differing(A, B, 1);
        if (e || d) rc = 1;
        printf(", \"evaluate\": {\"differing\": %zu}", d);
        rebx_b200_session_free(s);
        rebx_free(ra); reb_simulation_free(A);
        rebx_free(rb); reb_simulation_free(B);
    }
    {   /* the reference's operator entry: integrate_force (RK4) added with rebx_add_operator under a leapfrog host step
         * (half a step before, half after), against the same scheme applied by hand through a session */
        struct rebx_extras *ra, *rb;
        struct rebx_force *fa, *fb;
        struct reb_simulation* A = make_sim(n, &ra, &fa);
        struct reb_simulation* B = make_sim(n, &rb, &fb);
        rebx_remove_force(ra, fa);                       /* the force acts through the operator only */
        fa = rebx_load_force(ra, "radiation_forces");
        rebx_set_param_double(ra, &fa->ap, "c", 3.e8);
        struct rebx_operator* op = rebx_load_operator(ra, "integrate_force");
        rebx_set_param_pointer(ra, &op->ap, "force", fa);
        rebx_set_param_int(ra, &op->ap, "integrator", REBX_INTEGRATOR_RK4);
        if (!rebx_add_operator(ra, op)) rc = 5;
        if (A->pre_timestep_modifications != rebx_pre_timestep_modifications || A->post_timestep_modifications != rebx_post_timestep_modifications) rc = 6;
        const int k = 5;
        reb_shim_integrate(A, k, 0);                    /* pre: RK4(dt/2) -- drift kick drift under gravity -- post: RK4(dt/2) */
        /* by hand on B: the same sequence with the session's integrate_force; gravity steps by the host harness */
        rebx_remove_force(rb, fb);
        fb = rebx_load_force(rb, "radiation_forces");
        rebx_set_param_double(rb, &fb->ap, "c", 3.e8);
        for (int q = 0; q < k; q++){
            for (int half = 0; half < 2; half++){
                if (half == 1) reb_shim_integrate(B, 1, 0);
                rebx_add_force(rb, fb);
                rebx_b200_session* s = rebx_b200_session_create(B, rb);
                int e = rebx_b200_session_integrate_force(s, 0.5*B->dt, REBX_INTEGRATOR_RK4);
                for (size_t i = 0; i < B->N; i++){ B->particles[i].ax = 0.; B->particles[i].ay = 0.; B->particles[i].az = 0.; }
                e |= rebx_b200_session_sync_to_host(s);
                rebx_b200_session_free(s);
                /* keep the force OUT of the host step's additional_forces, as on A */
                struct rebx_node* n0 = rb->additional_forces; rb->additional_forces = NULL; free(n0);
                if (e) rc = 7;
            }
        }
        size_t d = 0;
        for (size_t i = 0; i < A->N; i++){
            const struct reb_particle *p = &A->particles[i], *q = &B->particles[i];
            d += (p->x != q->x) + (p->y != q->y) + (p->z != q->z) + (p->vx != q->vx) + (p->vy != q->vy) + (p->vz != q->vz);
        }
        if (d) rc = 1;
        printf(", \"operator_integrate_force_rk4\": {\"differing_coordinates\": %zu, \"messages\": %zu}", d, reb_shim_message_count(A));
        rebx_free(ra); reb_simulation_free(A);
        rebx_free(rb); reb_simulation_free(B);
    }
    printf(", \"ok\": %s}\n", rc == 0 ? "true" : "false");
    return rc;
}

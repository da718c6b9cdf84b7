/* C user program for the multi-device dispatcher (GPU test, compiled and run by tests/test_gpu_parity.py):
 * BASELINE config 2 (star + N dust grains, radiation_forces with a beta per particle) and the massive ring of
 * config 3 (gravitational_harmonics + gr_potential, whose back-reaction sums cross the devices) are evaluated
 * through sim->additional_forces(sim) twice -- on ONE device, and spread by index range over ALL visible devices
 * (REBX_B200_DEVICES=all, read when the extras' device context is created) -- and compared bit for bit.
 * One process, one host thread, no Python, no torch.  Prints one line of JSON with the timings.
 *   argv[1] = number of test particles (default 2^21 + 37), argv[2] = timed calls (default 5). */
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
static unsigned long long rng_state = 88172645463325252ULL;
static double uniform(void){ rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (double)(rng_state >> 11)/9007199254740992.0; }

static struct reb_simulation* make_sim(size_t n, int ring){
    struct reb_simulation* sim = reb_simulation_create();
    sim->G = 6.674e-11;
    struct reb_particle p = {0};
    p.m = ring ? 5.68e26 : 1.99e30;
    reb_simulation_add(sim, p);
    rng_state = 88172645463325252ULL;
    for (size_t i = 0; i < n; i++){
        const double a = ring ? 7.0e7 + 7.0e7*uniform() : 1.5e11*(100. + 20.*uniform());
        const double th = 6.283185307179586*uniform(), inc = 0.01*uniform();
        const double v = sqrt(sim->G*p.m/a)*(1.0 + 0.05*(uniform() - 0.5));
        struct reb_particle q = {0};
        q.x = a*cos(th); q.y = a*sin(th)*cos(inc); q.z = a*sin(th)*sin(inc);
        q.vx = -v*sin(th); q.vy = v*cos(th)*cos(inc); q.vz = v*cos(th)*sin(inc);
        q.m = ring ? 1e10 + 1e14*uniform() : 0.0;
        reb_simulation_add(sim, q);
    }
    return sim;
}

static struct rebx_extras* attach(struct reb_simulation* sim, int ring){
    struct rebx_extras* rebx = rebx_attach(sim);
    const size_t n = sim->N - 1;
    if (!ring){
        struct rebx_force* rad = rebx_load_force(rebx, "radiation_forces");
        rebx_set_param_double(rebx, &rad->ap, "c", 3.e8);
        double* beta = malloc(n*sizeof *beta);
        int64_t* idx = malloc(n*sizeof *idx);
        rng_state = 1234567ULL;
        for (size_t i = 0; i < n; i++){ beta[i] = 0.01 + 0.49*uniform(); idx[i] = (int64_t)i + 1; }
        rebx_set_param_double_array(rebx, "beta", idx, beta, n);
        free(beta); free(idx);
        rebx_add_force(rebx, rad);
    } else {
        struct rebx_force* gh = rebx_load_force(rebx, "gravitational_harmonics");
        rebx_set_param_double(rebx, (struct rebx_node**)&sim->particles[0].ap, "J2", 1.6e-2);
        rebx_set_param_double(rebx, (struct rebx_node**)&sim->particles[0].ap, "J4", -9.4e-4);
        rebx_set_param_double(rebx, (struct rebx_node**)&sim->particles[0].ap, "R_eq", 6.03e7);
        rebx_add_force(rebx, gh);
        struct rebx_force* gr = rebx_load_force(rebx, "gr_potential");
        rebx_set_param_double(rebx, &gr->ap, "c", 3.e8);
        rebx_add_force(rebx, gr);
    }
    return rebx;
}

/* evaluates `calls` times; returns seconds per call of the last calls-1 (the call alone, not the reset of the
 * accelerations before it); copies the accelerations of the FIRST call */
static double run(struct reb_simulation* sim, int calls, double* acc){
    double spent = 0;
    for (int c = 0; c < calls; c++){
        for (size_t i = 0; i < sim->N; i++){ sim->particles[i].ax = 1e-3*sim->particles[i].x; sim->particles[i].ay = 0.; sim->particles[i].az = -1e-9; }
        const double t0 = now();
        sim->additional_forces(sim);
        if (c > 0) spent += now() - t0;
        if (c == 0) for (size_t i = 0; i < sim->N; i++){ acc[3*i] = sim->particles[i].ax; acc[3*i+1] = sim->particles[i].ay; acc[3*i+2] = sim->particles[i].az; }
    }
    return calls > 1 ? spent/(calls - 1) : 0.;
}

int main(int argc, char** argv){
    const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : ((size_t)1 << 21) + 37;
    const int calls = argc > 2 ? atoi(argv[2]) + 1 : 6;
    const int ndev = rebx_b200_device_count();
    if (ndev < 1){ printf("{\"skipped\": \"no CUDA device\"}\n"); return 0; }
    int rc = 0;
    printf("{\"devices\": %d, \"particles\": %zu", ndev, n);
    for (int ring = 0; ring < 2; ring++){
        double* a1 = malloc(3*(n + 1)*sizeof *a1);
        double* aN = malloc(3*(n + 1)*sizeof *aN);
        unsetenv("REBX_B200_DEVICES");
        struct reb_simulation* s1 = make_sim(n, ring);
        struct rebx_extras* r1 = attach(s1, ring);
        const double t1 = run(s1, calls, a1);
        if (reb_shim_message_count(s1)){ fprintf(stderr, "%s\n", reb_shim_message(s1, 0)); rc = 2; }
        rebx_free(r1); reb_simulation_free(s1);
        setenv("REBX_B200_DEVICES", "all", 1);
        struct reb_simulation* sN = make_sim(n, ring);
        struct rebx_extras* rN = attach(sN, ring);
        const double tN = run(sN, calls, aN);
        if (reb_shim_message_count(sN)){ fprintf(stderr, "%s\n", reb_shim_message(sN, 0)); rc = 2; }
        uint64_t st[5];
        rebx_b200_get_stats(rN, st);
        rebx_free(rN); reb_simulation_free(sN);
        /* test particles: bit for bit.  The central body of the ring: its back-reaction is a sum over chunks whose
         * boundaries move with the device count, so it is compared to 1e-12 of the sum of absolute terms' scale. */
        size_t diff = 0;
        for (size_t i = 3; i < 3*(n + 1); i++) if (memcmp(&a1[i], &aN[i], sizeof(double))) diff++;
        double body = 0.;
        for (int k = 0; k < 3; k++){ const double d = fabs(a1[k] - aN[k]), s = fabs(a1[k]) + 1e-300; if (d/s > body) body = d/s; }
        if (diff != 0 || body > 1e-10) rc = 1;
        printf(", \"%s\": {\"ms_per_call_1_device\": %.3f, \"ms_per_call_all_devices\": %.3f, \"differing_components\": %zu, \"central_body_rel_diff\": %.3g, \"h2d_bytes\": %llu}",
               ring ? "ring_massive" : "debris_disk", 1e3*t1, 1e3*tN, diff, body, (unsigned long long)st[2]);
        free(a1); free(aN);
    }
    printf(", \"ok\": %s}\n", rc == 0 ? "true" : "false");
    return rc;
}

/* A C99 translation unit written the way a REBOUNDx user writes problem.c (reference
 * the example problem.c files): includes "rebound.h" and "reboundx.h", attaches, loads and adds forces,
 * sets parameters with the documented casts, adds a custom C force and lets the host call
 * sim->additional_forces.  Compiled by tests/test_c_api.py against include/ and the built library;
 * proves the public headers are valid C and the symbols link.  No GPU is needed: only the custom
 * (host) force is evaluated, the device-backed force is removed again before the dispatch. */
#include <stdio.h>
#include <string.h>
#include "rebound.h"
#include "reboundx.h"
#include "rebx_b200.h"

static int calls = 0;
static void my_force(struct reb_simulation* const sim, struct rebx_force* const force, struct reb_particle* const particles, const int N){
    (void)sim; (void)force;
    calls++;
    for (int i = 0; i < N; i++) particles[i].ax += 1.0 + i;
}

int main(void){
    struct reb_simulation* sim = reb_simulation_create();
    struct reb_particle p = {0};
    p.m = 1.0; reb_simulation_add(sim, p);
    p.m = 1e-3; p.x = 1.0; p.vy = 1.0; reb_simulation_add(sim, p);

    struct rebx_extras* rebx = rebx_attach(sim);
    if (!rebx || sim->extras != rebx) return 1;

    struct rebx_force* rad = rebx_load_force(rebx, "radiation_forces");
    if (!rad || rad->force_type != REBX_FORCE_VEL) return 2;
    rebx_set_param_double(rebx, &rad->ap, "c", 3.e8);
    rebx_set_param_double(rebx, (struct rebx_node**)&sim->particles[1].ap, "beta", 0.1);
    rebx_set_param_int(rebx, AP_PTR(sim->particles[0]), "radiation_source", 1);
    double* c = rebx_get_param(rebx, rad->ap, "c");
    double* beta = rebx_get_param(rebx, sim->particles[1].ap, "beta");
    if (!c || *c != 3.e8 || !beta || *beta != 0.1) return 3;
    if (rebx_get_param(rebx, sim->particles[0].ap, "beta") != NULL) return 4;
    if (!rebx_add_force(rebx, rad) || sim->additional_forces != rebx_additional_forces) return 5;
    if (!sim->force_is_velocity_dependent) return 6;
    if (rebx_get_force(rebx, "radiation_forces") != rad) return 7;
    if (!rebx_remove_force(rebx, rad)) return 8;            /* frees it; keep the dispatch host-only */

    struct rebx_force* mine = rebx_create_force(rebx, "my_force");
    mine->force_type = REBX_FORCE_POS;
    mine->update_accelerations = my_force;
    if (!rebx_add_force(rebx, mine)) return 9;
    sim->additional_forces(sim);
    if (calls != 1 || sim->particles[0].ax != 1.0 || sim->particles[1].ax != 2.0) return 10;

    if (rebx_load_force(rebx, "no_such_force") != NULL) return 11;          /* queues a REBOUND error */
    if (reb_shim_message_count(sim) == 0) return 12;
    if (rebx_rad_calc_beta(6.674e-11, 3e8, 1.99e30, 3.85e26, 1e-5, 1000., 1.) <= 0.) return 13;
    if (rebx_b200_workspace_bytes() == 0) return 14;                          /* device-layer header is C too */
    if (strcmp(rebx_version_str, "5.1.0") != 0) return 15;

    rebx_free(rebx);
    if (sim->extras != NULL || sim->additional_forces != NULL) return 16;
    reb_simulation_free(sim);
    printf("c api ok\n");
    return 0;
}

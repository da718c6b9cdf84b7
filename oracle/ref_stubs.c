/* oracle/ref_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * The reference's src/core.c takes the address of every built-in effect (core.c:282-350,
 * 400-444).  _ref/libreboundx_ref.so is built from the hot-path sources only, so the
 * effects that are out of scope are given trapping stubs
 * here to let the library load.  None of them is ever called by the oracle. */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void){ fprintf(stderr, "oracle/_ref: out-of-scope effect '%s' called\n", #name); abort(); }
STUB(rebx_gr_full) STUB(rebx_stochastic_forces)
STUB(rebx_tides_constant_time_lag) STUB(rebx_tides_dynamical) STUB(rebx_tides_spin)
STUB(rebx_modify_mass) STUB(rebx_modify_orbits_direct) STUB(rebx_track_min_distance)
STUB(rebx_ias15_step) STUB(rebx_kepler_step) STUB(rebx_jump_step)
STUB(rebx_interaction_step) STUB(rebx_drift_step) STUB(rebx_kick_step)

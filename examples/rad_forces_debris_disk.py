"""Radiation pressure + Poynting-Robertson drag on a debris disk -- the set-up of the reference's
examples/rad_forces_debris_disk/problem.c (Sun + 1000 dust grains, a in [100,120] AU, e = 0.1,
beta = 0.1, SI units), driven through the reference's own API calls:

    rebx_attach -> rebx_load_force("radiation_forces") -> rebx_set_param_double("c") ->
    rebx_add_force -> rebx_set_param_double(particle, "beta") -> [REBOUND calls sim->additional_forces]

REBOUND is not available in this environment; `reboundx_b200.host.Simulation` is a stand-in whose
`integrate()` is a plain drift-kick-drift leapfrog (NOT WHFast).  Needs a CUDA device.

    python examples/rad_forces_debris_disk.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reboundx_b200 import workloads                         # noqa: E402
from reboundx_b200.host import Extras, Simulation           # noqa: E402

AU, YR = 1.5e11, 3.15e7
w = workloads.debris_disk(1000, seed=3)

sim = Simulation(G=w["G"], integrator="whfast")
sim.set_particles(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["m"], w["r"])
sim.dt = 1e8
sim.N_active = 1

rebx = Extras(sim)                                          # rebx_attach
rad = rebx.load_force("radiation_forces")                   # rebx_load_force
rad.params["c"] = 3.0e8                                     # rebx_set_param_double(&rad->ap, "c", ...)
rebx.add_force(rad)                                         # rebx_add_force
for i in range(1, sim.N):
    rebx.particle_params(i)["beta"] = 0.1                   # rebx_set_param_double(&particles[i].ap, "beta", ...)


def semi_major_axes():
    x, y, z, vx, vy, vz = sim.posvel()
    mu = w["G"]*w["m"][0]*(1.0 - 0.1)          # radiation pressure reduces the effective solar mass by beta
    r = np.sqrt(x[1:]**2 + y[1:]**2 + z[1:]**2)
    v2 = vx[1:]**2 + vy[1:]**2 + vz[1:]**2
    return -mu/(v2 - 2*mu/r)/AU


a0 = semi_major_axes()
steps = 2000
sim.integrate(steps, "leapfrog")
warnings = {m for k, m in sim.messages()}                   # the reference's WHFast + velocity-dependent-force warning
a1 = semi_major_axes()
print(f"{steps} steps of dt = 1e8 s ({steps*1e8/YR:.0f} yr); device path stats: {rebx.stats()}")
print(f"mean semi-major axis (about the beta-reduced Sun) {a0.mean():.5f} AU -> {a1.mean():.5f} AU   (Poynting-Robertson drag shrinks the orbits)")
print(f"{len(warnings)} distinct warning(s) queued, e.g.: {sorted(warnings)[0][:90]}..." if warnings else "no warnings")
assert np.isfinite(a1).all() and a1.mean() < a0.mean()

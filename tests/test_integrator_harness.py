"""The host integrator harness of the REBOUND stand-in (test infrastructure for the trajectory criterion:
REBOUND's own IAS15 / WHFast are not in the reference tree nor in this image).  The two schemes added for
north-star criterion (ii) are checked against the analytic two-body solution:
  gauss_radau    15th-order Gauss-Radau predictor-corrector, fixed step (the scheme behind IAS15)
  wisdom_holman  Kepler drift about particle 0 + kicks by additional_forces (the splitting behind WHFast)
CPU only."""
import numpy as np
import pytest


def kepler_sim(e=0.3, m1=0.0, dt=0.1):
    from reboundx_b200.host import Simulation
    sim = Simulation(G=1.0)
    a = 1.0
    x = np.array([0.0, a*(1 - e)]); y = np.zeros(2); z = np.zeros(2)
    vy = np.array([0.0, np.sqrt((1 + m1)*(1 + e)/(a*(1 - e)))])
    sim.set_particles(x, y, z, vx=np.zeros(2), vy=vy, vz=np.zeros(2), m=np.array([1.0, m1]))
    sim.dt = dt
    sim.N_active = 1 if m1 == 0.0 else 2
    return sim


def state(sim):
    x, y, z, vx, vy, vz = sim.posvel()
    return np.array([x[1] - x[0], y[1] - y[0], z[1] - z[0], vx[1] - vx[0], vy[1] - vy[0], vz[1] - vz[0]])


def energy(s, mu=1.0):
    return 0.5*np.dot(s[3:], s[3:]) - mu/np.linalg.norm(s[:3])


@pytest.mark.parametrize("scheme,steps_per_orbit,tol", [("gauss_radau", 40, 2e-13), ("wisdom_holman", 7, 2e-13)])
def test_two_body_returns_to_start(built, scheme, steps_per_orbit, tol):
    sim = kepler_sim(e=0.3, dt=2*np.pi/steps_per_orbit)
    s0 = state(sim)
    sim.integrate(3*steps_per_orbit, scheme=scheme)            # three full periods
    s1 = state(sim)
    assert np.max(np.abs(s1 - s0)) < tol*3*steps_per_orbit, (s1 - s0)
    assert abs(energy(s1) - energy(s0)) < 1e-13
    sim.free()


def test_gauss_radau_is_high_order(built):
    errs = []
    for n in (6, 12):
        sim = kepler_sim(e=0.2, dt=2*np.pi/n)
        s0 = state(sim)
        sim.integrate(n, scheme="gauss_radau")
        errs.append(np.max(np.abs(state(sim) - s0)))
        sim.free()
    assert errs[1] < 1e-9 and errs[0]/errs[1] > 2.0**10, errs             # ~ 2^15 per halving until round-off


def test_wisdom_holman_with_a_massive_companion_and_kick(built):
    """mu = G (m0 + m1) in the drift; with no additional force a step is the exact two-body flow."""
    sim = kepler_sim(e=0.5, m1=1e-3, dt=0.37)
    s0 = state(sim)
    sim.integrate(200, scheme="wisdom_holman")
    s1 = state(sim)
    assert abs(energy(s1, 1.001) - energy(s0, 1.001)) < 1e-13
    # same elapsed time in one step of the same map
    sim2 = kepler_sim(e=0.5, m1=1e-3, dt=0.37*200)
    sim2.integrate(1, scheme="wisdom_holman")
    assert np.max(np.abs(state(sim2) - s1)) < 1e-11
    sim.free(); sim2.free()

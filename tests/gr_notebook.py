"""The reference's published known answers for `gr` (1PN, src/gr.c), produced with real REBOUND + IAS15:

  ipython_examples/GeneralRelativity.ipynb cells 1-9: Sun + Mercury-like planet (m = 1.66013e-07, a = 0.387098,
  e = 0.205630, G = 1), 10 time units without the force, then 100 with it:
      pomega = 0.0000332591781488      Rate of change of pomega = 43.1048 [arcsec / Julian century]
  ipython_examples/SavingAndLoadingSimulations.ipynb cells 1-3, 13: the same system with `gr` from t = 0:
      t = 10.0, pomega = 3.4181522883806095e-06        t = 20.0, pomega = 6.317073080790619e-06

REBOUND derives pomega through acos of a cosine that differs from 1 by pomega^2/2, so the PRINTED values carry
an absolute error of about eps/pomega (3e-11 rad at 3.4e-6, 3e-12 at 3.3e-5); the comparisons allow for that.
Here the trajectory comes from the stand-in's fixed-step Gauss-Radau scheme (the scheme behind IAS15) and pomega
from the eccentricity vector."""
import numpy as np

C = 10065.32012038219                 # reboundx.constants.C, printed in SavingAndLoadingSimulations.ipynb cell 11
M1, A, E = 1.66013e-07, 0.387098, 0.205630
MU = 1.0 + M1
NB_10_PLUS_100 = 0.0000332591781488
NB_RATE = 43.1048
NB_T10, NB_T20 = 3.4181522883806095e-06, 6.317073080790619e-06


def initial_state():
    x = np.array([0.0, A*(1 - E)])
    vy = np.array([0.0, np.sqrt(MU*(1 + E)/(A*(1 - E)))])
    m = np.array([1.0, M1])
    x -= (m*x).sum()/m.sum()              # sim.move_to_com()
    vy -= (m*vy).sum()/m.sum()
    z = np.zeros(2)
    return {"x": x, "y": z.copy(), "z": z.copy(), "vx": z.copy(), "vy": vy, "vz": z.copy(), "m": m, "r": z.copy()}


def pomega(posvel):
    X, Y, Z, VX, VY, VZ = posvel
    d = np.array([X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]])
    v = np.array([VX[1] - VX[0], VY[1] - VY[0], VZ[1] - VZ[0]])
    ev = ((v.dot(v) - MU/np.linalg.norm(d))*d - d.dot(v)*v)/MU
    return float(np.arctan2(ev[1], ev[0]))


def run(lib, steps_per_orbit=100):
    """Returns (pomega after 10 + 100 [GeneralRelativity.ipynb], pomega at t = 10, at t = 20 [SavingAndLoading...])."""
    from reboundx_b200 import scenarios
    period = 2*np.pi*np.sqrt(A**3/MU)
    n10 = int(round(10.0/(period/steps_per_orbit)))

    def build(forces):
        case = {"G": 1.0, "integrator": "ias15", "state": initial_state(), "acc0": None, "forces": forces, "particle_params": {}}
        sim, rebx, _ = scenarios.build_case(case, lib)
        sim.dt = 10.0/n10
        return sim, rebx

    sim, rebx = build([])
    sim.integrate(n10, "gauss_radau")
    gr = rebx.load_force("gr"); gr.params["c"] = C; rebx.add_force(gr)
    sim.integrate(10*n10, "gauss_radau")
    assert not [m for k, m in sim.messages() if k == "error"]
    p110 = pomega(np.array(sim.posvel()))
    rebx.detach(); sim.free()

    sim, rebx = build([{"name": "gr", "params": {"c": C}}])
    sim.integrate(n10, "gauss_radau")
    p10 = pomega(np.array(sim.posvel()))
    sim.integrate(n10, "gauss_radau")
    p20 = pomega(np.array(sim.posvel()))
    assert not [m for k, m in sim.messages() if k == "error"]
    rebx.detach(); sim.free()
    return p110, p10, p20


def check(p110, p10, p20):
    eps = np.finfo(float).eps
    assert abs(p110/NB_10_PLUS_100 - 1.0) < 1e-7, p110                 # measured 3e-8 (7e-13 rad; acos noise of the print: 3e-12)
    assert "%.4f" % (p110/100.0*628.33195/4.8481368e-06) == "%.4f" % NB_RATE      # the notebook's own derived quantity, as printed
    assert abs(p10 - NB_T10) < 2*eps/NB_T10, (p10, NB_T10)              # measured 1.5e-11 rad
    assert abs(p20 - NB_T20) < 2*eps/NB_T20, (p20, NB_T20)              # measured 8e-12 rad

"""The reference's own published known answers for yarkovsky_effect: reference
ipython_examples/YarkovskyEffect.ipynb, cells 4-12 (full version) and 16-20 (simple versions, flags +1 / -1):
one star, asteroids on circular orbits at 0.5 (and 0.75) AU, WHFast with dt = 0.05 yr for 1e5 yr, printed

    CHANGE IN SEMI-MAJOR AXIS: 1.922647061824989e-05 AU                     (cell 12)
    CHANGE IN SEMI-MAJOR AXIS(Asteroid 1): 4.1780726501072785e-05 AU        (cell 20)
    CHANGE IN SEMI-MAJOR AXIS(Asteroid 2): -3.411492135974026e-05 AU        (cell 20)

Shared by the CPU test (oracle) and the GPU test (this library)."""
import numpy as np

# rebound.units: G_SI * msun * yr^2 / au^3 with msun = 1.3271244004193938e20 / G_SI  (sim.units = ('yr', 'AU', 'Msun'))
G = 1.3271244004193938e20*31557600.0**2/149597870700.0**3
NOTEBOOK_FULL = 1.922647061824989e-05
NOTEBOOK_SIMPLE = (4.1780726501072785e-05, -3.411492135974026e-05)
C, STEF_BOLTZ, LSTAR = 63241.077, 8.96e-16, 2.7e-4                                   # cell 6
DENSITY, RADIUS, GAMMA, PROT, ALBEDO, EMISSIVITY, K = 5.05e6, 6.68e-09, 8.72e-10, 4.9e-4, .017, .9, .25
DT, STEPS = 0.05, 2_000_000                                                         # sim.dt, tmax = 1e5 yr


def case(full):
    a0 = [0.5] if full else [0.5, 0.75]
    n = 1 + len(a0)
    z = np.zeros(n)
    st = {"x": np.array([0.0] + a0), "y": z.copy(), "z": z.copy(), "vx": z.copy(), "vy": np.array([0.0] + [np.sqrt(G/a) for a in a0]),
          "vz": z.copy(), "m": np.array([1.0] + [0.0]*len(a0)), "r": np.array([0.0] + [RADIUS]*len(a0))}
    idx = np.arange(1, n, dtype=np.int64)
    pp = {"ye_body_density": (idx, np.full(n - 1, DENSITY)), "ye_albedo": (idx, np.full(n - 1, ALBEDO))}
    fparams = {"ye_c": C, "ye_lstar": LSTAR}
    if full:
        fparams["ye_stef_boltz"] = STEF_BOLTZ
        one = lambda v: (idx, np.array([v]))
        pp.update({"ye_flag": (idx, np.array([0], dtype=np.int32)), "ye_emissivity": one(EMISSIVITY), "ye_k": one(K),
                   "ye_thermal_inertia": one(GAMMA), "ye_rotation_period": one(PROT),
                   "ye_spin_axis_x": one(0.0), "ye_spin_axis_y": one(0.0), "ye_spin_axis_z": one(1.0)})
    else:
        pp["ye_flag"] = (idx, np.array([1, -1], dtype=np.int32))
    return {"G": G, "integrator": "whfast", "state": st, "acc0": None,
            "forces": [{"name": "yarkovsky_effect", "params": fparams}], "particle_params": pp}, a0


def semi_major_axes(posvel):
    """ps[i].a of the notebook: osculating semi-major axis about the star (massless asteroids)."""
    x, y, z, vx, vy, vz = posvel
    out = []
    for i in range(1, len(x)):
        d = np.array([x[i] - x[0], y[i] - y[0], z[i] - z[0]])
        v = np.array([vx[i] - vx[0], vy[i] - vy[0], vz[i] - vz[0]])
        out.append(1.0/(2.0/np.linalg.norm(d) - v.dot(v)/G))
    return np.array(out)


def drift(full, lib, steps=STEPS):
    """Change of the semi-major axes after `steps` Wisdom-Holman steps of 0.05 yr with `lib`'s force."""
    from reboundx_b200 import scenarios
    c, a0 = case(full)
    sim, rebx, _ = scenarios.build_case(c, lib)
    sim.dt, sim.N_active = DT, 1
    sim.integrate(steps, "wisdom_holman")
    errs = [m for k, m in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    pv = np.array(sim.posvel())
    rebx.detach(); sim.free()
    return semi_major_axes(pv) - np.array(a0), pv

"""Analytic laws the reference publishes for its damping forces (ipython_examples, produced with real REBOUND):

  Migration.ipynb cells 1-9        modify_orbits_forces (default frame: JACOBI), two planets of 1e-6 at a = 1 and 10 with
                                   e = inc = 0.1, tau_a = -500 and -1000: "a(t) = a0 e^(t/tau_a)" over tmax = 1e3 (IAS15)
  InnerDiskEdge.ipynb cells 3-11   modify_orbits_forces with its inner disc edge: a planet of 1e-4 at 0.5 AU, tau_a = -1e3, tau_e = -1e2,
                                   edge at 0.1 AU of width 0.02, WHFast dt = 0.1^1.5/20 for 3000 yr: a(t) = max(0.5 e^(-t/1e3), 0.1)
  GasDampingTimescale.ipynb        gas_damping_timescale: an Earth mass at 0.1 AU, e = 0.05, inc = 5 deg, d_factor = 5, cs_coeff = 0.272,
                                   tau_coeff = 0.003, 500 yr: "e'/e = -1/tau, i'/i = -1/(2 tau)" with tau of Dawson et al. 2016 eq. 16
                                   (cells 0-1, 7), integrated here as an ODE in (e, i)
  TypeIMigration.ipynb cells 2-15  type_I_migration with its inner disc edge: a planet of 1e-4 on a circular orbit at 1 AU,
                                   h0 = 0.03, Sigma0 = 1.1e-4, alpha = 1, flaring 0.25, edge at 0.1 AU of width h0 0.1^0.25,
                                   WHFast dt = 0.1^1.5/20 for 4000 yr:
                                   a(t) = max(1 - t/tau~, 0.1) with tau~ = h0^2/((2.7 + 1.1 alpha) m Sigma0 sqrt G)

The notebooks show these as plots (no printed digits), so they pin the damping forces -- and with them the restated
reb_orbit_from_particle elements and rebx_com_force frames the forces are built on -- at the accuracy of the laws
themselves (first order in e^2, in the planet mass and in the width of the trap), not at round-off.
Shared by the CPU test (oracle = the reference's own sources) and the GPU test (this library; `stop_after` = only the
first so many samples, the host-facing call costs ~30 us per force evaluation at these sizes)."""
import numpy as np

G_YR_AU_MSUN = 1.3271244004193938e20*31557600.0**2/149597870700.0**3       # sim.units = ('yr', 'AU', 'Msun')


def _kepler_state(G, M, bodies):
    """sim.add(m, a, e, inc) at pericentre, one after the other as REBOUND does (each about the centre of mass of the
    bodies already present -- with planets of 1e-6 .. 1e-4 the star itself to that order), then sim.move_to_com()."""
    m = np.array([M] + [b[0] for b in bodies])
    n = len(m)
    x, y, z, vx, vy, vz = (np.zeros(n) for _ in range(6))
    for i, (mp, a, e, inc) in enumerate(bodies, start=1):
        mt = m[:i].sum()
        cx, cvy = (m[:i]*x[:i]).sum()/mt, (m[:i]*vy[:i]).sum()/mt
        cvz = (m[:i]*vz[:i]).sum()/mt
        r = a*(1 - e)
        v = np.sqrt(G*(mt + mp)*(1 + e)/(a*(1 - e)))
        x[i] = cx + r
        vy[i] = cvy + v*np.cos(inc)
        vz[i] = cvz + v*np.sin(inc)
    for q in (x, y, z, vx, vy, vz):
        q -= (m*q).sum()/m.sum()
    return {"x": x, "y": y, "z": z, "vx": vx, "vy": vy, "vz": vz, "m": m, "r": np.zeros(n)}


def semi_major_axes(posvel, m, G):
    X, Y, Z, VX, VY, VZ = posvel
    out = []
    for i in range(1, len(m)):
        d = np.array([X[i] - X[0], Y[i] - Y[0], Z[i] - Z[0]])
        v = np.array([VX[i] - VX[0], VY[i] - VY[0], VZ[i] - VZ[0]])
        out.append(1.0/(2.0/np.linalg.norm(d) - v.dot(v)/(G*(m[0] + m[i]))))
    return np.array(out)


def migration(lib, steps_per_inner_orbit=60, samples=5, stop_after=None):
    """Migration.ipynb: returns (times, a[2][samples], predicted a[2][samples])."""
    from reboundx_b200 import scenarios
    tmax, a0 = 1.0e3, (1.0, 10.0)
    st = _kepler_state(1.0, 1.0, [(1e-6, a0[0], 0.1, 0.1), (1e-6, a0[1], 0.1, 0.1)])
    idx = np.array([1, 2], dtype=np.int64)
    tau = np.array([-tmax/2.0, -tmax])
    case = {"G": 1.0, "integrator": "ias15", "state": st, "acc0": None,
            "forces": [{"name": "modify_orbits_forces", "params": {}}], "particle_params": {"tau_a": (idx, tau)}}
    sim, rebx, _ = scenarios.build_case(case, lib)
    per_sample = int(round(tmax/samples/(2*np.pi/steps_per_inner_orbit)))
    sim.dt = tmax/samples/per_sample
    times, got = [], []
    for s in range(samples if stop_after is None else stop_after):
        sim.integrate(per_sample, "gauss_radau")
        times.append((s + 1)*tmax/samples)
        got.append(semi_major_axes(np.array(sim.posvel()), st["m"], 1.0))
    errs = [msg for k, msg in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    rebx.detach(); sim.free()
    times = np.array(times)
    pred = np.array([a0[k]*np.exp(times/tau[k]) for k in range(2)])
    return times, np.array(got).T, pred


def type_I(lib, samples=8, stop_after=None):
    """TypeIMigration.ipynb: returns (times, a[samples], predicted a[samples], edge position, edge width)."""
    from reboundx_b200 import scenarios
    G, mp, h0, sd0, alpha, flare, edge = G_YR_AU_MSUN, 1.0e-4, 0.03, 1.1e-4, 1.0, 0.25, 0.1
    width = h0*edge**flare
    st = _kepler_state(G, 1.0, [(mp, 1.0, 0.0, 0.0)])
    case = {"G": G, "integrator": "whfast", "state": st, "acc0": None, "particle_params": {},
            "forces": [{"name": "type_I_migration", "params": {"tIm_scale_height_1": h0, "tIm_surface_density_1": sd0,
                                                                 "tIm_surface_density_exponent": alpha, "tIm_flaring_index": flare,
                                                                 "ide_position": edge, "ide_width": width}}]}
    sim, rebx, _ = scenarios.build_case(case, lib)
    tmax = 4.0e3
    dt = edge**1.5/20.0
    per_sample = int(round(tmax/samples/dt))
    sim.dt = tmax/samples/per_sample
    sim.N_active = 1                     # the stand-in's Wisdom-Holman scheme: Kepler drifts about the star + kicks by the force
    times, got = [], []
    for s in range(samples if stop_after is None else stop_after):
        sim.integrate(per_sample, "wisdom_holman")
        times.append((s + 1)*tmax/samples)
        got.append(semi_major_axes(np.array(sim.posvel()), st["m"], G)[0])
    errs = [msg for k, msg in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    rebx.detach(); sim.free()
    times = np.array(times)
    tau_tilde = h0**2/((2.7 + 1.1*alpha)*mp*sd0*np.sqrt(G))
    pred = np.maximum(1.0 - times/tau_tilde, edge)
    return times, np.array(got), pred, edge, width


def inner_disk_edge(lib, samples=6, stop_after=None):
    """InnerDiskEdge.ipynb: returns (times, a[samples], predicted a[samples])."""
    from reboundx_b200 import scenarios
    G, mp, edge, width = G_YR_AU_MSUN, 1.0e-4, 0.1, 0.02
    st = _kepler_state(G, 1.0, [(mp, 0.5, 0.0, 0.0)])
    one = np.array([1], dtype=np.int64)
    case = {"G": G, "integrator": "whfast", "state": st, "acc0": None,
            "particle_params": {"tau_a": (one, np.array([-1.0e3])), "tau_e": (one, np.array([-1.0e2]))},
            "forces": [{"name": "modify_orbits_forces", "params": {"ide_position": edge, "ide_width": width}}]}
    sim, rebx, _ = scenarios.build_case(case, lib)
    tmax = 3.0e3
    per_sample = int(round(tmax/samples/(edge**1.5/20.0)))
    sim.dt = tmax/samples/per_sample
    sim.N_active = 1
    times, got = [], []
    for s in range(samples if stop_after is None else stop_after):
        sim.integrate(per_sample, "wisdom_holman")
        times.append((s + 1)*tmax/samples)
        got.append(semi_major_axes(np.array(sim.posvel()), st["m"], G)[0])
    errs = [msg for k, msg in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    rebx.detach(); sim.free()
    times = np.array(times)
    return times, np.array(got), np.maximum(0.5*np.exp(-times/1.0e3), edge)


def _ecc_inc(posvel, m, G):
    X, Y, Z, VX, VY, VZ = posvel
    d = np.array([X[1] - X[0], Y[1] - Y[0], Z[1] - Z[0]])
    v = np.array([VX[1] - VX[0], VY[1] - VY[0], VZ[1] - VZ[0]])
    mu = G*(m[0] + m[1])
    ev = ((v.dot(v) - mu/np.linalg.norm(d))*d - d.dot(v)*v)/mu
    h = np.cross(d, v)
    return float(np.linalg.norm(ev)), float(np.arccos(h[2]/np.linalg.norm(h)))


def gas_damping(lib, samples=5, stop_after=None):
    """GasDampingTimescale.ipynb: returns (times, e[samples], inc[samples], e and inc of the documented law)."""
    from reboundx_b200 import scenarios
    G, mp, a0, e0, i0, d, cs_coeff, tau_coeff = G_YR_AU_MSUN, 3.0e-6, 0.1, 0.05, 5*np.pi/180, 5.0, 0.272, 0.003
    st = _kepler_state(G, 1.0, [(mp, a0, e0, i0)])
    case = {"G": G, "integrator": "whfast", "state": st, "acc0": None,
            "particle_params": {"d_factor": (np.array([1], dtype=np.int64), np.array([d]))},
            "forces": [{"name": "gas_damping_timescale", "params": {"cs_coeff": cs_coeff, "tau_coeff": tau_coeff}}]}
    sim, rebx, _ = scenarios.build_case(case, lib)
    tmax = 500.0
    period = 2*np.pi*np.sqrt(a0**3/(G*(1.0 + mp)))
    per_sample = int(round(tmax/samples/(period/20.0)))
    sim.dt = tmax/samples/per_sample
    sim.N_active = 1
    times, ecc, inc = [], [], []
    for s in range(samples if stop_after is None else stop_after):
        sim.integrate(per_sample, "wisdom_holman")
        times.append((s + 1)*tmax/samples)
        e, i = _ecc_inc(np.array(sim.posvel()), st["m"], G)
        ecc.append(e); inc.append(i)
    errs = [msg for k, msg in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    rebx.detach(); sim.free()

    def tau(e, i):                      # the notebook's calc_tau (cell 7) with the force's own coefficients
        cs, vk = cs_coeff*a0**-0.25, np.sqrt(G/a0)
        v = np.sqrt(e*e + i*i)*vk
        t = tau_coeff*d*a0*a0/mp
        if v <= cs:
            return t
        return t*(v/cs)**3 if i < cs/vk else t*(v/cs)**4

    def rhs(y):
        t = tau(y[0], y[1])
        return np.array([-y[0]/t, -y[1]/(2*t)])

    y, law, k, nsub = np.array([e0, i0]), [], 0, 400
    h = tmax/samples/nsub
    for s in range(samples if stop_after is None else stop_after):
        for _ in range(nsub):           # classical Runge-Kutta on (e, i)
            k1 = rhs(y); k2 = rhs(y + 0.5*h*k1); k3 = rhs(y + 0.5*h*k2); k4 = rhs(y + h*k3)
            y = y + h*(k1 + 2*k2 + 2*k3 + k4)/6
        law.append(y.copy())
    law = np.array(law)
    return np.array(times), np.array(ecc), np.array(inc), law[:, 0], law[:, 1]

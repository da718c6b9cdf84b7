"""Seeded synthetic ensembles for the BASELINE.json configurations.

Pure numpy; everything is generated from `numpy.random.default_rng(seed)` so a test, the
bench and the oracle see bit-identical inputs.  Each builder returns a dict with the SoA
particle state (`x y z vx vy vz m r`, index 0 is the central body), the gravitational
constant `G`, per-particle parameter arrays and force-level parameters.
"""
import numpy as np

AU_SI = 1.5e11
TWO_PI = 2.0*np.pi


def kepler_to_cartesian(mu, a, e, inc, Omega, omega, f):
    """Vectorised orbital elements -> position/velocity relative to the primary."""
    cosf, sinf = np.cos(f), np.sin(f)
    r = a*(1.0 - e*e)/(1.0 + e*cosf)
    v0 = np.sqrt(mu/(a*(1.0 - e*e)))
    cO, sO = np.cos(Omega), np.sin(Omega)
    co, so = np.cos(omega), np.sin(omega)
    ci, si = np.cos(inc), np.sin(inc)
    cof, sof = co*cosf - so*sinf, so*cosf + co*sinf          # cos/sin(omega+f)
    x = r*(cO*cof - sO*sof*ci)
    y = r*(sO*cof + cO*sof*ci)
    z = r*(sof*si)
    vx = v0*((e + cosf)*(-ci*co*sO - cO*so) - sinf*(co*cO - ci*so*sO))
    vy = v0*((e + cosf)*(ci*co*cO - sO*so) - sinf*(co*sO + ci*so*cO))
    vz = v0*((e + cosf)*co*si - sinf*si*so)
    return x, y, z, vx, vy, vz


def _assemble(central_mass, x, y, z, vx, vy, vz, m=None, r=None):
    n = x.shape[0] + 1
    out = {}
    for name, arr in (("x", x), ("y", y), ("z", z), ("vx", vx), ("vy", vy), ("vz", vz)):
        full = np.zeros(n)
        full[1:] = arr
        out[name] = full
    mm = np.zeros(n)
    mm[0] = central_mass
    if m is not None:
        mm[1:] = m
    out["m"] = mm
    rr = np.zeros(n)
    if r is not None:
        rr[1:] = r
    out["r"] = rr
    return out


def gravity_like_acc(w, rng):
    """A deterministic non-zero starting acceleration (stands in for REBOUND's gravity):
    Keplerian pull of the central body plus a small random part, so that a += f is tested
    against a realistic accumulation base."""
    dx, dy, dz = w["x"] - w["x"][0], w["y"] - w["y"][0], w["z"] - w["z"][0]
    r2 = dx*dx + dy*dy + dz*dz
    r2[0] = 1.0
    pref = -w["G"]*w["m"][0]/(r2*np.sqrt(r2))
    pref[0] = 0.0
    jitter = 1.0 + 1e-3*rng.standard_normal(dx.shape[0])
    return pref*dx*jitter, pref*dy*jitter, pref*dz*jitter


def debris_disk(n, seed=3):
    """Config 2: star + n massless dust grains, radiation_forces with beta per particle
    (shape of reference examples/rad_forces_debris_disk/problem.c:21-74, SI units)."""
    rng = np.random.default_rng(seed)
    G, c, mstar = 6.674e-11, 3.0e8, 1.99e30
    a = rng.uniform(100.0, 120.0, n)*AU_SI
    e = np.full(n, 0.1)
    inc = rng.uniform(0.0, np.arcsin(0.1), n)
    Om, om, f = (rng.uniform(0.0, TWO_PI, n) for _ in range(3))
    w = _assemble(mstar, *kepler_to_cartesian(G*mstar, a, e, inc, Om, om, f))
    w.update(G=G, c=c, beta=rng.uniform(0.01, 0.5, n), name="debris_disk")
    return w


def circumplanetary_ring(n, seed=5, tilted=False, massive=False):
    """Config 3: oblate planet (J2, J4) + n ring particles; gravitational_harmonics + gr_potential."""
    rng = np.random.default_rng(seed)
    G, c, mp = 6.674e-11, 3.0e8, 5.68e26
    a = rng.uniform(7.0e7, 1.4e8, n)
    e = rng.uniform(0.0, 0.01, n)
    inc = rng.uniform(0.0, 0.01, n)
    Om, om, f = (rng.uniform(0.0, TWO_PI, n) for _ in range(3))
    m = rng.uniform(1e10, 1e14, n) if massive else None
    w = _assemble(mp, *kepler_to_cartesian(G*mp, a, e, inc, Om, om, f), m=m)
    w.update(G=G, c=c, J2=1.6e-2, J4=-9.4e-4, R_eq=6.03e7,
             Omega=(0.3, -0.2, 0.9) if tilted else None, name="circumplanetary_ring")
    return w


def asteroid_ensemble(n, seed=7, full_fraction=0.8):
    """Config 4: star + n asteroids with per-particle Yarkovsky parameters (units AU, yr, Msun;
    force-level values from reference ipython_examples/YarkovskyEffect.ipynb cell 6)."""
    rng = np.random.default_rng(seed)
    G = 4.0*np.pi**2
    a = rng.uniform(2.1, 3.3, n)
    e = rng.uniform(0.0, 0.3, n)
    inc = rng.uniform(0.0, 0.3, n)
    Om, om, f = (rng.uniform(0.0, TWO_PI, n) for _ in range(3))
    radius = 10.0**rng.uniform(-9.0, -7.0, n)
    m = rng.uniform(1e-18, 1e-15, n)
    w = _assemble(1.0, *kepler_to_cartesian(G*1.0, a, e, inc, Om, om, f), m=m, r=radius)
    s = rng.standard_normal((3, n))
    s /= np.linalg.norm(s, axis=0)
    u = rng.uniform(0.0, 1.0, n)
    flag = np.where(u < full_fraction, 0, np.where(u < full_fraction + (1 - full_fraction)/2, 1, -1)).astype(np.int32)
    w.update(G=G, c=63241.077, ye_lstar=2.7e-4, ye_c=63241.077, ye_stef_boltz=8.96e-16,
             ye_body_density=np.full(n, 5.05e6), ye_albedo=rng.uniform(0.01, 0.3, n),
             ye_emissivity=np.full(n, 0.9), ye_k=np.full(n, 0.25),
             ye_thermal_inertia=rng.uniform(5e-10, 1e-9, n),
             ye_rotation_period=10.0**rng.uniform(-4.0, -2.0, n),
             ye_spin_axis_x=s[0], ye_spin_axis_y=s[1], ye_spin_axis_z=s[2], ye_flag=flag,
             name="asteroid_ensemble")
    return w


def planetesimal_disk(n, seed=11):
    """Config 5: star + n low-mass planetesimals; modify_orbits_forces + gas_damping_timescale +
    type_I_migration with the inner-disk-edge planet trap (parameter values from the reference's
    three examples)."""
    rng = np.random.default_rng(seed)
    G = 4.0*np.pi**2
    a = rng.uniform(0.3, 3.0, n)
    e = rng.uniform(0.0, 0.05, n)
    inc = rng.uniform(0.0, 0.05, n)
    Om, om, f = (rng.uniform(0.0, TWO_PI, n) for _ in range(3))
    m = 10.0**rng.uniform(-10.0, -8.0, n)
    w = _assemble(1.0, *kepler_to_cartesian(G*1.0, a, e, inc, Om, om, f), m=m)
    w.update(G=G, tau_a=np.full(n, -1e6), tau_e=np.full(n, -1e4), tau_inc=np.full(n, -1e4),
             d_factor=np.full(n, 5.0), ide_position=0.3, ide_width=0.02, cs_coeff=0.272, tau_coeff=0.003,
             tIm_flaring_index=0.25, tIm_surface_density_exponent=1.0, tIm_surface_density_1=1.1255e-4,
             tIm_scale_height_1=0.03, name="planetesimal_disk")
    return w

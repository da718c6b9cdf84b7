"""Definitions of the golden cases (inputs only).  make_golden.py runs the reference-compiled
oracle (oracle/_ref) on each and stores inputs + outputs as <name>.npz; tests load the .npz
files, never this module's generators, so the fixtures pin themselves."""
import numpy as np

from reboundx_b200 import workloads as W


def _state(w):
    return {k: np.array(w[k], dtype=np.float64) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}


def _acc0(w, seed):
    return tuple(W.gravity_like_acc(w, np.random.default_rng(seed)))


def _idx(n, lo=1):
    return np.arange(lo, n, dtype=np.int64)


def rad_debris_disk_example():
    """Config 1a: shape of reference examples/rad_forces_debris_disk/problem.c:21-74
    (Sun + 1000 dust grains, a in [100,120] AU, e = 0.1, beta = 0.1, c = 3e8, WHFast)."""
    w = W.debris_disk(1000, seed=3)
    n = 1001
    return {"G": w["G"], "integrator": "whfast", "state": _state(w), "acc0": _acc0(w, 1),
            "forces": [{"name": "radiation_forces", "params": {"c": 3.0e8}}],
            "particle_params": {"beta": (_idx(n), np.full(n - 1, 0.1))}}


def rad_circumplanetary_example():
    """Config 1b: shape of reference examples/rad_forces_circumplanetary/problem.c:24-97
    (Sun + Saturn + 2 dust grains, IAS15, Sun flagged as radiation_source, beta from grain size)."""
    G, c = 6.674e-11, 3.0e8
    msun, msat, asat = 1.99e30, 5.68e26, 1.43e12
    vsat = np.sqrt(G*(msun + msat)/asat)
    rg = 3.0e8
    x = np.array([0.0, asat, asat + rg, asat - 0.7*rg])
    y = np.array([0.0, 0.0, 0.0, 1.0e7])
    z = np.array([0.0, 0.0, 2.0e6, -1.0e6])
    vorb = np.sqrt(G*msat/rg)
    st = {"x": x, "y": y, "z": z, "vx": np.array([0.0, 0.0, 0.0, 30.0]), "vy": np.array([0.0, vsat, vsat + vorb, vsat - 1.2*vorb]),
          "vz": np.array([0.0, 0.0, 10.0, -5.0]), "m": np.array([msun, msat, 0.0, 0.0]), "r": np.zeros(4)}
    L, rho, Qpr = 3.85e26, 1000.0, 1.0
    beta = [3.*L*Qpr/(16.*np.pi*G*msun*c*rho*s) for s in (1.0e-5, 1.0e-4)]
    return {"G": G, "integrator": "ias15", "state": st, "acc0": None,
            "forces": [{"name": "radiation_forces", "params": {"c": c}}],
            "particle_params": {"beta": (np.array([2, 3]), np.array(beta)), "radiation_source": (np.array([0]), np.array([1]))}}


def rad_presence_and_two_sources():
    """beta absent on some grains (skipped, radiation_forces.c:78); two flagged sources, one of
    them not index 0 (one sweep per source in index order, :108-114)."""
    w = W.debris_disk(36, seed=21)
    n = 37
    w["m"][5] = 0.5e30
    have = np.array([i for i in range(1, n) if i % 4 != 0], dtype=np.int64)
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 2),
            "forces": [{"name": "radiation_forces", "params": {"c": 3.0e8}}],
            "particle_params": {"beta": (have, np.linspace(0.05, 0.45, len(have))),
                                "radiation_source": (np.array([0, 5]), np.array([1, 1]))}}


def harmonics_presence_edges():
    """Two oblate bodies (index 0 tilted with J4; index 7 with J2 and R_eq only), a body with
    J2 == 0 and J4 set (skipped, gravitational_harmonics.c:145-147), a body with J2 but no R_eq
    (skipped, :150-152), a spin along -z (fac == 0 branch of uvw, :121-129); massive particles."""
    w = W.circumplanetary_ring(40, seed=22, massive=True)
    n = 41
    w["m"][7] = 3.0e22
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 3),
            "forces": [{"name": "gravitational_harmonics", "params": {}}],
            "particle_params": {"J2": (np.array([0, 7, 9, 11]), np.array([1.6e-2, 3.0e-3, 0.0, 2.0e-3])),
                                "J4": (np.array([0, 9]), np.array([-9.4e-4, 1.0e-4])),
                                "R_eq": (np.array([0, 7, 9]), np.array([6.03e7, 2.5e6, 1.0e6]))},
            "particle_vec3": {"Omega": (np.array([0, 7]), np.array([[0.3, -0.2, 0.9], [0.0, 0.0, -2.0]]))}}


def gr_potential_massive():
    w = W.circumplanetary_ring(32, seed=23, massive=True)
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 4),
            "forces": [{"name": "gr_potential", "params": {"c": 3.0e8}}]}


def stacked_rad_gr_harmonics():
    """Three stacked effects, executed in reverse add order (linkedlist.c:33-37)."""
    w = W.circumplanetary_ring(63, seed=24, massive=True)
    n = 64
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 5),
            "forces": [{"name": "radiation_forces", "params": {"c": 3.0e8}},
                       {"name": "gr_potential", "params": {"c": 3.0e8}},
                       {"name": "gravitational_harmonics", "params": {}}],
            "particle_params": {"beta": (_idx(n), np.linspace(0.01, 0.3, n - 1)),
                                "J2": (np.array([0]), np.array([1.6e-2])), "J4": (np.array([0]), np.array([-9.4e-4])),
                                "R_eq": (np.array([0]), np.array([6.03e7]))}}


def yarkovsky_all_modes():
    """ye_flag -1 / 0 / +1, a grain without density (skipped), a grain with r == 0 (skipped);
    stacked with gr_potential (config 4 shape)."""
    w = W.asteroid_ensemble(47, seed=25)
    n = 48
    w["r"][3] = 0.0
    pp = {p: (_idx(n), w[p]) for p in ("ye_rotation_period", "ye_thermal_inertia", "ye_albedo", "ye_emissivity", "ye_k",
                                       "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z")}
    keep = np.array([i for i in range(1, n) if i != 5], dtype=np.int64)
    pp["ye_body_density"] = (keep, w["ye_body_density"][keep - 1])
    flag = w["ye_flag"].copy()
    flag[:9] = [0, 1, -1, 0, 0, 1, -1, 0, 0]
    pp["ye_flag"] = (_idx(n), flag)
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 6),
            "forces": [{"name": "yarkovsky_effect", "params": {"ye_lstar": w["ye_lstar"], "ye_c": w["ye_c"], "ye_stef_boltz": w["ye_stef_boltz"]}},
                       {"name": "gr_potential", "params": {"c": w["c"]}}],
            "particle_params": pp}


def _trio_case(n_p, seed, coordinates, d_on_star=False):
    w = W.planetesimal_disk(n_p, seed=seed)
    n = n_p + 1
    # place three planetesimals around the inner disk edge so every branch of the planet trap
    # (inner_disk_edge.c:61-77) is exercised: inside the edge, in the transition band, outside
    G = w["G"]
    for i, a in ((1, 0.28), (2, 0.3001), (3, 0.2995), (4, 0.31)):
        v = np.sqrt(G*(1.0 + w["m"][i])/a)
        w["x"][i], w["y"][i], w["z"][i] = a, 0.0, 0.0
        w["vx"][i], w["vy"][i], w["vz"][i] = 0.0, v*np.cos(0.01*i), v*np.sin(0.01*i)
    lo = 0 if d_on_star else 1
    pp = {"tau_a": (_idx(n), w["tau_a"]), "tau_e": (_idx(n), w["tau_e"]), "tau_inc": (_idx(n), w["tau_inc"]),
          "d_factor": (_idx(n, lo), np.full(n - lo, 5.0))}
    if coordinates == 2:
        pp["primary"] = (np.array([0]), np.array([1]))
    common = {"coordinates": coordinates}
    forces = [{"name": "modify_orbits_forces", "params": dict(common, ide_position=0.3, ide_width=0.02)},
              {"name": "gas_damping_timescale", "params": dict(common, cs_coeff=0.272, tau_coeff=0.003)},
              {"name": "type_I_migration", "params": dict(common, ide_position=0.3, ide_width=0.02, tIm_flaring_index=0.25,
                                                          tIm_surface_density_exponent=1.0, tIm_surface_density_1=1.1255e-4,
                                                          tIm_scale_height_1=0.03)}]
    return {"G": G, "state": _state(w), "acc0": _acc0(w, seed), "forces": forces, "particle_params": pp}


def trio_particle():
    return _trio_case(39, 26, 2)


def trio_jacobi():
    return _trio_case(23, 27, 0)


def trio_barycentric():
    return _trio_case(23, 28, 1, d_on_star=True)


def modify_orbits_partial_params():
    """tau_a / tau_e / tau_inc each present on a different subset (absent => 1/tau_a = 0,
    tau_e = tau_inc = inf, modify_orbits_forces.c:78-112); no planet trap."""
    w = W.planetesimal_disk(30, seed=29)
    n = 31
    i = _idx(n)
    return {"G": w["G"], "state": _state(w), "acc0": _acc0(w, 7),
            "forces": [{"name": "modify_orbits_forces", "params": {"coordinates": 2}}],
            "particle_params": {"tau_a": (i[i % 2 == 0], np.full(int(np.sum(i % 2 == 0)), -1e5)),
                                "tau_e": (i[i % 3 == 0], np.full(int(np.sum(i % 3 == 0)), -1e3)),
                                "tau_inc": (i[i % 5 == 0], np.full(int(np.sum(i % 5 == 0)), -2e3)),
                                "primary": (np.array([0]), np.array([1]))}}


def single_particle():
    """N = 1: nothing but the source itself; every loop is empty."""
    st = {k: np.zeros(1) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}
    st["m"][0] = 1.0
    return {"G": 1.0, "state": st, "acc0": (np.array([0.5]), np.array([-0.25]), np.array([0.125])),
            "forces": [{"name": "radiation_forces", "params": {"c": 1e4}}, {"name": "gr_potential", "params": {"c": 1e4}}]}


ALL = [rad_debris_disk_example, rad_circumplanetary_example, rad_presence_and_two_sources, harmonics_presence_edges,
       gr_potential_massive, stacked_rad_gr_harmonics, yarkovsky_all_modes, trio_particle, trio_jacobi, trio_barycentric,
       modify_orbits_partial_params, single_particle]

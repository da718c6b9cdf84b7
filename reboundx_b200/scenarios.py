"""Turns a synthetic workload (workloads.py) into a live simulation with REBOUNDx forces
attached through the reference's own API calls (rebx_attach / rebx_load_force /
rebx_set_param_* / rebx_add_force), for whichever REBOUNDx-compatible library is given.
"""
import numpy as np

from .host import COORDINATES, Extras, Simulation

# per-particle double parameters each force reads (reference Appendix B of SURVEY.md)
_PARTICLE_DOUBLES = {
    "radiation_forces": ["beta"],
    "yarkovsky_effect": ["ye_body_density", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo",
                         "ye_emissivity", "ye_k", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z"],
    "modify_orbits_forces": ["tau_a", "tau_e", "tau_inc"],
    "gas_damping_timescale": ["d_factor"],
    "exponential_migration": ["em_tau_a", "em_aini", "em_afin"],
}
_FORCE_DOUBLES = {
    "radiation_forces": ["c"],
    "gr_potential": ["c"],
    "gr": ["c"],
    "lense_thirring": ["lt_c"],
    "gas_dynamical_friction": ["gas_df_rhog", "gas_df_alpha_rhog", "gas_df_cs", "gas_df_alpha_cs", "gas_df_xmin", "gas_df_hr", "gas_df_Qd"],
    "yarkovsky_effect": ["ye_lstar", "ye_c", "ye_stef_boltz"],
    "modify_orbits_forces": ["ide_position", "ide_width"],
    "gas_damping_timescale": ["cs_coeff", "tau_coeff"],
    "type_I_migration": ["ide_position", "ide_width", "tIm_flaring_index", "tIm_surface_density_exponent",
                         "tIm_surface_density_1", "tIm_scale_height_1"],
}
_COM_FORCES = ("modify_orbits_forces", "gas_damping_timescale", "type_I_migration", "exponential_migration")


def build(w, forces, lib=None, coordinates="PARTICLE", integrator="ias15", acc=None):
    """Returns (sim, rebx, {name: Force}).  `forces` are added in the given order, so they
    EXECUTE in reverse order (push-front list, reference linkedlist.c:33-37)."""
    sim = Simulation(G=w["G"], integrator=integrator)
    sim.set_particles(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["m"], w["r"])
    if acc is not None:
        sim.set_acc(*acc)
    rebx = Extras(sim, lib=lib)
    n = sim.N
    idx = np.arange(1, n, dtype=np.int64)
    done = set()
    handles = {}
    for name in forces:
        f = rebx.load_force(name)
        handles[name] = f
        for p in _FORCE_DOUBLES.get(name, []):
            if w.get(p) is not None:
                f.params[p] = w[p]
        if name in _COM_FORCES and coordinates is not None:
            f.params["coordinates"] = COORDINATES[coordinates]
            if coordinates == "PARTICLE" and "primary" not in done:
                rebx.particle_params(0)["primary"] = 1
                done.add("primary")
        for p in _PARTICLE_DOUBLES.get(name, []):
            if p in done or w.get(p) is None:
                continue
            rebx.set_particle_params(p, w[p], idx)
            done.add(p)
        if name == "yarkovsky_effect" and "ye_flag" not in done:
            rebx.set_particle_params("ye_flag", w["ye_flag"], idx)
            done.add("ye_flag")
        if name == "gravitational_harmonics" and "J2" not in done:
            p0 = rebx.particle_params(0)
            p0["J2"] = w["J2"]
            if w.get("J4") is not None:
                p0["J4"] = w["J4"]
            p0["R_eq"] = w["R_eq"]
            if w.get("Omega") is not None:
                p0["Omega"] = w["Omega"]
            done.add("J2")
        if name == "lense_thirring" and "I" not in done:
            p0 = rebx.particle_params(0)
            p0["I"] = w["I"]
            p0["Omega"] = w["Omega"]
            done.add("I")
        if name == "central_force" and "Acentral" not in done:
            p0 = rebx.particle_params(0)
            p0["Acentral"] = w["Acentral"]
            p0["gammacentral"] = w["gammacentral"]
            done.add("Acentral")
        rebx.add_force(f)
    if w.get("t") is not None:
        sim.t = w["t"]
    return sim, rebx, handles


# ---- general cases (arbitrary presence patterns), used by the golden fixtures ----------------------
def build_case(case, lib=None):
    """case = {"G", "integrator"?, "state": {x..r}, "acc0": (ax, ay, az) | None,
               "forces": [{"name", "params": {p: value}}, ...]   (ADD order),
               "particle_params": {p: (index array, value array)},   double or int by registry
               "particle_vec3": {p: (index array, [k,3] array)}}"""
    st = case["state"]
    sim = Simulation(G=case["G"], integrator=case.get("integrator", "ias15"))
    sim.set_particles(st["x"], st["y"], st["z"], st["vx"], st["vy"], st["vz"], st["m"], st["r"])
    if case.get("acc0") is not None:
        sim.set_acc(*case["acc0"])
    rebx = Extras(sim, lib=lib)
    for p, (idx, val) in case.get("particle_params", {}).items():
        if len(idx):
            rebx.set_particle_params(p, val, idx)
    for p, (idx, val) in case.get("particle_vec3", {}).items():
        for i, v in zip(idx, val):
            rebx.particle_params(int(i))[p] = tuple(float(c) for c in v)
    handles = []
    for f in case["forces"]:
        h = rebx.load_force(f["name"])
        for p, v in f.get("params", {}).items():
            h.params[p] = v
        rebx.add_force(h)
        handles.append(h)
    return sim, rebx, handles


def run_case(case, lib=None):
    """One dispatch; returns ((ax, ay, az), [(kind, message), ...])."""
    sim, rebx, _ = build_case(case, lib)
    sim.additional_forces()
    msgs = sim.messages()
    out = sim.acc()
    rebx.detach()
    sim.free()
    return out, msgs

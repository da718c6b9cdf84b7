"""Python front end of the oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Two oracles are offered on identical inputs:

* ``ref_apply``  -- kind "reference": the reference's own C sources, compiled in place into
  oracle/_ref/libreboundx_ref.so (oracle/Makefile), driven through the reference's public
  API by the same harness that drives the product (reboundx_b200/scenarios.py).
* ``port_apply`` -- kind "port": the plain-C restatement in oracle/oracle_forces.c.

Parity status: radiation_forces, gravitational_harmonics, gr_potential and
modify_orbits_forces without the planet trap use no third-party arithmetic, so the oracle is
the reference itself.  The other effects call REBOUND's reb_orbit_from_particle, restated in
reboundx_b200/rebound_shim (REBOUND is not under /root/reference): "parity unpinned" at that
boundary -- oracle and product share the restatement.
"""
import ctypes
import os
from ctypes import POINTER, c_double, c_int, c_int64, c_longdouble, c_ubyte, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libreboundx_ref.so")
PORT_PATH = os.path.join(_HERE, "liboracle_forces.so")

YARK_PARAMS = ["ye_body_density", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo", "ye_emissivity",
               "ye_k", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z"]
GAS_DF_PARAMS = ["gas_df_rhog", "gas_df_alpha_rhog", "gas_df_cs", "gas_df_alpha_cs", "gas_df_xmin", "gas_df_hr", "gas_df_Qd"]
KIND = {"modify_orbits_forces": 5, "gas_damping_timescale": 6, "type_I_migration": 7, "exponential_migration": 9}
COORD = {"JACOBI": 0, "BARYCENTRIC": 1, "PARTICLE": 2}


def have_ref():
    return os.path.exists(REF_PATH)


def ref_lib():
    from reboundx_b200 import _lib
    return _lib.load(REF_PATH)


def ref_apply(w, forces, acc0=None, coordinates="PARTICLE", reps=0, messages=None, n_active=None):
    """Runs the compiled reference on workload `w`; `forces` in ADD order (they execute reversed).
    Returns (ax, ay, az) or, with reps > 0, ((ax, ay, az), best seconds per dispatch)."""
    from reboundx_b200 import scenarios
    sim, rebx, _ = scenarios.build(w, forces, lib=ref_lib(), coordinates=coordinates, acc=acc0)
    if n_active is not None:
        sim.N_active = n_active
    if reps > 0:
        sim.additional_forces()                       # warm-up
        if acc0 is not None:
            sim.set_acc(*acc0)
        else:
            sim.set_acc()
        sim.additional_forces()
        out = sim.acc()
        t = sim.time_additional_forces(reps)
        sim.messages()
        rebx.detach()
        sim.free()
        return out, t
    sim.additional_forces()
    msgs = sim.messages()
    if messages is not None:
        messages.extend(msgs)
    out = sim.acc()
    rebx.detach()
    sim.free()
    return out


class _DampingParams(ctypes.Structure):
    _fields_ = [("tau_a", POINTER(c_double)), ("tau_e", POINTER(c_double)), ("tau_inc", POINTER(c_double)),
                ("has_tau_a", POINTER(c_ubyte)), ("has_tau_e", POINTER(c_ubyte)), ("has_tau_inc", POINTER(c_ubyte)),
                ("trap", c_int), ("dedge", c_double), ("hedge", c_double),
                ("d_factor", POINTER(c_double)), ("has_d_factor", POINTER(c_ubyte)),
                ("cs_coeff", c_double), ("tau_coeff", c_double),
                ("tIm_beta", c_double), ("tIm_h0", c_double), ("tIm_sd0", c_double), ("tIm_s", c_double),
                ("tIm_dedge", c_double), ("tIm_hedge", c_double),
                ("em_tau_a", POINTER(c_double)), ("em_aini", POINTER(c_double)), ("em_afin", POINTER(c_double)),
                ("has_em_tau_a", POINTER(c_ubyte)), ("has_em_aini", POINTER(c_ubyte)), ("has_em_afin", POINTER(c_ubyte)),
                ("t", c_double)]


def _port():
    lib = ctypes.CDLL(PORT_PATH)
    return lib


def _d(a):
    return a.ctypes.data_as(POINTER(c_double))


def _full(w, name, n):
    """Per-particle array of length n (index 0 = central body gets 0) + presence mask."""
    v = w.get(name)
    if v is None:
        return None, None
    full = np.zeros(n)
    full[1:] = v
    mask = np.ones(n, dtype=np.uint8)
    mask[0] = 0
    return full, mask


def port_apply(w, exec_order, acc0=None, coordinates="PARTICLE", n_active=-1, com="quadratic"):
    """Runs the C restatement on workload `w`; `exec_order` is the EXECUTION order of forces.
    com: "quadratic" = rebx_com_force's loops as the reference has them (O(n^2) in the centre-of-mass
    frames); "linear" = oracle_com_force_linear (the reference's COM recurrence, back-reaction summed in
    long double, O(n)); "linear_exact" = the same with exact prefix sums for COM_i (diagnostics only).
    Returns ((ax, ay, az), {force: [sum_x, sum_y, sum_z, abs_x, abs_y, abs_z]}): the back-reaction
    terms onto the body summed in long double, and the sum of their absolute values."""
    lib = _port()
    n = w["x"].shape[0]
    S = {k: np.ascontiguousarray(w[k], dtype=np.float64) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}
    acc = [np.zeros(n) if acc0 is None else np.array(a, dtype=np.float64) for a in (acc0 or (None,)*3)]
    br = {}
    keep = []
    for name in exec_order:
        out = (c_longdouble*6)()
        if name == "radiation_forces":
            beta, mask = _full(w, "beta", n)
            keep += [beta, mask]
            lib.oracle_radiation_forces(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                                        _d(S["m"]), _d(beta), mask.ctypes.data_as(POINTER(c_ubyte)),
                                        c_double(w["G"]), c_double(w["c"]), None, c_int64(0), _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name == "gravitational_harmonics":
            Om = (c_double*3)(*(w.get("Omega") or (0.0, 0.0, 1.0)))
            lib.oracle_gravitational_harmonics(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["m"]), c_int64(0),
                                               c_double(w["J2"]), c_double(w.get("J4") or 0.0), c_double(w["R_eq"]), Om,
                                               c_double(w["G"]), _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
            br[name] = np.array([float(v) for v in out])
        elif name == "gr_potential":
            lib.oracle_gr_potential(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["m"]), c_double(w["G"]), c_double(w["c"]),
                                    _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
            br[name] = np.array([float(v) for v in out])
        elif name == "central_force":
            lib.oracle_central_force(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["m"]), c_int64(0),
                                     c_double(w["Acentral"]), c_double(w["gammacentral"]), _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
            br[name] = np.array([float(v) for v in out])
        elif name == "gas_dynamical_friction":
            par = (c_double*7)(*[w[k] for k in GAS_DF_PARAMS])
            lib.oracle_gas_dynamical_friction(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                                              _d(S["m"]), _d(S["r"]), c_double(w["G"]), par, _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name == "lense_thirring":
            Om = (c_double*3)(*w["Omega"])
            lib.oracle_lense_thirring(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]), _d(S["m"]),
                                      c_double(w["G"]), c_double(w["lt_c"]), c_double(w["I"]), Om, _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
            br[name] = np.array([float(v) for v in out])
        elif name == "gr":
            lib.oracle_gr(c_int64(n), c_int64(n_active), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                          _d(S["m"]), c_double(w["G"]), c_double(w["c"]), c_int(w.get("max_iterations") or 10),
                          _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name == "yarkovsky_effect":
            pars, masks = [], []
            for p in YARK_PARAMS:
                full, mask = _full(w, p, n)
                if full is None:
                    full, mask = np.zeros(n), np.zeros(n, dtype=np.uint8)
                pars.append(full)
                masks.append(mask)
            flag = np.zeros(n, dtype=np.int32)
            flag[1:] = w["ye_flag"]
            fmask = np.ones(n, dtype=np.uint8)
            fmask[0] = 0
            keep += pars + masks + [flag, fmask]
            P = (POINTER(c_double)*9)(*[_d(a) for a in pars])
            M = (POINTER(c_ubyte)*9)(*[a.ctypes.data_as(POINTER(c_ubyte)) for a in masks])
            sb = w.get("ye_stef_boltz")
            lib.oracle_yarkovsky_effect(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                                        _d(S["m"]), _d(S["r"]), P, M, flag.ctypes.data_as(POINTER(c_int)),
                                        fmask.ctypes.data_as(POINTER(c_ubyte)), c_double(w["G"]), c_double(w["ye_lstar"]),
                                        c_double(w["ye_c"]), c_double(sb or 0.0), c_int(0 if sb is None else 1),
                                        _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name in KIND:
            dp = _DampingParams()
            for p in ("tau_a", "tau_e", "tau_inc", "d_factor", "em_tau_a", "em_aini", "em_afin"):
                full, mask = _full(w, p, n)
                if full is not None:
                    keep += [full, mask]
                    setattr(dp, p, _d(full))
                    setattr(dp, "has_" + p, mask.ctypes.data_as(POINTER(c_ubyte)))
            dp.t = w.get("t") or 0.0
            if name == "modify_orbits_forces":
                dp.trap = int(w.get("ide_position") is not None and w.get("ide_width") is not None)
                dp.dedge, dp.hedge = w.get("ide_position") or 0.0, w.get("ide_width") or 0.0
            dp.cs_coeff, dp.tau_coeff = w.get("cs_coeff") or 0.0, w.get("tau_coeff") or 0.0
            dp.tIm_beta = w.get("tIm_flaring_index") or 0.0
            dp.tIm_h0 = 0.01 if w.get("tIm_scale_height_1") is None else w["tIm_scale_height_1"]
            dp.tIm_sd0 = w.get("tIm_surface_density_1") or 0.0
            dp.tIm_s = w.get("tIm_surface_density_exponent") or 0.0
            dp.tIm_dedge, dp.tIm_hedge = w.get("ide_position") or 0.0, w.get("ide_width") or 0.0
            if com == "quadratic":
                lib.oracle_com_force.restype = c_int
                rc = lib.oracle_com_force(c_int(KIND[name]), c_int(COORD[coordinates]), c_int64(0), ctypes.byref(dp), c_double(w["G"]),
                                          c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]), _d(S["m"]),
                                          _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
                br[name] = np.array([float(v) for v in out])
            else:
                lib.oracle_com_force_linear.restype = c_int
                rc = lib.oracle_com_force_linear(c_int(KIND[name]), c_int(COORD[coordinates]), c_int64(0), ctypes.byref(dp), c_double(w["G"]),
                                                 c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]), _d(S["m"]),
                                                 _d(acc[0]), _d(acc[1]), _d(acc[2]), c_int(1 if com == "linear_exact" else 0))
            assert rc == 0
        else:
            raise ValueError(name)
    return acc, br


# ---- general cases (reboundx_b200.scenarios.build_case format) -------------------------------------
def _pp(case, name, n, dtype=np.float64):
    """(values[n], present[n]) of a per-particle parameter, or (None, zeros) when nobody has it."""
    ent = case.get("particle_params", {}).get(name)
    vals = np.zeros(n, dtype=dtype)
    mask = np.zeros(n, dtype=np.uint8)
    if ent is not None and len(ent[0]):
        idx = np.asarray(ent[0], dtype=np.int64)
        vals[idx] = np.asarray(ent[1], dtype=dtype)
        mask[idx] = 1
    return vals, mask


def port_case(case):
    """Runs the C restatement on a general case; returns (ax, ay, az).  Forces execute in
    reverse ADD order like the reference's push-front list."""
    lib = _port()
    lib.oracle_com_force.restype = c_int
    st = case["state"]
    n = st["x"].shape[0]
    S = {k: np.ascontiguousarray(st[k], dtype=np.float64) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}
    acc = [np.zeros(n) if case.get("acc0") is None else np.array(a, dtype=np.float64) for a in (case.get("acc0") or (None,)*3)]
    G = case["G"]
    ub = lambda a: a.ctypes.data_as(POINTER(c_ubyte))
    for f in case["forces"][::-1]:
        name, fp = f["name"], f.get("params", {})
        if name == "radiation_forces":
            if "c" not in fp:
                continue
            beta, mask = _pp(case, "beta", n)
            _, smask = _pp(case, "radiation_source", n, np.int32)
            src = np.nonzero(smask)[0].astype(np.int64)
            lib.oracle_radiation_forces(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                                        _d(S["m"]), _d(beta), ub(mask), c_double(G), c_double(fp["c"]),
                                        src.ctypes.data_as(POINTER(c_int64)) if len(src) else None, c_int64(len(src)),
                                        _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name == "gravitational_harmonics":
            J2, hJ2 = _pp(case, "J2", n)
            J4, hJ4 = _pp(case, "J4", n)
            Rq, hRq = _pp(case, "R_eq", n)
            om = case.get("particle_vec3", {}).get("Omega", ([], []))
            omega = {int(i): v for i, v in zip(om[0], om[1])}
            for i in range(n):
                if not hJ2[i] or J2[i] == 0.0 or not hRq[i]:
                    continue
                Om = (c_double*3)(*[float(c) for c in omega.get(i, (0.0, 0.0, 1.0))])
                lib.oracle_gravitational_harmonics(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["m"]), c_int64(i),
                                                   c_double(J2[i]), c_double(J4[i] if hJ4[i] else 0.0), c_double(Rq[i]), Om,
                                                   c_double(G), _d(acc[0]), _d(acc[1]), _d(acc[2]), None)
        elif name == "gr_potential":
            if "c" not in fp:
                continue
            lib.oracle_gr_potential(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["m"]), c_double(G), c_double(fp["c"]),
                                    _d(acc[0]), _d(acc[1]), _d(acc[2]), None)
        elif name == "yarkovsky_effect":
            if "ye_lstar" not in fp or "ye_c" not in fp:
                continue
            pars, masks = zip(*[_pp(case, p, n) for p in YARK_PARAMS])
            flag, fmask = _pp(case, "ye_flag", n, np.int32)
            P = (POINTER(c_double)*9)(*[_d(a) for a in pars])
            M = (POINTER(c_ubyte)*9)(*[ub(a) for a in masks])
            sb = fp.get("ye_stef_boltz")
            lib.oracle_yarkovsky_effect(c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]),
                                        _d(S["m"]), _d(S["r"]), P, M, flag.ctypes.data_as(POINTER(c_int)), ub(fmask),
                                        c_double(G), c_double(fp["ye_lstar"]), c_double(fp["ye_c"]), c_double(sb or 0.0),
                                        c_int(0 if sb is None else 1), _d(acc[0]), _d(acc[1]), _d(acc[2]))
        elif name in KIND:
            dp = _DampingParams()
            keep = []
            dp.t = case.get("t", 0.0)
            for p in ("tau_a", "tau_e", "tau_inc", "d_factor", "em_tau_a", "em_aini", "em_afin"):
                vals, mask = _pp(case, p, n)
                keep += [vals, mask]
                setattr(dp, p, _d(vals))
                setattr(dp, "has_" + p, ub(mask))
            coords = fp.get("coordinates", 0)
            _, pmask = _pp(case, "primary", n, np.int32)
            prim = np.nonzero(pmask)[0]
            primary = int(prim[0]) if len(prim) else -1
            if name == "modify_orbits_forces":
                dp.trap = int("ide_position" in fp and "ide_width" in fp)
                dp.dedge, dp.hedge = fp.get("ide_position", 0.0), fp.get("ide_width", 0.0)
            if name == "gas_damping_timescale" and ("cs_coeff" not in fp or "tau_coeff" not in fp):
                continue
            dp.cs_coeff, dp.tau_coeff = fp.get("cs_coeff", 0.0), fp.get("tau_coeff", 0.0)
            dp.tIm_beta, dp.tIm_h0 = fp.get("tIm_flaring_index", 0.0), fp.get("tIm_scale_height_1", 0.01)
            dp.tIm_sd0, dp.tIm_s = fp.get("tIm_surface_density_1", 0.0), fp.get("tIm_surface_density_exponent", 0.0)
            dp.tIm_dedge, dp.tIm_hedge = fp.get("ide_position", 0.0), fp.get("ide_width", 0.0)
            lib.oracle_com_force(c_int(KIND[name]), c_int(coords), c_int64(primary), ctypes.byref(dp), c_double(G),
                                 c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]), _d(S["m"]),
                                 _d(acc[0]), _d(acc[1]), _d(acc[2]), None)
        else:
            raise ValueError(name)
    return acc


# ---- the reference's integrate_force operator (src/integrate_force.c + src/integrator_*.c) ----------
INTEGRATORS = {"implicit_midpoint": 0, "rk4": 1, "euler": 2, "rk2": 3}       # reference reboundx.h enum rebx_integrator


def ref_integrate_force(w, force_name, dt, scheme, nsteps=1, lib=None, coordinates="PARTICLE"):
    """Velocities after `nsteps` applications of the reference's rebx_integrate_force(sim, operator, dt)
    with the reference's own force loop, all compiled in oracle/_ref.  Returns (vx, vy, vz).
    `lib`: another library with the same operator API (the tests drive the product through the same calls)."""
    from ctypes import c_char_p
    from reboundx_b200 import scenarios
    from reboundx_b200.host import RebxForce
    lib = lib or ref_lib()
    sim, rebx, handles = scenarios.build(w, [force_name], lib=lib, coordinates=coordinates)
    lib.rebx_load_operator.restype = POINTER(RebxForce)          # rebx_operator has the same leading layout
    op = lib.rebx_load_operator(rebx._ptr, c_char_p(b"integrate_force"))
    ap = ctypes.cast(ctypes.addressof(op.contents) + RebxForce.ap.offset, c_void_p)
    lib.rebx_set_param_pointer(rebx._ptr, ap, c_char_p(b"force"), handles[force_name]._ptr)
    lib.rebx_set_param_int(rebx._ptr, ap, c_char_p(b"integrator"), c_int(INTEGRATORS[scheme]))
    for _ in range(nsteps):
        lib.rebx_integrate_force(sim.ptr, op, c_double(dt))
    msgs = sim.messages()
    out = sim.posvel()[3:]
    rebx.detach()
    sim.free()
    return out, msgs

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
KIND = {"modify_orbits_forces": 5, "gas_damping_timescale": 6, "type_I_migration": 7}
COORD = {"JACOBI": 0, "BARYCENTRIC": 1, "PARTICLE": 2}


def have_ref():
    return os.path.exists(REF_PATH)


def ref_lib():
    from reboundx_b200 import _lib
    return _lib.load(REF_PATH)


def ref_apply(w, forces, acc0=None, coordinates="PARTICLE", reps=0, messages=None):
    """Runs the compiled reference on workload `w`; `forces` in ADD order (they execute reversed).
    Returns (ax, ay, az) or, with reps > 0, ((ax, ay, az), best seconds per dispatch)."""
    from reboundx_b200 import scenarios
    sim, rebx, _ = scenarios.build(w, forces, lib=ref_lib(), coordinates=coordinates, acc=acc0)
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
                ("tIm_dedge", c_double), ("tIm_hedge", c_double)]


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


def port_apply(w, exec_order, acc0=None, coordinates="PARTICLE"):
    """Runs the C restatement on workload `w`; `exec_order` is the EXECUTION order of forces.
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
            for p in ("tau_a", "tau_e", "tau_inc", "d_factor"):
                full, mask = _full(w, p, n)
                if full is not None:
                    keep += [full, mask]
                    setattr(dp, p, _d(full))
                    setattr(dp, "has_" + p, mask.ctypes.data_as(POINTER(c_ubyte)))
            if name == "modify_orbits_forces":
                dp.trap = int(w.get("ide_position") is not None and w.get("ide_width") is not None)
                dp.dedge, dp.hedge = w.get("ide_position") or 0.0, w.get("ide_width") or 0.0
            dp.cs_coeff, dp.tau_coeff = w.get("cs_coeff") or 0.0, w.get("tau_coeff") or 0.0
            dp.tIm_beta = w.get("tIm_flaring_index") or 0.0
            dp.tIm_h0 = 0.01 if w.get("tIm_scale_height_1") is None else w["tIm_scale_height_1"]
            dp.tIm_sd0 = w.get("tIm_surface_density_1") or 0.0
            dp.tIm_s = w.get("tIm_surface_density_exponent") or 0.0
            dp.tIm_dedge, dp.tIm_hedge = w.get("ide_position") or 0.0, w.get("ide_width") or 0.0
            lib.oracle_com_force.restype = c_int
            rc = lib.oracle_com_force(c_int(KIND[name]), c_int(COORD[coordinates]), c_int64(0), ctypes.byref(dp), c_double(w["G"]),
                                      c_int64(n), _d(S["x"]), _d(S["y"]), _d(S["z"]), _d(S["vx"]), _d(S["vy"]), _d(S["vz"]), _d(S["m"]),
                                      _d(acc[0]), _d(acc[1]), _d(acc[2]), out)
            assert rc == 0
            br[name] = np.array([float(v) for v in out])
        else:
            raise ValueError(name)
    return acc, br

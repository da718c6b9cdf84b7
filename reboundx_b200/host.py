"""Host-side mirror of the REBOUNDx Python interface for the force path.

The reference wraps its C library with ctypes (reference reboundx/extras.py, params.py)
and leans on the ``rebound`` package for the simulation object.  ``rebound`` is not
available here, so this module provides

* ``Simulation``: a thin handle on the stand-in host (reboundx_b200/rebound_shim), with
  numpy array I/O for particle state and REBOUND's error/warning queue semantics
  (queued errors are raised as ``RuntimeError`` like reference extras.py:239-243);
* ``Extras`` / ``Force`` / ``Params``: same method names and meaning as the reference's
  classes (``load_force``, ``add_force``, ``get_force``, ``remove_force``,
  ``register_param``, ``force.params["c"] = ...``, ``particle params``), bound to the same
  C symbols (``rebx_*``).

Every class takes the ctypes library to bind, so the same code drives this package's CUDA
library and -- in tests only -- the reference's own C sources compiled as the oracle.
"""
import ctypes
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int64, c_size_t,
                    c_uint32, c_uint64, c_void_p, cast)

import numpy as np

from . import _lib

REBX_TYPES = {"none": 0, "double": 1, "int": 2, "pointer": 3, "force": 4, "uint32": 5,
              "orbit": 6, "ode": 7, "vec3d": 8, "string": 9}          # reference reboundx.h:61-72
REBX_FORCE_TYPE = {"none": 0, "pos": 1, "vel": 2}                        # reference extras.py:10
COORDINATES = {"JACOBI": 0, "BARYCENTRIC": 1, "PARTICLE": 2}             # reference tools.py:4


class Vec3d(Structure):
    _fields_ = [("x", c_double), ("y", c_double), ("z", c_double)]


class RebParticle(Structure):
    _fields_ = [("x", c_double), ("y", c_double), ("z", c_double),
                ("vx", c_double), ("vy", c_double), ("vz", c_double),
                ("ax", c_double), ("ay", c_double), ("az", c_double),
                ("m", c_double), ("r", c_double), ("last_collision", c_double),
                ("c", c_void_p), ("name", c_char_p), ("ap", c_void_p), ("sim", c_void_p)]


class _Integrator(Structure):
    _fields_ = [("name", c_char_p), ("state", c_void_p)]


class RebSimulation(Structure):
    # leading fields of the stand-in struct reb_simulation (rebound_shim/rebound.h)
    _fields_ = [("t", c_double), ("G", c_double), ("softening", c_double), ("dt", c_double),
                ("dt_last_done", c_double), ("N", c_size_t), ("N_var", c_size_t), ("N_active", c_size_t),
                ("particles", POINTER(RebParticle)), ("force_is_velocity_dependent", c_int),
                ("integrator", _Integrator), ("extras", c_void_p),
                ("additional_forces", c_void_p), ("pre_timestep_modifications", c_void_p),
                ("post_timestep_modifications", c_void_p), ("free_particle_ap", c_void_p),
                ("extras_cleanup", c_void_p)]


class RebxNode(Structure):
    pass


RebxNode._fields_ = [("object", c_void_p), ("next", POINTER(RebxNode))]     # reference extras.py:257-258


class RebxParam(Structure):
    _fields_ = [("name", c_char_p), ("type", c_int), ("value", c_void_p)]    # reference extras.py:251-253


class RebxForce(Structure):
    _fields_ = [("name", c_char_p), ("ap", POINTER(RebxNode)), ("sim", c_void_p), ("force_type", c_int),
                ("update_accelerations", c_void_p), ("free_memory", c_void_p)]   # reference extras.py:319-324


class RebxExtras(Structure):
    _fields_ = [("sim", c_void_p), ("additional_forces", POINTER(RebxNode)),
                ("pre_timestep_modifications", POINTER(RebxNode)), ("post_timestep_modifications", POINTER(RebxNode)),
                ("registered_params", POINTER(RebxNode)), ("allocated_forces", POINTER(RebxNode)),
                ("allocated_operators", POINTER(RebxNode))]                   # reference extras.py:327-333


FORCE_FN = ctypes.CFUNCTYPE(None, c_void_p, POINTER(RebxForce), POINTER(RebParticle), c_int)   # reference extras.py:316


def _dptr(a):
    return None if a is None else a.ctypes.data_as(POINTER(c_double))


def _f64(a, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected {n} values, got shape {a.shape}")
    return a


class Simulation:
    """Handle on a stand-in `struct reb_simulation`."""

    def __init__(self, G=1.0, integrator="ias15"):
        self._shim = shim = _lib.load_shim()
        shim.reb_simulation_create.restype = POINTER(RebSimulation)
        shim.reb_shim_message.restype = c_char_p
        shim.reb_shim_message.argtypes = [c_void_p, c_size_t]
        shim.reb_shim_message_count.restype = c_size_t
        shim.reb_shim_particle_ap.restype = POINTER(POINTER(RebxNode))
        shim.reb_shim_particle_ap.argtypes = [c_void_p, c_size_t]
        shim.reb_shim_time_additional_forces.restype = c_double
        self._p = shim.reb_simulation_create()
        self._extras_ref = None
        self.G = G
        self.integrator = integrator

    # --- plain fields
    @property
    def ptr(self):
        return self._p

    @property
    def struct(self):
        return self._p.contents

    G = property(lambda self: self._p.contents.G, lambda self, v: setattr(self._p.contents, "G", float(v)))
    t = property(lambda self: self._p.contents.t, lambda self, v: setattr(self._p.contents, "t", float(v)))
    dt = property(lambda self: self._p.contents.dt, lambda self, v: setattr(self._p.contents, "dt", float(v)))
    N = property(lambda self: int(self._p.contents.N))
    N_var = property(lambda self: int(self._p.contents.N_var), lambda self, v: setattr(self._p.contents, "N_var", int(v)))
    N_active = property(lambda self: int(self._p.contents.N_active), lambda self, v: setattr(self._p.contents, "N_active", int(v)))
    force_is_velocity_dependent = property(lambda self: int(self._p.contents.force_is_velocity_dependent))

    @property
    def integrator(self):
        return self._p.contents.integrator.name.decode()

    @integrator.setter
    def integrator(self, name):
        self._shim.reb_shim_set_integrator(self._p, c_char_p(name.encode()))

    # --- particle state as arrays
    def set_particles(self, x, y, z, vx=None, vy=None, vz=None, m=None, r=None):
        x = _f64(x)
        n = x.shape[0]
        arrs = [x] + [_f64(a, n) for a in (y, z, vx, vy, vz, m, r)]
        self._shim.reb_shim_set_particles(self._p, c_size_t(n), *[_dptr(a) for a in arrs])

    def set_acc(self, ax=None, ay=None, az=None):
        n = self.N
        arrs = [_f64(a, n) for a in (ax, ay, az)]
        self._shim.reb_shim_set_acc(self._p, *[_dptr(a) for a in arrs])

    def acc(self):
        n = self.N
        out = [np.empty(n) for _ in range(3)]
        self._shim.reb_shim_get_acc(self._p, *[_dptr(a) for a in out])
        return out

    def posvel(self):
        n = self.N
        out = [np.empty(n) for _ in range(6)]
        self._shim.reb_shim_get_posvel(self._p, *[_dptr(a) for a in out])
        return out

    def particle(self, i):
        if not 0 <= i < self.N:
            raise IndexError(i)
        return self._p.contents.particles[i]

    def particle_ap(self, i):
        return self._shim.reb_shim_particle_ap(self._p, c_size_t(i))

    def add_particle(self, **fields):
        """reb_simulation_add of one particle (x, y, z, vx, vy, vz, m, r)."""
        self._shim.reb_simulation_add.argtypes = [c_void_p, RebParticle]
        self._shim.reb_simulation_add(self._p, RebParticle(**{k: float(v) for k, v in fields.items()}))

    def remove_particle(self, i, keep_sorted=True):
        self._shim.reb_simulation_remove_particle(self._p, c_size_t(i), c_int(1 if keep_sorted else 0))

    # --- messages
    def messages(self, clear=True):
        n = self._shim.reb_shim_message_count(self._p)
        out = []
        for i in range(n):
            s = self._shim.reb_shim_message(self._p, i).decode(errors="replace")
            out.append(("error" if s[0] == "E" else "warning", s[1:]))
        if clear:
            self._shim.reb_shim_clear_messages(self._p)
        return out

    def process_messages(self):
        """Raise queued errors like the reference's Extras.process_messages (extras.py:239-243)."""
        msgs = self.messages()
        errs = [m for k, m in msgs if k == "error"]
        if errs:
            raise RuntimeError(errs[0])
        return [m for k, m in msgs if k == "warning"]

    # --- the hot path entry REBOUND would call
    def additional_forces(self):
        return bool(self._shim.reb_shim_call_additional_forces(self._p))

    def time_additional_forces(self, reps=1):
        return float(self._shim.reb_shim_time_additional_forces(self._p, c_int(reps)))

    def integrate(self, nsteps, scheme="leapfrog"):
        """Host integrator harness of the stand-in (NOT REBOUND's WHFast/IAS15): `nsteps` fixed steps of
        `dt` with Newtonian gravity from the first N_active bodies + additional_forces.  Schemes: leapfrog,
        rk4, gauss_radau (15th-order, the scheme behind IAS15 without step control), wisdom_holman (Kepler
        drifts about particle 0 + kicks by additional_forces, the splitting behind WHFast)."""
        self._shim.reb_shim_integrate(self._p, c_int(nsteps), c_int({"leapfrog": 0, "rk4": 1, "gauss_radau": 2, "wisdom_holman": 3, "wisdom_holman_merged": 4}[scheme]))

    def free(self):
        if self._p is not None:
            self._shim.reb_simulation_free(self._p)
            self._p = None

    def __del__(self):
        try:
            if self._extras_ref is not None:
                self._extras_ref.detach()
            self.free()
        except Exception:
            pass


class Params:
    """Dictionary view of one object's parameter list (reference reboundx/params.py:12-100)."""

    def __init__(self, rebx, apptr):
        self._rebx = rebx
        self._apptr = apptr        # struct rebx_node**

    def _param(self, key):
        lib = self._rebx._lib
        p = lib.rebx_get_param_struct(self._rebx._ptr, self._apptr.contents, c_char_p(key.encode()))
        return p.contents if p else None

    def __contains__(self, key):
        return self._param(key) is not None

    def __getitem__(self, key):
        p = self._param(key)
        if p is None:
            raise AttributeError(f"REBOUNDx Error: Parameter '{key}' not found.")
        t = p.type
        if t == REBX_TYPES["double"]:
            return cast(p.value, POINTER(c_double)).contents.value
        if t == REBX_TYPES["int"]:
            return cast(p.value, POINTER(c_int)).contents.value
        if t == REBX_TYPES["uint32"]:
            return cast(p.value, POINTER(c_uint32)).contents.value
        if t == REBX_TYPES["vec3d"]:
            v = cast(p.value, POINTER(Vec3d)).contents
            return (v.x, v.y, v.z)
        return p.value

    def __setitem__(self, key, value):
        rebx, lib = self._rebx, self._rebx._lib
        t = lib.rebx_get_type(rebx._ptr, c_char_p(key.encode()))
        k = c_char_p(key.encode())
        if t == REBX_TYPES["none"]:
            raise AttributeError(f"REBOUNDx Error: Parameter '{key}' not registered. See registering parameters in the docs.")
        if t == REBX_TYPES["double"]:
            lib.rebx_set_param_double(rebx._ptr, self._apptr, k, c_double(value))
        elif t == REBX_TYPES["int"]:
            lib.rebx_set_param_int(rebx._ptr, self._apptr, k, c_int(int(value)))
        elif t == REBX_TYPES["uint32"]:
            lib.rebx_set_param_uint32(rebx._ptr, self._apptr, k, c_uint32(int(value)))
        elif t == REBX_TYPES["vec3d"]:
            lib.rebx_set_param_vec3d(rebx._ptr, self._apptr, k, Vec3d(*[float(c) for c in value]))
        else:
            raise TypeError(f"parameter type {t} is outside the force path")
        rebx.process_messages()


class Force:
    def __init__(self, rebx, ptr):
        self._rebx = rebx
        self._ptr = ptr            # POINTER(RebxForce)

    @property
    def name(self):
        return self._ptr.contents.name.decode()

    @property
    def force_type(self):
        return {v: k for k, v in REBX_FORCE_TYPE.items()}[self._ptr.contents.force_type]

    @force_type.setter
    def force_type(self, value):
        self._ptr.contents.force_type = REBX_FORCE_TYPE[value.lower()]

    @property
    def update_accelerations(self):
        return self._ptr.contents.update_accelerations

    @update_accelerations.setter
    def update_accelerations(self, func):
        self._cb = FORCE_FN(func)          # keep alive (reference extras.py:305-317)
        self._ptr.contents.update_accelerations = cast(self._cb, c_void_p)

    @property
    def params(self):
        ap_field = ctypes.cast(ctypes.addressof(self._ptr.contents) + RebxForce.ap.offset, POINTER(POINTER(RebxNode)))
        return Params(self._rebx, ap_field)


# reference reboundx/extras.py:13-29 (bit, is it an error) -- messages as in src/input.c:644-689
REBX_BINARY_WARNINGS = [
    (True, 1, "REBOUNDx: Cannot open binary file. Check filename."),
    (True, 2, "REBOUNDx: Binary file is unreadable. Please open an issue on Github mentioning the version of REBOUND and REBOUNDx you are using and include the binary file."),
    (True, 4, "REBOUNDx: Ran out of system memory."),
    (True, 8, "REBOUNDx: REBOUNDx structure couldn't be loaded."),
    (True, 16, "REBOUNDx: At least one registered parameter was not loaded."),
    (False, 32, "REBOUNDx: At least one force or operator parameter was not loaded from the binary file."),
    (False, 64, "REBOUNDx: At least one particle's parameters were not loaded from the binary file."),
    (False, 128, "REBOUNDx: At least one force was not loaded from the binary file."),
    (False, 256, "REBOUNDx: At least one operator was not loaded from the binary file."),
    (False, 512, "REBOUNDx: At least one operator step was not loaded from the binary file."),
    (False, 1024, "REBOUNDx: At least one force was not added to the simulation."),
    (False, 2048, "REBOUNDx: Unknown field found in binary file. Any unknown fields not loaded."),
    (False, 4096, "REBOUNDx: Unknown list in the REBOUNDx structure wasn't loaded."),
    (False, 8192, "REBOUNDx: The value of at least one parameter was not loaded."),
    (False, 16384, "REBOUNDx: Binary file was saved with a different version of REBOUNDx."),
    (False, 32768, "REBOUNDx: A force parameter failed to load from the list of REBOUNDx implemented forces."),
]


class Extras:
    """Mirror of reboundx.Extras for the force path, bound to `lib` (default: this package's
    CUDA library)."""

    def __init__(self, sim, filename=None, lib=None):
        """`filename`: recreate the state saved by `save()` (reference reboundx/extras.py:44-61): registered
        parameters, forces with their parameters, the additional_forces order and every particle's parameters
        come from the file; its warning bits become RuntimeError / RuntimeWarning like the reference's."""
        self._lib = lib = lib or _lib.load()
        self._sim = sim
        lib.rebx_attach.restype = POINTER(RebxExtras)
        lib.rebx_load_force.restype = POINTER(RebxForce)
        lib.rebx_create_force.restype = POINTER(RebxForce)
        lib.rebx_get_force.restype = POINTER(RebxForce)
        lib.rebx_get_param_struct.restype = POINTER(RebxParam)
        lib.rebx_get_param_struct.argtypes = [c_void_p, POINTER(RebxNode), c_char_p]
        lib.rebx_get_type.restype = c_int
        lib.rebx_get_type.argtypes = [c_void_p, c_char_p]
        lib.rebx_add_force.restype = c_int
        lib.rebx_remove_force.restype = c_int
        lib.rebx_len.restype = c_int
        if filename is None:
            self._ptr = lib.rebx_attach(sim.ptr)
        else:
            import warnings as _warnings
            # like rebx_create_extras_from_binary: no default registrations, the file brings its own
            self._owned = RebxExtras()
            self._ptr = ctypes.pointer(self._owned)
            lib.rebx_initialize(sim.ptr, self._ptr)
            w = c_int(0)
            lib.rebx_init_extras_from_binary(self._ptr, c_char_p(str(filename).encode()), ctypes.byref(w))
            for major, bit, message in REBX_BINARY_WARNINGS:
                if w.value & bit:
                    if major:
                        raise RuntimeError(message)
                    _warnings.warn(message, RuntimeWarning)
        sim._extras_ref = self
        self.process_messages()

    def save(self, filename):
        """Writes the REBOUNDx state to a binary file (reference reboundx/extras.py:151-156, src/output.c:272)."""
        self._lib.rebx_output_binary(self._ptr, c_char_p(str(filename).encode()))
        self.process_messages()

    @classmethod
    def from_pointer(cls, sim, ptr, lib=None):
        """Wraps a `struct rebx_extras*` created elsewhere (e.g. by rebx_create_extras_from_binary)."""
        self = cls.__new__(cls)
        self._lib = lib or _lib.load()
        self._sim = sim
        self._lib.rebx_get_param_struct.restype = POINTER(RebxParam)
        self._lib.rebx_get_param_struct.argtypes = [c_void_p, POINTER(RebxNode), c_char_p]
        self._lib.rebx_get_type.restype = c_int
        self._lib.rebx_get_type.argtypes = [c_void_p, c_char_p]
        self._lib.rebx_get_force.restype = POINTER(RebxForce)
        self._lib.rebx_len.restype = c_int
        self._ptr = ctypes.cast(ptr, POINTER(RebxExtras))
        sim._extras_ref = self
        return self

    @property
    def struct(self):
        return self._ptr.contents

    def process_messages(self):
        return self._sim.process_messages()

    def register_param(self, name, param_type):
        self._lib.rebx_register_param(self._ptr, c_char_p(name.encode()), c_int(REBX_TYPES[param_type.lower().replace("rebx_type_", "")]))
        self.process_messages()

    def load_force(self, name):
        ptr = self._lib.rebx_load_force(self._ptr, c_char_p(name.encode()))
        self.process_messages()
        return Force(self, ptr)

    def create_force(self, name):
        ptr = self._lib.rebx_create_force(self._ptr, c_char_p(name.encode()))
        self.process_messages()
        return Force(self, ptr)

    def add_force(self, force):
        if not isinstance(force, Force):
            raise TypeError("REBOUNDx Error: Object passed to rebx.add_force is not a reboundx.Force instance.")
        ok = self._lib.rebx_add_force(self._ptr, force._ptr)
        self.process_messages()
        return ok

    def get_force(self, name):
        ptr = self._lib.rebx_get_force(self._ptr, c_char_p(name.encode()))
        if not ptr:
            raise AttributeError(f"REBOUNDx Error: Force {name} passed to rebx.get_force not found.")
        return Force(self, ptr)

    def remove_force(self, force):
        ok = self._lib.rebx_remove_force(self._ptr, force._ptr)
        if not ok:
            raise AttributeError("REBOUNDx Error: Force passed to rebx.remove_force not found in simulation.")

    def force_order(self):
        """Names in execution order (the list is push-front, reference linkedlist.c:33-37)."""
        out, node = [], self._ptr.contents.additional_forces
        while node:
            out.append(cast(node.contents.object, POINTER(RebxForce)).contents.name.decode())
            node = node.contents.next
        return out

    def particle_params(self, i):
        return Params(self, self._sim.particle_ap(i))

    def set_particle_params(self, name, values, index=None):
        """Bulk set-up of one parameter on many particles."""
        lib = self._lib
        t = lib.rebx_get_type(self._ptr, c_char_p(name.encode()))
        idx = None if index is None else np.ascontiguousarray(index, dtype=np.int64)
        iptr = None if idx is None else idx.ctypes.data_as(POINTER(c_int64))
        k = c_char_p(name.encode())
        if t == REBX_TYPES["double"]:
            vals = np.ascontiguousarray(values, dtype=np.float64)
            if hasattr(lib, "rebx_set_param_double_array"):
                lib.rebx_set_param_double_array(self._ptr, k, iptr, vals.ctypes.data_as(POINTER(c_double)), c_size_t(vals.size))
            else:
                self._sim._shim.reb_shim_bulk_set_double(self._sim.ptr, cast(lib.rebx_set_param_double, c_void_p), self._ptr,
                                                         k, iptr, vals.ctypes.data_as(POINTER(c_double)), c_size_t(vals.size))
        elif t == REBX_TYPES["int"]:
            vals = np.ascontiguousarray(values, dtype=np.int32)
            if hasattr(lib, "rebx_set_param_int_array"):
                lib.rebx_set_param_int_array(self._ptr, k, iptr, vals.ctypes.data_as(POINTER(c_int)), c_size_t(vals.size))
            else:
                self._sim._shim.reb_shim_bulk_set_int(self._sim.ptr, cast(lib.rebx_set_param_int, c_void_p), self._ptr,
                                                      k, iptr, vals.ctypes.data_as(POINTER(c_int)), c_size_t(vals.size))
        else:
            raise TypeError(f"bulk set-up supports double and int parameters, not type {t} ('{name}')")
        self.process_messages()

    def mark_params_dirty(self):
        if hasattr(self._lib, "rebx_mark_params_dirty"):
            self._lib.rebx_mark_params_dirty(self._ptr)

    def stats(self):
        """Counters of the device path: calls, kernel launches, H2D bytes, D2H bytes, table builds."""
        out = (c_uint64*5)()
        self._lib.rebx_b200_get_stats(self._ptr, out)
        return dict(zip(("calls", "launches", "h2d_bytes", "d2h_bytes", "table_builds"), [int(v) for v in out]))

    def host_table(self, force, index):
        """SoA parameter table as the builder would upload it (host copy; no device needed)."""
        lib = self._lib
        lib.rebx_b200_host_table.restype = c_int64
        n = lib.rebx_b200_host_table(self._ptr, force._ptr, c_int(index), None, c_int64(0))
        if n < 0:
            raise ValueError("no such table")
        out = np.empty(max(n, 1))
        lib.rebx_b200_host_table(self._ptr, force._ptr, c_int(index), out.ctypes.data_as(POINTER(c_double)), c_int64(n))
        return out[:n]

    def detach(self):
        if getattr(self, "_ptr", None) is not None:
            if getattr(self, "_owned", None) is not None:     # Python-owned struct (loaded from a file): reference extras.py:63-66
                self._lib.rebx_free_pointers(self._ptr)
                self._lib.rebx_detach(self._ptr)
                self._owned = None
            else:
                self._lib.rebx_free(self._ptr)
            self._ptr = None
            self._sim._extras_ref = None

    def __del__(self):
        try:
            self.detach()
        except Exception:
            pass

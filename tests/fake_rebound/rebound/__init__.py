"""Minimal stand-in for the `rebound` Python package (TEST INFRASTRUCTURE ONLY).

The reference's ctypes wrapper (reference reboundx/*.py) imports `rebound` for the Simulation and
Particle structures.  REBOUND is not installable here, so this module exposes the stand-in host
(reboundx_b200/rebound_shim) under the names the wrapper uses; only what the wrapper touches on the
force path is provided.  Used by tests/test_reference_python_wrapper.py.
"""
import ctypes
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_size_t, c_void_p

from reboundx_b200 import _lib

_shim = _lib.load_shim()


class Vec3dBasic(Structure):
    _fields_ = [("x", c_double), ("y", c_double), ("z", c_double)]


class Vec3d:
    def __init__(self, *args):
        v = args[0] if len(args) == 1 else args
        if isinstance(v, Vec3dBasic):
            self._vec3d = Vec3dBasic(v.x, v.y, v.z)
        else:
            self._vec3d = Vec3dBasic(*[float(c) for c in v])

    def __iter__(self):
        return iter((self._vec3d.x, self._vec3d.y, self._vec3d.z))


class Orbit(Structure):
    _fields_ = [("d", c_double)]


class ODE(Structure):
    _fields_ = [("length", c_int)]


class Rotation(Structure):
    _fields_ = [("ix", c_double), ("iy", c_double), ("iz", c_double), ("r", c_double)]


class Simulation(Structure):
    pass


class Particle(Structure):
    _fields_ = [("x", c_double), ("y", c_double), ("z", c_double), ("vx", c_double), ("vy", c_double), ("vz", c_double),
                ("ax", c_double), ("ay", c_double), ("az", c_double), ("m", c_double), ("r", c_double),
                ("last_collision", c_double), ("c", c_void_p), ("name", c_char_p), ("ap", c_void_p),
                ("_sim", POINTER(Simulation))]


class _Integrator(Structure):
    _fields_ = [("name", c_char_p), ("state", c_void_p)]


Simulation._fields_ = [("t", c_double), ("G", c_double), ("softening", c_double), ("dt", c_double), ("dt_last_done", c_double),
                       ("N", c_size_t), ("N_var", c_size_t), ("N_active", c_size_t), ("_particles", POINTER(Particle)),
                       ("force_is_velocity_dependent", c_int), ("_integrator", _Integrator), ("extras", c_void_p),
                       ("_additional_forces", c_void_p), ("_pre_timestep_modifications", c_void_p),
                       ("_post_timestep_modifications", c_void_p), ("_free_particle_ap", c_void_p), ("_extras_cleanup", c_void_p)]


def _new_simulation():
    _shim.reb_simulation_create.restype = c_void_p
    ptr = _shim.reb_simulation_create()
    return Simulation.from_address(ptr)


def _add(self, m=0.0, x=0.0, y=0.0, z=0.0, vx=0.0, vy=0.0, vz=0.0, r=0.0, a=None, e=0.0):
    """sim.add(m=...) in Cartesian form, or sim.add(a=..., e=...): a planar orbit at pericentre about the centre of mass
    of the particles already there (rebound's default primary), mu = G (M + m)."""
    if a is not None:
        import math
        ps = [self._particles[i] for i in range(self.N)]
        M = sum(q.m for q in ps)
        com = [sum(q.m*getattr(q, k) for q in ps)/M for k in ("x", "y", "z", "vx", "vy", "vz")]
        mu = self.G*(M + m)
        x, y, z = com[0] + a*(1.0 - e), com[1], com[2]
        vx, vy, vz = com[3], com[4] + math.sqrt(mu*(1.0 + e)/(a*(1.0 - e))), com[5]
    p = Particle(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, m=m, r=r)
    _shim.reb_simulation_add.argtypes = [c_void_p, Particle]
    _shim.reb_simulation_add(c_void_p(ctypes.addressof(self)), p)


def _process_messages(self):
    _shim.reb_shim_message_count.restype = c_size_t
    _shim.reb_shim_message.restype = c_char_p
    _shim.reb_shim_message.argtypes = [c_void_p, c_size_t]
    n = _shim.reb_shim_message_count(c_void_p(ctypes.addressof(self)))
    msgs = [_shim.reb_shim_message(c_void_p(ctypes.addressof(self)), i).decode() for i in range(n)]
    _shim.reb_shim_clear_messages(c_void_p(ctypes.addressof(self)))
    for m in msgs:
        if m[0] == "E":
            raise RuntimeError(m[1:])


def _particles(self):
    return [self._particles[i] for i in range(self.N)]


def _call_additional_forces(self):
    _shim.reb_shim_call_additional_forces(c_void_p(ctypes.addressof(self)))


def _integrate(self, tmax):
    """sim.integrate(t): fixed-step 15th-order Gauss-Radau of the stand-in (the scheme behind IAS15), 100 steps per
    innermost orbit, gravity of all bodies + additional_forces."""
    import math
    span = tmax - self.t
    if span <= 0 or self.N < 2:
        return
    p0, p1 = self._particles[0], self._particles[1]
    d = math.sqrt((p1.x - p0.x)**2 + (p1.y - p0.y)**2 + (p1.z - p0.z)**2)
    period = 2*math.pi*math.sqrt(d**3/(self.G*(p0.m + p1.m)))
    n = max(1, int(math.ceil(span/(period/100))))
    self.dt = span/n
    _shim.reb_shim_integrate.argtypes = [c_void_p, c_int, c_int]
    _shim.reb_shim_integrate(c_void_p(ctypes.addressof(self)), n, 2)
    self.process_messages()


def _pomega(self):
    """Longitude of pericentre about particles[0] (planar orbits: the angle of the eccentricity vector)."""
    import math
    sim = self._sim.contents
    p0 = sim._particles[0]
    mu = sim.G*(p0.m + self.m)
    d = (self.x - p0.x, self.y - p0.y, self.z - p0.z)
    v = (self.vx - p0.vx, self.vy - p0.vy, self.vz - p0.vz)
    r = math.sqrt(sum(c*c for c in d)); v2 = sum(c*c for c in v); rv = sum(a*b for a, b in zip(d, v))
    ex, ey = ((v2 - mu/r)*d[0] - rv*v[0])/mu, ((v2 - mu/r)*d[1] - rv*v[1])/mu
    return math.atan2(ey, ex)


def _sim_new(cls, *args, **kwargs):
    _shim.reb_simulation_create.restype = c_void_p
    sim = cls.from_address(_shim.reb_simulation_create())
    sim._owner = True               # views made by ctypes (`pointer.contents`) do not own the C simulation
    return sim


def _sim_del(self):
    """Like rebound.Simulation.__del__: frees the C simulation, which tells an attached REBOUNDx instance through
    sim->extras_cleanup that its simulation is gone."""
    if getattr(self, "_owner", False):
        self._owner = False
        _shim.reb_simulation_free.argtypes = [c_void_p]
        _shim.reb_simulation_free(c_void_p(ctypes.addressof(self)))


Simulation.__new__ = staticmethod(lambda cls, *a, **k: _sim_new(cls))
Simulation.__init__ = lambda self, *a, **k: None
Simulation.__del__ = _sim_del
Simulation.integrate = _integrate
Particle.pomega = property(_pomega)
Simulation.add = _add
Simulation.process_messages = _process_messages
Simulation.particles = property(_particles)
Simulation.update_acceleration = _call_additional_forces
Simulation.__new_sim__ = staticmethod(_new_simulation)


class Simulationarchive:
    pass

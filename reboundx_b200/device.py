"""Resident-state front end of the device layer (include/rebx_b200.h).

A ``Shard`` keeps one contiguous index range of an ensemble as SoA tensors in HBM together
with the SoA parameter tables of the attached effects, and applies them with one call of
``rebx_b200_launch_pass`` per force evaluation.  torch is used for device memory, streams
and (in ``sync_bodies`` / ``reduce_backreaction``) torch.distributed plumbing only; every
kernel is this package's own CUDA code reached through the C-ABI with raw pointers.

Sharding: particles are split by index range across ranks; the O(1)
source / primary bodies are replicated in a small ``bodies`` tensor that rank 0 broadcasts
each evaluation; back-reaction sums (3 doubles per effect) are all-reduced and applied by
the rank that owns the body.
"""
import ctypes
from ctypes import POINTER, Structure, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

from . import _lib, sharding

MAX_EFFECTS, BODY_DOUBLES, MAX_TABLES, MAX_SCALARS = 4, 24, 9, 8
ABSENT_BITS = 0x7ff8b200dead0001
ABSENT = np.frombuffer(np.uint64(ABSENT_BITS).tobytes(), dtype=np.float64)[0]

KIND = {"radiation_forces": 1, "gravitational_harmonics": 2, "gr_potential": 3, "yarkovsky_effect": 4,
        "modify_orbits_forces": 5, "gas_damping_timescale": 6, "type_I_migration": 7}
FLAG_TRAP, FLAG_BARYCENTRIC = 1, 2
# algorithmic HBM bytes per particle of ONE fused pass: state read once,
# acceleration read + written, parameter tables read once
STATE_BYTES = {"pos": 24, "vel": 24, "m": 8, "r": 8, "acc_rw": 48}
NEEDS = {1: ("vel",), 2: ("m",), 3: ("m",), 4: ("vel", "m", "r"), 5: ("vel", "m"), 6: ("vel", "m"), 7: ("vel", "m")}
TABLE_BYTES = {1: 8, 2: 0, 3: 0, 4: 9*8 + 4, 5: 24, 6: 8, 7: 0}


class CState(Structure):
    _fields_ = [("n", c_int64), ("index_base", c_int64)] + [(k, c_void_p) for k in
                ("x", "y", "z", "vx", "vy", "vz", "m", "r", "ax", "ay", "az")]


class CEffect(Structure):
    _fields_ = [("kind", c_int32), ("flags", c_int32), ("body", c_void_p), ("tab", c_void_p*MAX_TABLES),
                ("itab", c_void_p), ("scal", c_double*MAX_SCALARS), ("backreaction", c_void_p)]


class CPass(Structure):
    _fields_ = [("st", CState), ("n_effects", c_int32), ("reserved", c_int32), ("fx", CEffect*MAX_EFFECTS)]


def algorithmic_bytes_per_particle(kinds):
    needs = {"pos", "acc_rw"}
    for k in kinds:
        needs.update(NEEDS[k])
    return sum(STATE_BYTES[s] for s in needs) + sum(TABLE_BYTES[k] for k in kinds)


def spin_frame(Omega):
    """Body-fixed frame of an oblate body (reference src/gravitational_harmonics.c:104-134)."""
    ox, oy, oz = (float(v) for v in Omega)
    omega = np.sqrt(ox*ox + oy*oy + oz*oz)
    sx, sy, sz = ox/omega, oy/omega, oz/omega
    fac = np.sqrt(sx*sx + sy*sy)
    u = (-sy/fac, sx/fac, 0.0) if fac != 0.0 else (1.0, 0.0, 0.0)
    w = (sx, sy, sz)
    v = (-(u[1]*w[2] - u[2]*w[1]), -(u[2]*w[0] - u[0]*w[2]), -(u[0]*w[1] - u[1]*w[0]))
    return u, v, w


class Shard:
    """Particles [lo, hi) of a workload, resident on `device`, with `exec_order` effects."""

    def __init__(self, w, exec_order, device, lo=0, hi=None, coordinates="PARTICLE", acc0=None, index_base=None, standalone=False,
                 math="exact"):
        import torch
        self.standalone = standalone        # the whole ensemble is this one shard: evaluate() runs no collective
        if math not in ("exact", "tolerance"):
            raise ValueError("math must be 'exact' (bit-identical to the reference where + - * / sqrt) or 'tolerance' (kernels_tol.cu)")
        self.math = math
        self.torch = torch
        self.lib = lib = _lib.load()
        lib.rebx_b200_launch_pass.restype = c_int
        lib.rebx_b200_apply_backreaction.restype = c_int
        lib.rebx_b200_workspace_bytes.restype = ctypes.c_size_t
        lib.rebx_b200_error_string.restype = ctypes.c_char_p
        if not torch.cuda.is_available() or lib.rebx_b200_device_count() < 1:
            raise RuntimeError("reboundx_b200.device needs a CUDA device: the force path has no CPU fallback")
        self.device = torch.device(device)
        n_total = w["x"].shape[0]
        hi = n_total if hi is None else hi
        self.lo, self.hi, self.n = lo, hi, hi - lo
        # global index of element `lo` (differs from `lo` when `w` is itself one rank's slice of a
        # larger ensemble that replicates the central body at local index 0)
        self.index_base = lo if index_base is None else index_base
        self.exec_order = list(exec_order)
        if not 1 <= len(self.exec_order) <= MAX_EFFECTS:
            raise ValueError("1..4 stacked effects per pass")
        self.kinds = [KIND[name] for name in self.exec_order]
        self.coordinates = coordinates

        def dev(a, dtype=torch.float64):
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, dtype=dtype)

        self.state = {k: dev(w[k][lo:hi]) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}
        for i, k in enumerate(("ax", "ay", "az")):
            self.state[k] = dev(acc0[i][lo:hi]) if acc0 is not None else torch.zeros(self.n, dtype=torch.float64, device=self.device)

        def table(name):
            full = np.full(n_total, ABSENT)
            full[1:] = w[name]
            return dev(full[lo:hi])

        # body records (host master copy, refreshed from the global state by the owner rank)
        self.bodies_host = np.zeros((len(self.kinds), BODY_DOUBLES))
        self.tables, self.itables, self.scalars, self.flags = [], [], [], []
        for e, (name, kind) in enumerate(zip(self.exec_order, self.kinds)):
            tabs, itab, scal, flags = [], None, [w["G"]] + [0.0]*(MAX_SCALARS - 1), 0
            self.bodies_host[e, 0] = 0.0
            self.bodies_host[e, 1:8] = [w[k][0] for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
            if kind == 1:
                tabs = [table("beta")]
                scal[1] = w["c"]
            elif kind == 2:
                u, v, ww = spin_frame(w.get("Omega") or (0.0, 0.0, 1.0))
                self.bodies_host[e, 8:11] = [w["J2"], w.get("J4") or 0.0, w["R_eq"]]
                self.bodies_host[e, 11:14], self.bodies_host[e, 14:17], self.bodies_host[e, 17:20] = u, v, ww
            elif kind == 3:
                scal[1] = w["c"]*w["c"]
            elif kind == 4:
                names = ["ye_body_density", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo", "ye_emissivity",
                         "ye_k", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z"]
                tabs = [table(p) for p in names]
                code = np.full(n_total, 2, dtype=np.int32)
                code[1:] = w["ye_flag"]
                itab = dev(code[lo:hi], dtype=torch.int32)
                scal[1:4] = [w["ye_lstar"], w["ye_c"], w["ye_stef_boltz"]]
            elif kind == 5:
                tabs = [table(p) for p in ("tau_a", "tau_e", "tau_inc")]
                if w.get("ide_position") is not None and w.get("ide_width") is not None:
                    flags |= FLAG_TRAP
                    scal[1:3] = [w["ide_position"], w["ide_width"]]
            elif kind == 6:
                tabs = [table("d_factor")]
                scal[1:3] = [w["cs_coeff"], w["tau_coeff"]]
            elif kind == 7:
                scal[1:7] = [w.get("tIm_flaring_index") or 0.0,
                             0.01 if w.get("tIm_scale_height_1") is None else w["tIm_scale_height_1"],
                             w.get("tIm_surface_density_1") or 0.0, w.get("tIm_surface_density_exponent") or 0.0,
                             w.get("ide_position") or 0.0, w.get("ide_width") or 0.0]
            if coordinates != "PARTICLE":
                if kind not in (5, 6, 7):
                    raise ValueError("centre-of-mass coordinates apply to the rebx_com_force effects (damping trio) only")
                if coordinates == "BARYCENTRIC":
                    flags |= FLAG_BARYCENTRIC
                elif coordinates != "JACOBI":
                    raise ValueError("coordinates must be PARTICLE, BARYCENTRIC or JACOBI")
            self.tables.append(tabs)
            self.itables.append(itab)
            self.scalars.append(scal)
            self.flags.append(flags)
        self.bodies = dev(self.bodies_host)
        self.backreaction = torch.zeros((len(self.kinds), 3), dtype=torch.float64, device=self.device)
        self.workspace = torch.zeros(lib.rebx_b200_workspace_bytes(), dtype=torch.uint8, device=self.device)   # zeroed once: CTA ticket counter
        self.launches = 0
        # The replicated body records are re-broadcast from their owner only when they may have changed (construction,
        # a stepping call that moves the body on its owner only, mark_bodies_stale()): a loop of force evaluations over a
        # fixed state runs NO collective for a stack without back-reaction.  Every rank makes the same sequence of calls,
        # so the flag agrees across ranks.
        self._bodies_stale = True
        if coordinates != "PARTICLE":
            if len(self.kinds) > 3:
                raise ValueError("at most 3 stacked effects in a centre-of-mass frame")
            lib.rebx_b200_scan_scratch_bytes.restype = ctypes.c_size_t
            self.scan_scratch = torch.empty(lib.rebx_b200_scan_scratch_bytes(c_int64(self.n)), dtype=torch.uint8, device=self.device)
            self.moments = torch.zeros(8, dtype=torch.float64, device=self.device)     # sum m, sum m r, sum m v of this shard, status
            self.tsum = torch.zeros(3, dtype=torch.float64, device=self.device)        # JACOBI: sum of t_i of this shard
            self.lead_mass = dev(np.array([w["m"][0]]))                                # mass of global particle 0 (replicated)
        # REBX_B200_COM_WHOLE: this shard is the whole ensemble (the literal mass recurrences are available as a fallback)
        self._com_flags = 1 if (lo == 0 and hi == n_total and self.index_base == 0 and standalone) else 0
        self._pass = self._make_pass()
        # the same pass with REBX_B200_PASS_APPLY_BACKREACTION: a standalone shard's sums are final when its sweep ends,
        # so the sweep's last CTA adds them to the body itself (one launch per evaluation instead of three)
        self._pass_apply = self._make_pass()
        self._pass_apply.reserved |= 4

    # ------------------------------------------------------------------------------------------
    def _make_pass(self):
        P = CPass()
        P.st.n, P.st.index_base = self.n, self.index_base
        for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r", "ax", "ay", "az"):
            setattr(P.st, k, self.state[k].data_ptr())
        P.n_effects = len(self.kinds)
        P.reserved = 1 | (2 if self.math == "tolerance" else 0)   # REBX_B200_PASS_SHARED_BODY (every effect of a Shard refers to the
                                                                  # central body, index 0) | REBX_B200_PASS_TOLERANCE
        for e, kind in enumerate(self.kinds):
            fx = P.fx[e]
            fx.kind, fx.flags = kind, self.flags[e]
            fx.body = self.bodies.data_ptr() + e*BODY_DOUBLES*8
            for j, t in enumerate(self.tables[e]):
                fx.tab[j] = t.data_ptr()
            fx.itab = self.itables[e].data_ptr() if self.itables[e] is not None else None
            for j, v in enumerate(self.scalars[e]):
                fx.scal[j] = float(v)
            fx.backreaction = self.backreaction.data_ptr() + e*3*8
        return P

    @property
    def algorithmic_bytes(self):
        """HBM bytes one evaluation must move for this shard (the roofline numerator): one pass over the
        state, plus 16 B/particle of scan scratch in the Jacobi frame."""
        return (algorithmic_bytes_per_particle(self.kinds) + (16 if self.coordinates == "JACOBI" else 0))*self.n

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.lib.rebx_b200_error_string(c_int(rc)).decode()}")

    def sync_bodies(self, group=None, src=0):
        """Active-body state broadcast (NCCL over NVLink when a process group is up)."""
        if not self.standalone:
            sharding.broadcast_bodies(self.bodies, src=src, group=group)
        self._bodies_stale = False

    def mark_bodies_stale(self):
        """The owner's copy of a body record was changed from outside (every rank must call this at the same point)."""
        self._bodies_stale = True

    def launch(self, stream=None, apply=False):
        """One force evaluation of the shard: a += sum of the stacked effects (asynchronous); the back-reaction sums are
        left in self.backreaction, and with apply=True also added to the body's acceleration by the same launch."""
        s = self.torch.cuda.current_stream(self.device) if stream is None else stream
        n = c_int(0)
        rc = self.lib.rebx_b200_launch_pass(ctypes.byref(self._pass_apply if apply else self._pass), c_void_p(self.workspace.data_ptr()),
                                            c_void_p(s.cuda_stream), ctypes.byref(n))
        self._check(rc, "rebx_b200_launch_pass")
        self.launches += n.value
        return n.value

    def reduce_backreaction(self, group=None):
        if not self.standalone:
            sharding.allreduce_backreaction(self.backreaction, group=group)

    def apply_backreaction(self, stream=None):
        s = self.torch.cuda.current_stream(self.device) if stream is None else stream
        n = c_int(0)
        rc = self.lib.rebx_b200_apply_backreaction(ctypes.byref(self._pass), c_void_p(s.cuda_stream), ctypes.byref(n))
        self._check(rc, "rebx_b200_apply_backreaction")
        self.launches += n.value
        return n.value

    # ---- centre-of-mass frames of rebx_com_force (reference src/rebxtools.c:66-147), sharded ----------
    # Three local stages with one small exchange between them; `evaluate` wires them to
    # torch.distributed, a test may wire several shards on one device by hand.
    def _libcall(self, name, *args, stream=None):
        n = c_int(0)
        fn = getattr(self.lib, name)
        fn.restype = c_int
        rc = fn(*args, c_void_p(self._stream(stream).cuda_stream), ctypes.byref(n))
        self._check(rc, name)
        self.launches += n.value
        return n.value

    def com_moments(self, stream=None):
        """Stage 1: this shard's mass moments -> self.moments (JACOBI also keeps the in-shard prefix)."""
        st, scr = ctypes.byref(self._pass.st), c_void_p(self.scan_scratch.data_ptr())
        # the mass of global particle 0 (slot 7 of the replicated body record) fixes the grid on which the masses
        # are summed, so that every shard reproduces the reference's sequentially rounded interior masses
        lead = c_void_p(self.lead_mass.data_ptr())
        if self.coordinates == "JACOBI":
            return self._libcall("rebx_b200_jacobi_moments", st, scr, c_void_p(self.moments.data_ptr()), lead, c_int(self._com_flags), stream=stream)
        return self._libcall("rebx_b200_mass_moments", st, c_void_p(self.moments.data_ptr()), scr, lead, c_int(self._com_flags), stream=stream)

    def com_sweep(self, moments, stream=None):
        """Stage 2.  BARYCENTRIC: `moments` = the all-reduced totals; the sweep leaves -sum (m_i/M) f_i of this
        shard in self.backreaction.  JACOBI: `moments` = the carry (sum over the shards in front, or None);
        leaves this shard's sum of t_i in self.tsum."""
        if self.coordinates == "JACOBI":
            carry = c_void_p(moments.data_ptr()) if moments is not None else None
            return self._libcall("rebx_b200_jacobi_sweep", ctypes.byref(self._pass), c_void_p(self.scan_scratch.data_ptr()), carry,
                                 c_void_p(self.tsum.data_ptr()), c_int(self._com_flags), stream=stream)
        k = self._libcall("rebx_b200_com_bodies", c_void_p(moments.data_ptr()), c_void_p(self.bodies.data_ptr()), c_int(len(self.kinds)), stream=stream)
        return k + self.launch(stream)

    def com_apply(self, carry, stream=None):
        """Stage 3.  BARYCENTRIC: `carry` = the all-reduced (n_effects, 3) back-reaction, added to every
        particle.  JACOBI: `carry` = sum of tsum over the shards behind (or None)."""
        st = ctypes.byref(self._pass.st)
        if self.coordinates == "JACOBI":
            c = c_void_p(carry.data_ptr()) if carry is not None else None
            return self._libcall("rebx_b200_jacobi_apply", st, c_void_p(self.scan_scratch.data_ptr()), c, stream=stream)
        k = 0
        for e in range(len(self.kinds)):
            k += self._libcall("rebx_b200_add_uniform", st, c_void_p(carry.data_ptr() + e*3*8), stream=stream)
        return k

    def _evaluate_com(self, group=None):
        k = self.com_moments()
        if self.coordinates == "JACOBI":
            self._carry7 = sharding.exchange_prefix(self.moments, group)        # kept alive until the kernels ran
            k += self.com_sweep(self._carry7)
            self._carry3 = sharding.exchange_suffix(self.tsum, group)
            return k + self.com_apply(self._carry3)
        sharding.allreduce_sum(self.moments, group)
        k += self.com_sweep(self.moments)
        sharding.allreduce_sum(self.backreaction, group)
        return k + self.com_apply(self.backreaction)

    def launch_local(self, stream=None, staged=False):
        """Every kernel of one evaluation of this shard, without the cross-shard exchanges (what a
        kernel-only timing wants): the sweep in PARTICLE coordinates; in the centre-of-mass frames the staged
        pipeline (moments -> sweep -> apply), or -- JACOBI, the whole ensemble on this device, `staged` False -- the
        fused one (rebx_b200_launch_jacobi)."""
        if self.coordinates == "PARTICLE":
            return self.launch(stream)
        if self.coordinates == "JACOBI" and self.standalone and self._com_flags == 1 and not staged:
            # the whole ensemble on this device: the fused pipeline (one cooperative launch, jacobi_fused.cuh)
            return self._libcall("rebx_b200_launch_jacobi", ctypes.byref(self._pass), c_void_p(self.scan_scratch.data_ptr()), stream=stream)
        k = self.com_moments(stream)
        if self.coordinates == "JACOBI":
            return k + self.com_sweep(None, stream) + self.com_apply(None, stream)
        return k + self.com_sweep(self.moments, stream) + self.com_apply(self.backreaction, stream)

    def evaluate(self, group=None):
        """Full evaluation incl. the multi-GPU exchange steps; returns kernels launched."""
        if self.standalone:
            if self.coordinates == "PARTICLE":
                return self.launch(apply=True)
            return self.launch_local()
        if self.coordinates != "PARTICLE":
            return self._evaluate_com(group)
        if self._bodies_stale:
            self.sync_bodies(group)
        k = self.launch()
        if any(kind != 1 and kind != 4 for kind in self.kinds):
            self.reduce_backreaction(group)
            k += self.apply_backreaction()
        return k

    # ---- resident stepping around the path -------------------------------
    def _stream(self, stream):
        return self.torch.cuda.current_stream(self.device) if stream is None else stream

    def _call(self, name, *args, stream=None):
        n = c_int(0)
        fn = getattr(self.lib, name)
        fn.restype = c_int
        rc = fn(ctypes.byref(self._pass.st), *args, c_void_p(self._stream(stream).cuda_stream), ctypes.byref(n))
        self._check(rc, name)
        self.launches += n.value
        return n.value

    def drift(self, dt, stream=None):
        """x += dt v (reference rebx_drift_step, src/steppers.c:109-118)."""
        return self._call("rebx_b200_drift", c_double(dt), stream=stream)

    def kick(self, dt, stream=None):
        """v += dt a (reference rebx_kick_step's update loop, src/steppers.c:120-130)."""
        return self._call("rebx_b200_kick", c_double(dt), stream=stream)

    def gravity(self, G, n_active=1, softening=0.0, stream=None):
        """a = Newtonian pull of the first `n_active` body records (overwrites a)."""
        return self._call("rebx_b200_point_mass_gravity", c_void_p(self.bodies.data_ptr()), c_int(n_active), c_double(G),
                          c_double(softening), stream=stream)

    def refresh_bodies(self, group=None, stream=None):
        """Re-read the body records from the (moved) particles this shard owns, then broadcast."""
        k = self._call("rebx_b200_gather_bodies", c_void_p(self.bodies.data_ptr()), c_int(len(self.kinds)), stream=stream)
        self.sync_bodies(group)
        return k

    def leapfrog_step(self, dt, G, group=None):
        """One drift-kick-drift step entirely on the device: the same scheme as the host harness
        (rebound_shim.c: reb_shim_integrate scheme 0) with N_active = 1."""
        k = self.drift(0.5*dt)
        k += self.refresh_bodies(group)
        k += self.gravity(G, 1)
        k += self.launch()
        if any(kind != 1 and kind != 4 for kind in self.kinds):
            self.reduce_backreaction(group)
            k += self.apply_backreaction()
        k += self.kick(dt)
        k += self.drift(0.5*dt)
        return k

    def kepler_drift(self, dt, G, stream=None):
        """Two-body drift of the shard about the central body (rebx_b200_kepler_drift)."""
        return self._call("rebx_b200_kepler_drift", c_void_p(self.bodies.data_ptr()), c_double(G), c_double(dt), stream=stream)

    def wh_step(self, dt, G, group=None):
        """One Wisdom-Holman drift-kick-drift step entirely on the device -- Kepler drift about the central
        body for dt/2, kick by the stacked effects alone, Kepler drift for dt/2: the scheme of the host harness
        `wisdom_holman` (rebound_shim.c), i.e. the splitting behind WHFast with N_active = 1."""
        k = self.kepler_drift(0.5*dt, G)
        k += self.refresh_bodies(group)
        k += self._call("rebx_b200_zero_acc")
        k += self.launch()
        if any(kind != 1 and kind != 4 for kind in self.kinds):
            self.reduce_backreaction(group)
            k += self.apply_backreaction()
        k += self.kick(dt)
        k += self.refresh_bodies(group)
        k += self.kepler_drift(0.5*dt, G)
        k += self.refresh_bodies(group)
        return k

    def wh_steps(self, nsteps, dt, G, group=None, fused=True):
        """`nsteps` Wisdom-Holman steps with the half drifts between consecutive kicks merged into one Kepler
        drift of dt (what WHFast does between outputs; host harness scheme `wisdom_holman_merged`):
        drift(dt/2) [kick drift(dt)]^(n-1) kick drift(dt/2).  One Kepler solve per particle and step.  A single
        effect without back-reaction (radiation_forces, yarkovsky_effect) takes its kick + drift in one fused
        sweep (rebx_b200_wh_kick_drift) unless `fused` is False; the results are the same bits either way."""
        if nsteps <= 0:
            return 0
        reduces = any(kind != 1 and kind != 4 for kind in self.kinds)
        fused = fused and len(self.kinds) == 1 and not reduces       # rebx_b200_wh_kick_drift: one effect, no back-reaction
        k = self.kepler_drift(0.5*dt, G)
        k += self.refresh_bodies(group)
        for s in range(nsteps):
            drift = dt if s + 1 < nsteps else 0.5*dt
            if fused:
                n = c_int(0)
                self.lib.rebx_b200_wh_kick_drift.restype = c_int
                rc = self.lib.rebx_b200_wh_kick_drift(ctypes.byref(self._pass), c_double(G), c_double(dt), c_double(drift),
                                                      c_void_p(self._stream(None).cuda_stream), ctypes.byref(n))
                self._check(rc, "rebx_b200_wh_kick_drift")
                self.launches += n.value
                k += n.value
            else:
                k += self._call("rebx_b200_zero_acc")
                k += self.launch()
                if reduces:
                    self.reduce_backreaction(group)
                    k += self.apply_backreaction()
                k += self.kick(dt)
                k += self.refresh_bodies(group)
                k += self.kepler_drift(drift, G)
            k += self.refresh_bodies(group)
        return k

    def fused_leapfrog_step(self, dt, G, softening=0.0, stream=None, group=None):
        """The same step as `leapfrog_step` in one sweep (rebx_b200_leapfrog_step): the acceleration
        never touches HBM.  Sharded (a process group with more than one rank): every rank sweeps its own
        particles, the back-reaction sums are all-reduced -- no exchange at all for a stack without
        back-reaction such as radiation_forces -- and every rank advances its replica of the central
        body identically (rebx_b200_leapfrog_sweep / _body)."""
        n = c_int(0)
        s = c_void_p(self._stream(stream).cuda_stream)
        ws = c_void_p(self.workspace.data_ptr())
        if self.standalone or sharding._world(group)[0] == 1:
            self.lib.rebx_b200_leapfrog_step.restype = c_int
            rc = self.lib.rebx_b200_leapfrog_step(ctypes.byref(self._pass), c_double(dt), c_double(G), c_double(softening), ws, s, ctypes.byref(n))
            self._check(rc, "rebx_b200_leapfrog_step")
        else:
            self.lib.rebx_b200_leapfrog_sweep.restype = c_int
            self.lib.rebx_b200_leapfrog_body.restype = c_int
            rc = self.lib.rebx_b200_leapfrog_sweep(ctypes.byref(self._pass), c_double(dt), c_double(G), c_double(softening), ws, s, ctypes.byref(n))
            self._check(rc, "rebx_b200_leapfrog_sweep")
            if any(kind != 1 and kind != 4 for kind in self.kinds):
                self.reduce_backreaction(group)
            rc = self.lib.rebx_b200_leapfrog_body(ctypes.byref(self._pass), c_double(dt), s, ctypes.byref(n))
            self._check(rc, "rebx_b200_leapfrog_body")
        self.launches += n.value
        return n.value

    def capture_steps(self, nsteps, dt, G, scheme="leapfrog"):
        """CUDA graph of `nsteps` resident steps (scheme "leapfrog": fused_leapfrog_step; "whfast": wh_steps with
        merged drifts) for launch-bound ensembles: at N ~ 10^3 a step is 2-3 kernels of a few microseconds each,
        so the launches cost more than the work; replaying a captured graph issues them all with one call.
        Single shard (no collectives inside the capture).  Returns the torch.cuda.CUDAGraph; `.replay()` advances
        the state by `nsteps` steps -- the same bits as the eager calls."""
        torch = self.torch
        if not self.standalone and sharding._world(None)[0] > 1:
            raise RuntimeError("capture_steps records one shard's kernels; sharded stepping has collectives between them")
        step = (lambda: self.wh_steps(nsteps, dt, G)) if scheme == "whfast" else (lambda: [self.fused_leapfrog_step(dt, G) for _ in range(nsteps)])
        # every kernel must have been launched once before a capture (module loading is not capturable): do a
        # throw-away run on the live state and put the state back
        keep = {k: v.clone() for k, v in self.state.items()}
        bodies = self.bodies.clone()
        step()
        for k, v in keep.items():
            self.state[k].copy_(v)
        self.bodies.copy_(bodies)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.graph(graph, stream=side):
            step()
        return graph

    def integrate_force(self, dt, scheme="rk4", stream=None):
        """Velocities across `dt` under this shard's stacked effects alone (positions fixed): the
        reference's integrate_force operator (src/integrate_force.c) with its euler / rk2 / rk4 /
        implicit_midpoint schemes, on the device."""
        code = {"implicit_midpoint": 0, "rk4": 1, "euler": 2, "rk2": 3}[scheme]
        if getattr(self, "_scratch", None) is None:
            self._scratch = self.torch.empty(12*self.n, dtype=self.torch.float64, device=self.device)
        n, it = c_int(0), c_int(-1)
        self.lib.rebx_b200_integrate_force.restype = c_int
        rc = self.lib.rebx_b200_integrate_force(ctypes.byref(self._pass), c_double(dt), c_int(code), c_void_p(self._scratch.data_ptr()),
                                                c_void_p(self.workspace.data_ptr()), c_void_p(self._stream(stream).cuda_stream),
                                                ctypes.byref(n), ctypes.byref(it))
        self._check(rc, "rebx_b200_integrate_force")
        self.launches += n.value
        return it.value

    def posvel(self):
        return [self.state[k].cpu().numpy() for k in ("x", "y", "z", "vx", "vy", "vz")]

    def acc(self):
        return [self.state[k].cpu().numpy() for k in ("ax", "ay", "az")]

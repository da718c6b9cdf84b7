"""Host logic of the drop-in library (CPU only; no compute calls into CUDA).

Mirrors the reference's own API tests: error paths (reference reboundx/tests/test_effects.py:12-83,
test_gr.py:17-21), parameter semantics (test_params.py:35-200), custom force callbacks
(test_effects.py:19-27), plus this build's additions (SoA table builder, dirty tracking).
"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def simrebx(built):
    from reboundx_b200.host import Extras, Simulation
    sim = Simulation(G=1.0)
    sim.set_particles(x=np.array([0.0, 1.0, 2.0, 3.0]), y=np.zeros(4), z=np.zeros(4),
                      vx=np.zeros(4), vy=np.array([0.0, 1.0, 0.7, 0.5]), vz=np.zeros(4), m=np.array([1.0, 1e-3, 0.0, 0.0]))
    rebx = Extras(sim)
    yield sim, rebx
    rebx.detach()
    sim.free()


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(rebx_[a-zA-Z0-9_]+)\s*\(", text))
    names -= {"rebx_b200_kind"}
    return sorted(names)


@pytest.mark.parametrize("header", ["reboundx.h", "rebx_b200.h"])
def test_library_exports_every_declared_symbol(built, header):
    from reboundx_b200 import load
    lib = load()
    missing = [n for n in _declared(header) if not hasattr(lib, n)]
    assert not missing, missing
    for data in ("rebx_version_str", "rebx_build_str", "rebx_githash_str"):
        assert ctypes.c_char_p.in_dll(lib, data).value
    assert ctypes.c_char_p.in_dll(lib, "rebx_version_str").value == b"5.1.0"


def test_exported_symbols_keep_the_rebx_prefix(built):
    """reference .github/workflows/c.yml:33-34: every exported symbol starts with rebx_."""
    import subprocess
    from reboundx_b200 import LIB_PATH
    # the reference's own check: ! nm -g --defined-only libreboundx.so | cut -d ' ' -f3 | grep -v "^rebx_"
    out = subprocess.run(["nm", "-g", "--defined-only", LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert len(names) > 60
    bad = [n for n in names if not n.startswith("rebx_")]
    assert not bad, bad[:10]


def test_struct_layouts_are_the_reference_abi(built):
    from reboundx_b200 import host
    assert ctypes.sizeof(host.RebxNode) == 16 and ctypes.sizeof(host.RebxParam) == 24
    assert ctypes.sizeof(host.RebxForce) == 48 and ctypes.sizeof(host.RebxExtras) == 56
    assert ctypes.sizeof(host.RebParticle) == 128
    assert host.RebxForce.ap.offset == 8 and host.RebxForce.update_accelerations.offset == 32


def test_default_registry(simrebx):
    sim, rebx = simrebx
    lib = rebx._lib
    assert lib.rebx_len(rebx.struct.registered_params) == 107          # reference src/core.c:51-159
    t = lambda n: lib.rebx_get_type(rebx._ptr, n.encode())
    assert t("beta") == 1 and t("ye_flag") == 2 and t("Omega") == 8 and t("force") == 4 and t("nope") == 0
    with pytest.raises(RuntimeError, match="already in registered list"):
        rebx.register_param("beta", "double")
    rebx.register_param("my_custom", "double")
    assert t("my_custom") == 1 and lib.rebx_len(rebx.struct.registered_params) == 108


def test_param_semantics(simrebx):
    sim, rebx = simrebx
    p = rebx.particle_params(1)
    assert "beta" not in p
    with pytest.raises(AttributeError):
        p["beta"]
    with pytest.raises(AttributeError, match="not registered"):
        p["unregistered_name"] = 1.0
    v = 0.25
    p["beta"] = v
    v = 9.0                                     # values are copied in (reference test_params.py:55-60)
    assert p["beta"] == 0.25
    p["beta"] = 0.5                             # update in place: no second node
    from reboundx_b200.host import RebxNode, RebxParam
    assert p["beta"] == 0.5
    assert rebx._lib.rebx_len(ctypes.cast(sim.particle(1).ap, ctypes.POINTER(RebxNode))) == 1
    p["primary"] = 1
    p["Omega"] = (1.0, 2.0, 3.0)
    assert p["primary"] == 1 and p["Omega"] == (1.0, 2.0, 3.0)
    # newest parameter sits at the head of the list (push-front, reference linkedlist.c:33-37)
    head = ctypes.cast(sim.particle(1).ap, ctypes.POINTER(RebxNode)).contents
    assert ctypes.cast(head.object, ctypes.POINTER(RebxParam)).contents.name == b"Omega"
    assert rebx._lib.rebx_len(ctypes.cast(sim.particle(1).ap, ctypes.POINTER(RebxNode))) == 3


def test_force_lifecycle_and_lifo_order(simrebx):
    sim, rebx = simrebx
    with pytest.raises(RuntimeError, match="not found in REBOUNDx library"):
        rebx.load_force("no_such_force")                        # reference test_effects.py:15-17
    a = rebx.load_force("radiation_forces")
    b = rebx.load_force("gr_potential")
    assert a.force_type == "vel" and b.force_type == "pos"
    assert sim.force_is_velocity_dependent == 0
    rebx.add_force(b)
    assert sim.force_is_velocity_dependent == 0
    rebx.add_force(a)
    assert sim.force_is_velocity_dependent == 1                  # reference src/core.c:481-483
    assert rebx.force_order() == ["radiation_forces", "gr_potential"]   # reverse of add order
    assert rebx.get_force("gr_potential").name == "gr_potential"
    with pytest.raises(AttributeError):
        rebx.get_force("gr")
    rebx.remove_force(b)
    assert rebx.force_order() == ["radiation_forces"]
    empty = rebx.create_force("custom")
    with pytest.raises(RuntimeError, match="update_accelerations"):
        rebx.add_force(empty)


def test_whfast512_is_rejected(built):
    from reboundx_b200.host import Extras, Simulation
    sim = Simulation(integrator="whfast512")
    sim.set_particles(x=np.zeros(2), y=np.zeros(2), z=np.zeros(2))
    rebx = Extras(sim)
    f = rebx.load_force("gr_potential")
    with pytest.raises(RuntimeError, match="WHFast512"):
        rebx.add_force(f)


def test_custom_host_force_runs_through_the_dispatcher(simrebx):
    """A user callback (reference test_effects.py:19-27) is not device-backed: the dispatcher
    calls it on the host with (sim, force, sim->particles, N) and emits the reference's warnings."""
    sim, rebx = simrebx
    seen = []

    def f(simp, forcep, particles, N):
        seen.append(N)
        for i in range(N):
            particles[i].ax += 0.5*(i + 1)

    force = rebx.create_force("mine")
    force.force_type = "vel"
    force.update_accelerations = f
    rebx.add_force(force)
    sim.integrator = "whfast"
    sim.N_var = 0
    assert sim.additional_forces()
    msgs = sim.messages()
    assert seen == [4]
    assert np.array_equal(sim.acc()[0], [0.5, 1.0, 1.5, 2.0])
    assert any("velocity-dependent force to WHFAST" in m for k, m in msgs if k == "warning")
    sim.integrator = "ias15"
    sim.N_var = 2
    sim.additional_forces()
    assert any("Variational particles" in m for k, m in sim.messages())
    sim.N_var = 0


def test_missing_speed_of_light_is_an_error(simrebx):
    """reference test_gr.py:17-21 (same message path for radiation_forces / gr_potential)."""
    sim, rebx = simrebx
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs the device path to reach the parameter check")
    rebx.add_force(rebx.load_force("gr_potential"))
    sim.additional_forces()
    with pytest.raises(RuntimeError, match="speed of light"):
        sim.process_messages()


def test_no_gpu_means_loud_error_not_cpu_fallback(simrebx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sim, rebx = simrebx
    f = rebx.load_force("radiation_forces")
    f.params["c"] = 1e4
    rebx.particle_params(1)["beta"] = 0.1
    rebx.add_force(f)
    sim.set_acc(np.full(4, 7.0), np.full(4, 8.0), np.full(4, 9.0))
    sim.additional_forces()
    with pytest.raises(RuntimeError, match="GPU only"):
        sim.process_messages()
    ax, ay, az = sim.acc()
    assert np.all(ax == 7.0) and np.all(ay == 8.0) and np.all(az == 9.0)      # untouched, not silently computed


def test_soa_table_builder_presence_semantics(built):
    """The SoA tables that replace the per-particle rebx_get_param walks (reference
    src/core.c:725-746): absent != zero, presence flags, newest-first lookup."""
    from reboundx_b200 import device
    from reboundx_b200.host import Extras, Simulation
    n = 70001                                   # > one builder thread's range
    sim = Simulation()
    z = np.zeros(n)
    sim.set_particles(x=np.arange(n, dtype=float), y=z, z=z, vx=z, vy=z, vz=z, m=z, r=np.ones(n))
    rebx = Extras(sim)
    rad = rebx.load_force("radiation_forces"); rad.params["c"] = 1.0; rebx.add_force(rad)
    idx = np.arange(1, n, 3, dtype=np.int64)
    rebx.set_particle_params("beta", 0.001*idx, idx)
    rebx.particle_params(5)["radiation_source"] = 1
    rebx.particle_params(60000)["radiation_source"] = 0        # presence counts, the value does not
    beta = rebx.host_table(rad, 0)
    bits = beta.view(np.uint64)
    present = np.zeros(n, dtype=bool); present[idx] = True
    assert np.all(bits[~present] == device.ABSENT_BITS)
    assert np.array_equal(beta[present], 0.001*idx)
    assert list(rebx.host_table(rad, -2)) == [5.0, 60000.0]

    gh = rebx.load_force("gravitational_harmonics"); rebx.add_force(gh)
    for i, (j2, req) in {3: (0.1, 1.0), 9: (0.0, 1.0), 11: (0.2, None), 40000: (0.3, 2.0)}.items():
        rebx.particle_params(i)["J2"] = j2
        if req is not None:
            rebx.particle_params(i)["R_eq"] = req
    assert list(rebx.host_table(gh, -2)) == [3.0, 40000.0]     # J2 == 0 and missing R_eq are skipped

    ye = rebx.load_force("yarkovsky_effect")
    ye.params["ye_lstar"] = 1.0; ye.params["ye_c"] = 1.0; ye.params["ye_stef_boltz"] = 1.0
    rebx.add_force(ye)
    full = ["ye_rotation_period", "ye_thermal_inertia", "ye_emissivity", "ye_k", "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z"]
    for i, flag in ((1, 1), (2, -1), (3, 0), (4, 0), (6, 7)):
        pp = rebx.particle_params(i)
        pp["ye_body_density"] = 1.0; pp["ye_albedo"] = 0.1; pp["ye_flag"] = flag
    for name in full:
        rebx.particle_params(3)[name] = 1.0
    rebx.particle_params(7)["ye_flag"] = 1                      # no density/albedo -> not eligible
    code = rebx.host_table(ye, -1)
    assert list(code[:8]) == [2, 1, -1, 0, 2, 2, 2, 2]          # 4: incomplete full version; 6: undefined flag value
    rebx.detach(); sim.free()


def test_free_order_either_way(built):
    """C users may free the simulation or the extras first (reference src/core.c:226-236)."""
    from reboundx_b200.host import Extras, Simulation
    for sim_first in (True, False):
        sim = Simulation()
        sim.set_particles(x=np.zeros(3), y=np.zeros(3), z=np.zeros(3))
        rebx = Extras(sim)
        rebx.particle_params(1)["beta"] = 0.3
        rebx.add_force(rebx.load_force("gr_potential"))
        if sim_first:
            sim._extras_ref = None
            sim.free()
            assert not rebx.struct.sim                   # extras_cleanup hook nulled it
            rebx.detach()
        else:
            rebx.detach()
            assert not sim.struct.extras and not sim.struct.additional_forces
            sim.free()


def test_rad_calc_helpers(built):
    from reboundx_b200 import load
    lib = load()
    lib.rebx_rad_calc_beta.restype = ctypes.c_double
    lib.rebx_rad_calc_particle_radius.restype = ctypes.c_double
    d = ctypes.c_double
    G, c, M, L, rho, Q = 6.674e-11, 3e8, 1.99e30, 3.85e26, 1000.0, 1.0
    beta = lib.rebx_rad_calc_beta(d(G), d(c), d(M), d(L), d(1e-5), d(rho), d(Q))
    assert beta == 3.*L*Q/(16.*np.pi*G*M*c*rho*1e-5)
    r = lib.rebx_rad_calc_particle_radius(d(G), d(c), d(M), d(L), d(beta), d(rho), d(Q))
    assert abs(r - 1e-5) < 1e-19


def test_rad_calc_beta_published_value(built):
    """reference ipython_examples/Radiation_Forces_Circumplanetary_Dust.ipynb cell 11 prints
    "Particle 2's beta parameter = 0.05767021830984005" for rebx.rad_calc_beta(G, c, Msun, L, 1e-5 m, 1000 kg/m^3, 1)."""
    import ctypes
    from reboundx_b200 import load
    lib = load()
    lib.rebx_rad_calc_beta.restype = ctypes.c_double
    lib.rebx_rad_calc_beta.argtypes = [ctypes.c_double]*7
    beta = lib.rebx_rad_calc_beta(6.674e-11, 3.e8, 1.99e30, 3.85e26, 1.e-5, 1000., 1.)
    assert beta == 0.05767021830984005
    lib.rebx_rad_calc_particle_radius.restype = ctypes.c_double
    lib.rebx_rad_calc_particle_radius.argtypes = [ctypes.c_double]*7
    assert abs(lib.rebx_rad_calc_particle_radius(6.674e-11, 3.e8, 1.99e30, 3.85e26, beta, 1000., 1.)/1.e-5 - 1.0) < 1e-15


def test_scan_scratch_holds_the_staged_and_the_fused_jacobi_layouts(built):
    """rebx_b200_scan_scratch_bytes (a pure host function): the buffer must hold the staged pipeline's rows, totals, t
    vector and literal masses AND the fused pipeline's per-CTA flags (one 8-byte and one 4-byte entry per CTA, at most one
    CTA per 256-particle tile) -- kernels.cu: JacobiScratch / launch_jacobi_fused."""
    import ctypes
    from reboundx_b200 import _lib
    lib = _lib.load()
    lib.rebx_b200_scan_scratch_bytes.restype = ctypes.c_size_t
    lib.rebx_b200_scan_scratch_bytes.argtypes = [ctypes.c_int64]
    last = 0
    for n in (1, 2, 255, 256, 257, 3001, 20001, 1_000_001, 10_000_001):
        ntiles = (n + 255)//256
        got = lib.rebx_b200_scan_scratch_bytes(n)
        staged = ((ntiles + 1)*10 + 32 + 4*(n + 2))*8
        fused_flags = (ntiles + 1)*8 + ntiles*4
        assert got >= staged + fused_flags, (n, got, staged, fused_flags)
        assert got % 8 == 0 and got > last
        last = got


"""Drop-in evidence (CPU only): the reference's OWN binary writer and loader (src/output.c, src/input.c,
compiled unmodified into oracle/_ref/librebx_io_bridge.so) run on top of this library's registry /
parameter / force API and struct layouts -- INTEGRATION.md section 3 ("link the reference's untouched
objects next to this library").  A state written by reference code from our lists and read back by
reference code through our API must reproduce every force and parameter."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "librebx_io_bridge.so")


def test_reference_binary_io_round_trip_through_this_library(built, tmp_path):
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/librebx_io_bridge.so not built (no reference tree on this machine)")
    from reboundx_b200 import load
    from reboundx_b200.host import Extras, RebxExtras, Simulation
    load()
    bridge = ctypes.CDLL(BRIDGE)
    n = 12
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(3, n))

    def new_sim():
        sim = Simulation(G=2.5)
        sim.set_particles(pos[0], pos[1], pos[2], m=np.linspace(1.0, 0.1, n))
        return sim

    sim = new_sim()
    rebx = Extras(sim)
    rad = rebx.load_force("radiation_forces"); rad.params["c"] = 3.0e8; rebx.add_force(rad)
    mof = rebx.load_force("modify_orbits_forces"); mof.params["coordinates"] = 2
    mof.params["ide_position"] = 0.3; mof.params["ide_width"] = 0.02; rebx.add_force(mof)
    gh = rebx.load_force("gravitational_harmonics"); rebx.add_force(gh)
    rebx.register_param("my_custom", "double")
    for i in range(1, n):
        pp = rebx.particle_params(i)
        pp["beta"] = 0.01*i
        if i % 2:
            pp["tau_a"] = -1e5*i
        if i % 3 == 0:
            pp["my_custom"] = 7.0 + i
    p0 = rebx.particle_params(0)
    p0["primary"] = 1; p0["J2"] = 1e-2; p0["R_eq"] = 0.1; p0["Omega"] = (0.1, 0.2, 0.9)
    sim.process_messages()

    path = str(tmp_path / "state.rebx").encode()
    bridge.rebx_output_binary(rebx._ptr, ctypes.c_char_p(path))               # reference code walks OUR lists
    assert os.path.getsize(path) > 64
    with open(path, "rb") as f:
        assert f.read(20) == b"REBOUNDx Binary File"                             # reference header, src/output.c:272-293

    sim2 = new_sim()
    bridge.rebx_create_extras_from_binary.restype = ctypes.POINTER(RebxExtras)
    ptr = bridge.rebx_create_extras_from_binary(sim2.ptr, ctypes.c_char_p(path))   # reference code calls OUR API
    assert ptr, sim2.messages()
    errs = [m for k, m in sim2.messages() if k == "error"]
    assert not errs, errs
    rebx2 = Extras.from_pointer(sim2, ptr)

    assert rebx2.force_order() == rebx.force_order() == ["gravitational_harmonics", "modify_orbits_forces", "radiation_forces"]
    assert rebx2.get_force("radiation_forces").params["c"] == 3.0e8
    m2 = rebx2.get_force("modify_orbits_forces").params
    assert (m2["coordinates"], m2["ide_position"], m2["ide_width"]) == (2, 0.3, 0.02)
    assert rebx2.get_force("radiation_forces").force_type == "vel" and sim2.force_is_velocity_dependent == 1
    assert rebx2._lib.rebx_get_type(rebx2._ptr, b"my_custom") == 1                # custom registration survived
    for i in range(n):
        a, b = rebx.particle_params(i), rebx2.particle_params(i)
        for name in ("beta", "tau_a", "my_custom", "primary", "J2", "R_eq", "Omega"):
            assert (name in a) == (name in b), (i, name)
            if name in a:
                assert a[name] == b[name], (i, name)
    # the loaded forces are the device-backed entry points of this library
    from reboundx_b200.host import RebxForce
    f1 = rebx.get_force("radiation_forces")._ptr.contents.update_accelerations
    f2 = rebx2.get_force("radiation_forces")._ptr.contents.update_accelerations
    assert f1 == f2 and f1
    sim2._extras_ref = None
    rebx2._lib.rebx_free(rebx2._ptr); rebx2._ptr = None
    sim2.free()
    rebx.detach(); sim.free()

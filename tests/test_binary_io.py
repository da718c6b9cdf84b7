"""Binary state files: this library's own writer / loader
(reboundx_b200/csrc/binary_io.cpp) against the reference's src/output.c + src/input.c, which are
compiled unmodified into oracle/_ref/librebx_io_bridge.so and run on the same lists.

* the two writers produce the same file (same tree, same bytes up to struct padding);
* each loader reads the other writer's file and reproduces forces, order, registry and parameters;
* the reference's warning / error bits for missing, truncated, foreign-version and unknown-force files;
* the Python mirror of reboundx.Extras(sim, filename) / Extras.save.
CPU only: no force is evaluated.
"""
import ctypes
import os
import struct
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "librebx_io_bridge.so")
CONTAINERS = {1, 2, 3, 4, 13, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26}      # field types closed by an END field
END = 8


def parse(buf, pos, stop):
    """Independent reader of the nested field format -> [(type, payload bytes | [children])]."""
    out = []
    while pos < stop:
        ftype, = struct.unpack_from("<i", buf, pos)
        size, = struct.unpack_from("<q", buf, pos + 8)
        pos += 16
        if ftype == END:
            return out, pos
        if ftype in CONTAINERS:
            kids, after = parse(buf, pos, pos + size)
            assert after == pos + size, (ftype, after, pos + size)          # size spans children + END
            out.append((ftype, kids))
        else:
            out.append((ftype, bytes(buf[pos:pos + size])))
        pos += size
    return out, pos


def tree(path):
    buf = open(path, "rb").read()
    assert buf[:31] == b"REBOUNDx Binary File. Version: "
    t, pos = parse(buf, 64, len(buf))
    assert pos == len(buf)
    return buf[:64], t


def build_state(n=12, extra_force=None):
    from reboundx_b200.host import Extras, Simulation
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(3, n))
    sim = Simulation(G=2.5)
    sim.set_particles(pos[0], pos[1], pos[2], m=np.linspace(1.0, 0.1, n))
    rebx = Extras(sim)
    rad = rebx.load_force("radiation_forces"); rad.params["c"] = 3.0e8; rebx.add_force(rad)
    mof = rebx.load_force("modify_orbits_forces"); mof.params["coordinates"] = 2
    mof.params["ide_position"] = 0.3; mof.params["ide_width"] = 0.02; rebx.add_force(mof)
    gh = rebx.load_force("gravitational_harmonics"); rebx.add_force(gh)
    unused = rebx.load_force("gr_potential"); unused.params["c"] = 1.0e4          # allocated, never added
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
    return sim, rebx, pos


def fresh_sim(pos):
    from reboundx_b200.host import Simulation
    sim = Simulation(G=2.5)
    n = pos.shape[1]
    sim.set_particles(pos[0], pos[1], pos[2], m=np.linspace(1.0, 0.1, n))
    return sim


def assert_same_state(a, b, n):
    assert b.force_order() == a.force_order() == ["gravitational_harmonics", "modify_orbits_forces", "radiation_forces"]
    assert b.get_force("radiation_forces").params["c"] == 3.0e8
    assert b.get_force("gr_potential").params["c"] == 1.0e4
    m2 = b.get_force("modify_orbits_forces").params
    assert (m2["coordinates"], m2["ide_position"], m2["ide_width"]) == (2, 0.3, 0.02)
    assert b._lib.rebx_get_type(b._ptr, b"my_custom") == 1
    assert b._lib.rebx_get_type(b._ptr, b"beta") == 1 and b._lib.rebx_get_type(b._ptr, b"Omega") == 8
    for i in range(n):
        pa, pb = a.particle_params(i), b.particle_params(i)
        for name in ("beta", "tau_a", "my_custom", "primary", "J2", "R_eq", "Omega"):
            assert (name in pa) == (name in pb), (i, name)
            if name in pa:
                assert pa[name] == pb[name], (i, name)


def bridge():
    if not os.path.exists(BRIDGE):
        pytest.skip("oracle/_ref/librebx_io_bridge.so not built (no reference tree on this machine)")
    return ctypes.CDLL(BRIDGE)


def test_native_writer_produces_the_reference_file(built, tmp_path):
    br = bridge()
    sim, rebx, _ = build_state()
    ours, theirs = str(tmp_path / "ours.rebx"), str(tmp_path / "ref.rebx")
    rebx.save(ours)
    br.rebx_output_binary(rebx._ptr, ctypes.c_char_p(theirs.encode()))
    h1, t1 = tree(ours)
    h2, t2 = tree(theirs)
    assert h1 == h2 and len(h1) == 64
    assert t1 == t2                                                   # same nested fields, same payload bytes
    assert os.path.getsize(ours) == os.path.getsize(theirs)
    a, b = bytearray(open(ours, "rb").read()), bytearray(open(theirs, "rb").read())
    assert len(t1) == 1 and t1[0][0] == 26                            # one SNAPSHOT
    kinds = [k for k, _ in t1[0][1]]
    assert kinds == [3, 24]                                           # REBX_STRUCTURE, PARTICLES
    assert len(t1[0][1][1][1]) == 12                                  # one PARTICLE record per particle
    rebx.detach(); sim.free()


@pytest.mark.parametrize("writer,reader", [("native", "reference"), ("reference", "native"), ("native", "native")])
def test_cross_loading(built, tmp_path, writer, reader):
    from reboundx_b200 import load
    from reboundx_b200.host import Extras, RebxExtras
    br = bridge() if "reference" in (writer, reader) else None
    lib = load()
    sim, rebx, pos = build_state()
    path = str(tmp_path / "state.rebx")
    (lib if writer == "native" else br).rebx_output_binary(rebx._ptr, ctypes.c_char_p(path.encode()))
    sim2 = fresh_sim(pos)
    src = lib if reader == "native" else br
    src.rebx_create_extras_from_binary.restype = ctypes.POINTER(RebxExtras)
    ptr = src.rebx_create_extras_from_binary(sim2.ptr, ctypes.c_char_p(path.encode()))
    assert ptr
    msgs = sim2.messages()
    assert not msgs, msgs
    rebx2 = Extras.from_pointer(sim2, ptr)
    assert_same_state(rebx, rebx2, pos.shape[1])
    assert sim2.force_is_velocity_dependent == 1                       # radiation_forces is REBX_FORCE_VEL
    sim2._extras_ref = None
    lib.rebx_free(rebx2._ptr); rebx2._ptr = None
    sim2.free()
    rebx.detach(); sim.free()


def test_python_save_and_load(built, tmp_path):
    from reboundx_b200.host import Extras
    sim, rebx, pos = build_state()
    path = tmp_path / "state.rebx"
    rebx.save(path)
    sim2 = fresh_sim(pos)
    rebx2 = Extras(sim2, str(path))                                   # reference: reboundx.Extras(sim, filename)
    assert_same_state(rebx, rebx2, pos.shape[1])
    # saving the reloaded state gives the same file again
    again = tmp_path / "again.rebx"
    rebx2.save(again)
    assert open(path, "rb").read() == open(again, "rb").read()
    rebx.detach(); sim.free()


def test_missing_truncated_and_foreign_files(built, tmp_path):
    from reboundx_b200 import load
    from reboundx_b200.host import Extras
    lib = load()
    sim, rebx, pos = build_state()
    path = tmp_path / "state.rebx"
    rebx.save(path)
    data = open(path, "rb").read()

    def load_bits(blob):
        p = tmp_path / "case.rebx"
        p.write_bytes(blob)
        s = fresh_sim(pos)
        from reboundx_b200.host import RebxExtras
        ex = RebxExtras()
        lib.rebx_initialize(s.ptr, ctypes.byref(ex))
        w = ctypes.c_int(0)
        lib.rebx_init_extras_from_binary(ctypes.byref(ex), ctypes.c_char_p(str(p).encode()), ctypes.byref(w))
        s.messages()
        lib.rebx_free_pointers(ctypes.byref(ex))
        s.free()
        return w.value

    assert load_bits(data) == 0
    with pytest.raises(RuntimeError, match="Cannot open binary file"):
        Extras(fresh_sim(pos), str(tmp_path / "nope.rebx"))
    for cut in (0, 10, 63, 64, 70, 80, 200, len(data)//2, len(data) - 16, len(data) - 1):
        assert load_bits(data[:cut]) & 2, cut                         # REBX_INPUT_BINARY_ERROR_CORRUPT, and no crash
    other = bytearray(data); other[31:36] = b"9.9.9"
    assert load_bits(bytes(other)) == 16384                           # ..._WARNING_VERSION only: everything still loads
    # a force this library does not implement: warning bits, the rest of the file still loads
    renamed = data.replace(b"gr_potential\0", b"gr_potentiaX\0")
    bits = load_bits(renamed)
    assert bits & 128 and not bits & 2                                # ..._WARNING_FORCE_NOT_LOADED
    # an unknown field type inside the structure is skipped by its size
    t = bytearray(data)
    idx = data.index(struct.pack("<i", 20) + b"\0\0\0\0")             # ALLOCATED_OPERATORS header -> unknown type 99
    t[idx:idx + 4] = struct.pack("<i", 99)
    bits = load_bits(bytes(t))
    assert bits & 2048 and not bits & 2                               # ..._WARNING_FIELD_UNKNOWN
    rebx.detach(); sim.free()


def test_bulk_round_trip_and_loader_speed(built, tmp_path):
    """2*10^5 particles, two parameters each: the native loader reproduces every value and is not slower than
    the reference's fread-per-field loader running on the same library."""
    from reboundx_b200 import load
    from reboundx_b200.host import Extras, RebxExtras, Simulation
    lib = load()
    n = 200_000
    rng = np.random.default_rng(11)
    sim = Simulation(G=1.0)
    sim.set_particles(rng.normal(size=n), rng.normal(size=n), rng.normal(size=n), m=np.zeros(n))
    rebx = Extras(sim)
    rad = rebx.load_force("radiation_forces"); rad.params["c"] = 3.0e8; rebx.add_force(rad)
    beta, taua = rng.uniform(0.01, 0.5, n), -rng.uniform(1e3, 1e6, n)
    idx = np.arange(n, dtype=np.int64)
    lib.rebx_set_param_double_array(rebx._ptr, b"beta", idx.ctypes.data_as(ctypes.c_void_p), beta.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    lib.rebx_set_param_double_array(rebx._ptr, b"tau_a", idx.ctypes.data_as(ctypes.c_void_p), taua.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
    path = str(tmp_path / "big.rebx")
    t0 = time.perf_counter(); rebx.save(path); t_save = time.perf_counter() - t0
    t_save_ref = None
    if os.path.exists(BRIDGE):
        ref_path = str(tmp_path / "big_ref.rebx")
        t0 = time.perf_counter(); ctypes.CDLL(BRIDGE).rebx_output_binary(rebx._ptr, ctypes.c_char_p(ref_path.encode())); t_save_ref = time.perf_counter() - t0
        assert os.path.getsize(ref_path) == os.path.getsize(path)
    timings = {}
    for who in ("native", "reference"):
        if who == "reference" and not os.path.exists(BRIDGE):
            continue
        src = lib if who == "native" else ctypes.CDLL(BRIDGE)
        s2 = Simulation(G=1.0)
        s2.set_particles(np.zeros(n), np.zeros(n), np.zeros(n), m=np.zeros(n))
        src.rebx_create_extras_from_binary.restype = ctypes.POINTER(RebxExtras)
        t0 = time.perf_counter()
        ptr = src.rebx_create_extras_from_binary(s2.ptr, ctypes.c_char_p(path.encode()))
        timings[who] = time.perf_counter() - t0
        assert not s2.messages()
        r2 = Extras.from_pointer(s2, ptr)
        for i in (0, 1, 77, n//2, n - 1):
            pp = r2.particle_params(i)
            assert pp["beta"] == beta[i] and pp["tau_a"] == taua[i]
        s2._extras_ref = None
        lib.rebx_free(r2._ptr); r2._ptr = None
        s2.free()
    print(f"save native {t_save*1e3:.0f} ms" + (f", reference {t_save_ref*1e3:.0f} ms" if t_save_ref else "") + f"; load {', '.join(f'{k} {v*1e3:.0f} ms' for k, v in timings.items())} for {n} particles x 2 params")
    if "reference" in timings:
        assert timings["native"] <= 1.5*timings["reference"] + 0.05
    rebx.detach(); sim.free()


def test_loader_survives_random_corruption(built, tmp_path):
    """Bounds-checked cursor: 400 copies of a valid file with random bytes overwritten (field types, sizes,
    payloads) load without crashing; whatever was readable is kept, the rest is reported through the bits."""
    from reboundx_b200 import load
    from reboundx_b200.host import RebxExtras
    lib = load()
    sim, rebx, pos = build_state()
    path = tmp_path / "state.rebx"
    rebx.save(path)
    data = bytearray(open(path, "rb").read())
    rng = np.random.default_rng(5)
    flagged = 0
    for trial in range(400):
        blob = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(64, len(blob)))
            blob[at:at + int(rng.integers(1, 9))] = bytes(rng.integers(0, 256, size=8, dtype=np.uint8))[:max(0, min(8, len(blob) - at))][:int(rng.integers(1, 9))]
        p = tmp_path / "fuzz.rebx"
        p.write_bytes(bytes(blob[:len(data)]))
        s = fresh_sim(pos)
        ex = RebxExtras()
        lib.rebx_initialize(s.ptr, ctypes.byref(ex))
        w = ctypes.c_int(0)
        lib.rebx_init_extras_from_binary(ctypes.byref(ex), ctypes.c_char_p(str(p).encode()), ctypes.byref(w))
        flagged += bool(w.value)
        s.messages()
        lib.rebx_free_pointers(ctypes.byref(ex))
        s.free()
    assert flagged > 100            # most corruptions are noticed; none may crash
    rebx.detach(); sim.free()


def test_threaded_loading_of_a_large_file(built, tmp_path):
    """From 5*10^5 particle records on the loader works on several threads (records with increasing indices touch
    disjoint particles): every value of a 6*10^5-particle file comes back, through the SoA table of the force."""
    from reboundx_b200 import load
    from reboundx_b200.host import Extras, Simulation
    lib = load()
    n = 600_000
    z = np.zeros(n)
    beta = np.random.default_rng(2).uniform(0.01, 0.5, n)
    beta[::7] = np.nan                                                    # marks "not set" below
    sim = Simulation(G=1.0); sim.set_particles(z, z, z, m=z)
    rebx = Extras(sim)
    rad = rebx.load_force("radiation_forces"); rad.params["c"] = 3.0e8; rebx.add_force(rad)
    have = np.flatnonzero(~np.isnan(beta)).astype(np.int64)
    vals = beta[have].copy()
    lib.rebx_set_param_double_array(rebx._ptr, b"beta", have.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(have)))
    path = tmp_path / "large.rebx"
    rebx.save(path)
    sim2 = Simulation(G=1.0); sim2.set_particles(z, z, z, m=z)
    rebx2 = Extras(sim2, str(path))
    table = rebx2.host_table(rebx2.get_force("radiation_forces"), 0)      # the SoA beta table built from the loaded lists
    assert table.shape[0] == n
    assert np.array_equal(table[have], vals)
    absent = np.ones(n, dtype=bool); absent[have] = False
    assert np.all(np.isnan(table[absent]))                                # REBX_B200_ABSENT_BITS is a NaN payload
    rebx2.detach(); sim2.free(); rebx.detach(); sim.free()


def test_golden_file_written_by_the_reference(built, tmp_path):
    """tests/golden/state_reference_writer.rebx was produced by the reference's src/output.c (make_rebx_golden.py):
    this library's loader rebuilds the state from it, and its writer turns that state back into the same tree --
    on any machine, reference tree or not."""
    from reboundx_b200.host import Extras
    golden = os.path.join(ROOT, "tests", "golden", "state_reference_writer.rebx")
    sim, rebx, pos = build_state()                      # what the reference writer was given
    sim2 = fresh_sim(pos)
    rebx2 = Extras(sim2, golden)
    assert_same_state(rebx, rebx2, pos.shape[1])
    again = tmp_path / "again.rebx"
    rebx2.save(again)
    h1, t1 = tree(golden)
    h2, t2 = tree(str(again))
    assert t1 == t2
    assert h1[:36] == h2[:36] == b"REBOUNDx Binary File. Version: 5.1.0"
    rebx.detach(); sim.free()

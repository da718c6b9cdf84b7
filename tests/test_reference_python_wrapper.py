"""Drop-in evidence (CPU only, dev container only): the reference's UNMODIFIED Python wrapper
(/root/reference/reboundx/*.py, imported in place) driving this library.

The wrapper is pointed at libreboundx_b200.so by intercepting the one `cdll.LoadLibrary` call it makes
(INTEGRATION.md section 3); `rebound` is replaced by tests/fake_rebound (a view of the stand-in host).
Everything the wrapper does on the force path -- Extras.__init__ on Python-owned memory, load_force,
add_force, get_force, remove_force, register_param, Params on forces and particles, a custom Python force
called back through rebx_additional_forces -- then runs through this library's symbols and struct layouts.
"""
import ctypes
import importlib
import os
import sys

import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def reboundx(built):
    if not os.path.isdir(os.path.join(REF, "reboundx")):
        pytest.skip("reference tree not present on this machine")
    from reboundx_b200 import LIB_PATH, load
    load()
    sys.path.insert(0, os.path.join(HERE, "fake_rebound"))
    sys.path.insert(0, REF)
    real = ctypes.cdll.LoadLibrary

    def redirect(path):
        return real(LIB_PATH if "libreboundx" in os.path.basename(path) else path)

    ctypes.cdll.LoadLibrary = redirect
    try:
        mod = importlib.import_module("reboundx")
    finally:
        ctypes.cdll.LoadLibrary = real
    yield mod
    for name in [n for n in sys.modules if n == "reboundx" or n.startswith("reboundx.") or n == "rebound"]:
        if "reboundx_b200" not in name:
            del sys.modules[name]
    sys.path.remove(REF)
    sys.path.remove(os.path.join(HERE, "fake_rebound"))


def _sim():
    import rebound
    sim = rebound.Simulation.__new_sim__()
    sim.add(m=1.0)
    sim.add(m=1e-3, x=1.0, vy=1.0)
    sim.add(m=0.0, x=2.0, vy=0.7)
    return sim


def test_wrapper_reports_this_library(reboundx):
    assert reboundx.__version__ == "5.1.0"
    assert reboundx.__githash__.startswith("b200native")


def test_forces_and_params_through_the_reference_wrapper(reboundx):
    sim = _sim()
    rebx = reboundx.Extras(sim)                          # rebx_initialize + rebx_register_default_params on Python-owned memory
    rad = rebx.load_force("radiation_forces")
    rad.params["c"] = 3.0e8                              # reference Params.__setitem__ -> rebx_set_param_double
    assert rad.params["c"] == 3.0e8
    assert rad.force_type == 2                           # REBX_FORCE_VEL (the reference getter returns the enum value)
    rebx.add_force(rad)
    assert sim.force_is_velocity_dependent == 1
    gh = rebx.load_force("gravitational_harmonics")
    rebx.add_force(gh)
    ps = sim.particles
    ps[2].params["beta"] = 0.25                          # the property the wrapper monkeypatches onto rebound.Particle
    ps[0].params["J2"] = 1e-3
    ps[0].params["Omega"] = [0.0, 0.1, 1.0]
    ps[0].params["primary"] = 1
    assert ps[2].params["beta"] == 0.25 and ps[0].params["primary"] == 1
    assert list(ps[0].params["Omega"]) == [0.0, 0.1, 1.0]
    assert len(ps[0].params) == 3
    with pytest.raises(AttributeError):
        ps[1].params["beta"]                             # not set on that particle
    with pytest.raises(AttributeError):
        ps[1].params["no_such_parameter"] = 1.0          # not registered
    rebx.register_param("my_param", "REBX_TYPE_DOUBLE")
    ps[1].params["my_param"] = 4.5
    assert ps[1].params["my_param"] == 4.5
    assert rebx.get_force("radiation_forces").params["c"] == 3.0e8
    with pytest.raises(RuntimeError):
        rebx.load_force("not_a_force")                   # reference tests/test_effects.py:15-17
    rebx.remove_force(gh)
    with pytest.raises(AttributeError):
        rebx.remove_force(gh)
    import gc
    sim._extras_ref = None                               # the wrapper parks a reference on the simulation (extras.py:42)
    del rebx, rad, gh, ps
    gc.collect()                                         # Extras.__del__: rebx_free_pointers + rebx_detach
    assert not sim.extras


def test_custom_python_force_is_dispatched(reboundx):
    """reference tests/test_effects.py:19-27: a Python force registered through the wrapper's FORCEFUNCPTR is
    called by this library's rebx_additional_forces with (sim, force, particles, N)."""
    sim = _sim()
    rebx = reboundx.Extras(sim)
    calls = []

    def myforce(simp, forcep, particles, N):
        calls.append(N)
        particles[1].ax += 0.125

    f = rebx.create_force("myforce")
    f.force_type = "pos"
    f.update_accelerations = myforce
    rebx.add_force(f)
    sim.update_acceleration()
    sim.process_messages()
    assert calls == [3]
    assert sim.particles[1].ax == 0.125


def test_the_references_own_force_tests_pass_unmodified(reboundx):
    """reference reboundx/tests/test_effects.py, class TestForces (12 tests: load / create / get / remove forces, the
    error cases of add_force, a custom Python force integrated for 10 time units), imported in place and run
    UNMODIFIED: `import reboundx` resolves to the reference's wrapper bound to this library, `import rebound` to
    tests/fake_rebound.  (The other classes of that file need operators, the other files a GPU for the built-in
    forces -- those are ported one by one in tests/test_gpu_parity.py.)"""
    import importlib.util
    import unittest
    path = os.path.join(REF, "reboundx", "tests", "test_effects.py")
    spec = importlib.util.spec_from_file_location("reference_test_effects", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite = unittest.TestLoader().loadTestsFromTestCase(mod.TestForces)
    assert suite.countTestCases() == 12
    result = unittest.TestResult()
    suite.run(result)
    problems = [(str(t), tb) for t, tb in result.errors + result.failures]
    assert not problems, problems[0]
    assert result.testsRun == 12


def test_the_references_own_parameter_tests_pass_unmodified(reboundx):
    """reference reboundx/tests/test_params.py, class TestParams (23 tests of the parameter API through
    reboundx.params.Params: add / update / copy semantics of doubles and ints, force-valued and custom-struct
    parameters, unregistered names, re-registration, len / iter / del), imported in place and run UNMODIFIED against
    this library.  Its setUp also loads the OPERATOR "modify_mass" as a third parameter holder; that built-in operator is out of this
    library's scope (only integrate_force / drift / kick are built in), so for this test `Extras.load_operator` hands back a force object of that name instead -- the same
    `.params` interface over the same list API, which is what the tests exercise."""
    import importlib.util
    import unittest
    path = os.path.join(REF, "reboundx", "tests", "test_params.py")
    spec = importlib.util.spec_from_file_location("reference_test_params", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    saved = reboundx.Extras.load_operator
    reboundx.Extras.load_operator = lambda self, name: self.create_force(name)
    try:
        suite = unittest.TestLoader().loadTestsFromTestCase(mod.TestParams)
        n = suite.countTestCases()
        result = unittest.TestResult()
        suite.run(result)
    finally:
        reboundx.Extras.load_operator = saved
    problems = [(str(t), tb) for t, tb in result.errors + result.failures]
    assert not problems, (len(problems), problems[0])
    assert n == 23 and result.testsRun == 23


def test_the_references_lifetime_tests_that_need_no_force_evaluation(reboundx):
    """reference reboundx/tests/test_segfaults.py: test_dropsim (the Simulation is garbage collected while the Extras
    object lives: the next call must raise AttributeError, not crash -- rebx_extras_cleanup) and
    test_rebx_not_attached, unmodified.  (The other three integrate with `gr`, i.e. need the GPU.)"""
    import importlib.util
    import unittest
    path = os.path.join(REF, "reboundx", "tests", "test_segfaults.py")
    spec = importlib.util.spec_from_file_location("reference_test_segfaults", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite = unittest.TestSuite([mod.TestSegFaults("test_dropsim"), mod.TestSegFaults("test_rebx_not_attached")])
    result = unittest.TestResult()
    suite.run(result)
    problems = [(str(t), tb) for t, tb in result.errors + result.failures]
    assert not problems, problems[0]
    assert result.testsRun == 2

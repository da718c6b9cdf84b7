"""Pins the oracle (CPU, no GPU needed).

* the C restatement (oracle/oracle_forces.c, kind "port") against the committed golden vectors,
  which were produced by the reference's own C sources (tests/golden/make_golden.py);
* when oracle/_ref is present, the reference-compiled library against the same vectors and
  against the restatement on larger seeded ensembles.
The reference itself ships no acceleration-level known answers; the
fixtures cover its two radiation examples (config 1) and the edge cases of its presence logic.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import golden_io  # noqa: E402

# effects whose arithmetic is + - * / sqrt only are bit-reproducible on any IEEE host
EXACT = {"rad_debris_disk_example", "rad_circumplanetary_example", "rad_presence_and_two_sources",
         "harmonics_presence_edges", "gr_potential_massive", "stacked_rad_gr_harmonics",
         "modify_orbits_partial_params", "single_particle"}


def test_fixture_inventory():
    assert set(golden_io.names()) >= EXACT | {"yarkovsky_all_modes", "trio_particle", "trio_jacobi", "trio_barycentric"}


@pytest.mark.parametrize("name", golden_io.names())
def test_port_matches_golden(built, name):
    from oracle import oracle
    case, want, _ = golden_io.load(name)
    got = np.array(oracle.port_case(case))
    if name in EXACT:
        assert np.array_equal(got, want)
    else:       # libm transcendentals may differ between hosts by an ulp
        scale = np.maximum(np.max(np.abs(want), axis=0), 1e-300)
        assert np.max(np.abs(got - want)/scale) <= 1e-14


@pytest.mark.parametrize("name", golden_io.names())
def test_reference_build_matches_golden(built, name):
    from oracle import oracle
    from reboundx_b200 import scenarios
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    case, want, msgs = golden_io.load(name)
    got, got_msgs = scenarios.run_case(case, oracle.ref_lib())
    got = np.array(got)
    if name in EXACT:
        assert np.array_equal(got, want)
    else:
        scale = np.maximum(np.max(np.abs(want), axis=0), 1e-300)
        assert np.max(np.abs(got - want)/scale) <= 1e-14
    assert sorted(set(got_msgs)) == sorted(msgs)


@pytest.mark.parametrize("which", ["debris_disk", "ring", "ring_massive_tilted", "asteroids", "planetesimals"])
def test_port_equals_reference_on_seeded_ensembles(built, which):
    from oracle import oracle
    from reboundx_b200 import workloads as W
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    w, add = {"debris_disk": (W.debris_disk(5000), ["radiation_forces"]),
              "ring": (W.circumplanetary_ring(5000), ["gravitational_harmonics", "gr_potential"]),
              "ring_massive_tilted": (W.circumplanetary_ring(5000, tilted=True, massive=True), ["gr_potential", "gravitational_harmonics"]),
              "asteroids": (W.asteroid_ensemble(5000), ["yarkovsky_effect", "gr_potential"]),
              "planetesimals": (W.planetesimal_disk(5000), ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"])}[which]
    acc0 = W.gravity_like_acc(w, np.random.default_rng(0))
    ref = oracle.ref_apply(w, add, acc0)
    port, _ = oracle.port_apply(w, add[::-1], acc0)
    for k in range(3):
        assert np.array_equal(np.array(ref[k]), port[k])


def test_planet_trap_branches(built):
    """reference src/inner_disk_edge.c:61-77"""
    import ctypes
    from oracle import oracle
    lib = ctypes.CDLL(oracle.PORT_PATH)
    lib.oracle_planet_trap.restype = ctypes.c_double
    f = lambda r: lib.oracle_planet_trap(ctypes.c_double(r), ctypes.c_double(0.3), ctypes.c_double(0.02))
    assert f(0.5) == 1.0 and f(0.2) == -10.0
    mid = f(0.303)
    assert -4.5 < mid < 1.0
    assert abs(mid - (5.5*np.cos(((0.3*1.02 - 0.303)*2*np.pi)/(4*0.02*0.3)) - 4.5)) < 1e-14


def test_restated_orbit_elements_roundtrip(built):
    """The REBOUND restatement (rebound_shim.c) inverts the element->Cartesian map of workloads.py
    (independent implementation), pinning a, e, inc and P at the 1e-12 level."""
    import ctypes
    from reboundx_b200 import _lib, workloads as W
    from reboundx_b200.host import RebParticle
    shim = _lib.load_shim()

    class Orbit(ctypes.Structure):
        _fields_ = [(k, ctypes.c_double) for k in ("d", "v", "h", "P", "n", "a", "e", "inc", "Omega", "omega", "pomega", "f", "M", "l",
                                                   "theta", "T", "rhill", "pal_h", "pal_k", "pal_ix", "pal_iy", "hx", "hy", "hz", "ex", "ey", "ez")]
    shim.reb_orbit_from_particle.restype = Orbit
    shim.reb_orbit_from_particle.argtypes = [ctypes.c_double, RebParticle, RebParticle]
    rng = np.random.default_rng(5)
    G, M = 4*np.pi**2, 1.0
    for _ in range(50):
        a, e, inc = rng.uniform(0.5, 5), rng.uniform(0, 0.6), rng.uniform(0.01, 1.5)
        x, y, z, vx, vy, vz = W.kepler_to_cartesian(G*M, np.array([a]), np.array([e]), np.array([inc]),
                                                    *(rng.uniform(0, 2*np.pi, (3, 1))))
        p = RebParticle(x=x[0], y=y[0], z=z[0], vx=vx[0], vy=vy[0], vz=vz[0], m=0.0)
        o = shim.reb_orbit_from_particle(G, p, RebParticle(m=M))
        assert abs(o.a - a) < 1e-11*a and abs(o.e - e) < 1e-11 and abs(o.inc - inc) < 1e-11
        assert abs(o.P - 2*np.pi*np.sqrt(a**3/(G*M))) < 1e-11*o.P


def test_yarkovsky_notebook_known_answers(built):
    """Pins the oracle -- the reference's yarkovsky_effect.c compiled in place, on top of the restated
    reb_orbit_from_particle, under the stand-in's Wisdom-Holman scheme -- to the numbers the reference itself
    publishes (ipython_examples/YarkovskyEffect.ipynb, produced with real REBOUND + WHFast): the semi-major-axis
    drift after 2*10^6 steps agrees to 1e-6 relative (measured 1.3e-8, 7e-9 and 1.4e-7, i.e. 5e-12 AU on a = 0.75 AU).  This is a
    trajectory-level known answer for the one effect whose parity is otherwise unpinned at the REBOUND boundary."""
    from oracle import oracle
    import yarkovsky_notebook as nb
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    da, _ = nb.drift(True, oracle.ref_lib())
    assert abs(da[0]/nb.NOTEBOOK_FULL - 1.0) < 1e-6, da
    da, _ = nb.drift(False, oracle.ref_lib())
    assert abs(da[0]/nb.NOTEBOOK_SIMPLE[0] - 1.0) < 1e-6, da
    assert abs(da[1]/nb.NOTEBOOK_SIMPLE[1] - 1.0) < 1e-6, da


def test_gr_notebook_known_answers(built):
    """Pins the oracle's `gr` -- the reference's gr.c compiled in place on top of the restated REBOUND Jacobi
    transformations -- to the perihelion advance the reference publishes (GeneralRelativity.ipynb,
    SavingAndLoadingSimulations.ipynb; real REBOUND + IAS15): see tests/gr_notebook.py."""
    from oracle import oracle
    import gr_notebook as nb
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    nb.check(*nb.run(oracle.ref_lib()))


def test_damping_notebook_laws(built):
    """Pins the oracle's damping forces -- the reference's modify_orbits_forces.c, gas_damping_timescale.c,
    type_I_migration.c and inner_disk_edge.c compiled in place, on top of the RESTATED reb_orbit_from_particle elements
    and rebx_com_force frames (default: JACOBI) -- to the analytic laws the reference publishes for them
    (ipython_examples/Migration.ipynb, InnerDiskEdge.ipynb, TypeIMigration.ipynb, GasDampingTimescale.ipynb; plotted there
    against real REBOUND, no printed digits): tests/damping_notebooks.py.  Trajectory-level, at the accuracy of the laws
    (measured: 7e-8 ... 1e-4), for the effects whose parity is otherwise unpinned at the REBOUND boundary."""
    from oracle import oracle
    import damping_notebooks as nb
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    lib = oracle.ref_lib()
    # a(t) = a0 e^(t/tau_a), two planets; the outer one is sampled within its 5 orbits (osculating a wobbles by ~e P/tau)
    _, a, pred = nb.migration(lib)
    assert np.max(np.abs(a[0]/pred[0] - 1.0)) < 5e-4, a[0]/pred[0] - 1.0                 # measured 1.4e-4
    assert np.max(np.abs(a[1]/pred[1] - 1.0)) < 1e-2, a[1]/pred[1] - 1.0                 # measured 4.6e-3
    # modify_orbits_forces + planet trap: e-fold decay, then the planet stops inside the edge's band
    _, a, pred = nb.inner_disk_edge(lib)
    moving = pred > 0.1
    assert np.max(np.abs(a[moving]/pred[moving] - 1.0)) < 1e-6, a/pred - 1.0             # measured 7e-8
    assert np.all(np.abs(a[~moving] - 0.1) <= 0.1*0.02), a                                # stops at 0.1012
    # type_I_migration + planet trap: constant da/dt, then the planet stops inside the edge's band
    _, a, pred, edge, width = nb.type_I(lib)
    moving = pred > edge
    assert np.max(np.abs(a[moving] - pred[moving])) < 3e-4, a - pred                     # measured 1.2e-4 AU
    assert np.all(np.abs(a[~moving] - edge) <= edge*width), a                             # stops at 0.1010
    # gas_damping_timescale: e'/e = -1/tau, i'/i = -1/(2 tau), tau(e, i) of Dawson et al. 2016 eq. 16
    _, e, inc, e_law, i_law = nb.gas_damping(lib)
    assert np.max(np.abs(e/e_law - 1.0)) < 3e-4 and np.max(np.abs(inc/i_law - 1.0)) < 3e-4, (e/e_law - 1.0, inc/i_law - 1.0)   # measured 1e-4
    assert e[-1] < 0.97*0.05 and inc[-1] < 0.99*5*np.pi/180                               # and it did damp


@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_linear_com_oracle_against_the_reference_loops(built, coordinates, base):
    """oracle_com_force_linear is the full-size (10^6) oracle of the centre-of-mass frames: the reference's own
    COM recurrence + the back-reaction summed once in long double instead of one particle at a time.  Pinned here
    against the reference's O(N^2) loops (oracle/_ref when built, and their restatement) at N = 2*10^4: they
    differ only by the reference's accumulated rounding of the n sequential subtractions, <= sqrt(n) eps of |a|."""
    from oracle import oracle
    from parity import errors
    from reboundx_b200 import workloads as W
    trio = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]
    n = 20000
    w = W.planetesimal_disk(n)
    if coordinates == "BARYCENTRIC":
        w["d_factor"] = np.full(n, 5.0)
    acc0 = None if base == "zero" else W.gravity_like_acc(w, np.random.default_rng(3))
    quad, _ = oracle.port_apply(w, trio[::-1], acc0, coordinates)
    lin, _ = oracle.port_apply(w, trio[::-1], acc0, coordinates, com="linear")
    if oracle.have_ref():
        ref = oracle.ref_apply(w, trio, acc0, coordinates)
        for k in range(3):
            assert np.array_equal(np.array(ref[k]), quad[k])
    e = errors(lin, quad, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(lin, quad)
    assert e["max_normwise"] <= 2e-14, e            # measured 5e-15 (JACOBI), 6e-15 (BARYCENTRIC)
    if coordinates == "BARYCENTRIC":
        # the star: a tiny force that is nearly all back-reaction, i.e. the sum the two functions round differently
        star = errors([a[:1] for a in lin], [a[:1] for a in quad])
        assert star["max_normwise"] <= 1e-11, star


def test_exact_com_differs_from_the_reference_by_the_interior_mass(built):
    """Why the device scans the masses on the reference's rounding grid (kernels_jacobi.cuh): with EXACT prefix
    sums for COM_i the Jacobi forces move by ~2e-12 norm-wise against the reference -- its interior mass
    carries ~sqrt(N) ulps of sequential rounding and enters mu = G (m + M_i)."""
    from oracle import oracle
    from parity import errors
    from reboundx_b200 import workloads as W
    trio = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]
    w = W.planetesimal_disk(20000)
    lin, _ = oracle.port_apply(w, trio[::-1], None, "JACOBI", com="linear")
    exact, _ = oracle.port_apply(w, trio[::-1], None, "JACOBI", com="linear_exact")
    e = errors(exact, lin)
    assert 1e-13 < e["max_normwise"] < 1e-10, e

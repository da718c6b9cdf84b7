"""GPU parity tests proper: the CUDA path, called through the drop-in C-ABI exactly as REBOUND
would call it (sim->additional_forces on host AoS particles), against the oracle on the same
seeded inputs.  Oracle = the reference's own sources compiled in place (oracle/_ref) when
present, else the C restatement (bit-identical to it, tests/test_oracle.py)."""
import os

import numpy as np
import pytest

from parity import TOL, errors, guard

pytestmark = pytest.mark.gpu

TRIO = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]


def oracle_acc(w, add_order, acc0, coordinates="PARTICLE"):
    from oracle import oracle
    want, br = oracle.port_apply(w, add_order[::-1], acc0, coordinates)
    if oracle.have_ref():
        ref = oracle.ref_apply(w, add_order, acc0, coordinates)
        for k in range(3):          # the restatement IS the reference, bit for bit
            assert np.array_equal(np.array(ref[k]), want[k])
    return want, br


def device_acc(w, add_order, acc0, coordinates="PARTICLE"):
    from reboundx_b200 import scenarios
    sim, rebx, handles = scenarios.build(w, add_order, coordinates=coordinates, acc=acc0)
    sim.additional_forces()
    sim.process_messages()
    out = sim.acc()
    stats = rebx.stats()
    rebx.detach()
    sim.free()
    assert stats["launches"] >= 1, "no CUDA kernel was launched"
    return out, stats


def acc_variants(w, seed):
    from reboundx_b200 import workloads
    return {"zero": None, "gravity": workloads.gravity_like_acc(w, np.random.default_rng(seed))}


@pytest.mark.parametrize("n", [1, 2, 1000, 40001])
@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_radiation_forces_bit_exact(built, n, base):
    """Config 1/2 shape: star + dust, beta per particle (reference src/radiation_forces.c)."""
    from reboundx_b200 import workloads
    w = workloads.debris_disk(n)
    acc0 = acc_variants(w, 1)[base]
    want, _ = oracle_acc(w, ["radiation_forces"], acc0)
    got, _ = device_acc(w, ["radiation_forces"], acc0)
    e = errors(got, want)
    assert e["bit_identical"], e


@pytest.mark.parametrize("order", [["gravitational_harmonics", "gr_potential"], ["gr_potential", "gravitational_harmonics"]])
@pytest.mark.parametrize("tilted", [False, True])
def test_harmonics_gr_potential_test_particles(built, order, tilted):
    """Config 3: oblate planet + massless ring, J2/J4 fused with gr_potential, both LIFO orders."""
    from reboundx_b200 import workloads
    w = workloads.circumplanetary_ring(30001, tilted=tilted)
    for base, acc0 in acc_variants(w, 2).items():
        want, _ = oracle_acc(w, order, acc0)
        got, stats = device_acc(w, order, acc0)
        e = errors(got, want)
        assert e["bit_identical"], (base, e)


def test_harmonics_gr_potential_massive_ring_backreaction(built):
    """Massive ring particles: per-particle results stay bit-exact; the back-reaction summed onto
    the planet is a parallel reduction and is compared with the exactly (long double) summed
    terms, tolerance relative to the sum of |terms|."""
    from reboundx_b200 import workloads
    order = ["gravitational_harmonics", "gr_potential"]
    w = workloads.circumplanetary_ring(30001, tilted=True, massive=True)
    acc0 = acc_variants(w, 3)["gravity"]
    want, br = oracle_acc(w, order, acc0)
    got, _ = device_acc(w, order, acc0)
    e = errors(got, want, skip=(0,))
    assert e["bit_identical"], e
    exact = np.array([acc0[k][0] for k in range(3)]) + br["gravitational_harmonics"][:3] + br["gr_potential"][:3]
    scale = br["gravitational_harmonics"][3:] + br["gr_potential"][3:] + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= TOL*scale[k], (k, got[k][0], exact[k], scale[k])
        assert abs(want[k][0] - exact[k]) <= 1e-11*scale[k]      # the reference's own sequential sum


@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_yarkovsky_gr_potential(built, base):
    """Config 4: all three ye_flag modes, per-particle parameters, fused with gr_potential."""
    from reboundx_b200 import workloads
    order = ["yarkovsky_effect", "gr_potential"]
    w = workloads.asteroid_ensemble(200001)
    acc0 = acc_variants(w, 4)[base]
    want, _ = oracle_acc(w, order, acc0)
    got, _ = device_acc(w, order, acc0)
    e = errors(got, want, skip=(0,))
    # libdevice pow/atan/sin/cos vs glibc: not bit-identical; norm-wise within tolerance always, the per-component
    # picture pinned by the recorded guard (tests/parity.py)
    guard(f"yarkovsky_gr_potential/{base}", e)
    simple = np.concatenate([[False], w["ye_flag"] != 0])
    es = errors([g[simple] for g in got], [x[simple] for x in want])
    assert es["bit_identical"], es          # simple versions use + - * / sqrt only


@pytest.mark.parametrize("add_order", [TRIO, TRIO[::-1], ["modify_orbits_forces"], ["type_I_migration", "gas_damping_timescale"]])
def test_damping_trio_particle_coordinates(built, add_order):
    """Config 5 in REBX_COORDINATES_PARTICLE (the O(N) mode of the reference)."""
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(30001)
    acc0 = acc_variants(w, 5)["gravity"]
    want, br = oracle_acc(w, add_order, acc0)
    got, _ = device_acc(w, add_order, acc0)
    e = errors(got, want, skip=(0,))
    guard("trio_particle/gravity/" + "+".join(add_order), e, frac=4e-5)       # 30001 x 3 components: at most three outliers
    exact = np.array([acc0[k][0] for k in range(3)]) + sum(br[f][:3] for f in add_order)
    scale = sum(br[f][3:] for f in add_order) + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= 1e-12*scale[k], (k, got[k][0], exact[k], scale[k])


def test_modify_orbits_without_trap_bit_exact(built):
    """modify_orbits_forces without the planet trap uses + - * / only: bit-identical."""
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(20001)
    w["ide_position"] = None
    w["ide_width"] = None
    acc0 = acc_variants(w, 6)["gravity"]
    want, _ = oracle_acc(w, ["modify_orbits_forces"], acc0)
    got, _ = device_acc(w, ["modify_orbits_forces"], acc0)
    assert errors(got, want, skip=(0,))["bit_identical"]


def test_resident_shards_match_host_path(built):
    """The resident SoA entry (reboundx_b200.device.Shard) over two index-range shards gives the
    same accelerations as the host-facing call over the whole ensemble."""
    import torch
    from reboundx_b200 import device, workloads
    w = workloads.debris_disk(50001)
    acc0 = acc_variants(w, 7)["gravity"]
    want, _ = oracle_acc(w, ["radiation_forces"], acc0)
    n = w["x"].shape[0]
    cut = 20000
    parts = []
    for lo, hi in ((0, cut), (cut, n)):
        sh = device.Shard(w, ["radiation_forces"], "cuda:0", lo=lo, hi=hi, acc0=acc0)
        sh.evaluate()
        torch.cuda.synchronize()
        parts.append(sh.acc())
    got = [np.concatenate([p[k] for p in parts]) for k in range(3)]
    assert errors(got, want)["bit_identical"]


def test_resident_backreaction_two_shards(built):
    """Back-reaction across shards: per-shard sums added, applied by the shard owning the body."""
    import torch
    from reboundx_b200 import device, workloads
    order = ["gravitational_harmonics", "gr_potential"]
    w = workloads.circumplanetary_ring(30001, massive=True)
    acc0 = acc_variants(w, 8)["gravity"]
    want, br = oracle_acc(w, order, acc0)
    n = w["x"].shape[0]
    shards = [device.Shard(w, order[::-1], "cuda:0", lo=lo, hi=hi, acc0=acc0) for lo, hi in ((0, 15000), (15000, n))]
    for sh in shards:
        sh.launch()
    total = shards[0].backreaction + shards[1].backreaction       # what the NCCL all-reduce computes
    for sh in shards:
        sh.backreaction.copy_(total)
        sh.apply_backreaction()
    torch.cuda.synchronize()
    got = [np.concatenate([sh.acc()[k] for sh in shards]) for k in range(3)]
    assert errors(got, want, skip=(0,))["bit_identical"]
    exact = np.array([acc0[k][0] for k in range(3)]) + sum(br[f][:3] for f in order)
    scale = sum(br[f][3:] for f in order) + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= TOL*scale[k]


def test_dirty_tracking_and_custom_host_force(built):
    """Tables are rebuilt when (and only when) a parameter changes; a custom host force keeps
    running through the same dispatcher between device sweeps."""
    from reboundx_b200 import scenarios, workloads
    from reboundx_b200.host import RebParticle
    w = workloads.debris_disk(5000)
    sim, rebx, handles = scenarios.build(w, ["radiation_forces"])
    calls = []

    def kick(simp, forcep, particles, N):
        calls.append(N)
        particles[1].ax += 1.0

    custom = rebx.create_force("kick")
    custom.force_type = "pos"
    custom.update_accelerations = kick
    rebx.add_force(custom)
    sim.additional_forces(); sim.process_messages()
    s1 = rebx.stats()
    a1 = sim.acc()
    sim.set_acc()
    sim.additional_forces(); sim.process_messages()
    s2 = rebx.stats()
    a2 = sim.acc()
    assert calls == [sim.N, sim.N]
    assert s2["table_builds"] == s1["table_builds"], "tables rebuilt although nothing changed"
    assert np.array_equal(a1[0], a2[0])
    rebx.particle_params(7)["beta"] = 0.9
    sim.set_acc()
    sim.additional_forces(); sim.process_messages()
    s3 = rebx.stats()
    a3 = sim.acc()
    assert s3["table_builds"] == s2["table_builds"] + 1
    assert a3[0][7] != a2[0][7] and np.array_equal(np.delete(a3[0], 7), np.delete(a2[0], 7))
    w2 = dict(w); w2["beta"] = w["beta"].copy(); w2["beta"][6] = 0.9
    want, _ = oracle_acc(w2, ["radiation_forces"], None)
    want[0][1] += 1.0
    assert errors(a3, want)["bit_identical"]


def test_multi_chunk_pipeline(built, monkeypatch):
    """Small chunk size forces the double-buffered H2D/kernel/D2H pipeline through many chunks
    (odd sizes, ragged tail) and must not change a single bit."""
    from reboundx_b200 import workloads
    monkeypatch.setenv("REBX_B200_CHUNK", "4096")
    order = ["gravitational_harmonics", "gr_potential"]
    w = workloads.circumplanetary_ring(33333, massive=True)
    acc0 = acc_variants(w, 9)["gravity"]
    want, br = oracle_acc(w, order, acc0)
    got, stats = device_acc(w, order, acc0)
    assert errors(got, want, skip=(0,))["bit_identical"]
    exact = np.array([acc0[k][0] for k in range(3)]) + sum(br[f][:3] for f in order)
    scale = sum(br[f][3:] for f in order) + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= TOL*scale[k]


@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
@pytest.mark.parametrize("add_order", [TRIO, ["gas_damping_timescale", "modify_orbits_forces"], ["type_I_migration"]])
def test_damping_trio_centre_of_mass_frames(built, coordinates, add_order):
    """REBX_COORDINATES_JACOBI (the reference's default) and _BARYCENTRIC: O(N^2) sequential loops in
    the reference (src/rebxtools.c:92-128), scans / reductions on the device.  The evaluation order
    of the sums differs, so the bar is the north-star tolerance, not bit identity."""
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(3000)
    if coordinates == "BARYCENTRIC":
        w["d_factor"] = np.full(3000, 5.0)
    acc0 = acc_variants(w, 10)["gravity"]
    want, _ = oracle_acc(w, add_order, acc0, coordinates)
    from reboundx_b200 import scenarios
    sim, rebx, _ = scenarios.build(w, add_order, coordinates=coordinates, acc=acc0)
    sim.additional_forces()
    msgs = sim.messages()
    got = sim.acc()
    assert rebx.stats()["launches"] >= 3
    rebx.detach(); sim.free()
    errs = [m for k, m in msgs if k == "error"]
    if coordinates == "BARYCENTRIC" and "gas_damping_timescale" in add_order:
        assert errs and all("d_factor" in m for m in errs)       # the star has no d_factor: same error as the reference
    else:
        assert not errs, errs
    if coordinates == "BARYCENTRIC":
        # The star itself is evaluated against the barycentre it sits 3.5e-8 AU away from: its
        # relative position is the difference of two AU-scale numbers, so ANY change in the order
        # the centre of mass is summed in (sequential fold in REBOUND, tree reduction here) moves
        # it by ~eps*|x|/|x0 - com|.  That one body gets a conditioning-aware bound; all others
        # must meet the tolerance.
        e = errors(got, want, skip=(0,))
        M = w["m"].sum()
        off = np.sqrt(sum((w[k][0] - (w[k]*w["m"]).sum()/M)**2 for k in ("x", "y", "z")))
        cond = np.max(np.abs(w["x"]))/off
        star = errors([g[:1] for g in got], [x[:1] for x in want])
        assert star["max_normwise"] <= 16*np.finfo(float).eps*cond, (star, cond)
    else:
        e = errors(got, want)
    guard(f"com_frames_3000/{coordinates}/" + "+".join(add_order), e, frac=1.2e-4 if coordinates == "JACOBI" else 1e-3)


def _com_device(w, add_order, coordinates, acc0):
    from reboundx_b200 import scenarios
    sim, rebx, _ = scenarios.build(w, add_order, coordinates=coordinates, acc=acc0)
    sim.additional_forces()
    sim.messages()
    got = sim.acc()
    rebx.detach(); sim.free()
    return got


@pytest.mark.parametrize("math", ["exact", "tolerance"])
def test_jacobi_frame_ragged_sizes(built, monkeypatch, math):
    """The fused Jacobi pipeline on ensembles around its granularities (32-particle warp tiles, 256-particle CTA rows, more
    warps than tiles, a single planetesimal): against the reference's own loops, norm-wise at the north-star tolerance."""
    from reboundx_b200 import workloads
    if math == "tolerance":
        monkeypatch.setenv("REBX_B200_MATH", "tolerance")
    for n in (1, 2, 31, 32, 33, 63, 255, 256, 257, 1023, 2049):
        w = workloads.planetesimal_disk(n)
        acc0 = acc_variants(w, 50 + n)["gravity"]
        want, _ = oracle_acc(w, TRIO, acc0, "JACOBI")
        got = _com_device(w, TRIO, "JACOBI", acc0)
        e = errors(got, want)
        assert all(np.isfinite(a).all() for a in got), n
        assert e["max_normwise"] <= TOL, (n, e)


_STAGED_SNIPPET = """
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from reboundx_b200 import scenarios, workloads
w = workloads.planetesimal_disk(50001)
acc0 = workloads.gravity_like_acc(w, np.random.default_rng(77))
trio = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]
sim, rebx, _ = scenarios.build(w, trio, coordinates="JACOBI", acc=acc0)
sim.additional_forces(); sim.messages()
np.save({out!r}, np.array(sim.acc()))
print(rebx.stats()["launches"])
"""


@pytest.mark.parametrize("math", ["exact", "tolerance"])
def test_jacobi_fused_pipeline_against_the_staged_one(built, tmp_path, math):
    """rebx_b200_launch_jacobi runs ONE cooperative launch (jacobi_fused.cuh); REBX_B200_JACOBI=staged (read once per
    process, hence the subprocesses) keeps the six-launch pipeline a sharded caller drives.  Same interior masses, the
    position / velocity moments summed in another order: the two agree far inside the tolerance, and the launch
    counters tell which one ran."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("fused", "staged"):
        out = str(tmp_path / f"{mode}.npy")
        env = dict(os.environ)
        env.pop("REBX_B200_JACOBI", None)
        if mode == "staged":
            env["REBX_B200_JACOBI"] = "staged"
        if math == "tolerance":
            env["REBX_B200_MATH"] = "tolerance"
        run = subprocess.run([sys.executable, "-c", _STAGED_SNIPPET.format(root=root, tests=os.path.join(root, "tests"), out=out)],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
        assert run.returncode == 0, run.stderr[-2000:]
        res[mode] = (np.load(out), int(run.stdout.strip().splitlines()[-1]))
    assert res["fused"][1] < res["staged"][1], (res["fused"][1], res["staged"][1])      # unpack + 1 + pack against unpack + 6 + pack
    e = errors(list(res["fused"][0]), list(res["staged"][0]))
    assert e["max_normwise"] <= 1e-14, e
    assert not np.array_equal(res["fused"][0], np.zeros_like(res["fused"][0]))


@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_centre_of_mass_frames_20000_against_the_reference(built, coordinates, base):
    """The harshest view the reference can still produce in seconds: N = 2*10^4, the PURE damping force (zero
    base) and the gravity base, against the reference's own O(N^2) loops.  Round 1 missed the tolerance here
    (2e-12 norm-wise): the interior mass of rebx_com_force enters mu = G (m + M_i) and the reference's M_i
    carries ~sqrt(N) ulps of sequential rounding.  The device now reproduces that rounding exactly
    (kernels_jacobi.cuh: scan on the mass grid), so the bar is the plain north-star tolerance."""
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(20000)
    if coordinates == "BARYCENTRIC":
        w["d_factor"] = np.full(20000, 5.0)
    acc0 = acc_variants(w, 21)[base]
    want, _ = oracle_acc(w, TRIO, acc0, coordinates)
    got = _com_device(w, TRIO, coordinates, acc0)
    # BARYCENTRIC: the star's own force is the difference of AU-scale numbers 3e-8 AU apart (see the N = 3000 test)
    e = errors(got, want, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got, want)
    # Per component the reference itself is the noisier side here: it subtracts the back-reaction of every particle
    # behind j from a_j one at a time (up to N roundings of eps |a_j|), the device once.  The O(N) oracle, which sums
    # them in long double, differs from the reference's loops by exactly this much (1.1e-12 per component, 1.3e-4 of
    # the components above 1e-13, 5.2e-15 norm-wise at N = 2*10^4: tests/test_oracle.py) -- so the outlier fraction is
    # bounded at 5e-4 here and the per-component maximum by the recorded guard.
    # BARYCENTRIC adds one more ill-conditioned ingredient: the star is evaluated against the barycentre it sits
    # 3e-8 AU from, and (m_0/M) f_0 ~ f_0 dominates the shift EVERY particle receives (rebxtools.c:108-115).  That
    # shift agrees to 2.6e-14 relative, i.e. ~1.5e-16 absolute on every component -- 5e-14 norm-wise on the pure
    # force, and a per-component outlier wherever a component nearly cancels against the shift (0.3 % of them on
    # the gravity base; the O(N) oracle vs the reference's loops: same picture, tests/test_oracle.py).
    guard(f"com_frames_20000/{coordinates}/{base}", e, frac=5e-4 if coordinates == "JACOBI" else 1e-2)
    # Against the reference's own COM recurrence with the back-reaction summed once (the O(N) oracle) the Jacobi
    # frame meets the tolerance on EVERY component, on the pure force too (measured 2.2e-14):
    from oracle import oracle
    lin, _ = oracle.port_apply(w, TRIO[::-1], acc0, coordinates, com="linear")
    el = errors(got, lin, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got, lin)
    if coordinates == "JACOBI":
        assert el["max_component"] <= TOL, el
    guard(f"com_frames_20000_vs_linear_oracle/{coordinates}/{base}", el, frac=1e-5 if coordinates == "JACOBI" else 1e-2)


@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
def test_config5_full_size_centre_of_mass_frames(built, coordinates):
    """BASELINE config 5 at its full 10^6 planetesimals in the reference's default frame (JACOBI) and in
    BARYCENTRIC: the reference's O(N^2) loops would need ~40 min per force, so the oracle is
    oracle_com_force_linear -- the reference's own COM recurrence (bit-identical interior masses and
    centres of mass) with the back-reaction summed in long double, pinned against the O(N^2) loops and
    oracle/_ref at N <= 2*10^4 by tests/test_oracle.py."""
    from oracle import oracle
    from reboundx_b200 import workloads
    n = 1_000_000
    w = workloads.planetesimal_disk(n)
    if coordinates == "BARYCENTRIC":
        w["d_factor"] = np.full(n, 5.0)
    acc0 = acc_variants(w, 22)["gravity"]
    want, _ = oracle.port_apply(w, TRIO[::-1], acc0, coordinates, com="linear")
    got = _com_device(w, TRIO, coordinates, acc0)
    e = errors(got, want, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got, want)
    # 10^6 sequential steps of the reference's COM fold (reproduced by the oracle) leave ~3e-19 AU of rounding in
    # COM_i, the device's tree sums ~1e-20: a handful of nearly circular, nearly planar orbits amplify that above
    # 1e-13 on single small components (JACOBI, zero base: 3e-5 of the components); norm-wise all is <= 6e-14.
    fr = 1e-4 if coordinates == "JACOBI" else 2e-2
    guard(f"com_frames_1e6/{coordinates}/gravity", e, frac=fr)
    want0, _ = oracle.port_apply(w, TRIO[::-1], None, coordinates, com="linear")
    got0 = _com_device(w, TRIO, coordinates, None)
    e0 = errors(got0, want0, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got0, want0)
    if coordinates == "JACOBI":
        guard(f"com_frames_1e6/{coordinates}/zero", e0, frac=fr)
    else:
        # BARYCENTRIC, pure force: every particle receives the SAME shift -sum_i (m_i/M) f_i, which the star's own term
        # dominates -- the star sits 3e-8 AU from the barycentre, so its force carries the 10^6 roundings of the
        # reference's sequential centre-of-mass fold (reproduced by the oracle, not by the device's tree sum).  The
        # disagreement is therefore ONE 3-vector, the same on every particle (~9e-14 of |a|): norm-wise inside the
        # tolerance, per component meaningless (2/3 of all components sit above 1e-13 because of it).  Asserted as
        # such: norm-wise <= 1e-13, and with the common shift removed every particle agrees to 1e-15 of |a|.
        guard(f"com_frames_1e6/{coordinates}/zero", e0, frac=1.0)
        d = np.array(got0)[:, 1:] - np.array(want0)[:, 1:]
        common = np.median(d, axis=1, keepdims=True)
        scale = np.max(np.abs(np.array(want0)[:, 1:]), axis=0)
        assert np.max(np.abs(d - common)/scale) <= 1e-15, float(np.max(np.abs(d - common)/scale))


@pytest.mark.parametrize("case", ["lead_0.999_crosses_binade", "exact_ties", "massless_in_front", "tiny_masses_power_of_two_lead",
                                  "lead_just_below_a_power_of_two"])
@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
def test_interior_mass_edge_cases(built, case, coordinates):
    """Mass distributions on which the exactness argument of the mass-grid scan does NOT hold (running sum
    leaves the binade of the leading mass, exact ties, peel at the bottom of the binade): the launch must
    notice and redo the reference's two recurrences literally.  Against the reference's O(N^2) loops."""
    from reboundx_b200 import workloads
    n = 3000
    w = workloads.planetesimal_disk(n)
    rng = np.random.default_rng(5)
    if case == "lead_0.999_crosses_binade":
        w["m"][0] = 0.999
        w["m"][1:] = 10.0**rng.uniform(-7.0, -5.5, n)         # ~2e-3 in all: the running sum crosses 1.0
    elif case == "exact_ties":
        w["m"][1:] = (rng.integers(1, 1000, n) + 0.5)*2.0**-52   # every addition is a tie
    elif case == "massless_in_front":
        w["m"][1:40] = 0.0
    elif case == "tiny_masses_power_of_two_lead":
        w["m"][0] = 2.0
        w["m"][1:] = 10.0**rng.uniform(-18.0, -14.0, n)       # many below half a grid unit (2.2e-16)
    elif case == "lead_just_below_a_power_of_two":
        w["m"][0] = 1.0 - 2.0**-40                             # the running sum leaves the binade [0.5, 1) within a few particles
    if coordinates == "BARYCENTRIC":
        w["d_factor"] = np.full(n, 5.0)
    acc0 = acc_variants(w, 23)["gravity"]
    add = ["modify_orbits_forces", "type_I_migration"] if case == "massless_in_front" else TRIO
    want, _ = oracle_acc(w, add, acc0, coordinates)
    got = _com_device(w, add, coordinates, acc0)
    e = errors(got, want, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got, want)
    guard(f"interior_mass_edge/{case}/{coordinates}", e, frac=1.2e-4 if coordinates == "JACOBI" else 2e-2)


@pytest.mark.parametrize("coordinates", ["JACOBI", "BARYCENTRIC"])
@pytest.mark.parametrize("cuts", [(), (1000,), (2, 1537, 2998)])
def test_sharded_centre_of_mass_frames(built, coordinates, cuts):
    """the centre-of-mass frames over index-range SHARDS.  Each shard runs the three
    local stages (rebx_b200_jacobi_moments/_sweep/_apply, or mass_moments / sweep / add_uniform); between
    them only 7 + 3 doubles per shard are exchanged -- here by hand between shards resident on one
    device, exactly what Shard.evaluate() does with torch.distributed all-gather / all-reduce."""
    import torch
    from reboundx_b200 import device, sharding, workloads
    w = workloads.planetesimal_disk(3000)
    acc0 = acc_variants(w, 10)["gravity"]
    want, _ = oracle_acc(w, TRIO, acc0, coordinates)
    n = w["x"].shape[0]
    edges = [0] + list(cuts) + [n]
    shards = [device.Shard(w, TRIO[::-1], "cuda:0", lo=lo, hi=hi, coordinates=coordinates, acc0=acc0) for lo, hi in zip(edges, edges[1:])]
    for sh in shards:
        sh.com_moments()
    rows = [sh.moments for sh in shards]
    keep = []
    if coordinates == "JACOBI":
        for r, sh in enumerate(shards):
            keep.append(sharding.exclusive_prefix(rows, r))
            sh.com_sweep(keep[-1])
        rows3 = [sh.tsum for sh in shards]
        for r, sh in enumerate(shards):
            keep.append(sharding.exclusive_suffix(rows3, r))
            sh.com_apply(keep[-1])
    else:
        total = sum(rows[1:], rows[0].clone())
        for sh in shards:
            sh.com_sweep(total)
        br = sum((sh.backreaction for sh in shards[1:]), shards[0].backreaction.clone())
        for sh in shards:
            sh.com_apply(br)
    torch.cuda.synchronize()
    got = [np.concatenate([sh.acc()[k] for sh in shards]) for k in range(3)]
    e = errors(got, want, skip=(0,)) if coordinates == "BARYCENTRIC" else errors(got, want)
    guard(f"sharded_com_frames/{coordinates}/{len(cuts)}cuts", e, frac=1.2e-4 if coordinates == "JACOBI" else 1e-3)
    if not cuts:        # one shard == the host-facing call's resident chunk
        from reboundx_b200 import scenarios
        sim, rebx, _ = scenarios.build(w, TRIO, coordinates=coordinates, acc=acc0)
        sim.additional_forces()
        sim.messages()              # BARYCENTRIC: the star's missing d_factor is reported, like the reference does
        host = sim.acc()
        rebx.detach(); sim.free()
        if coordinates == "JACOBI":
            # the host-facing call runs the FUSED pipeline (jacobi_fused.cuh: contiguous tile ranges per CTA), the shard
            # above the staged one: the position / velocity moments are summed in another order, the masses are the same
            eh = errors(got, host)
            assert eh["max_normwise"] <= 1e-14, eh
            # and the fused pipeline is what a standalone shard runs: same kernel, same bits as the host-facing call
            one = device.Shard(w, TRIO[::-1], "cuda:0", coordinates=coordinates, acc0=acc0, standalone=True)
            one.evaluate()
            torch.cuda.synchronize()
            assert errors(one.acc(), host)["bit_identical"]
        else:
            assert errors(got, host)["bit_identical"]


def test_nccl_sharded_evaluation(built, tmp_path):
    """World-size-2 NCCL run of the sharded evaluation (tests/nccl_sharded_check.py) when the box has two
    GPUs: per-particle results bit-identical in PARTICLE coordinates, north-star tolerance in the
    centre-of-mass frames (the cross-shard carry changes the summation order)."""
    import json, os, socket, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "report.json"
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", str(port), os.path.join(here, "nccl_sharded_check.py"), str(out)], check=True, timeout=600)
    report = json.loads(out.read_text())
    assert len(report) == 9
    for r in report:
        if r["expect"] == "bits":           # no cross-shard arithmetic per particle: identical bits
            assert r["bit_identical"], r
            assert r["body_max_normwise"] <= TOL, r
        elif r["expect"] == "tol":          # centre-of-mass frames: the cross-shard carry changes the summation order
            assert r["max_normwise"] <= TOL, r
        else:                               # trajectories under a back-reaction summed per shard: north-star criterion (ii)
            assert r["max_normwise"] <= 1e-10, r


def test_jacobi_is_the_default_frame(built):
    """No 'coordinates' parameter => REBX_COORDINATES_JACOBI (reference src/modify_orbits_forces.c:131-135)."""
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(1500)
    want, _ = oracle_acc(w, ["modify_orbits_forces"], None, "JACOBI")
    got, _ = device_acc(w, ["modify_orbits_forces"], None, coordinates=None)
    assert errors(got, want)["max_normwise"] <= TOL


# ---- criterion (ii): trajectories after 10^3 steps --------------------------------------------------
def _trajectory(case_or_w, lib, nsteps, scheme, dt, n_active, add_order=None, coordinates="PARTICLE"):
    from reboundx_b200 import scenarios
    if add_order is None:
        sim, rebx, _ = scenarios.build_case(case_or_w, lib)
    else:
        sim, rebx, _ = scenarios.build(case_or_w, add_order, lib=lib, coordinates=coordinates)
    sim.dt = dt
    sim.N_active = n_active
    sim.integrate(nsteps, scheme)
    errs = [m for k, m in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    out = np.array(sim.posvel())
    rebx.detach(); sim.free()
    return out


def _assert_trajectories_agree(a, b, tol=1e-10):
    pos_scale = np.sqrt((b[:3]**2).sum(axis=0))
    vel_scale = np.sqrt((b[3:]**2).sum(axis=0))
    dpos = np.sqrt(((a[:3] - b[:3])**2).sum(axis=0))
    dvel = np.sqrt(((a[3:] - b[3:])**2).sum(axis=0))
    live = pos_scale > 0
    assert np.max(dpos[live]/pos_scale[live]) <= tol, np.max(dpos[live]/pos_scale[live])
    live = vel_scale > 0
    assert np.max(dvel[live]/vel_scale[live]) <= tol, np.max(dvel[live]/vel_scale[live])


@pytest.mark.parametrize("scheme", ["leapfrog", "rk4", "wisdom_holman", "gauss_radau"])
def test_trajectory_debris_disk_1000_steps(built, scheme):
    """Config 1a shape (Sun + 1000 dust grains, beta = 0.1, dt = 1e8 s): 10^3 steps with the CUDA force
    vs the reference-compiled force, everything else identical (host harness, not REBOUND).
    wisdom_holman = the example's own scheme class (WHFast: Kepler drifts + kicks by the additional
    force); gauss_radau = the 15th-order scheme behind IAS15 with a fixed step."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import golden_io
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    case, _, _ = golden_io.load("rad_debris_disk_example")
    ours = _trajectory(case, None, 1000, scheme, 1e8, 1)
    ref = _trajectory(case, oracle.ref_lib(), 1000, scheme, 1e8, 1)
    _assert_trajectories_agree(ours, ref)
    assert np.array_equal(ours, ref)          # bit-exact forces => bit-exact trajectories


def test_trajectory_circumplanetary_example_gauss_radau(built):
    """Config 1b shape (reference examples/rad_forces_circumplanetary: Sun + Saturn + 2 grains, IAS15, the Sun
    flagged as radiation source, N_active = 2): 10^3 fixed Gauss-Radau steps of 1/80 of the grains' orbit."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import golden_io
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    case, _, _ = golden_io.load("rad_circumplanetary_example")
    ours = _trajectory(case, None, 1000, "gauss_radau", 2.0e3, 2)
    ref = _trajectory(case, oracle.ref_lib(), 1000, "gauss_radau", 2.0e3, 2)
    _assert_trajectories_agree(ours, ref)
    assert np.array_equal(ours, ref)


@pytest.mark.parametrize("which", ["ring", "asteroids", "planetesimals_particle", "planetesimals_jacobi"])
@pytest.mark.parametrize("scheme", ["rk4", "gauss_radau", "wisdom_holman"])
def test_trajectories_other_configs_1000_steps(built, which, scheme):
    from oracle import oracle
    from reboundx_b200 import workloads as W
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    if which == "ring":
        w, add, coords, dt = W.circumplanetary_ring(400, massive=True), ["gravitational_harmonics", "gr_potential"], "PARTICLE", 50.0
    elif which == "asteroids":
        w, add, coords, dt = W.asteroid_ensemble(400), ["yarkovsky_effect", "gr_potential"], "PARTICLE", 0.01
    elif which == "planetesimals_particle":
        w, add, coords, dt = W.planetesimal_disk(400), TRIO, "PARTICLE", 0.002
    else:
        w, add, coords, dt = W.planetesimal_disk(400), TRIO, "JACOBI", 0.002
    ours = _trajectory(w, None, 1000, scheme, dt, 1, add, coords)
    ref = _trajectory(w, oracle.ref_lib(), 1000, scheme, dt, 1, add, coords)
    _assert_trajectories_agree(ours, ref)


def test_gr_notebook_known_answers_with_this_library(built):
    """The reference's published perihelion advance under `gr` (tests/gr_notebook.py) with THIS library's `gr`:
    two active bodies, so the work is the from-scratch host part for the actives (dispatch.cpp: gr_host_actives)
    around the device reduction and sweep."""
    import gr_notebook as nb
    nb.check(*nb.run(None))


@pytest.mark.parametrize("full", [True, False])
def test_yarkovsky_notebook_known_answer_on_the_gpu(built, full):
    """The reference's published Yarkovsky drift (ipython_examples/YarkovskyEffect.ipynb cells 12 and 20, see
    tests/yarkovsky_notebook.py) with THIS library's force: the first 10 % of the notebook's run (2*10^5
    Wisdom-Holman steps through sim->additional_forces) must follow the reference-compiled force to 1e-10
    (criterion (ii)) and give 10 % of the published drift (the drift is linear in time: measured 2e-5 off)."""
    import yarkovsky_notebook as nb
    from oracle import oracle
    steps = nb.STEPS//10
    da, pv = nb.drift(full, None, steps)
    want = np.array([nb.NOTEBOOK_FULL] if full else nb.NOTEBOOK_SIMPLE)/10
    assert np.all(np.abs(da/want - 1.0) < 1e-3), (da, want)
    if oracle.have_ref():
        da_ref, pv_ref = nb.drift(full, oracle.ref_lib(), steps)
        _assert_trajectories_agree(pv, pv_ref)
        assert np.all(np.abs(da/da_ref - 1.0) < 1e-6), (da, da_ref)


def test_damping_notebook_laws_with_this_library(built):
    """The analytic laws the reference publishes for its damping forces (tests/damping_notebooks.py; the CPU suite pins
    the oracle to them over the full spans) with the CUDA force through the drop-in call: the first sample of each law,
    against the law itself and against the oracle's trajectory (criterion (ii))."""
    from oracle import oracle
    from reboundx_b200 import _lib
    import damping_notebooks as nb
    lib = _lib.load()
    ref = oracle.ref_lib() if oracle.have_ref() else None
    _, a, pred = nb.migration(lib, stop_after=1)
    assert abs(a[0][0]/pred[0][0] - 1.0) < 5e-4 and abs(a[1][0]/pred[1][0] - 1.0) < 1e-2, (a, pred)
    if ref is not None:
        _, ar, _ = nb.migration(ref, stop_after=1)
        assert np.max(np.abs(a/ar - 1.0)) < 1e-10, a/ar - 1.0
    _, a, pred = nb.inner_disk_edge(lib, stop_after=1)
    assert abs(a[0]/pred[0] - 1.0) < 1e-6, (a, pred)
    _, a, pred, edge, width = nb.type_I(lib, stop_after=1)
    assert abs(a[0] - pred[0]) < 3e-4, (a, pred)
    _, e, inc, e_law, i_law = nb.gas_damping(lib, stop_after=1)
    assert abs(e[0]/e_law[0] - 1.0) < 3e-4 and abs(inc[0]/i_law[0] - 1.0) < 3e-4, (e, e_law, inc, i_law)
    if ref is not None:
        _, er, incr, _, _ = nb.gas_damping(ref, stop_after=1)
        assert abs(e[0]/er[0] - 1.0) < 1e-10 and abs(inc[0]/incr[0] - 1.0) < 1e-10, (e, er, inc, incr)


@pytest.mark.parametrize("case", ["one_active_massive_asteroids", "three_active_massless", "stacked_with_yarkovsky"])
def test_gr_1pn(built, case):
    """`gr` (reference src/gr.c): device reduction + host part for the <= 8 active bodies + device sweep.
    Massless test particles: bit-identical.  Massive ones change the order in which their pull on the
    active bodies is summed, so the bar is the tolerance."""
    from oracle import oracle
    from reboundx_b200 import scenarios, workloads
    w = workloads.asteroid_ensemble(20001)
    forces, n_active = ["gr"], 1
    if case == "three_active_massless":
        w["m"][1], w["m"][2] = 1e-3, 3e-4
        w["m"][3:] = 0.0
        n_active = 3
    elif case == "stacked_with_yarkovsky":
        forces = ["yarkovsky_effect", "gr"]
    acc0 = acc_variants(w, 11)["gravity"]
    want, _ = oracle.port_apply(w, forces[::-1], acc0, n_active=n_active)
    if oracle.have_ref():
        ref = oracle.ref_apply(w, forces, acc0, n_active=n_active)
        for k in range(3):
            assert np.array_equal(np.array(ref[k]), want[k])
    sim, rebx, _ = scenarios.build(w, forces, acc=acc0)
    sim.N_active = n_active
    sim.additional_forces()
    sim.process_messages()
    got = sim.acc()
    rebx.detach(); sim.free()
    e = errors(got, want)
    if case == "three_active_massless":
        assert e["bit_identical"], e
    else:
        guard(f"gr_1pn/{case}", e, frac=2e-5)


def test_gr_n_active_unset_and_many_active_bodies(built):
    """The reference's common use of `gr`: N_active unset (every particle active), e.g. the Sun and eight planets.  The
    active bodies' mutual terms are O(N_active^2) host work, the test particles' sweep is on the device with the
    actives tiled through shared memory: up to 4096 active bodies.  Against the reference (oracle/_ref)."""
    from oracle import oracle
    from reboundx_b200 import scenarios, workloads
    rng = np.random.default_rng(8)
    for n, n_active, massive in ((9, None, True), (300, None, True), (20001, 40, False), (5001, 600, True)):
        w = workloads.asteroid_ensemble(n - 1)
        A = n if n_active is None else n_active
        w["m"][1:A] = 10.0**rng.uniform(-7, -4, A - 1)
        if not massive:
            w["m"][A:] = 0.0
        acc0 = acc_variants(w, 12)["gravity"]
        want, _ = oracle.port_apply(w, ["gr"], acc0, n_active=(-1 if n_active is None else n_active))
        if oracle.have_ref():
            ref = oracle.ref_apply(w, ["gr"], acc0, n_active=n_active)
            for k in range(3):
                assert np.array_equal(np.array(ref[k]), want[k])
        sim, rebx, _ = scenarios.build(w, ["gr"], acc=acc0)
        if n_active is not None:
            sim.N_active = n_active
        sim.additional_forces()
        sim.process_messages()
        got = sim.acc()
        rebx.detach(); sim.free()
        e = errors(got, want)
        if not massive:
            assert e["bit_identical"], (n, n_active, e)
        else:
            # massive test particles: their pull on the actives is summed in another order than the reference's loop
            assert e["max_normwise"] <= TOL, (n, n_active, e)


def test_gr_refuses_an_unbounded_number_of_active_bodies(built):
    from reboundx_b200 import scenarios, workloads
    w = workloads.asteroid_ensemble(5000)
    sim, rebx, _ = scenarios.build(w, ["gr"])           # N_active unset: 5001 active bodies, O(N^2) in the reference
    sim.additional_forces()
    with pytest.raises(RuntimeError, match="N_active"):
        sim.process_messages()
    rebx.detach(); sim.free()


def test_resident_leapfrog_matches_host_harness_bitwise(built):
    """drift/kick/gravity kept on the device around the force sweep.  1000
    drift-kick-drift steps of a resident shard (no PCIe traffic) reproduce, bit for bit, the host
    harness stepping the reference-compiled radiation force through sim->additional_forces."""
    import os, sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import golden_io
    from oracle import oracle
    from reboundx_b200 import device, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    w = workloads.debris_disk(1000, seed=3)
    w["beta"] = np.full(1000, 0.1)
    case = {"G": w["G"], "integrator": "whfast", "state": {k: w[k] for k in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}, "acc0": None,
            "forces": [{"name": "radiation_forces", "params": {"c": w["c"]}}],
            "particle_params": {"beta": (np.arange(1, 1001, dtype=np.int64), w["beta"])}}
    ref = _trajectory(case, oracle.ref_lib(), 1000, "leapfrog", 1e8, 1)
    sh = device.Shard(w, ["radiation_forces"], "cuda:0")
    for _ in range(1000):
        sh.leapfrog_step(1e8, w["G"])
    torch.cuda.synchronize()
    got = np.array(sh.posvel())
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("which", ["debris_disk", "ring_massive"])
def test_resident_wisdom_holman_matches_host_harness(built, which):
    """BASELINE config 2 as quoted ("radiation_forces under WHFast"), resident: 1000 Wisdom-Holman steps
    (Kepler drift about the star on the device, rebx_b200_kepler_drift; kick by the stacked effects) against
    the host harness `wisdom_holman` stepping the reference-compiled force.  The Kepler arithmetic is the
    same source on both sides (csrc/kepler_core.h), so radiation is bit for bit; with a massive ring the
    back-reaction on the planet is summed in another order (tolerance of criterion (ii))."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    if which == "debris_disk":
        w, add, dt = workloads.debris_disk(1000, seed=3), ["radiation_forces"], 1e8
    else:
        w, add, dt = workloads.circumplanetary_ring(1000, massive=True), ["gravitational_harmonics", "gr_potential"], 50.0
    ref = _trajectory(w, oracle.ref_lib(), 1000, "wisdom_holman", dt, 1, add)
    sh = device.Shard(w, add[::-1], "cuda:0")
    for _ in range(1000):
        sh.wh_step(dt, w["G"])
    torch.cuda.synchronize()
    got = np.array(sh.posvel())
    if which == "debris_disk":
        assert np.array_equal(got, ref)
    else:
        _assert_trajectories_agree(got, ref)
    # the merged form (one Kepler drift of dt between consecutive kicks), against the host harness doing the same
    ref = _trajectory(w, oracle.ref_lib(), 1000, "wisdom_holman_merged", dt, 1, add)
    for fused in (True, False):      # radiation: kick + drift in one sweep (rebx_b200_wh_kick_drift) or as separate kernels
        sh = device.Shard(w, add[::-1], "cuda:0")
        sh.wh_steps(1000, dt, w["G"], fused=fused)
        torch.cuda.synchronize()
        got = np.array(sh.posvel())
        if which == "debris_disk":
            assert np.array_equal(got, ref)
        else:
            _assert_trajectories_agree(got, ref)


@pytest.mark.parametrize("scheme", ["leapfrog", "whfast"])
def test_cuda_graph_of_resident_steps(built, scheme):
    """device.Shard.capture_steps: a CUDA graph of 20 resident steps of the 1000-grain example, replayed 10 times,
    gives the bits of 200 eager steps (and the launches no longer dominate: timed in tools/latency_probe.py)."""
    import torch
    from reboundx_b200 import device, workloads
    w = workloads.debris_disk(1000, seed=3)
    eager = device.Shard(w, ["radiation_forces"], "cuda:0")
    if scheme == "whfast":
        for _ in range(10):
            eager.wh_steps(20, 1e8, w["G"])
    else:
        for _ in range(200):
            eager.fused_leapfrog_step(1e8, w["G"])
    torch.cuda.synchronize()
    graphed = device.Shard(w, ["radiation_forces"], "cuda:0")
    g = graphed.capture_steps(20, 1e8, w["G"], scheme)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(np.array(graphed.posvel()), np.array(eager.posvel()))


@pytest.mark.parametrize("which", ["debris_disk", "ring_massive", "asteroids", "planetesimals"])
def test_fused_leapfrog_step_equals_separate_sweeps(built, which):
    """rebx_b200_leapfrog_step (one sweep, accelerations in registers) against the five separate
    sweeps of Shard.leapfrog_step: bit-identical state after 100 steps, central body included."""
    import torch
    from reboundx_b200 import device, workloads as W
    if which == "debris_disk":
        w, order, dt = W.debris_disk(20001), ["radiation_forces"], 1e8
    elif which == "ring_massive":
        w, order, dt = W.circumplanetary_ring(20001, massive=True, tilted=True), ["gr_potential", "gravitational_harmonics"], 50.0
    elif which == "asteroids":
        w, order, dt = W.asteroid_ensemble(20001), ["gr_potential", "yarkovsky_effect"], 0.01
    else:
        w, order, dt = W.planetesimal_disk(20001), ["type_I_migration", "gas_damping_timescale", "modify_orbits_forces"], 0.002
    a = device.Shard(w, order, "cuda:0")
    b = device.Shard(w, order, "cuda:0")
    for _ in range(100):
        a.leapfrog_step(dt, w["G"])
        b.fused_leapfrog_step(dt, w["G"])
    torch.cuda.synchronize()
    pa, pb = np.array(a.posvel()), np.array(b.posvel())
    if which == "debris_disk":
        assert np.array_equal(pa, pb)
        assert torch.equal(a.bodies[:, :8], b.bodies[:, :8])
    else:
        # massive particles: the back-reaction onto the central body is a reduction whose partial sums
        # follow the grid, and the two kernels run different grids -> last-bit differences in the
        # body's kick.  Everything else is the same arithmetic.
        _assert_trajectories_agree(pa, pb, tol=1e-11)


# ---- tolerance-mode arithmetic (kernels_tol.cu): FMA, shared reciprocals; default stays bit-exact -------------
TOL_CASES = {
    "debris_disk": (["radiation_forces"], {}),
    "circumplanetary_ring": (["gravitational_harmonics", "gr_potential"], {"tilted": True, "massive": True}),
    "asteroid_ensemble": (["yarkovsky_effect", "gr_potential"], {}),
    "planetesimal_disk": (TRIO, {}),
}


@pytest.mark.parametrize("name", sorted(TOL_CASES))
@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_tolerance_mode_against_the_reference(built, monkeypatch, name, base):
    """REBX_B200_MATH=tolerance through the drop-in call, against the reference's own loops: the north-star bar
    (<= 1e-13) norm-wise on the pure force AND on the gravity base, the per-component picture pinned by the
    recorded guards.  The kernels must really be the tolerance ones (results differ from the exact build's)."""
    from reboundx_b200 import workloads
    add, kw = TOL_CASES[name]
    w = getattr(workloads, name)(200001, **kw)
    acc0 = acc_variants(w, 31)[base]
    want, br = oracle_acc(w, add, acc0)
    exact, _ = device_acc(w, add, acc0)
    monkeypatch.setenv("REBX_B200_MATH", "tolerance")
    got, _ = device_acc(w, add, acc0)
    assert not errors(got, exact)["bit_identical"], "REBX_B200_MATH=tolerance did not select the tolerance kernels"
    e = errors(got, want, skip=(0,))
    # Tolerance mode is a NORM-WISE promise (a few ulp per operation, ~1e-15 of |a|): a component that is orders of
    # magnitude below |a| -- the tilted ring's small components after the rotation back from the body frame, the
    # damping force of a nearly planar orbit -- carries that absolute error, so 0.03 % (ring) and 0.4 % (pure damping
    # force) of the components sit above 1e-13 relative to themselves.  The per-component bar is the exact build's job.
    guard(f"tolerance_mode/{name}/{base}", e, frac={"circumplanetary_ring": 1e-3, "planetesimal_disk": 1e-2}.get(name, 2e-5))
    # the central body: back-reaction sums against the long-double sum of the oracle's terms
    reducing = [f for f in add if f in br]
    if reducing:
        a0 = [0.0, 0.0, 0.0] if acc0 is None else [acc0[k][0] for k in range(3)]
        exact0 = np.array(a0) + sum(br[f][:3] for f in reducing)
        scale = sum(br[f][3:] for f in reducing) + np.abs(a0)
        for k in range(3):
            assert abs(got[k][0] - exact0[k]) <= 1e-12*scale[k] + 1e-300, (k, got[k][0], exact0[k], scale[k])


@pytest.mark.parametrize("n", [20000, 300001])
@pytest.mark.parametrize("base", ["zero", "gravity"])
def test_tolerance_mode_jacobi_frame(built, monkeypatch, n, base):
    """REBX_B200_MATH=tolerance in the reference's default frame of the damping forces (JACOBI): the centre-of-mass
    pipeline with jacobi_sweep_tol_kernel (kernels_tol.cu) -- the interior masses still carry the reference's own
    rounding (mass grid), the forces are damping_fast with per-particle body-mass constants.  Norm-wise inside the
    north-star tolerance against the O(N) oracle (the reference's own COM recurrence, pinned against its O(N^2) loops
    at N <= 2*10^4 by tests/test_oracle.py), and against the reference's loops themselves at N = 2*10^4."""
    from oracle import oracle
    from reboundx_b200 import workloads
    w = workloads.planetesimal_disk(n)
    acc0 = acc_variants(w, 41)[base]
    exact = _com_device(w, TRIO, "JACOBI", acc0)
    monkeypatch.setenv("REBX_B200_MATH", "tolerance")
    got = _com_device(w, TRIO, "JACOBI", acc0)
    assert not errors(got, exact)["bit_identical"], "REBX_B200_MATH=tolerance did not select the tolerance Jacobi sweep"
    lin, _ = oracle.port_apply(w, TRIO[::-1], acc0, "JACOBI", com="linear")
    e = errors(got, lin)
    assert e["max_normwise"] <= TOL, e
    assert all(np.isfinite(a).all() for a in got)
    # the per-component outliers are those of the PARTICLE-frame tolerance sweep (components far below |a|)
    assert e["frac_component_above_tol"] <= 1e-2, e
    if n <= 20000:
        want, _ = oracle_acc(w, TRIO, acc0, "JACOBI")
        er = errors(got, want)
        assert er["max_normwise"] <= TOL, er
    ee = errors(got, exact)
    assert ee["max_normwise"] <= TOL, ee


def test_tolerance_mode_resident_full_size(built):
    """device.Shard(math="tolerance") at the configurations' full sizes: finite everywhere, and within the tolerance of
    the bit-exact sweep on every particle (norm-wise; the exact sweep itself is pinned to the oracle above)."""
    import torch
    from reboundx_b200 import device, workloads
    for name, (n, add, _) in FULL.items():
        w = getattr(workloads, name)(n)
        acc0 = workloads.gravity_like_acc(w, np.random.default_rng(98))
        res = {}
        for math in ("exact", "tolerance"):
            sh = device.Shard(w, add[::-1], "cuda:0", acc0=acc0, math=math, standalone=True)
            sh.evaluate()
            torch.cuda.synchronize()
            res[math] = sh.acc()
            del sh
        e = errors(res["tolerance"], res["exact"])
        assert all(np.isfinite(a).all() for a in res["tolerance"])
        assert e["max_normwise"] <= TOL, (name, e)
        assert not e["bit_identical"], name


# ---- BASELINE.json full sizes: size-independent properties ---------------------------------------------
def _subsample(w, idx, per_particle):
    """Workload restricted to the central body + the particles `idx` (global indices >= 1)."""
    sub = {}
    for k, v in w.items():
        if isinstance(v, np.ndarray) and v.shape[0] == w["x"].shape[0]:
            sub[k] = np.concatenate([v[:1], v[idx]])
        elif isinstance(v, np.ndarray) and k in per_particle:
            sub[k] = v[idx - 1]
        else:
            sub[k] = v
    return sub


FULL = {
    "debris_disk": (10_000_000, ["radiation_forces"], ["beta"]),
    "circumplanetary_ring": (10_000_000, ["gravitational_harmonics", "gr_potential"], []),
    "asteroid_ensemble": (1_000_000, ["yarkovsky_effect", "gr_potential"],
                          ["ye_body_density", "ye_rotation_period", "ye_thermal_inertia", "ye_albedo", "ye_emissivity", "ye_k",
                           "ye_spin_axis_x", "ye_spin_axis_y", "ye_spin_axis_z", "ye_flag"]),
    "planetesimal_disk": (1_000_000, TRIO, ["tau_a", "tau_e", "tau_inc", "d_factor"]),
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_against_oracle_on_a_random_subsample(built, name):
    """At the configuration's full size the oracle cannot be run on everything in seconds, but every
    particle's force depends only on itself and the central body: the oracle evaluated on
    (body + 50 000 random particles) must reproduce the full device result at those indices --
    bit for bit where the arithmetic is + - * / sqrt, to the tolerance otherwise.  Massless
    ensembles only (a massive ensemble couples through the back-reaction sum, tested at small N)."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    n, add, per_particle = FULL[name]
    w = getattr(workloads, name)(n)
    w["m"][1:] = 0.0 if name in ("debris_disk", "circumplanetary_ring") else w["m"][1:]
    rng = np.random.default_rng(99)
    acc0 = workloads.gravity_like_acc(w, rng)
    sh = device.Shard(w, add[::-1], "cuda:0", acc0=acc0)
    sh.launch()
    torch.cuda.synchronize()
    got = sh.acc()
    idx = np.sort(rng.choice(np.arange(1, n + 1), size=50_000, replace=False))
    sub = _subsample(w, idx, per_particle)
    sub_acc0 = [np.concatenate([a[:1], a[idx]]) for a in acc0]
    want, _ = oracle.port_apply(sub, add[::-1], sub_acc0)
    g = [a[idx] for a in got]
    ww = [a[1:] for a in want]
    e = errors(g, ww)
    if name in ("debris_disk", "circumplanetary_ring"):
        assert e["bit_identical"], e
    else:
        guard(f"full_size_subsample/{name}", e)
    assert all(np.isfinite(a).all() for a in got)


def test_fused_wisdom_holman_sweep_with_yarkovsky(built):
    """rebx_b200_wh_kick_drift's second instantiation (yarkovsky_effect alone: all three modes in the ensemble):
    the fused kick + drift sweep gives the bits of the separate kernels, and follows the host harness stepping the
    reference-compiled force to the tolerance of criterion (ii)."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    w = workloads.asteroid_ensemble(3000)
    dt, steps = 0.01, 300
    out = []
    for fused in (True, False):
        sh = device.Shard(w, ["yarkovsky_effect"], "cuda:0")
        sh.wh_steps(steps, dt, w["G"], fused=fused)
        torch.cuda.synchronize()
        out.append(np.array(sh.posvel()))
    assert np.array_equal(out[0], out[1])
    if oracle.have_ref():
        ref = _trajectory(w, oracle.ref_lib(), steps, "wisdom_holman_merged", dt, 1, ["yarkovsky_effect"])
        _assert_trajectories_agree(out[0], ref)


@pytest.mark.parametrize("which", ["debris_disk", "asteroid_ensemble"])
def test_tolerance_mode_fused_wisdom_holman_sweep(built, which):
    """device.Shard(math="tolerance").wh_steps: the fused kick + Kepler drift sweep in tolerance arithmetic
    (wh_kick_drift_tol_kernel: FMA, Newton reciprocals, the force's reciprocal distance reused by the drift) follows the
    bit-exact sweep -- which IS the host harness + reference force for radiation -- to criterion (ii) after 10^3 steps,
    and really is another kernel (results differ in the last bits)."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    if which == "debris_disk":
        w, add, dt, steps = workloads.debris_disk(20000, seed=3), ["radiation_forces"], 1e8, 1000
    else:
        w, add, dt, steps = workloads.asteroid_ensemble(3000), ["yarkovsky_effect"], 0.01, 300
    out = {}
    for math in ("exact", "tolerance"):
        sh = device.Shard(w, add, "cuda:0", math=math)
        sh.wh_steps(steps, dt, w["G"])
        torch.cuda.synchronize()
        out[math] = np.array(sh.posvel())
        del sh
    assert np.isfinite(out["tolerance"]).all()
    assert not np.array_equal(out["tolerance"], out["exact"]), "math='tolerance' did not select the tolerance sweep"
    _assert_trajectories_agree(out["tolerance"], out["exact"])
    if oracle.have_ref():
        ref = _trajectory(w, oracle.ref_lib(), steps, "wisdom_holman_merged", dt, 1, add)
        _assert_trajectories_agree(out["tolerance"], ref)


@pytest.mark.parametrize("which", ["debris_disk", "asteroid_ensemble"])
def test_tolerance_mode_fused_leapfrog_step(built, which):
    """device.Shard(math="tolerance").fused_leapfrog_step: the drift-kick-drift step in one sweep in tolerance arithmetic
    (step_tol_kernel; one effect without back-reaction) follows the bit-exact fused step to criterion (ii) after 10^3
    steps, the central body included, and really is another kernel."""
    import torch
    from reboundx_b200 import device, workloads
    if which == "debris_disk":
        w, add, dt, steps = workloads.debris_disk(20000, seed=3), ["radiation_forces"], 1e8, 1000
    else:
        w, add, dt, steps = workloads.asteroid_ensemble(3000), ["yarkovsky_effect"], 0.01, 300
    out = {}
    for math in ("exact", "tolerance"):
        sh = device.Shard(w, add, "cuda:0", math=math)
        for _ in range(steps):
            sh.fused_leapfrog_step(dt, w["G"])
        torch.cuda.synchronize()
        out[math] = np.array(sh.posvel())
        del sh
    assert np.isfinite(out["tolerance"]).all()
    assert not np.array_equal(out["tolerance"], out["exact"]), "math='tolerance' did not select the tolerance step"
    assert np.array_equal(out["tolerance"][:, 0], out["exact"][:, 0])       # the central body: same step_body arithmetic
    _assert_trajectories_agree(out["tolerance"], out["exact"])


def test_full_size_resident_wisdom_holman_against_the_host_harness_on_a_subsample(built):
    """BASELINE config 2 at its full size, "radiation_forces under WHFast": 50 resident Wisdom-Holman steps of 10^7
    grains (fused kick + Kepler drift sweeps), and -- test particles being independent -- the host harness stepping
    the reference-compiled force on a random subsample of 2000 of them must land on the same bits."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    n, steps, dt = 10_000_000, 50, 1e8
    w = workloads.debris_disk(n, seed=3)
    sh = device.Shard(w, ["radiation_forces"], "cuda:0")
    sh.wh_steps(steps, dt, w["G"])
    torch.cuda.synchronize()
    got = np.array(sh.posvel())
    idx = np.sort(np.random.default_rng(17).choice(np.arange(1, n + 1), size=2000, replace=False))
    sub = _subsample(w, idx, ["beta"])
    ref = _trajectory(sub, oracle.ref_lib(), steps, "wisdom_holman_merged", dt, 1, ["radiation_forces"])
    assert np.array_equal(got[:, np.concatenate([[0], idx])], ref)
    moved = np.abs(got[0, 1:] - w["x"][1:])
    assert np.median(moved) > 1e6                     # metres: the grains did move


def test_full_size_radiation_is_linear_in_beta_and_shard_independent(built):
    """Config 2 at 10^7: doubling every beta doubles every force exactly (a power-of-two scaling
    commutes with each rounding), and cutting the ensemble into three ragged shards changes no bit."""
    import torch
    from reboundx_b200 import device, workloads
    n = 10_000_000
    w = workloads.debris_disk(n)
    one = device.Shard(w, ["radiation_forces"], "cuda:0")
    one.launch()
    w2 = dict(w); w2["beta"] = 2.0*w["beta"]
    two = device.Shard(w2, ["radiation_forces"], "cuda:0")
    two.launch()
    torch.cuda.synchronize()
    for k in ("ax", "ay", "az"):
        assert torch.equal(two.state[k], 2.0*one.state[k])
        assert bool(torch.any(one.state[k] != 0))
    del two
    cuts = [0, 3_333_333, 6_000_001, n + 1]
    parts = []
    for lo, hi in zip(cuts, cuts[1:]):
        sh = device.Shard(w, ["radiation_forces"], "cuda:0", lo=lo, hi=hi)
        sh.launch()
        parts.append(sh)
    torch.cuda.synchronize()
    for k in ("ax", "ay", "az"):
        assert torch.equal(torch.cat([p.state[k] for p in parts]), one.state[k])


def test_full_size_host_facing_call(built):
    """The exact call bench.py's e2e leg times -- sim->additional_forces(sim) on 10^7 host AoS particles,
    chunked three-buffer pipeline -- checked against the oracle on a random subsample, bit for bit."""
    from oracle import oracle
    from reboundx_b200 import scenarios, workloads
    n = 10_000_000
    w = workloads.debris_disk(n)
    rng = np.random.default_rng(5)
    acc0 = workloads.gravity_like_acc(w, rng)
    sim, rebx, _ = scenarios.build(w, ["radiation_forces"], acc=acc0)
    sim.additional_forces()
    sim.process_messages()
    got = sim.acc()
    stats = rebx.stats()
    rebx.detach(); sim.free()
    assert stats["h2d_bytes"] >= n*128 and stats["d2h_bytes"] >= n*128 and stats["launches"] >= 3*(n >> 19)
    idx = np.sort(rng.choice(np.arange(1, n + 1), size=50_000, replace=False))
    sub = _subsample(w, idx, ["beta"])
    want, _ = oracle.port_apply(sub, ["radiation_forces"], [np.concatenate([a[:1], a[idx]]) for a in acc0])
    assert errors([a[idx] for a in got], [a[1:] for a in want])["bit_identical"]
    assert all(got[k][0] == acc0[k][0] for k in range(3))          # the source itself feels nothing


def test_raw_pointer_writes_are_seen(built):
    """The reference's documented C idiom (doc/c_quickstart.rst:49-54): keep the pointer rebx_get_param returned and
    write through it between evaluations -- no API call announces the write.  The library remembers which values'
    pointers it handed out and re-reads those (only those) at every dispatch: the new value is used at once, without
    rebx_mark_params_dirty; an untouched pointer costs no rebuild."""
    import ctypes
    from reboundx_b200 import scenarios, workloads
    w = workloads.debris_disk(3000)
    sim, rebx, handles = scenarios.build(w, ["radiation_forces"])
    lib = rebx._lib
    lib.rebx_get_param.restype = ctypes.POINTER(ctypes.c_double)
    lib.rebx_get_param.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p]
    ptr = lib.rebx_get_param(rebx._ptr, sim.particle(9).ap, b"beta")
    cptr = lib.rebx_get_param(rebx._ptr, handles["radiation_forces"]._ptr.contents.ap, b"c")
    sim.additional_forces(); sim.process_messages()
    a1 = sim.acc()[0][9]
    builds = rebx.stats()["table_builds"]
    sim.set_acc(); sim.additional_forces(); sim.process_messages()
    assert rebx.stats()["table_builds"] == builds              # nothing written: nothing rebuilt
    ptr[0] = 2.0*ptr[0]                                          # raw write through the kept pointer, no API call
    sim.set_acc(); sim.additional_forces(); sim.process_messages()
    assert sim.acc()[0][9] == 2.0*a1
    assert rebx.stats()["table_builds"] == builds + 1
    a2 = sim.acc()[0][11]
    cptr[0] = 0.5*cptr[0]                                        # a force-level parameter, same idiom
    sim.set_acc(); sim.additional_forces(); sim.process_messages()
    assert sim.acc()[0][11] != a2
    rebx.detach(); sim.free()


def test_particle_removal_invalidates_the_tables(built):
    """REBOUND calls sim->free_particle_ap for a removed particle and then moves the particles behind it, even when the
    removed one carried no parameters: every per-index table entry and cached body index shifts.  Remove a parameter-less
    particle, add one at the end (N unchanged), and compare with the oracle on the new ordering."""
    from oracle import oracle
    from reboundx_b200 import scenarios, workloads
    w = workloads.debris_disk(3000)
    sim, rebx, _ = scenarios.build(w, ["radiation_forces"])
    sim.additional_forces(); sim.process_messages()             # tables built for the original ordering
    k = 5
    # strip particle k's parameters first (then the removal hook sees a parameter-less particle: the case the
    # advisor flagged), remove it the way REBOUND does (hook, then the particles behind it shift down) ...
    lib = rebx._lib
    lib.rebx_free_ap(sim.particle_ap(k))
    sim.remove_particle(k, keep_sorted=True)
    # ... and add a particle at the end, so that N is what it was
    last = {q: w[q][-1] for q in ("x", "y", "z", "vx", "vy", "vz", "m", "r")}
    sim.add_particle(**{q: float(v)*1.01 for q, v in last.items()})
    rebx.particle_params(sim.N - 1)["beta"] = 0.3
    idx = np.r_[0:k, k + 1:w["x"].shape[0]]
    w2 = {q: (np.append(v[idx], v[-1]*1.01) if isinstance(v, np.ndarray) and v.shape[0] == w["x"].shape[0] else v) for q, v in w.items()}
    w2["beta"] = np.append(np.delete(w["beta"], k - 1), 0.3)
    sim.set_acc(); sim.additional_forces(); sim.process_messages()
    got = sim.acc()
    want, _ = oracle.port_apply(w2, ["radiation_forces"], None)
    for c in range(3):
        assert np.array_equal(got[c], want[c])
    rebx.detach(); sim.free()


def test_missing_parameters_report_like_the_reference(built):
    """Error paths of the effects (reference tests/test_gr.py:17-21 and the effects' own checks)."""
    from reboundx_b200 import scenarios, workloads
    from reboundx_b200.host import Extras, Simulation
    w = workloads.debris_disk(100)
    for name, pattern in (("radiation_forces", "speed of light"), ("gr_potential", "speed of light"), ("gr", "speed of light")):
        sim = Simulation(G=w["G"])
        sim.set_particles(w["x"], w["y"], w["z"], w["vx"], w["vy"], w["vz"], w["m"], w["r"])
        sim.N_active = 1
        rebx = Extras(sim)
        rebx.add_force(rebx.load_force(name))
        before = sim.acc()
        sim.additional_forces()
        with pytest.raises(RuntimeError, match=pattern):
            sim.process_messages()
        assert all(np.array_equal(a, b) for a, b in zip(before, sim.acc()))
        rebx.detach(); sim.free()
    w5 = workloads.planetesimal_disk(100)
    sim, rebx, _ = scenarios.build(w5, ["modify_orbits_forces"], coordinates=None)
    rebx.get_force("modify_orbits_forces").params["coordinates"] = 2          # PARTICLE, but nobody is 'primary'
    sim.additional_forces()
    with pytest.raises(RuntimeError, match="primary param was not found"):
        sim.process_messages()
    rebx.detach(); sim.free()


@pytest.mark.parametrize("effect", ["gr_potential", "gravitational_harmonics"])
def test_force_is_minus_gradient_of_the_reference_potential(built, effect):
    """The reference's conservation tests (reboundx/tests/test_conservation.py:44-56,73-87) pin these two
    effects through their potentials.  Independent of the force oracle: m_i a_i computed on the device
    must equal -dU/dx_i, with U the REFERENCE's own potential function (compiled in oracle/_ref)
    differentiated by 4th-order central differences.  Same 3-body set-up as the reference's test."""
    import ctypes
    from oracle import oracle
    from reboundx_b200 import workloads as W
    from reboundx_b200.host import Extras, Simulation
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    G = 1.0
    x1, y1, z1, vx1, vy1, vz1 = (float(v[0]) for v in W.kepler_to_cartesian(G*1.0001, np.array([1.0]), np.array([0.1]), np.array([0.0]), np.array([0.3]), np.array([0.7]), np.array([1.1])))
    x2, y2, z2, vx2, vy2, vz2 = (float(v[0]) for v in W.kepler_to_cartesian(G*1.0002, np.array([2.0]), np.array([0.1]), np.array([0.2]), np.array([1.3]), np.array([0.2]), np.array([2.5])))
    state = {"x": np.array([0.0, x1, x2]), "y": np.array([0.0, y1, y2]), "z": np.array([0.0, z1, z2]),
             "vx": np.array([0.0, vx1, vx2]), "vy": np.array([0.0, vy1, vy2]), "vz": np.array([0.0, vz1, vz2]),
             "m": np.array([1.0, 1e-4, 1e-4])}

    def build(lib, pos):
        sim = Simulation(G=G)
        sim.set_particles(pos["x"], pos["y"], pos["z"], state["vx"], state["vy"], state["vz"], state["m"])
        rebx = Extras(sim, lib=lib)
        f = rebx.load_force(effect)
        if effect == "gr_potential":
            f.params["c"] = 1.0e2
        else:
            p0 = rebx.particle_params(0)
            p0["J2"] = 1e-3; p0["J4"] = 1e-3; p0["R_eq"] = 1e-1; p0["Omega"] = (0.2, -0.1, 1.0)
        rebx.add_force(f)
        return sim, rebx, f

    ref = oracle.ref_lib()
    ref.rebx_gr_potential_potential.restype = ctypes.c_double
    ref.rebx_gravitational_harmonics_potential.restype = ctypes.c_double

    def potential(pos):
        sim, rebx, f = build(ref, pos)
        U = ref.rebx_gr_potential_potential(rebx._ptr, f._ptr) if effect == "gr_potential" else ref.rebx_gravitational_harmonics_potential(rebx._ptr)
        rebx.detach(); sim.free()
        return U

    sim, rebx, _ = build(None, state)
    sim.additional_forces(); sim.process_messages()
    acc = sim.acc()
    rebx.detach(); sim.free()
    h = 1e-4
    for i in range(3):
        for c, key in enumerate(("x", "y", "z")):
            def U(delta):
                pos = {k: state[k].copy() for k in ("x", "y", "z")}
                pos[key][i] += delta
                return potential(pos)
            dU = (-U(2*h) + 8*U(h) - 8*U(-h) + U(-2*h))/(12*h)
            force = state["m"][i]*acc[c][i]
            scale = max(abs(state["m"][i]*a[i]) for a in acc)
            assert abs(force + dU) <= 2e-7*scale + 1e-18, (i, key, force, -dU)


@pytest.mark.parametrize("scheme", ["euler", "rk2", "rk4", "implicit_midpoint"])
def test_device_integrate_force_against_the_reference_operator(built, scheme):
    """the reference's integrate_force operator and its four schemes
    (src/integrate_force.c, src/integrator_*.c, compiled unmodified into oracle/_ref together with the
    reference's radiation loop) against rebx_b200_integrate_force on a resident shard."""
    import torch
    from oracle import oracle
    from reboundx_b200 import device, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    w = workloads.debris_disk(20001)
    w["beta"] = w["beta"]*50.0                  # a strong force, so every stage matters
    dt = 3.0e8
    (vx, vy, vz), msgs = oracle.ref_integrate_force(w, "radiation_forces", dt, scheme, nsteps=3)
    assert not [m for k, m in msgs if k == "error"]
    sh = device.Shard(w, ["radiation_forces"], "cuda:0")
    for _ in range(3):
        sh.integrate_force(dt, scheme)
    torch.cuda.synchronize()
    got = sh.posvel()
    assert np.array_equal(np.array(got[:3]), np.array([w["x"], w["y"], w["z"]]))       # positions untouched
    want = np.array([vx, vy, vz])
    if scheme == "implicit_midpoint":
        # the convergence test sums over all particles; the device sums in another order, so the
        # iteration may stop one step apart in the last bit
        assert np.max(np.abs(np.array(got[3:]) - want)/np.max(np.abs(want))) <= 1e-15
    else:
        assert np.array_equal(np.array(got[3:]), want)
    assert np.any(want != np.array([w["vx"], w["vy"], w["vz"]]))


# ---- same-shape neighbouring forces --------------------------------------------
@pytest.mark.parametrize("gamma", [-1.0, -1.7, 2.0])
def test_central_force(built, gamma):
    """reference src/central_force.c: a = A r^(gamma-1) d from a body carrying Acentral and gammacentral,
    with recoil of that body.  pow() is not bit-reproducible between libdevice and glibc."""
    from reboundx_b200 import workloads
    w = workloads.circumplanetary_ring(30001, massive=True)
    w["Acentral"], w["gammacentral"] = 3.0e-9*(1e8)**(-1.7 - gamma), gamma
    acc0 = acc_variants(w, 12)["gravity"]
    want, br = oracle_acc(w, ["central_force"], acc0)
    got, _ = device_acc(w, ["central_force"], acc0)
    e = errors(got, want, skip=(0,))
    guard(f"central_force/{gamma}", e, frac=4e-5)
    exact = np.array([acc0[k][0] for k in range(3)]) + br["central_force"][:3]
    scale = br["central_force"][3:] + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= 1e-12*scale[k]


@pytest.mark.parametrize("coordinates", ["PARTICLE", "JACOBI", "BARYCENTRIC"])
def test_exponential_migration(built, coordinates):
    """reference src/exponential_migration.c under rebx_com_force, all three coordinate systems;
    the force depends on sim->t."""
    from reboundx_b200 import workloads
    n = 3000
    w = workloads.planetesimal_disk(n)
    w.update(em_tau_a=np.full(n, 2.0e4), em_aini=np.full(n, 1.0), em_afin=np.full(n, 1.5), t=321.5)
    acc0 = acc_variants(w, 13)["gravity"]
    want, _ = oracle_acc(w, ["exponential_migration"], acc0, coordinates)
    got, _ = device_acc(w, ["exponential_migration"], acc0, coordinates)
    e = errors(got, want, skip=(0,) if coordinates == "BARYCENTRIC" else ())
    assert e["max_normwise"] <= TOL, e
    assert np.max(np.abs(np.array(got) - np.array(acc0))) > 0


def test_lense_thirring_bit_exact(built):
    """reference src/lense_thirring.c: + - * / sqrt only, so per-particle results are bit-identical."""
    from reboundx_b200 import workloads
    w = workloads.asteroid_ensemble(30001)
    w.update(lt_c=w["c"], I=0.07*1.0*(4.65e-3)**2, Omega=(0.3, -0.4, 90.0))
    acc0 = acc_variants(w, 14)["gravity"]
    want, br = oracle_acc(w, ["lense_thirring"], acc0)
    got, _ = device_acc(w, ["lense_thirring"], acc0)
    assert errors(got, want, skip=(0,))["bit_identical"]
    exact = np.array([acc0[k][0] for k in range(3)]) + br["lense_thirring"][:3]
    scale = br["lense_thirring"][3:] + np.abs([acc0[k][0] for k in range(3)])
    for k in range(3):
        assert abs(got[k][0] - exact[k]) <= 1e-12*scale[k]
    assert np.any(np.array(got)[:, 1:] != np.array(acc0)[:, 1:])


@pytest.mark.parametrize("which", ["gr_potential", "gravitational_harmonics", "central_force", "central_force_log"])
def test_device_potentials_match_the_reference(built, which):
    """rebx_gr_potential_potential / rebx_gravitational_harmonics_potential /
    rebx_central_force_potential reduced on the device, against the reference's functions in oracle/_ref.
    The reference sums sequentially; the tolerance is relative to the sum of |terms| (all terms of these
    potentials have one sign for gr_potential / central_force, so that is |H| itself there)."""
    import ctypes
    from oracle import oracle
    from reboundx_b200 import scenarios, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    w = workloads.circumplanetary_ring(20001, massive=True, tilted=True)
    name = "central_force" if which.startswith("central") else which
    if name == "central_force":
        w["Acentral"], w["gammacentral"] = 3.0e-9, (-1.0 if which.endswith("log") else -1.7)
    vals = []
    for lib in (oracle.ref_lib(), None):
        sim, rebx, handles = scenarios.build(w, [name], lib=lib)
        L = rebx._lib
        for fn in ("rebx_gr_potential_potential", "rebx_gravitational_harmonics_potential", "rebx_central_force_potential"):
            getattr(L, fn).restype = ctypes.c_double
        if name == "gr_potential":
            H = L.rebx_gr_potential_potential(rebx._ptr, handles[name]._ptr)
        elif name == "gravitational_harmonics":
            H = L.rebx_gravitational_harmonics_potential(rebx._ptr)
        else:
            H = L.rebx_central_force_potential(rebx._ptr)
        sim.process_messages()
        vals.append(H)
        rebx.detach(); sim.free()
    ref, got = vals
    assert ref != 0.0
    # harmonics terms change sign with latitude; their absolute sum is bounded by ~|J2 term| sum ~ a few |H|
    assert abs(got - ref) <= (1e-13 if name != "gravitational_harmonics" else 1e-11)*abs(ref), (ref, got)


@pytest.mark.parametrize("name", ["gr_potential", "central_force", "gravitational_harmonics"])
def test_reference_conservation_tests(built, name):
    """The reference's own test file for these effects, reboundx/tests/test_conservation.py:44-87, ported to this
    library: star + two 1e-4 planets (a = 1, e = 0.1; a = 2, e = 0.1, inc = 0.2), the force on the device, and
    H = Newtonian energy + the effect's potential (reduced on the device) conserved to the reference's bar of
    1e-12 -- under the fixed-step Gauss-Radau scheme standing in for IAS15, over 300 time units (48 inner orbits; the
    reference integrates 10^4; 10^3 was run once with the same result, 300 keeps the GPU suite short)."""
    import ctypes
    from reboundx_b200 import workloads as W
    from reboundx_b200.host import Extras, Simulation
    G = 1.0
    p1 = [float(v[0]) for v in W.kepler_to_cartesian(G*1.0001, np.array([1.0]), np.array([0.1]), np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
    p2 = [float(v[0]) for v in W.kepler_to_cartesian(G*1.0002, np.array([2.0]), np.array([0.1]), np.array([0.2]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
    m = np.array([1.0, 1e-4, 1e-4])
    st = [np.array([0.0, p1[k], p2[k]]) for k in range(6)]
    for k in range(6):                              # sim.move_to_com()
        st[k] = st[k] - (m*st[k]).sum()/m.sum()
    sim = Simulation(G=G)
    sim.set_particles(st[0], st[1], st[2], st[3], st[4], st[5], m)
    rebx = Extras(sim)
    force = rebx.load_force(name)
    rebx.add_force(force)
    L = rebx._lib
    if name == "gr_potential":
        force.params["c"] = 1.0e4
        L.rebx_gr_potential_potential.restype = ctypes.c_double
        potential = lambda: L.rebx_gr_potential_potential(rebx._ptr, force._ptr)
    elif name == "central_force":
        p0 = rebx.particle_params(0)
        p0["Acentral"] = 1.0e-4; p0["gammacentral"] = -1.0
        L.rebx_central_force_potential.restype = ctypes.c_double
        potential = lambda: L.rebx_central_force_potential(rebx._ptr)
    else:
        p0 = rebx.particle_params(0)
        p0["J2"] = 1.0e-3; p0["J4"] = 1.0e-3; p0["R_eq"] = 1.0e-3
        L.rebx_gravitational_harmonics_potential.restype = ctypes.c_double
        potential = lambda: L.rebx_gravitational_harmonics_potential(rebx._ptr)

    def energy():                                   # sim.energy(): kinetic + pairwise Newtonian
        x, y, z, vx, vy, vz = (np.array(a) for a in sim.posvel())
        E = 0.5*np.sum(m*(vx*vx + vy*vy + vz*vz))
        for i in range(3):
            for j in range(i + 1, 3):
                E -= G*m[i]*m[j]/np.sqrt((x[i] - x[j])**2 + (y[i] - y[j])**2 + (z[i] - z[j])**2)
        return E

    H0 = energy() + potential()
    period = 2*np.pi/np.sqrt(G*1.0001)
    nsteps = int(round(3.0e2/(period/100)))
    sim.dt = 3.0e2/nsteps
    sim.integrate(nsteps, "gauss_radau")
    H = energy() + potential()
    sim.process_messages()
    rebx.detach(); sim.free()
    assert abs((H - H0)/H0) < 1.0e-12, (H0, H)


def _two_planets():
    """reference reboundx/tests/test_machine_independent.py:13-19 / test_conservation.py:8-14"""
    from reboundx_b200 import workloads as W
    G = 1.0
    p1 = [float(v[0]) for v in W.kepler_to_cartesian(G*1.0001, np.array([1.0]), np.array([0.1]), np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
    p2 = [float(v[0]) for v in W.kepler_to_cartesian(G*1.0002, np.array([2.0]), np.array([0.1]), np.array([0.2]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
    m = np.array([1.0, 1e-4, 1e-4])
    st = [np.array([0.0, p1[k], p2[k]]) for k in range(6)]
    return [a - (m*a).sum()/m.sum() for a in st], m


@pytest.mark.parametrize("name", ["modify_orbits_forces", "gr", "gr_potential", "radiation_forces", "central_force", "gravitational_harmonics"])
def test_reference_machine_independent_tests(built, tmp_path, name):
    """reference reboundx/tests/test_machine_independent.py (one test per effect): set the effect up, save the
    REBOUNDx state, integrate; then rebuild the extras from the saved file on a fresh copy of the initial
    simulation, integrate again, and require particles[0].x to be EQUAL.  Here: this library's binary writer and
    loader, the CUDA force, the Wisdom-Holman stand-in for the reference's WHFast (dt = P/100, 2*10^3 time units)."""
    from reboundx_b200.host import Extras, Simulation
    st, m = _two_planets()

    def fresh():
        sim = Simulation(G=1.0, integrator="whfast")
        sim.set_particles(st[0], st[1], st[2], st[3], st[4], st[5], m)
        sim.dt = 2*np.pi/np.sqrt(1.0001)/100
        return sim

    sim = fresh()
    rebx = Extras(sim)
    force = rebx.load_force(name)
    rebx.add_force(force)
    ps = [rebx.particle_params(i) for i in range(3)]
    if name == "modify_orbits_forces":
        ps[1]["tau_a"] = -1e4; ps[1]["tau_e"] = -1e3; ps[2]["tau_e"] = -1e3; ps[2]["tau_inc"] = -1e3
    elif name in ("gr", "gr_potential"):
        force.params["c"] = 1.0e4
    elif name == "radiation_forces":
        force.params["c"] = 1.0e4; ps[1]["beta"] = 0.5; ps[2]["beta"] = 0.3
    elif name == "central_force":
        ps[0]["Acentral"] = 1.0e-3; ps[0]["gammacentral"] = -1.0
    else:
        ps[0]["J2"] = 1.0e-3; ps[0]["J4"] = 1.0e-3; ps[0]["R_eq"] = 1.0e-3
    path = tmp_path / (name + ".rebx")
    rebx.save(path)
    nsteps = int(2.0e3/sim.dt)
    sim.integrate(nsteps, "wisdom_holman")
    errs = [msg for k, msg in sim.messages() if k == "error"]
    assert not errs, errs[:2]
    first = np.array(sim.posvel())
    rebx.detach(); sim.free()

    sim2 = fresh()
    rebx2 = Extras(sim2, str(path))
    assert rebx2.force_order() == [name]
    sim2.integrate(nsteps, "wisdom_holman")
    second = np.array(sim2.posvel())
    rebx2.detach(); sim2.free()
    assert second[0][0] == first[0][0]                  # the reference's assertion: particles[0].x
    assert np.array_equal(first, second)                # and everything else
    x0 = np.array(st[:3])[:, 1:]
    assert np.max(np.abs(first[:3, 1:] - x0)) > 1e-3    # the planets did move


def test_reference_gr_nactive(built):
    """reference reboundx/tests/test_gr.py:62-92 (test_Nactive): star + 1e-3 planet + two massless bodies under `gr`;
    with N_active unset every body is "active" (here: the host part of this library's gr), with N_active = 2 the two
    massless ones are test particles (here: the device sweep); after 100 time units particles[3].x must agree to
    1e-12.  IAS15 -> the fixed-step Gauss-Radau stand-in, dt = P/100."""
    import gr_notebook as nb
    from reboundx_b200 import workloads as W
    from reboundx_b200.host import Extras, Simulation
    m = np.array([1.0, 1.0e-3, 0.0, 0.0])
    cols = [np.zeros(4) for _ in range(6)]
    msum = 1.0
    for i, a in ((1, 1.0), (2, 2.0), (3, 3.0)):          # sim.add(a=..., e=0.1): about the centre of mass of what is there
        el = [float(v[0]) for v in W.kepler_to_cartesian(msum + m[i], np.array([a]), np.array([0.1]), np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
        com = [(m[:i]*cols[k][:i]).sum()/msum for k in range(6)]
        for k in range(6):
            cols[k][i] = com[k] + el[k]
        msum += m[i]
    for k in range(6):
        cols[k] = cols[k] - (m*cols[k]).sum()/m.sum()     # move_to_com
    out = []
    for n_active in (None, 2):
        sim = Simulation(G=1.0)
        sim.set_particles(cols[0], cols[1], cols[2], cols[3], cols[4], cols[5], m)
        rebx = Extras(sim)
        gr = rebx.load_force("gr"); rebx.add_force(gr); gr.params["c"] = nb.C
        if n_active is not None:
            sim.N_active = n_active
        period = 2*np.pi/np.sqrt(1.001)
        nsteps = int(round(100.0/(period/100)))
        sim.dt = 100.0/nsteps
        sim.integrate(nsteps, "gauss_radau")
        sim.process_messages()
        out.append(np.array(sim.posvel())[0][3])
        launches = rebx.stats()["launches"]
        rebx.detach(); sim.free()
    assert launches > 0
    assert abs((out[1] - out[0])/out[0]) < 1.0e-12, out


def test_reference_gr_simple(built):
    """reference reboundx/tests/test_rebx.py:29-34: gr with c = 100 turns the pericentre by more than 1e-4 in 10 time units."""
    from reboundx_b200 import workloads as W
    from reboundx_b200.host import Extras, Simulation
    p1 = [float(v[0]) for v in W.kepler_to_cartesian(1.0, np.array([1.0]), np.array([0.2]), np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([0.0]))]
    sim = Simulation(G=1.0)
    sim.set_particles(*[np.array([0.0, p1[k]]) for k in range(6)], np.array([1.0, 0.0]))
    rebx = Extras(sim)
    gr = rebx.load_force("gr"); rebx.add_force(gr); gr.params["c"] = 1.0e2
    sim.dt = 2*np.pi/200
    sim.integrate(int(10.0/sim.dt), "gauss_radau")
    sim.process_messages()
    x, y, z, vx, vy, vz = (np.array(a) for a in sim.posvel())
    d = np.array([x[1] - x[0], y[1] - y[0], z[1] - z[0]]); v = np.array([vx[1] - vx[0], vy[1] - vy[0], vz[1] - vz[0]])
    ev = (v.dot(v) - 1.0/np.linalg.norm(d))*d - d.dot(v)*v
    assert np.arctan2(ev[1], ev[0]) > 1.0e-4
    rebx.detach(); sim.free()


def test_gas_dynamical_friction(built):
    """reference src/gas_dynamical_friction.c (units of its example: G = 4 pi^2, disc around a central mass)."""
    from reboundx_b200 import workloads
    w = workloads.asteroid_ensemble(30001)
    w["m"][1:] = 10.0**np.random.default_rng(2).uniform(-8, -5, 30001)
    w["r"][1:] = 1e-5
    w.update(gas_df_rhog=0.2, gas_df_alpha_rhog=-1.5, gas_df_cs=0.6, gas_df_alpha_cs=-0.5, gas_df_xmin=0.045, gas_df_hr=0.01, gas_df_Qd=5.0)
    acc0 = acc_variants(w, 15)["gravity"]
    want, _ = oracle_acc(w, ["gas_dynamical_friction"], acc0)
    got, _ = device_acc(w, ["gas_dynamical_friction"], acc0)
    e = errors(got, want)
    guard("gas_dynamical_friction", e, frac=4e-5)
    assert np.any(np.array(got)[:, 1:] != np.array(acc0)[:, 1:])


def _compile_c(tmp_path, source, name):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / name)
    lib = os.path.join(root, "reboundx_b200", "lib")
    shim = os.path.join(root, "reboundx_b200", "rebound_shim")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O2", "-I", os.path.join(root, "include"), "-I", shim,
           os.path.join(root, "tests", source), "-L", lib, "-lreboundx_b200", "-L", shim, "-lrebound_shim", "-lm",
           f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{shim}", "-o", exe]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    return exe


def test_c_program_multi_device_dispatch(built, tmp_path):
    """tests/c_multi_device.c: a plain C program evaluates config 2 and the massive ring of config 3 through
    sim->additional_forces on ONE device and spread over ALL visible devices (REBX_B200_DEVICES=all) -- test particles
    bit for bit, the central body's back-reaction to rounding.  With a single GPU the second run degenerates to the
    first (still exercises the partitioned code path's bookkeeping)."""
    import json
    import subprocess
    exe = _compile_c(tmp_path, "c_multi_device.c", "c_multi_device")
    run = subprocess.run([exe, str((1 << 21) + 37), "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    rep = json.loads(run.stdout.strip().splitlines()[-1])
    assert rep["ok"] and rep["debris_disk"]["differing_components"] == 0 and rep["ring_massive"]["differing_components"] == 0, rep


def test_c_program_resident_session(built, tmp_path):
    """tests/c_resident_session.c: BASELINE config 1a integrated for 10^3 steps by a resident session driven from C
    (rebx_b200_session_step_wh / _step_leapfrog: one upload, one download) -- bit for bit the trajectory of the host
    harness that calls sim->additional_forces every step; the session's evaluation against the dispatcher's; and the
    reference's operator entry rebx_load_operator("integrate_force") + rebx_add_operator against the same scheme by hand."""
    import json
    import subprocess
    exe = _compile_c(tmp_path, "c_resident_session.c", "c_resident_session")
    run = subprocess.run([exe, "1000", "1000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    rep = json.loads(run.stdout.strip().splitlines()[-1])
    assert rep["ok"], rep
    assert rep["wisdom_holman_merged"]["differing_coordinates"] == 0 and rep["leapfrog"]["differing_coordinates"] == 0, rep
    assert rep["evaluate"]["differing"] == 0 and rep["operator_integrate_force_rk4"]["differing_coordinates"] == 0, rep


@pytest.mark.parametrize("scheme", ["euler", "rk2", "rk4", "implicit_midpoint"])
@pytest.mark.parametrize("force,coordinates", [("radiation_forces", "PARTICLE"), ("modify_orbits_forces", "JACOBI")])
def test_operator_api_integrate_force_against_the_reference_operator(built, scheme, force, coordinates):
    """The reference's operator entry points, driven identically on both libraries: rebx_load_operator("integrate_force"),
    "force" / "integrator" parameters on the operator, rebx_integrate_force(sim, operator, dt) -- the product's runs
    the scheme on the device for a device-backed force about particle 0 (radiation_forces) and the host schemes
    around device sweeps otherwise (modify_orbits_forces in the Jacobi frame)."""
    from oracle import oracle
    from reboundx_b200 import _lib, workloads
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref")
    if force == "radiation_forces":
        w = workloads.debris_disk(20001)
        w["beta"] = w["beta"]*50.0
        dt = 3.0e8
    else:
        w = workloads.planetesimal_disk(2000)
        w["tau_a"] = np.full(2000, -1e3); w["tau_e"] = np.full(2000, -1e2); w["tau_inc"] = np.full(2000, -1e2)
        dt = 0.05
    want, msgs = oracle.ref_integrate_force(w, force, dt, scheme, nsteps=3, coordinates=coordinates)
    assert not [m for k, m in msgs if k == "error"]
    got, msgs2 = oracle.ref_integrate_force(w, force, dt, scheme, nsteps=3, lib=_lib.load(), coordinates=coordinates)
    assert not [m for k, m in msgs2 if k == "error"], msgs2
    want, got = np.array(want), np.array(got)
    assert np.any(want != np.array([w["vx"], w["vy"], w["vz"]]))
    if force == "radiation_forces" and scheme != "implicit_midpoint":
        assert np.array_equal(got, want)
    else:
        # implicit midpoint: the convergence sums run in another order on the device; Jacobi frame: O(N) scans vs the
        # reference's O(N^2) loops -- tolerance relative to the largest velocity component
        assert np.max(np.abs(got - want)/np.max(np.abs(want))) <= 1e-14

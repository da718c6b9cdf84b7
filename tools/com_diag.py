"""Diagnostic (GPU box): where do device and oracle differ per component in the centre-of-mass frames?"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle
from parity import errors
from reboundx_b200 import scenarios, workloads as W
TRIO = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]

def dev(w, add, coords, acc0):
    sim, rebx, _ = scenarios.build(w, add, coordinates=coords, acc=acc0)
    sim.additional_forces(); sim.messages()
    got = sim.acc(); rebx.detach(); sim.free()
    return got

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for coords in ("BARYCENTRIC", "JACOBI"):
    w = W.planetesimal_disk(n)
    if coords == "BARYCENTRIC":
        w["d_factor"] = np.full(n, 5.0)
    for base in ("zero", "gravity"):
        acc0 = None if base == "zero" else W.gravity_like_acc(w, np.random.default_rng(21))
        lin, _ = oracle.port_apply(w, TRIO[::-1], acc0, coords, com="linear")
        got = dev(w, TRIO, coords, acc0)
        g = np.array(got); l = np.array(lin)
        print(coords, base, "device vs linear:", errors(got, lin, skip=(0,)))
        for add in (["modify_orbits_forces"], ["gas_damping_timescale"], ["type_I_migration"]):
            l1, _ = oracle.port_apply(w, add, acc0, coords, com="linear")
            g1 = dev(w, add, coords, acc0)
            print("    ", add[0], errors(g1, l1, skip=(0,)))
        d = np.abs(g - l)[:, 1:]; ref = np.abs(l)[:, 1:]
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.where(d == 0, 0, d/ref)
        worst = np.argsort(rel.ravel())[-5:]
        for idx in worst:
            c, i = np.unravel_index(idx, rel.shape)
            print("   worst: comp", c, "particle", i + 1, "rel %.3g" % rel[c, i], "abs %.3g" % d[c, i], "a", l[:, i + 1],
                  "a0", None if acc0 is None else [acc0[k][i + 1] for k in range(3)])

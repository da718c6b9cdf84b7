"""Parity report (test infrastructure; uses the oracle): per configuration, the measured error metrics of
the CUDA path against the reference-compiled oracle on identical seeded inputs, through the
host-facing call.  Run on the GPU box:   python tests/parity_report.py > gpurun_out/parity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle                                     # noqa: E402
from parity import errors                                     # noqa: E402
from reboundx_b200 import scenarios, workloads as W           # noqa: E402

TRIO = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]
CASES = [
    ("config1a rad_forces_debris_disk (N=1001)", lambda: W.debris_disk(1000, seed=3), ["radiation_forces"], "PARTICLE", None),
    ("config2 debris disk, radiation_forces", lambda: W.debris_disk(200000), ["radiation_forces"], "PARTICLE", None),
    ("config3 ring, J2/J4 + gr_potential (massless)", lambda: W.circumplanetary_ring(200000, tilted=True), ["gravitational_harmonics", "gr_potential"], "PARTICLE", None),
    ("config3 ring, massive particles (back-reaction)", lambda: W.circumplanetary_ring(200000, tilted=True, massive=True), ["gravitational_harmonics", "gr_potential"], "PARTICLE", None),
    ("config4 asteroids, yarkovsky + gr_potential", lambda: W.asteroid_ensemble(200000), ["yarkovsky_effect", "gr_potential"], "PARTICLE", None),
    ("config4 asteroids, yarkovsky + gr (N_active=1)", lambda: W.asteroid_ensemble(200000), ["yarkovsky_effect", "gr"], "PARTICLE", 1),
    ("config5 planetesimals, damping trio, PARTICLE", lambda: W.planetesimal_disk(200000), TRIO, "PARTICLE", None),
    ("config5 planetesimals, damping trio, JACOBI (N=20000)", lambda: W.planetesimal_disk(20000), TRIO, "JACOBI", None),
    ("config5 planetesimals, damping trio, BARYCENTRIC (N=20000)", lambda: W.planetesimal_disk(20000), TRIO, "BARYCENTRIC", None),
]

if __name__ == "__main__":
    out = []
    for name, make, add, coords, n_active in CASES:
        w = make()
        if coords == "BARYCENTRIC":
            w["d_factor"] = np.full(w["x"].shape[0] - 1, 5.0)
        for base in ("zero", "gravity"):
            acc0 = None if base == "zero" else W.gravity_like_acc(w, np.random.default_rng(1))
            want = oracle.ref_apply(w, add, acc0, coords, n_active=n_active) if oracle.have_ref() else \
                oracle.port_apply(w, add[::-1], acc0, coords, n_active=(n_active or -1))[0]
            sim, rebx, _ = scenarios.build(w, add, coordinates=coords, acc=acc0)
            if n_active is not None:
                sim.N_active = n_active
            sim.additional_forces()
            sim.messages()
            got = sim.acc()
            rebx.detach(); sim.free()
            e_all = errors(got, want)
            e_nobody = errors(got, want, skip=(0,))
            out.append({"case": name, "accumulation_base": base, "N": int(w["x"].shape[0]), "oracle": "reference" if oracle.have_ref() else "port",
                        "all_particles": e_all, "without_central_body": e_nobody})
            print(json.dumps(out[-1]), flush=True)

"""(De)serialisation of golden cases: one .npz per case, inputs + reference outputs."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, case, out, messages):
    arrays = {"out": np.array(out)}
    for k, v in case["state"].items():
        arrays["state_" + k] = np.asarray(v, dtype=np.float64)
    if case.get("acc0") is not None:
        arrays["acc0"] = np.array(case["acc0"])
    for p, (idx, val) in case.get("particle_params", {}).items():
        arrays["pp_idx_" + p] = np.asarray(idx, dtype=np.int64)
        arrays["pp_val_" + p] = np.asarray(val)
    for p, (idx, val) in case.get("particle_vec3", {}).items():
        arrays["pv_idx_" + p] = np.asarray(idx, dtype=np.int64)
        arrays["pv_val_" + p] = np.asarray(val, dtype=np.float64)
    meta = {"G": case["G"], "integrator": case.get("integrator", "ias15"), "forces": case["forces"], "messages": messages}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)


def names():
    return sorted(f[:-4] for f in os.listdir(HERE) if f.endswith(".npz"))


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    case = {"G": meta["G"], "integrator": meta["integrator"], "forces": meta["forces"],
            "state": {k[6:]: z[k] for k in z.files if k.startswith("state_")},
            "acc0": tuple(z["acc0"]) if "acc0" in z.files else None,
            "particle_params": {k[7:]: (z[k], z["pp_val_" + k[7:]]) for k in z.files if k.startswith("pp_idx_")},
            "particle_vec3": {k[7:]: (z[k], z["pv_val_" + k[7:]]) for k in z.files if k.startswith("pv_idx_")}}
    return case, z["out"], [tuple(m) for m in meta["messages"]]

"""tools/tol_sweep.py (GPU box) -- tolerance-mode sweeps: kernel time and error against the bit-exact sweep, per
BASELINE workload.  REBX_B200_LIB selects an occupancy variant (tools/tol_variants.sh), REBX_B200_TOL_V the
particles per thread.  One JSON line per workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from parity import errors
from reboundx_b200 import device, workloads as W

WL = {"debris_disk": (["radiation_forces"], 10_000_000), "circumplanetary_ring": (["gravitational_harmonics", "gr_potential"], 10_000_000),
      "asteroid_ensemble": (["yarkovsky_effect", "gr_potential"], 1_000_000),
      "planetesimal_disk": (["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"], 1_000_000)}
names = sys.argv[1:] or list(WL)
for name in names:
    add, n = WL[name]
    w = getattr(W, name)(n)
    out = {"workload": name, "lib": os.environ.get("REBX_B200_LIB", "default").split("/")[-2] if os.environ.get("REBX_B200_LIB") else "lib",
           "V": os.environ.get("REBX_B200_TOL_V", "auto")}
    res = {}
    for base in ("zero", "gravity"):
        acc0 = None if base == "zero" else W.gravity_like_acc(w, np.random.default_rng(1))
        got = {}
        for math in ("exact", "tolerance"):
            sh = device.Shard(w, add[::-1], "cuda:0", acc0=acc0, math=math, standalone=True)
            sh.launch(); torch.cuda.synchronize()
            got[math] = sh.acc()
            if base == "zero":
                for _ in range(3): sh.launch()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(20): sh.launch()
                e1.record(); torch.cuda.synchronize()
                out[math + "_ms"] = e0.elapsed_time(e1)/20
                out["frac_" + math] = sh.algorithmic_bytes/(out[math + "_ms"]*1e-3)/1e9/6541.1
            del sh
        e = errors(got["tolerance"], got["exact"], skip=(0,))
        res[base] = {k: (v if isinstance(v, bool) else float("%.3g" % v)) for k, v in e.items()}
        star = [abs(got["tolerance"][k][0] - got["exact"][k][0]) for k in range(3)]
        res[base]["star_abs"] = [float("%.3g" % s) for s in star]
    out["vs_exact"] = res
    print(json.dumps(out), flush=True)

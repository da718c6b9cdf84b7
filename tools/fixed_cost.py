"""tools/fixed_cost.py (GPU box) -- the fixed cost of one sweep launch: time per launch, back to back on one stream, for
ensembles so small that the per-particle work vanishes (n = 256 ... 10^5), every BASELINE stack in both arithmetic builds."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from reboundx_b200 import device, workloads as W

WL = {"debris_disk": ["radiation_forces"],
      "circumplanetary_ring": ["gravitational_harmonics", "gr_potential"],
      "asteroid_ensemble": ["yarkovsky_effect", "gr_potential"],
      "planetesimal_disk": ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]}
for name, add in WL.items():
    for math in ("exact", "tolerance"):
        for n in (256, 10_000, 100_000):
            w = getattr(W, name)(n)
            sh = device.Shard(w, add[::-1], "cuda:0", math=math, standalone=True)
            for _ in range(10): sh.launch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(200): sh.launch()
            e1.record(); torch.cuda.synchronize()
            print(json.dumps({"workload": name, "math": math, "n": n, "us_per_launch": round(e0.elapsed_time(e1)*1e3/200, 2)}), flush=True)
            del sh

"""tools/size_scaling.py (GPU box) -- sweep time against ensemble size: separates the per-particle cost of a sweep from
its fixed cost (launch gap, per-CTA prologue, tail imbalance, last-CTA reduction).  One line per (workload, math, n)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from reboundx_b200 import device, workloads as W

WL = {"asteroid_ensemble": ["yarkovsky_effect", "gr_potential"],
      "planetesimal_disk": ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"],
      "circumplanetary_ring": ["gravitational_harmonics", "gr_potential"]}
names = sys.argv[1:] or list(WL)
for name in names:
    for math in ("exact", "tolerance"):
        for n in (125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000):
            w = getattr(W, name)(n)
            sh = device.Shard(w, WL[name][::-1], "cuda:0", math=math, standalone=True)
            for _ in range(5): sh.launch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): sh.launch()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)/20
            print(json.dumps({"workload": name, "math": math, "n": n, "ms": round(ms, 5), "ns_per_particle": round(ms*1e6/n, 3),
                              "frac_hbm": round(sh.algorithmic_bytes/(ms*1e-3)/1e9/6541.1, 4)}), flush=True)
            del sh; torch.cuda.empty_cache()

"""The same physics for 10^7 grains with the state resident in HBM; nothing crosses PCIe between steps.

    python examples/resident_debris_disk.py [n_particles] [steps] [leapfrog|whfast]

leapfrog: every drift-kick-drift step is ONE fused sweep (rebx_b200_leapfrog_step).
whfast:   Wisdom-Holman steps -- Kepler drift about the star + kick by radiation_forces, the splitting the
          reference example integrates with (sim.integrator = "whfast") -- one fused kick + drift sweep per
          step (device.Shard.wh_steps -> rebx_b200_wh_kick_drift).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                # noqa: E402
from reboundx_b200 import device, workloads                 # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
scheme = sys.argv[3] if len(sys.argv) > 3 else "leapfrog"
w = workloads.debris_disk(n, seed=3)
shard = device.Shard(w, ["radiation_forces"], "cuda:0")
if scheme == "whfast":
    run = lambda k: shard.wh_steps(k, 1e8, w["G"])
else:
    run = lambda k: [shard.fused_leapfrog_step(1e8, w["G"]) for _ in range(k)]
run(1)                                                      # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
run(steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
x, y, z, vx, vy, vz = shard.posvel()
mu = w["G"]*w["m"][0]*(1.0 - w["beta"])                 # radiation pressure reduces the effective solar mass per grain
a = -mu/((vx[1:]**2 + vy[1:]**2 + vz[1:]**2) - 2*mu/np.sqrt(x[1:]**2 + y[1:]**2 + z[1:]**2))/1.5e11
print(f"{n} grains, {steps} {scheme} steps: {dt/steps*1e3:.3f} ms/step, {n*steps/dt:.3e} particle-steps/s; mean a (about the beta-reduced Sun) = {a.mean():.4f} AU")

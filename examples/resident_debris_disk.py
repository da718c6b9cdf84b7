"""The same physics for 10^7 grains with the state resident in HBM: every drift-kick-drift step is ONE
fused sweep (rebx_b200_leapfrog_step); nothing crosses PCIe between steps.

    python examples/resident_debris_disk.py [n_particles] [steps]
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
w = workloads.debris_disk(n, seed=3)
shard = device.Shard(w, ["radiation_forces"], "cuda:0")
shard.fused_leapfrog_step(1e8, w["G"])                      # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    shard.fused_leapfrog_step(1e8, w["G"])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
x, y, z, vx, vy, vz = shard.posvel()
mu = w["G"]*w["m"][0]*(1.0 - w["beta"])                 # radiation pressure reduces the effective solar mass per grain
a = -mu/((vx[1:]**2 + vy[1:]**2 + vz[1:]**2) - 2*mu/np.sqrt(x[1:]**2 + y[1:]**2 + z[1:]**2))/1.5e11
print(f"{n} grains, {steps} steps: {dt/steps*1e3:.3f} ms/step, {n*steps/dt:.3e} particle-steps/s; mean a (about the beta-reduced Sun) = {a.mean():.4f} AU")

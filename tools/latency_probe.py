"""Per-call latency of the host-facing path (sim->additional_forces on host AoS particles) vs ensemble size;
prints one line per size: best-of-200 wall time of one call.  REBX_B200_ZEROCOPY_MAX=0 disables the
zero-copy path of small ensembles.  Usage: python tools/latency_probe.py [sizes...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reboundx_b200 import scenarios, workloads

sizes = [int(a) for a in sys.argv[1:]] or [1000, 2000, 4000, 8000, 16000, 32000, 100000]
for n in sizes:
    w = workloads.debris_disk(n, seed=3)
    sim, rebx, _ = scenarios.build(w, ["radiation_forces"])
    for _ in range(5):
        sim.additional_forces()
    t = sim.time_additional_forces(200)
    st = rebx.stats()
    print(f"N = {n:7d}: {t*1e6:8.1f} us per call   (zero-copy max {os.environ.get('REBX_B200_ZEROCOPY_MAX', 'default')})", flush=True)
    rebx.detach(); sim.free()


# resident stepping of a small ensemble: eager launches vs a replayed CUDA graph (device.Shard.capture_steps)
import time
import torch
from reboundx_b200 import device
for n in (1000, 10000):
    w = workloads.debris_disk(n, seed=3)
    for scheme, eager in (("leapfrog", lambda sh: [sh.fused_leapfrog_step(1e8, w["G"]) for _ in range(100)]),
                          ("whfast", lambda sh: sh.wh_steps(100, 1e8, w["G"]))):
        sh = device.Shard(w, ["radiation_forces"], "cuda:0")
        eager(sh); torch.cuda.synchronize()
        t0 = time.perf_counter(); eager(sh); torch.cuda.synchronize(); te = (time.perf_counter() - t0)/100
        g = sh.capture_steps(100, 1e8, w["G"], scheme)
        g.replay(); torch.cuda.synchronize()
        t0 = time.perf_counter(); g.replay(); torch.cuda.synchronize(); tg = (time.perf_counter() - t0)/100
        print(f"resident {scheme:8s} N = {n:6d}: eager {te*1e6:6.1f} us/step, CUDA graph {tg*1e6:6.1f} us/step", flush=True)

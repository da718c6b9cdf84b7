"""Run under torchrun (one rank per GPU, NCCL): the sharded evaluation of reboundx_b200.device.Shard
-- body broadcast, back-reaction all-reduce, and the prefix/suffix exchanges of the centre-of-mass
frames -- against the same ensemble evaluated as ONE shard on rank 0's GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/nccl_sharded_check.py OUT.json

Started by tests/test_gpu_parity.py::test_nccl_sharded_evaluation when the box has >= 2 GPUs.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRIO = ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"]
CASES = [("debris_disk", ["radiation_forces"], "PARTICLE", 200_001),
         ("circumplanetary_ring_massive", ["gravitational_harmonics", "gr_potential"], "PARTICLE", 100_001),
         ("planetesimal_disk", TRIO, "PARTICLE", 100_001),
         ("planetesimal_disk", TRIO, "BARYCENTRIC", 100_001),
         ("planetesimal_disk", TRIO, "JACOBI", 100_001)]


def main(out_path):
    from parity import errors
    from reboundx_b200 import device, sharding, workloads
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    report = []
    for name, add_order, coordinates, n in CASES:
        w = workloads.circumplanetary_ring(n, massive=True) if name.endswith("_massive") else getattr(workloads, name)(n)
        acc0 = workloads.gravity_like_acc(w, np.random.default_rng(5))
        ntot = w["x"].shape[0]
        lo, hi = sharding.shard_range(ntot, world, rank)
        sh = device.Shard(w, add_order[::-1], dev, lo=lo, hi=hi, coordinates=coordinates, acc0=acc0)
        sh.evaluate(None)
        torch.cuda.synchronize()
        mine = torch.stack([sh.state[k] for k in ("ax", "ay", "az")])
        per = sharding.shard_range(ntot, world, 0)[1]
        pad = torch.zeros((3, per), dtype=torch.float64, device=dev)
        pad[:, :hi - lo] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        if rank == 0:
            got = [np.concatenate([parts[r][k, :sharding.shard_range(ntot, world, r)[1] - sharding.shard_range(ntot, world, r)[0]].cpu().numpy()
                                   for r in range(world)]) for k in range(3)]
            one = device.Shard(w, add_order[::-1], dev, coordinates=coordinates, acc0=acc0, standalone=True)
            one.evaluate()              # standalone: no collective (only rank 0 builds this shard)
            torch.cuda.synchronize()
            want = one.acc()
            e = errors(got, want, skip=(0,))
            body = errors([g[:1] for g in got], [x[:1] for x in want])
            report.append({"case": name, "coordinates": coordinates, "n": ntot, "world": world,
                           "expect": "bits" if coordinates == "PARTICLE" else "tol",
                           "bit_identical": bool(e["bit_identical"]), "max_normwise": float(e["max_normwise"]),
                           "body_max_normwise": float(body["max_normwise"])})
        dist.barrier()
    # sharded fused drift-kick-drift steps (rebx_b200_leapfrog_sweep / all-reduce / _body) vs one shard
    for name, add_order, n, nsteps, scheme in [("debris_disk", ["radiation_forces"], 200_001, 20, "leapfrog"),
                                               ("circumplanetary_ring_massive", ["gravitational_harmonics", "gr_potential"], 100_001, 20, "leapfrog"),
                                               ("debris_disk", ["radiation_forces"], 200_001, 20, "whfast"),
                                               ("circumplanetary_ring_massive", ["gravitational_harmonics", "gr_potential"], 100_001, 20, "whfast")]:
        w = workloads.circumplanetary_ring(n, massive=True) if name.endswith("_massive") else getattr(workloads, name)(n)
        ntot = w["x"].shape[0]
        dt = 1e-3*2*np.pi*np.sqrt(float(np.hypot(w["x"][1], w["y"][1]))**3/(w["G"]*w["m"][0]))
        lo, hi = sharding.shard_range(ntot, world, rank)
        sh = device.Shard(w, add_order[::-1], dev, lo=lo, hi=hi)
        if scheme == "leapfrog":
            for _ in range(nsteps):
                sh.fused_leapfrog_step(dt, w["G"])
        else:
            sh.wh_steps(nsteps, dt, w["G"])
        torch.cuda.synchronize()
        mine = torch.stack([sh.state[k] for k in ("x", "y", "z", "vx", "vy", "vz")])
        per = sharding.shard_range(ntot, world, 0)[1]
        pad = torch.zeros((6, per), dtype=torch.float64, device=dev)
        pad[:, :hi - lo] = mine
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        if rank == 0:
            sizes = [sharding.shard_range(ntot, world, r)[1] - sharding.shard_range(ntot, world, r)[0] for r in range(world)]
            got = [np.concatenate([parts[r][k, :sizes[r]].cpu().numpy() for r in range(world)]) for k in range(6)]
            one = device.Shard(w, add_order[::-1], dev, standalone=True)
            if scheme == "leapfrog":
                for _ in range(nsteps):
                    one.fused_leapfrog_step(dt, w["G"])
            else:
                one.wh_steps(nsteps, dt, w["G"])
            torch.cuda.synchronize()
            want = one.posvel()
            ep, ev = errors(got[:3], want[:3]), errors(got[3:], want[3:])
            report.append({"case": name + ("_fused_leapfrog" if scheme == "leapfrog" else "_wisdom_holman"), "coordinates": "PARTICLE", "n": ntot, "world": world, "steps": nsteps,
                           "expect": "bits" if name == "debris_disk" else "traj",
                           "bit_identical": bool(ep["bit_identical"] and ev["bit_identical"]),
                           "max_normwise": float(max(ep["max_normwise"], ev["max_normwise"])), "body_max_normwise": 0.0})
        dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(report, f)
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])

"""Multi-rank plumbing on CPU: world_size 2 over gloo.

Each rank owns an index range of one ensemble; the body record is broadcast from rank 0 and the
back-reaction sums are all-reduced, exactly the exchange reboundx_b200.device.Shard performs
over NCCL.  The per-shard arithmetic here is the oracle's C restatement (tests may use it); on
the GPU box tests/test_gpu_parity.py runs the same exchange with the CUDA kernels.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from reboundx_b200 import sharding, workloads as W
    w = W.circumplanetary_ring(n, massive=True)
    ntot = n + 1
    lo, hi = sharding.shard_range(ntot, world, rank)
    # body record: owner (rank 0 holds index 0) broadcasts position/mass; others start from garbage
    body = torch.zeros(8, dtype=torch.float64)
    if rank == sharding.owner_of(0, ntot, world):
        body[:] = torch.tensor([0.0] + [w[k][0] for k in ("x", "y", "z", "vx", "vy", "vz", "m")])
    sharding.broadcast_bodies(body, src=sharding.owner_of(0, ntot, world))
    # local ensemble = [body] + own slice (without the body itself)
    sl = slice(max(lo, 1), hi)
    loc = {k: np.concatenate([[body[1 + j].item()], w[k][sl]]) for j, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "m"))}
    loc["r"] = np.zeros(loc["x"].shape[0])
    loc.update(G=w["G"], c=w["c"])
    acc, br = oracle.port_apply(loc, ["gr_potential"], None)
    back = torch.tensor(br["gr_potential"][:3].copy())
    sharding.allreduce_backreaction(back)
    res = {k: acc[i][1:] for i, k in enumerate(("ax", "ay", "az"))}
    if rank == sharding.owner_of(0, ntot, world):
        res["body_acc"] = back.numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_evaluation(built, tmp_path):
    from oracle import oracle
    from reboundx_b200 import sharding, workloads as W
    n, world = 4001, 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    w = W.circumplanetary_ring(n, massive=True)
    want, br = oracle.port_apply(w, ["gr_potential"], None)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert parts[0]["lo"] == 0 and parts[0]["hi"] == parts[1]["lo"] and parts[1]["hi"] == n + 1
    for i, k in enumerate(("ax", "ay", "az")):
        got = np.concatenate([p[k] for p in parts])
        assert np.array_equal(got, want[i][1:])                 # per-particle results: bit-identical
    body = parts[0]["body_acc"]
    for k in range(3):                                          # body: sum of per-shard sums
        assert abs(body[k] - br["gr_potential"][k]) <= 1e-13*br["gr_potential"][3 + k]


def test_shard_ranges_cover_and_align():
    from reboundx_b200 import sharding
    for n in (0, 1, 2, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            assert all(lo % 2 == 0 for lo, hi in edges if lo < n)
            if n:
                assert sharding.owner_of(n - 1, n, world) == max(r for r, (lo, hi) in enumerate(edges) if hi > lo)


def _scan_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from reboundx_b200 import sharding
    moments = torch.arange(7, dtype=torch.float64)*(rank + 1) + 0.1*rank       # this rank's 7 mass moments
    tsum = torch.tensor([1.0, 10.0, 100.0], dtype=torch.float64)*(rank + 1)      # this rank's sum of t_i
    pre = sharding.exchange_prefix(moments)
    suf = sharding.exchange_suffix(tsum)
    tot = sharding.allreduce_sum(moments.clone())
    np.savez(os.path.join(out_dir, f"scan{rank}.npz"), pre=pre.numpy(), suf=suf.numpy(), tot=tot.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_jacobi_prefix_suffix_exchange(world, tmp_path):
    """The cross-shard scans of the sharded Jacobi frame: exclusive prefix of the
    mass moments, exclusive suffix of the back-reaction sums, over gloo."""
    mp.spawn(_scan_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    moments = [np.arange(7)*(r + 1) + 0.1*r for r in range(world)]
    tsum = [np.array([1.0, 10.0, 100.0])*(r + 1) for r in range(world)]
    for r in range(world):
        got = np.load(tmp_path / f"scan{r}.npz")
        assert np.array_equal(got["pre"], sum(moments[:r], np.zeros(7)))
        assert np.array_equal(got["suf"], sum(tsum[r + 1:][::-1], np.zeros(3)))
        assert np.allclose(got["tot"], sum(moments), rtol=1e-15)


def test_prefix_suffix_pure_functions():
    from reboundx_b200 import sharding
    rows = [torch.tensor([float(i), 2.0*i]) for i in range(1, 5)]
    assert sharding.exclusive_prefix(rows, 0).tolist() == [0.0, 0.0]
    assert sharding.exclusive_prefix(rows, 3).tolist() == [6.0, 12.0]
    assert sharding.exclusive_suffix(rows, 3).tolist() == [0.0, 0.0]
    assert sharding.exclusive_suffix(rows, 0).tolist() == [9.0, 18.0]

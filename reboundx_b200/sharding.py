"""Index-range sharding of an ensemble across ranks.

The force path has no particle-particle coupling beyond the O(1) source / primary bodies, so
ranks own contiguous index ranges and exchange only
  * the body records (broadcast from the owner, a few hundred bytes), and
  * the back-reaction sums onto those bodies (all-reduce of 3 doubles per effect).
These helpers are backend-agnostic torch.distributed calls: NCCL over NVLink on the GPU box,
gloo in the CPU tests.
"""


def shard_range(n, world, rank):
    """Contiguous range [lo, hi) of rank `rank`; sizes differ by at most one, and every
    boundary is even so SoA slices stay 16-byte aligned for double2 access."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    per = -(-n // world)
    per += per & 1
    lo = min(n, rank*per)
    hi = min(n, lo + per)
    return lo, hi


def owner_of(index, n, world):
    for r in range(world):
        lo, hi = shard_range(n, world, r)
        if lo <= index < hi:
            return r
    raise IndexError(index)


def broadcast_bodies(bodies, src=0, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(bodies, src=src, group=group)
    return bodies


def allreduce_backreaction(backreaction, group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(backreaction, group=group)
    return backreaction

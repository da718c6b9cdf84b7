"""Index-range sharding of an ensemble across ranks.

The force path has no particle-particle coupling beyond the O(1) source / primary bodies, so
ranks own contiguous index ranges and exchange only
  * the body records (broadcast from the owner, a few hundred bytes), and
  * the back-reaction sums onto those bodies (all-reduce of 3 doubles per effect).
The centre-of-mass frames of rebx_com_force add one more exchange each (section 8e rows 3, 4):
  * BARYCENTRIC: all-reduce of the 7 mass moments before the sweep and of sum (m_i/M) f_i after it;
  * JACOBI: an exclusive PREFIX over ranks of the 7 mass moments (the centre of mass interior to a
    shard's first particle) and an exclusive SUFFIX over ranks of the 3-vector sum of t_i (what every
    particle in front must lose), each an all-gather of one small row per rank.
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


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allreduce_sum(t, group=None):
    """In-place sum over ranks (mass moments, barycentric back-reaction)."""
    import torch.distributed as dist
    if _world(group)[0] > 1:
        dist.all_reduce(t, group=group)
    return t


def gather_rows(row, group=None):
    """All ranks' copies of a small tensor, in rank order (a list of world_size tensors)."""
    import torch
    import torch.distributed as dist
    world, _ = _world(group)
    if world == 1:
        return [row]
    rows = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(rows, row.contiguous(), group=group)
    return rows


def exclusive_prefix(rows, rank):
    """Sum of rows[0..rank) added front to back (rank 0: zeros).  Fixed order => every rank that
    recomputes a neighbour's carry gets the same bits."""
    out = rows[0].new_zeros(rows[0].shape)
    for q in range(rank):
        out = out + rows[q]
    return out


def exclusive_suffix(rows, rank):
    """Sum of rows(rank..world) excluding rank's own, added back to front (last rank: zeros)."""
    out = rows[0].new_zeros(rows[0].shape)
    for q in range(len(rows) - 1, rank, -1):
        out = out + rows[q]
    return out


def _gathered(row, group=None):
    """All ranks' rows as ONE (world, ...) tensor with a single collective (NCCL); None on other backends."""
    import torch.distributed as dist
    world, _ = _world(group)
    if world == 1 or dist.get_backend(group) != "nccl":
        return None
    buf = row.new_empty((world,) + tuple(row.shape))
    dist.all_gather_into_tensor(buf, row.contiguous(), group=group)
    return buf


def exchange_prefix(row, group=None):
    """Exclusive prefix over ranks of a small row: one all-gather + one reduction kernel on NCCL (the list
    form, `exclusive_prefix(gather_rows(...))`, elsewhere -- same value up to the order of the additions)."""
    buf = _gathered(row, group)
    if buf is None:
        return exclusive_prefix(gather_rows(row, group), _world(group)[1])
    return buf[:_world(group)[1]].sum(dim=0)


def exchange_suffix(row, group=None):
    buf = _gathered(row, group)
    if buf is None:
        return exclusive_suffix(gather_rows(row, group), _world(group)[1])
    return buf[_world(group)[1] + 1:].sum(dim=0)

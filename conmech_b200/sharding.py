"""Multi-GPU driver of the batched solve: one process per GPU, the subset batch range-partitioned
over the ranks, model tables replicated, NO collective on the solve path.
The only exchange is the final gather of the per-subset results, outside the solve.

    torchrun --nproc-per-node 8 ...                       # RANK / LOCAL_RANK / WORLD_SIZE in the env
    sc = StiffnessChecker(model, checker_engine='cuda')   # device = LOCAL_RANK
    res = solve_many_sharded(sc._sc_ins, masks)           # same (B, ...) arrays on every rank
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous range [lo, hi) of rank `rank` when n items are split over `world` ranks; the first
    n % world ranks get one item more.  Ranges are disjoint, ordered by rank and cover 0..n."""
    if not (0 <= rank < world):
        raise ValueError('rank {} not in [0, {})'.format(rank, world))
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_many_sharded(engine, masks, want=('flag', 'status', 'max_trans', 'compliance'), group=None,
                       gather=True):
    """Solve this rank's contiguous share of ``masks`` ((B, words) uint32, identical on every rank)
    with ``engine.solve_many`` and, if ``gather``, all-gather the results so that every rank returns
    the full (B, ...) arrays in the original subset order.  Works with any initialised
    ``torch.distributed`` backend (NCCL ranks gather through host tensors with gloo semantics: the
    gather moves B*n_lc bytes per output, it is not on the solve path)."""
    import torch
    import torch.distributed as dist
    masks = np.ascontiguousarray(masks)
    B = masks.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(B, world, rank)
    local = engine.solve_many(masks[lo:hi], want=want) if hi > lo else {}
    if world == 1 or not gather:
        return local
    out = {}
    for name in want:
        mine = local.get(name)
        parts = [None] * world
        dist.all_gather_object(parts, None if mine is None else np.ascontiguousarray(mine), group=group)
        parts = [p for p in parts if p is not None and len(p)]
        out[name] = np.concatenate(parts, axis=0) if parts else np.zeros((0,))
    return out

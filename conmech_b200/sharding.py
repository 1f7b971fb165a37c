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
    from .masks import as_mask_matrix
    m = as_mask_matrix(np.asarray(masks), engine.info['mask_words'])
    if m is None:
        raise ValueError('solve_many_sharded takes a (B, mask_words) uint32/int32 mask matrix')
    masks = m
    B = masks.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(B, world, rank)
    local = engine.solve_many(masks[lo:hi], want=want) if hi > lo else {}
    if world == 1 or not gather:
        return local
    # final gather: one all_gather per output over equally padded tensors (shares differ by at most one
    # subset); NCCL ranks gather on their device, gloo ranks on the host.  Not on the solve path.
    use_cuda = dist.get_backend(group) == 'nccl'
    n_max = -(-B // world)
    out = {}
    if 'eR_rows' in want:
        # compact element rows: a rank's share is the rows of ITS subsets - (off[hi] - off[lo]) * n_lc of them, known to
        # every rank from the masks - so the shares are padded to the largest one and cut back after the gather
        from .cuda_stiffness import CudaStiffness
        off = CudaStiffness.element_row_offsets(masks)
        out['eR_row_offsets'] = off
    for name in want:
        if name == 'eR_row_offsets':
            continue
        mine = np.ascontiguousarray(local[name]) if hi > lo else None
        if name == 'eR_rows':
            shares = [shard_range(B, world, r) for r in range(world)]
            n_lc = [None]
            if rank == 0:
                n_lc = [int(mine.shape[0] // max(1, off[hi] - off[lo])) if off[hi] > off[lo] else 1]
            dist.broadcast_object_list(n_lc, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            L = n_lc[0]
            r_max = max(int(off[b] - off[a]) for a, b in shares) * L
            buf = np.zeros((max(r_max, 1), 12), dtype=np.float64)
            if mine is not None and mine.size:
                buf[:mine.shape[0]] = mine
            tns = torch.from_numpy(buf.reshape(-1))
            if use_cuda:
                tns = tns.cuda()
            parts = [torch.empty_like(tns) for _ in range(world)]
            dist.all_gather(parts, tns, group=group)
            out[name] = np.concatenate([p.cpu().numpy().reshape(-1, 12)[:int(off[b] - off[a]) * L]
                                        for p, (a, b) in zip(parts, shares)], axis=0)
            continue
        # every rank needs the trailing shape and dtype: rank 0 always has a non-empty share when B > 0
        meta = [None]
        if rank == 0:
            meta = [(mine.shape[1:], mine.dtype.str)]
        dist.broadcast_object_list(meta, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        tail, dt = meta[0]
        buf = np.zeros((n_max,) + tuple(tail), dtype=np.dtype(dt))
        if mine is not None:
            buf[:hi - lo] = mine
        tns = torch.from_numpy(buf.view(np.uint8).reshape(-1))
        if use_cuda:
            tns = tns.cuda()
        parts = [torch.empty_like(tns) for _ in range(world)]
        dist.all_gather(parts, tns, group=group)
        rows = []
        for r, p in enumerate(parts):
            rlo, rhi = shard_range(B, world, r)
            a = p.cpu().numpy().view(np.dtype(dt)).reshape((n_max,) + tuple(tail))
            rows.append(a[:rhi - rlo])
        out[name] = np.concatenate(rows, axis=0)
    return out

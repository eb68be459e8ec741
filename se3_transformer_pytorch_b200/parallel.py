"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

Point clouds in a batch are independent: every op is batched over `b` and neighbours never
cross clouds, so the hot path shards with NO data-path collective.  The only exchange is one all-gather of the
returned type-0 invariants (or whichever degree is returned) when the caller wants the whole batch on every rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """Contiguous split of `batch` clouds over `world_size` ranks; the first batch % world ranks take one extra."""
    base, rem = divmod(batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(t, world_size, rank):
    """Slice the leading (batch) axis of a tensor / dict of tensors; 2-D adjacency matrices ([n, n]) are shared."""
    if t is None:
        return None
    if isinstance(t, dict):
        return {k: shard_batch(v, world_size, rank) for k, v in t.items()}
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def all_gather_batch(local, batch, group=None):
    """All-gather along the batch axis (handles uneven shards by padding to the largest shard)."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if isinstance(local, dict):
        return {k: all_gather_batch(v, batch, group) for k, v in local.items()}
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < most:
        pad = torch.cat([local, local.new_zeros((most - local.shape[0],) + tuple(local.shape[1:]))], 0)
    pad = pad.contiguous()
    out = pad.new_empty((world * most,) + tuple(pad.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * most: r * most + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)


def sharded_forward(model, feats, coors, mask=None, *, group=None, gather=True, batch_kwargs=('edges', 'neighbor_mask', 'global_feats'),
                    **kwargs):
    """Run `model` on this rank's contiguous slice of the batch and (optionally) all-gather the outputs.
    Every rank is given the same full-batch inputs; per-batch keyword tensors named in `batch_kwargs` (and a batched
    adj_mat) are sliced as well."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    batch = coors.shape[0]
    kw = dict(kwargs)
    for name in batch_kwargs:
        if kw.get(name) is not None:
            kw[name] = shard_batch(kw[name], world, rank)
    if kw.get('adj_mat') is not None and kw['adj_mat'].dim() == 3:
        kw['adj_mat'] = shard_batch(kw['adj_mat'], world, rank)
    local = model(shard_batch(feats, world, rank), shard_batch(coors, world, rank), shard_batch(mask, world, rank), **kw)
    if not gather or world == 1:
        return local
    return all_gather_batch(local, batch, group)

"""world_size-2 gloo test of the batch-sharding host logic (no GPU): a stand-in per-cloud function is sharded and
all-gathered and must equal the single-process result, for even and uneven batches."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se3_transformer_pytorch_b200.parallel import shard_bounds, sharded_forward


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def fake_model(feats, coors, mask=None, return_type=None, edges=None):
    """independent per cloud, like the hot path: returns {'0': [b,n,d], '1': [b,n,d,3]}"""
    w = mask.float().unsqueeze(-1) if mask is not None else 1.0
    s = feats.cumsum(dim=1) * w + (0 if edges is None else edges.float().mean(dim=(1, 2), keepdim=True))
    out = {'0': s, '1': s.unsqueeze(-1) * coors.mean(dim=1, keepdim=True).unsqueeze(2)}
    return out if return_type is None else out[str(return_type)]


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    feats, coors = torch.randn(batch, 6, 4), torch.randn(batch, 6, 3)
    mask = torch.rand(batch, 6) > 0.2
    edges = torch.randint(0, 4, (batch, 6, 6))
    full = fake_model(feats, coors, mask, edges=edges)
    got = sharded_forward(fake_model, feats, coors, mask, edges=edges)
    ok = all(torch.allclose(got[k], full[k]) for k in full)
    got0 = sharded_forward(fake_model, feats, coors, mask, return_type=0)
    ok = ok and torch.allclose(got0, fake_model(feats, coors, mask, return_type=0))
    local = sharded_forward(fake_model, feats, coors, mask, gather=False)
    lo, hi = shard_bounds(batch, world, rank)
    ok = ok and local['0'].shape[0] == hi - lo
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _run(batch):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_shard_bounds_cover_batch():
    for batch in (1, 2, 5, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_sharded_forward_even_batch():
    _run(4)


def test_sharded_forward_uneven_batch():
    _run(5)

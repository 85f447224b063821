"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in nerfail_b200/dist.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerfail_b200 import dist as nd
    try:
        # view sharding covers every view exactly once
        mine = nd.shard_views(10, rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sorted(sum(gathered, [])) == list(range(10))
        # attack iteration: per-rank gradient of its own views, one all-reduce, identical sign step everywhere
        g = torch.Generator().manual_seed(0)
        P, H, W = 2, 6, 5
        init = torch.zeros(P, H, W, 4); init[..., 3] = (torch.rand(P, H, W, generator=g) > 0.3).float() * 255
        s = init.clone()
        per_view = torch.randn(4, P, H, W, 4, generator=g)              # gradient contribution of each of 4 views
        full = per_view.sum(0)
        local = per_view[nd.shard_views(4, rank, world)].sum(0)
        s = nd.attack_sign_step_(s, local, init, step=2.0, eps=3.0)
        want = init.clone()
        want[..., :3] = torch.clamp(init[..., :3] - 2.0 * torch.sign(full[..., :3]) * (init[..., 3:4] > 0), -3.0, 3.0)
        assert torch.allclose(s, want), "sign must be taken after the reduce"
        # exchanging and updating only the active rows' RGB (what the update consumes) gives the same table
        s2 = init.clone()
        idx = nd.active_rows(s2)
        assert idx.numel() == int((init[..., 3] > 0).sum())
        s2 = nd.attack_sign_step_(s2, per_view[nd.shard_views(4, rank, world)].sum(0), init, step=2.0, eps=3.0, active_idx=idx)
        assert torch.equal(s2, s)
        # data-parallel retraining: mean-of-ranks gradient equals the full-batch gradient
        torch.manual_seed(1)
        lin = torch.nn.Linear(7, 3)
        x, y = torch.randn(8, 7, generator=g), torch.randn(8, 3, generator=g)
        b, e = nd.shard_range(8, rank, world)
        loss = ((lin(x[b:e]) - y[b:e]) ** 2).mean()
        loss.backward()
        nd.allreduce_grads_(lin.parameters(), scale=1.0 / world)
        lin2 = torch.nn.Linear(7, 3); lin2.load_state_dict(lin.state_dict())
        ((lin2(x) - y) ** 2).mean().backward()
        assert torch.allclose(lin.weight.grad, lin2.weight.grad, atol=1e-6)
        fin = nd.allreduce_grads_(lin.parameters(), async_op=True)      # async bucket path
        fin()
        # gradients that are back-to-back views of one flat buffer (the fused training backward's layout) are reduced
        # in place: no flatten / scatter, the views see the reduced values
        flat = torch.arange(24, dtype=torch.float32) * (rank + 1)
        lin.weight.grad, lin.bias.grad = flat[:21].view(3, 7), flat[21:].view(3)
        assert nd._flat_view([lin.weight.grad, lin.bias.grad]).data_ptr() == flat.data_ptr()
        nd.allreduce_grads_(lin.parameters(), scale=0.5)
        want = torch.arange(24, dtype=torch.float32) * 3 * 0.5
        assert torch.equal(flat, want) and torch.equal(lin.bias.grad, want[21:])
        assert nd._flat_view([lin.bias.grad, lin.weight.grad]) is None  # out of order: falls back to flatten + scatter
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world))


def test_shard_range_is_a_partition():
    from nerfail_b200 import dist as nd
    for n in (0, 1, 7, 640000, 4096):
        for w in (1, 2, 3, 8):
            spans = [nd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_shard_pixels_is_a_balanced_partition():
    """dist.shard_pixels: the (view, pixel range) pieces of all ranks tile the batch exactly once, in order, cut on the
    quantum, and no rank has more than one quantum more than another (100 views over 8 ranks: 12.5 views each)."""
    from nerfail_b200 import dist as nd
    for V, px, q in ((100, 640000, 800), (100, 640000, 1), (3, 10, 1), (7, 24, 8), (1, 64, 16), (5, 30, 30)):
        for w in (1, 2, 3, 4, 8):
            pieces = [nd.shard_pixels(V, px, r, w, q) for r in range(w)]
            flat = [p for ps in pieces for p in ps]
            pos = 0
            for v, b, e in flat:
                assert 0 <= b < e <= px and v * px + b == pos, (V, px, q, w, v, b, e, pos)
                pos = v * px + e
            assert pos == V * px
            loads = [sum(e - b for _, b, e in ps) for ps in pieces]
            assert max(loads) - min(loads) <= q + (V * px) % q, (loads, q)
            if q > 1:
                assert all(b % q == 0 for _, b, _ in flat)
    assert [sum(e - b for _, b, e in nd.shard_pixels(100, 640000, r, 8, 800)) for r in range(8)] == [8000000] * 8

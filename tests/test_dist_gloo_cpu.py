"""CPU, world_size 2 over gloo: the ragged all-gather that repacks per-rank token outputs -- equal and
unequal per-rank image counts, the dealt (`deal_by_cost`) order restored, `index_down` carried along, and the two-phase
(start / finish) form the streaming pipeline uses."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _ragged(counts, rank, width=3, cap_extra=4, down_n=0):
    from setok_b200 import RaggedTokens
    total = sum(counts)
    data = torch.zeros(total + cap_extra, width)                      # capacity > live rows
    data[:total] = torch.arange(total * width, dtype=torch.float32).reshape(total, width) + 100 * rank
    offs = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)) if counts else [0], dtype=torch.int32)
    down = None
    if down_n:
        down = torch.full((len(counts), down_n), -1, dtype=torch.int64)
        for i, c in enumerate(counts):
            down[i, :c] = torch.arange(c) + 10 * i + 1000 * rank
    return RaggedTokens(data, offs, down)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from setok_b200.dist import RaggedAllGather, all_gather_ragged
    res = {}
    # (a) equal B, blocking form
    out = all_gather_ragged(_ragged([[2, 0, 5], [1, 4, 1]][rank], rank))
    res["equal"] = (out.counts, out.packed().clone())
    # (b) unequal B (3 vs 1 images), index_down carried, order given by a dealing of 4 images: rank 0 holds global images
    #     {0, 2, 3}, rank 1 holds {1}
    counts = [[2, 3, 1], [4]][rank]
    order = [[0, 2, 3], [1]][rank]
    out = all_gather_ragged(_ragged(counts, rank, down_n=6), order=order)
    res["unequal"] = (out.counts, out.packed().clone(), out.index_down.clone())
    # (c) two-phase form with a fixed batch capacity, two batches in flight, one rank with an empty batch of rows
    g = RaggedAllGather(batch_capacity=4)
    h1 = g.start(_ragged([[1, 1], [0, 0, 2]][rank], rank))
    h2 = g.start(_ragged([[3], [1, 1]][rank], rank + 2))
    o1, o2 = g.finish(h1), g.finish(h2)
    res["two_phase"] = (o1.counts, o1.packed().clone(), o2.counts, o2.packed().clone())
    q.put((rank, res))
    dist.destroy_process_group()


def test_all_gather_ragged_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    rows = lambda n, base: torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) + base
    for rank, r in res:
        counts, packed = r["equal"]
        assert counts == [2, 0, 5, 1, 4, 1]
        assert torch.equal(packed, torch.cat([rows(7, 0), rows(6, 100)]))
        # dealt order restored: global images 0, 1, 2, 3 = rank0[0], rank1[0], rank0[1], rank0[2]
        counts, packed, down = r["unequal"]
        assert counts == [2, 4, 3, 1]
        r0, r1 = rows(6, 0), rows(4, 100)
        assert torch.equal(packed, torch.cat([r0[0:2], r1, r0[2:5], r0[5:6]]))
        assert down.shape == (4, 6)
        assert down[0, :2].tolist() == [0, 1] and down[1, :4].tolist() == [1000, 1001, 1002, 1003]
        assert down[2, :3].tolist() == [10, 11, 12] and down[3, :2].tolist() == [20, -1]
        c1, p1, c2, p2 = r["two_phase"]
        assert c1 == [1, 1, 0, 0, 2] and torch.equal(p1, torch.cat([rows(2, 0), rows(2, 100)]))
        assert c2 == [3, 1, 1] and torch.equal(p2, torch.cat([rows(3, 200), rows(2, 300)]))


def test_deal_by_cost_balances_mixed_resolutions():
    """BASELINE config 5: equal thirds of {224, 336, 448}^2 images over 8 ranks: every image is dealt exactly once and the
    heaviest rank carries at most 1.2x the mean cost (a contiguous split would put all 448^2 images on three ranks)."""
    from setok_b200.dist import deal_by_cost, shard_batch
    sizes = [224] * 64 + [336] * 64 + [448] * 64
    parts = deal_by_cost(sizes, 8)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    cost = [sum(sizes[i] ** 2 for i in p) for p in parts]
    assert max(cost) <= 1.2 * sum(cost) / 8
    contiguous = [sum(s ** 2 for s in sizes[slice(*shard_batch(len(sizes), r, 8))]) for r in range(8)]
    assert max(contiguous) > 1.5 * sum(contiguous) / 8


def _ag_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from setok_b200.training import all_gather_with_grad
    x = (torch.arange(6, dtype=torch.float32).reshape(3, 2) + 10 * rank).requires_grad_()
    full = all_gather_with_grad(x)
    # a loss every rank weights differently: rank r contributes (r + 1) * sum(full * w)
    w = torch.arange(full.numel(), dtype=torch.float32).reshape(full.shape)
    ((rank + 1) * (full * w).sum()).backward()
    q.put((rank, full.detach().clone(), x.grad.clone()))
    dist.destroy_process_group()


def test_all_gather_with_grad_world2():
    """multilabel_constrastive.py:14-23 (diffdist all_gather): forward = concatenation over ranks, backward = the SUM over ranks of
    the gradients that flowed into this rank's slice."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ag_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    full_exp = torch.cat([torch.arange(6, dtype=torch.float32).reshape(3, 2), torch.arange(6, dtype=torch.float32).reshape(3, 2) + 10])
    w = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    for rank, full, grad in res:
        assert torch.equal(full, full_exp)
        assert torch.equal(grad, (1 + 2) * w[rank * 3:(rank + 1) * 3])

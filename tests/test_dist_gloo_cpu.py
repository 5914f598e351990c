"""CPU, world_size 2 over gloo: the ragged all-gather that repacks per-rank token outputs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from setok_b200 import RaggedTokens
    from setok_b200.dist import all_gather_ragged
    counts = [[2, 0, 5], [1, 4, 1]][rank]
    total = sum(counts)
    data = torch.zeros(12, 3)                       # capacity 12 rows, `total` live
    data[:total] = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3) + 100 * rank
    offs = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32)
    out = all_gather_ragged(RaggedTokens(data, offs))
    q.put((rank, out.counts, out.packed().clone()))
    dist.destroy_process_group()


def test_all_gather_ragged_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    [p.join(60) for p in procs]
    exp0 = torch.arange(21, dtype=torch.float32).reshape(7, 3)
    exp1 = torch.arange(18, dtype=torch.float32).reshape(6, 3) + 100
    for rank, counts, packed in res:
        assert counts == [2, 0, 5, 1, 4, 1]
        assert torch.equal(packed, torch.cat([exp0, exp1]))


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

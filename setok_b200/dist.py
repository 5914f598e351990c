"""Data-parallel repack of ragged token outputs (the single exchange step of the path).

Each rank tokenises its own slice of the batch; when the downstream LLM needs the global batch, one
all-gather of the per-image counts plus one all-gather of the (padded-to-max) packed rows rebuilds a global
RaggedTokens on every rank.  Backend-agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .ragged import RaggedTokens


def all_gather_ragged(local: RaggedTokens, group: Optional[dist.ProcessGroup] = None) -> RaggedTokens:
    world = dist.get_world_size(group)
    if world == 1:
        return local
    dev = local.data.device
    B = local.batch_size
    counts = (local.offsets[1:] - local.offsets[:-1]).to(torch.int32)
    all_counts = torch.empty(world * B, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
    per_rank = all_counts.view(world, B).sum(dim=1)
    max_rows = int(per_rank.max().item())            # one host sync: NCCL needs equal message sizes
    Cc = local.data.shape[-1]
    send = local.data[:max_rows]
    if send.shape[0] < max_rows:                     # capacity smaller than another rank's live rows
        pad = local.data.new_zeros(max_rows - send.shape[0], Cc)
        send = torch.cat([send, pad], 0)
    recv = torch.empty(world * max_rows, Cc, dtype=local.data.dtype, device=dev)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    # compact: drop each rank's padding
    rows = per_rank.tolist()
    parts = [recv[r * max_rows: r * max_rows + rows[r]] for r in range(world)]
    data = torch.cat(parts, 0)
    offsets = torch.zeros(world * B + 1, dtype=torch.int32, device=dev)
    offsets[1:] = torch.cumsum(all_counts, 0)
    return RaggedTokens(data, offsets)


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous split of a batch over ranks (first `n_items % world` ranks get one extra)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def deal_by_cost(sizes, world: int):
    """Mixed-resolution batches (BASELINE config 5): the per-image cost grows like N^2 + 24 N C, so a contiguous split can
    leave one rank with all the 448^2 images.  Longest-processing-time dealing: images sorted by cost, each goes to the
    currently lightest rank.  `sizes` = per-image side length (or token count); returns `world` index lists whose
    concatenation is a permutation of range(len(sizes))."""
    order = sorted(range(len(sizes)), key=lambda i: (-float(sizes[i]) ** 2, i))
    loads = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += float(sizes[i]) ** 2
    return [sorted(o) for o in out]

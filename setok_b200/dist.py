"""Data-parallel repack of ragged token outputs (the single exchange step of the path).

Each rank tokenises its own slice of the batch; when the downstream LLM needs the global batch, the ranks exchange
their ragged outputs: one small all-gather of a per-rank header (image count, per-image row counts, the images' global
positions) and one all-gather of the packed rows, sized by the largest rank's *live* row count (not by capacity), then
one row gather that drops the padding and puts the images in their global order.

The exchange is split in two halves so that it never stalls the stream the tokenizer runs on:

* ``RaggedAllGather.start(local)`` enqueues the header all-gather on a communication stream (after the kernels that
  produced ``local``) and returns at once;
* ``RaggedAllGather.finish(handle)`` reads the gathered header on the host -- the only host synchronisation, and it
  waits for that batch's header only -- and enqueues the row all-gather + repack on the communication stream.

A caller that pipelines (``pipeline.stream_tokenize``; ``bench.py``) launches the next batch's tokenizer between the two
halves, so the host read lands while the GPU is busy and the NCCL transfer overlaps the next batch's ViT.
``all_gather_ragged`` is the two halves back to back for callers that need the result now.

Ranks may hold different numbers of images (``shard_batch`` with a remainder, ``deal_by_cost``) and different image
resolutions; ``order`` carries each local image's position in the global batch so that a dealt batch comes back in its
original order.  Backend-agnostic: NCCL over NVLink on the GPU box, gloo on CPU tensors in the tests."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .ragged import RaggedTokens

_HDR = 4      # header words ahead of the per-image arrays: B_local, has_index_down, N (index_down width), row width check


class _Handle:
    __slots__ = ("local", "hdr_all", "event", "bcap")

    def __init__(self, local, hdr_all, event, bcap):
        self.local, self.hdr_all, self.event, self.bcap = local, hdr_all, event, bcap


class RaggedAllGather:
    """Two-phase all-gather of RaggedTokens over ``group``.  ``batch_capacity``: upper bound of any rank's image count
    (fixes the header size so that no collective is needed to agree on it)."""

    def __init__(self, batch_capacity: int, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.bcap = int(batch_capacity)
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == "cuda"
        self.comm = torch.cuda.Stream(self.device) if self.cuda else None

    # -- phase 1 -------------------------------------------------------------------------------
    def start(self, local: RaggedTokens, order: Optional[Sequence[int]] = None) -> _Handle:
        B = local.batch_size
        if B > self.bcap:
            raise ValueError(f"local batch {B} exceeds batch_capacity {self.bcap}")
        dev = local.data.device
        has_down = local.index_down is not None
        # header layout: [B, has_index_down, N, row width | order (bcap) | counts (bcap)]; the host-known part goes up in one
        # copy from pinned memory, the counts come from the device-resident offsets
        host = torch.full((_HDR + self.bcap,), -1, dtype=torch.int32, pin_memory=dev.type == "cuda")
        host[0], host[1], host[2], host[3] = B, int(has_down), (int(local.index_down.shape[1]) if has_down else 0), int(local.data.shape[-1])
        if order is not None:
            if len(order) != B:
                raise ValueError("order must name the global position of every local image")
            host[_HDR:_HDR + B] = torch.as_tensor(list(order), dtype=torch.int32)
        hdr = torch.full((_HDR + 2 * self.bcap,), -1, dtype=torch.int32, device=dev)
        hdr[:_HDR + self.bcap].copy_(host, non_blocking=True)
        hdr[_HDR + self.bcap:_HDR + self.bcap + B] = (local.offsets[1:] - local.offsets[:-1]).to(torch.int32)
        hdr_all = torch.empty(self.world * hdr.numel(), dtype=torch.int32, device=dev)
        event = None
        if dev.type == "cuda":
            produced = torch.cuda.Event()
            produced.record(torch.cuda.current_stream(dev))
            comm = self.comm or torch.cuda.Stream(dev)
            self.comm = comm
            comm.wait_event(produced)
            with torch.cuda.stream(comm):
                # (the packed rows are NOT record_stream'ed: a recorded block of capacity size -- 268 MB for a 256-image batch --
                # would not be reusable until the allocator has polled this stream, i.e. a cudaMalloc per step; the handle and
                # then the gathered result keep `local` alive until the row exchange that reads it has been ordered)
                for t in (hdr, hdr_all):
                    t.record_stream(comm)
                dist.all_gather_into_tensor(hdr_all, hdr, group=self.group)
                # pinned host copy of the gathered header, so that finish() waits on an event instead of the device
                host_all = torch.empty(hdr_all.shape, dtype=torch.int32, pin_memory=True)
                host_all.copy_(hdr_all, non_blocking=True)
                event = torch.cuda.Event()
                event.record(comm)
            hdr_all = host_all
        else:
            dist.all_gather_into_tensor(hdr_all, hdr, group=self.group)
        return _Handle(local, hdr_all, event, self.bcap)

    # -- phase 2 -------------------------------------------------------------------------------
    def finish(self, h: _Handle, wait: bool = True) -> RaggedTokens:
        """Returns the global RaggedTokens (every rank gets the same).  On CUDA the row exchange runs on the communication
        stream; with ``wait`` the current stream is made to wait for it (the result is then usable like any tensor), without
        it the caller orders its consumers after ``result.ready`` (a CUDA event) itself."""
        local = h.local
        dev = local.data.device
        if h.event is not None:
            h.event.synchronize()                      # this batch's header only
        hdr = h.hdr_all.numpy().reshape(self.world, _HDR + 2 * h.bcap)
        Bs = hdr[:, 0].astype(np.int64)
        if (hdr[:, 3] != hdr[0, 3]).any():
            raise ValueError(f"ranks disagree on the token width: {hdr[:, 3].tolist()}")
        orders = [hdr[r, _HDR:_HDR + Bs[r]].astype(np.int64) for r in range(self.world)]
        counts = [hdr[r, _HDR + h.bcap:_HDR + h.bcap + Bs[r]].astype(np.int64) for r in range(self.world)]
        per_rank = np.array([int(c.sum()) for c in counts], dtype=np.int64)
        max_rows = max(int(per_rank.max()), 1)
        n_img = int(Bs.sum())
        # global position of every (rank, local image): the senders' `order`, or rank-major when none was given
        base = np.cumsum(Bs) - Bs
        gpos = np.concatenate([np.where(orders[r] >= 0, orders[r], base[r] + np.arange(Bs[r])) for r in range(self.world)]) if n_img else np.zeros(0, np.int64)
        if n_img and sorted(gpos.tolist()) != list(range(n_img)):
            raise ValueError("the ranks' `order` lists are not a permutation of the global batch")
        cnt_flat = np.concatenate(counts) if n_img else np.zeros(0, np.int64)
        # source row (in the gathered, padded buffer) of the first row of every (rank, local image)
        src0 = np.concatenate([r * max_rows + np.cumsum(counts[r]) - counts[r] for r in range(self.world)]) if n_img else np.zeros(0, np.int64)
        inv = np.argsort(gpos, kind="stable")          # inv[g] = flat (rank, image) slot of global image g
        cnt_g, src_g = cnt_flat[inv], src0[inv]
        offs = np.concatenate([[0], np.cumsum(cnt_g)])
        total = int(offs[-1])
        rows_idx = np.repeat(src_g - offs[:-1], cnt_g) + np.arange(total)
        Cc = local.data.shape[-1]
        # index_down rides along when every rank has one of the same width
        down_ok = bool((hdr[:, 1] == 1).all() and (hdr[:, 2] == hdr[0, 2]).all()) and local.index_down is not None
        send = local.data[:max_rows]
        if send.shape[0] < max_rows:                   # capacity smaller than another rank's live rows
            send = torch.cat([send, local.data.new_zeros(max_rows - send.shape[0], Cc)], 0)
        ctx = torch.cuda.stream(self.comm) if dev.type == "cuda" else _null()
        with ctx:
            recv = torch.empty(self.world * max_rows, Cc, dtype=local.data.dtype, device=dev)
            dist.all_gather_into_tensor(recv, send.contiguous(), group=self.group)
            idx_t = torch.from_numpy(rows_idx)
            if dev.type == "cuda":
                idx_t = idx_t.pin_memory().to(dev, non_blocking=True)
            data = recv.index_select(0, idx_t) if total else recv[:0]
            offsets = torch.from_numpy(offs.astype(np.int32))
            offsets = offsets.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else offsets
            down = None
            if down_ok:
                N = int(hdr[0, 2])
                dsend = local.index_down.new_full((h.bcap, N), -1)
                dsend[:local.batch_size] = local.index_down
                drecv = torch.empty(self.world * h.bcap, N, dtype=local.index_down.dtype, device=dev)
                dist.all_gather_into_tensor(drecv, dsend, group=self.group)
                slot = np.concatenate([r * h.bcap + np.arange(Bs[r]) for r in range(self.world)])[inv]
                slot_t = torch.from_numpy(slot)
                slot_t = slot_t.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else slot_t
                down = drecv.index_select(0, slot_t)
            ready = None
            if dev.type == "cuda":
                ready = torch.cuda.Event()
                ready.record(self.comm)
        out = RaggedTokens(data, offsets, down)
        out._host = [int(v) for v in offs.tolist()]     # the host already knows the offsets: per-image slicing needs no sync
        out.ready = ready
        # the row all-gather reads `local` on the communication stream: `local` lives at least as long as the result.  With
        # wait=True the current stream waits for `ready`, so nothing launched afterwards can reuse its block early; with
        # wait=False the caller keeps the result until it has ordered itself after `ready` (pipeline._Readback does).
        out._keep = local
        if ready is not None and wait:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready)
            for t in (data, offsets) + ((down,) if down is not None else ()):
                t.record_stream(cur)
        return out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def all_gather_ragged(local: RaggedTokens, group: Optional[dist.ProcessGroup] = None, order: Optional[Sequence[int]] = None,
                      batch_capacity: Optional[int] = None) -> RaggedTokens:
    """Blocking form: both halves back to back.  Without ``batch_capacity`` the ranks first agree on the largest local image
    count with one tiny all-reduce (ranks may hold different numbers of images)."""
    world = dist.get_world_size(group)
    if world == 1 and order is None:
        return local
    dev = local.data.device
    if batch_capacity is None:
        b = torch.tensor([local.batch_size], dtype=torch.int32, device=dev)
        dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
        batch_capacity = int(b.item())
    g = RaggedAllGather(batch_capacity, group, dev)
    return g.finish(g.start(local, order))


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous split of a batch over ranks (first `n_items % world` ranks get one extra)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def deal_by_cost(sizes, world: int) -> List[List[int]]:
    """Mixed-resolution batches (BASELINE config 5): the per-image cost grows like N^2 + 24 N C, so a contiguous split can
    leave one rank with all the 448^2 images.  Longest-processing-time dealing: images sorted by cost, each goes to the
    currently lightest rank.  `sizes` = per-image side length (or token count); returns `world` index lists whose
    concatenation is a permutation of range(len(sizes)) -- pass a rank's list as `order` to the gather to get the global
    batch back in its original order."""
    order = sorted(range(len(sizes)), key=lambda i: (-float(sizes[i]) ** 2, i))
    loads = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += float(sizes[i]) ** 2
    return [sorted(o) for o in out]

"""Host <-> device streaming around the tokenizer (the end-to-end path a data loader drives).

`stream_tokenize` software-pipelines everything that would otherwise serialise a step:

* the H2D copy of the *next* batch (pinned host memory, on a copy stream) is issued before anything else happens to the
  current batch;
* the kernel launches of the current batch run on the caller's stream;
* with ``gather`` (a ``dist.RaggedAllGather``) the data-parallel exchange of batch i is split around the launch of batch
  i+1: its small header all-gather is enqueued right after batch i's kernels, the host reads that header -- and sizes /
  enqueues the row all-gather on the communication stream -- only after batch i+1 has been launched, so neither the host
  read nor the NCCL transfer ever idles the compute stream;
* the D2H read-back of a finished batch's ragged result runs on its own stream and is submitted right behind the batch's
  kernels; its row count is data-dependent, so the rows are copied speculatively (the previous batch's count + 25 %) and only a
  larger batch pays a second, host-issued hop.  The host sync waits only for that batch, while the GPU runs the next ones.

Every batch is still copied in and its result copied out; nothing is cached across steps."""
from __future__ import annotations

from collections import deque
from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch

from .ragged import RaggedTokens


class HostResult:
    """Host copy of one batch's output: packed tokens (sum K, C), offsets (B+1,), idx_cluster (B, N), score (B, 1, N)."""

    def __init__(self, tokens, offsets, idx_cluster, score, copied_rows: Optional[int] = None, hops: int = 1):
        self.tokens, self.offsets, self.idx_cluster, self.score = tokens, offsets, idx_cluster, score
        self.hops = hops                      # 2: the row count was not covered by the speculative copy (or there was none)
        # rows that actually crossed the bus (>= len(tokens) when the read-back was sized speculatively)
        self.copied_rows = int(tokens.shape[0]) if copied_rows is None else int(copied_rows)

    @property
    def nbytes(self) -> int:
        """Bytes copied device -> host for this batch (speculative surplus rows included)."""
        row = self.tokens.shape[1] * self.tokens.element_size() if self.tokens.dim() == 2 else 0
        return self.copied_rows * row + sum(t.numel() * t.element_size() for t in (self.offsets, self.idx_cluster, self.score))


class _Readback:
    """D2H of one batch's ragged result on a dedicated stream.  The row count is data-dependent: the host knows it either
    already (a gathered batch), or from the offsets that come back with a speculatively sized copy of the rows (`guess_rows`),
    or -- first batch, or a batch larger than the guess -- through a second hop.  Everything waits only for THIS batch's kernels
    (an event), never for the batch the main stream is already running, and lands in pinned memory.

    The device tensors are kept alive by this object until `result()` has synchronised with the copies (no
    `Tensor.record_stream`: a recorded block is not reusable until the allocator has polled the side stream's event, which
    made every step cudaMalloc a fresh capacity-sized buffer -- 134 MB for config 4's projected rows -- inside the step)."""

    def __init__(self, out, d2h: torch.cuda.Stream, after: torch.cuda.Event, guess_rows: int = 0):
        self.rt, self.idx, self.score = out
        self.d2h = d2h
        d2h.wait_event(after)
        known = self.rt._host is not None            # offsets already on the host (RaggedAllGather.finish)
        self.h_guess = None
        self.copied = 0
        with torch.cuda.stream(d2h):
            self.h_off = torch.empty(self.rt.offsets.shape, dtype=self.rt.offsets.dtype, pin_memory=True)
            self.h_off.copy_(self.rt.offsets, non_blocking=True)
            self.h_idx = torch.empty(self.idx.shape, dtype=self.idx.dtype, pin_memory=True)
            self.h_idx.copy_(self.idx, non_blocking=True)
            self.h_score = torch.empty(self.score.shape, dtype=self.score.dtype, pin_memory=True)
            self.h_score.copy_(self.score, non_blocking=True)
            self.h_tok = None
            if known:
                self.h_tok = self._rows(self.rt.total)
            elif guess_rows > 0:
                # Speculative single hop: a copy the host issues only once it knows the row count is submitted AFTER the next
                # batch's kernels and, on this platform, is served after them too (measured: the result of batch j came back
                # when batch j + 1 had finished, and the compute stream idled ~0.3 ms per batch waiting for the host).  So the rows
                # are copied now, sized by the previous batch's count plus a margin; result() falls back to the second hop only
                # when this batch turned out larger.
                n = min(int(guess_rows), int(self.rt.data.shape[0]))
                self.h_guess = self._rows(n)
            self.ev = torch.cuda.Event()
            self.ev.record(d2h)

    def _rows(self, total: int):
        h_tok = torch.empty((total, self.rt.data.shape[1]), dtype=self.rt.data.dtype, pin_memory=True)
        h_tok.copy_(self.rt.data[:total], non_blocking=True)
        self.copied += total
        return h_tok

    def _release(self):
        self.rt = self.idx = self.score = None        # the copies are complete: the blocks go back to their own stream's pool

    def result(self) -> HostResult:
        self.ev.synchronize()
        hops = 1
        if self.h_tok is None:
            total = int(self.h_off[-1])
            if self.h_guess is not None and total <= self.h_guess.shape[0]:
                self.h_tok = self.h_guess[:total]            # the speculative copy covered the batch: no second hop
            else:
                with torch.cuda.stream(self.d2h):
                    self.h_tok = self._rows(total)
                self.d2h.synchronize()
                hops = 2
        self._release()
        return HostResult(self.h_tok, self.h_off, self.h_idx, self.h_score, copied_rows=self.copied, hops=hops)

    def __del__(self):                                # abandoned before result(): the copies may still be reading the tensors
        try:
            if getattr(self, "rt", None) is not None:
                self.d2h.synchronize()
        except Exception:
            pass


_side_streams = {}


def _streams(dev: torch.device):
    """One copy stream and one read-back stream per device for the life of the process: the caching allocators keep a pool
    per stream, so a fresh stream per call would cudaMalloc its input buffers again on every call."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _side_streams:
        _side_streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _side_streams[key]


def stream_tokenize(tokenizer, host_batches: Iterable[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                    post: Optional[Callable] = None, gather=None, **forward_kwargs) -> Iterator[HostResult]:
    """host_batches yields (pinned images (B,3,H,W) float or uint8, pinned noise (B,N) or None).  Yields one HostResult per
    batch, in order.  `post(ragged, idx, score) -> (ragged, idx, score)` runs on the device right after the tokenizer (e.g.
    the projector); `gather` (dist.RaggedAllGather) then repacks the ranks' ragged outputs into the global batch."""
    dev = tokenizer.device
    copy_stream, d2h_stream = _streams(dev)
    main = torch.cuda.current_stream(dev)

    def upload(batch):
        imgs, noise = batch
        with torch.cuda.stream(copy_stream):
            d_img = imgs.to(dev, non_blocking=True)
            d_noise = None if noise is None else noise.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d_img, d_noise, ev

    state = {"guess": 0}

    def read(out, after):
        return _Readback(out, d2h_stream, after, guess_rows=state["guess"])

    def took(res: HostResult) -> HostResult:
        # next batches' speculative read-back: this batch's row count + 25 %, in whole blocks of 64 rows
        state["guess"] = max(64, (int(res.tokens.shape[0]) * 5 // 4 + 63) // 64 * 64)
        return res

    it = iter(host_batches)
    try:
        nxt = upload(next(it))
    except StopIteration:
        return
    pending_gather = None                # (handle, idx, score) of the batch whose row exchange is still to be enqueued
    reads = deque()                      # read-backs in flight, oldest first
    while nxt is not None:
        d_img, d_noise, ev = nxt
        if not ev.query():                  # the upload was issued a whole batch ago: normally complete, and then the compute
            main.wait_event(ev)             # stream needs no cross-stream wait in front of the batch's first kernel
        d_img.record_stream(main)
        if d_noise is not None:
            d_noise.record_stream(main)
        try:
            nxt = upload(next(it))          # next batch's H2D is in flight before this batch's kernels are even launched
        except StopIteration:
            nxt = None
        out = tokenizer(d_img, noise=d_noise, **forward_kwargs)
        if post is not None:
            out = post(*out)
        if gather is not None:
            handle = gather.start(out[0])   # header exchange enqueued behind this batch's kernels; no host wait
            if pending_gather is not None:
                reads.append(_finish_gather(gather, pending_gather, read))
            pending_gather = (handle, out[1], out[2])
        else:
            done = torch.cuda.Event()
            done.record(main)
            reads.append(read(out, done))
        while len(reads) > 1:
            yield took(reads.popleft().result())   # an older batch's D2H (and its host sync) overlaps the newer batches' kernels
    if pending_gather is not None:
        reads.append(_finish_gather(gather, pending_gather, read))
    while reads:
        yield took(reads.popleft().result())


def _finish_gather(gather, pending, read):
    handle, idx, score = pending
    rt = gather.finish(handle, wait=False)   # host reads that batch's header (long since landed), row exchange on the comm stream
    return read((rt, idx, score), rt.ready)

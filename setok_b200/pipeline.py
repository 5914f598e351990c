"""Host <-> device streaming around the tokenizer (the end-to-end path a data loader drives).

`stream_tokenize` software-pipelines three things that would otherwise serialise each step:
the H2D copy of the *next* batch (pinned host memory, on a copy stream), the kernel launches of the
current batch, and the D2H read-back of the *previous* batch's ragged result on its own stream (its row count is
data-dependent, so reading it needs a host sync — which waits only for that batch, while the GPU runs the next one).
Every batch is still copied in and its result copied out; nothing is cached across steps."""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch

from .ragged import RaggedTokens


class HostResult:
    """Host copy of one batch's output: packed tokens (sum K, C), offsets (B+1,), idx_cluster (B, N), score (B, 1, N)."""

    def __init__(self, tokens, offsets, idx_cluster, score):
        self.tokens, self.offsets, self.idx_cluster, self.score = tokens, offsets, idx_cluster, score

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.tokens, self.offsets, self.idx_cluster, self.score))


class _Readback:
    """D2H of one batch's ragged result on a dedicated stream.  The row count is data-dependent, so the read-back is two
    hops (offsets, then the packed rows); both wait only for THIS batch's kernels (an event), never for the batch the
    main stream is already running, and land in pinned host memory."""

    def __init__(self, out, d2h: torch.cuda.Stream, main: torch.cuda.Stream):
        self.rt, self.idx, self.score = out
        self.d2h = d2h
        done = torch.cuda.Event()
        done.record(main)
        d2h.wait_event(done)
        with torch.cuda.stream(d2h):
            for t in (self.rt.data, self.rt.offsets, self.idx, self.score):
                t.record_stream(d2h)
            self.h_off = torch.empty(self.rt.offsets.shape, dtype=self.rt.offsets.dtype, pin_memory=True)
            self.h_off.copy_(self.rt.offsets, non_blocking=True)
            self.h_idx = torch.empty(self.idx.shape, dtype=self.idx.dtype, pin_memory=True)
            self.h_idx.copy_(self.idx, non_blocking=True)
            self.h_score = torch.empty(self.score.shape, dtype=self.score.dtype, pin_memory=True)
            self.h_score.copy_(self.score, non_blocking=True)
            self.ev = torch.cuda.Event()
            self.ev.record(d2h)

    def result(self) -> HostResult:
        self.ev.synchronize()
        total = int(self.h_off[-1])
        with torch.cuda.stream(self.d2h):
            h_tok = torch.empty((total, self.rt.data.shape[1]), dtype=self.rt.data.dtype, pin_memory=True)
            h_tok.copy_(self.rt.data[:total], non_blocking=True)
        self.d2h.synchronize()
        return HostResult(h_tok, self.h_off, self.h_idx, self.h_score)


def stream_tokenize(tokenizer, host_batches: Iterable[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                    post: Optional[Callable] = None, **forward_kwargs) -> Iterator[HostResult]:
    """host_batches yields (pinned images (B,3,H,W), pinned noise (B,N) or None).  Yields one HostResult per batch,
    in order.  `post(ragged, idx, score) -> (ragged, idx, score)` runs on the device after the tokenizer (e.g. the
    projector or the data-parallel all-gather)."""
    dev = tokenizer.device
    copy_stream = torch.cuda.Stream(dev)
    d2h_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    def upload(batch):
        imgs, noise = batch
        with torch.cuda.stream(copy_stream):
            d_img = imgs.to(dev, non_blocking=True)
            d_noise = None if noise is None else noise.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d_img, d_noise, ev

    it = iter(host_batches)
    try:
        nxt = upload(next(it))
    except StopIteration:
        return
    prev = None
    while nxt is not None:
        d_img, d_noise, ev = nxt
        main.wait_event(ev)
        d_img.record_stream(main)
        if d_noise is not None:
            d_noise.record_stream(main)
        out = tokenizer(d_img, noise=d_noise, **forward_kwargs)
        if post is not None:
            out = post(*out)
        try:
            nxt = upload(next(it))          # next batch's H2D overlaps this batch's kernels
        except StopIteration:
            nxt = None
        cur = _Readback(out, d2h_stream, main)
        if prev is not None:
            yield prev.result()             # previous batch's D2H (and its host sync) overlaps this batch's kernels
        prev = cur
    if prev is not None:
        yield prev.result()

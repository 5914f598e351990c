"""Host <-> device streaming around the tokenizer (the end-to-end path a data loader drives).

`stream_tokenize` software-pipelines three things that would otherwise serialise each step:
the H2D copy of the *next* batch (pinned host memory, on a copy stream), the kernel launches of the
current batch, and the D2H read-back of the *previous* batch's ragged result (its row count is
data-dependent, so reading it needs a host sync — which now lands while the GPU is busy with the next batch).
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


def _to_host(out) -> HostResult:
    rt, idx, score = out
    return HostResult(rt.packed().cpu(), rt.offsets.cpu(), idx.cpu(), score.cpu())


def stream_tokenize(tokenizer, host_batches: Iterable[Tuple[torch.Tensor, Optional[torch.Tensor]]],
                    post: Optional[Callable] = None, **forward_kwargs) -> Iterator[HostResult]:
    """host_batches yields (pinned images (B,3,H,W), pinned noise (B,N) or None).  Yields one HostResult per batch,
    in order.  `post(ragged, idx, score) -> (ragged, idx, score)` runs on the device after the tokenizer (e.g. the
    projector or the data-parallel all-gather)."""
    dev = tokenizer.device
    copy_stream = torch.cuda.Stream(dev)
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
        if prev is not None:
            yield _to_host(prev)            # previous batch's D2H (and its host sync) overlaps too
        prev = out
    if prev is not None:
        yield _to_host(prev)

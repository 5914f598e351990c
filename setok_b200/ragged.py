"""Ragged (variable-K) batch of semantic tokens: packed rows + an offset vector.

This is the container `SetokTokenizer.forward` / `encode_images` return in place of the reference's
per-image tensor (tokenizer.py:182, setokim_arch.py:206-211).  It satisfies what the Setokim splice loop
needs from `image_features` (setokim_arch.py:262-301): ``features[i]`` is image i's ``(K_i, C)`` rows,
``len(features)`` is the batch size, ``.dim()`` is 3, ``.shape[-1]`` is the feature width.

``data`` has *capacity* rows (B*N); only the first ``offsets[B]`` are live.  Nothing here forces a
host sync until a Python-side consumer asks for per-image slices (``[i]``, ``packed()``, ``counts``).
"""
from __future__ import annotations

from typing import List, Optional

import torch


class RaggedTokens:
    def __init__(self, data: torch.Tensor, offsets: torch.Tensor, index_down: Optional[torch.Tensor] = None):
        if offsets.dim() != 1 or offsets.numel() < 1:
            raise ValueError("offsets must be a 1-D tensor of B+1 entries")
        self.data = data              # (capacity, C) on the device
        self.offsets = offsets        # (B+1,) int32 on the device
        self.index_down = index_down  # optional (B, N) int64, -1 padded: centre token of each cluster
        self._host: Optional[List[int]] = None

    # -- host view (one sync, cached) ----------------------------------------------------------
    def _h(self) -> List[int]:
        if self._host is None:
            self._host = [int(v) for v in self.offsets.detach().cpu().tolist()]
        return self._host

    @property
    def batch_size(self) -> int:
        return self.offsets.numel() - 1

    @property
    def counts(self) -> List[int]:
        h = self._h()
        return [h[i + 1] - h[i] for i in range(len(h) - 1)]

    @property
    def total(self) -> int:
        return self._h()[-1]

    def packed(self) -> torch.Tensor:
        """(sum K, C) view of the live rows."""
        return self.data[: self.total]

    # -- tensor-like surface used by the reference's callers -------------------------------------
    def __len__(self) -> int:
        return self.batch_size

    def __getitem__(self, i: int) -> torch.Tensor:
        h = self._h()
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.batch_size))]
        if i < 0:
            i += self.batch_size
        if not 0 <= i < self.batch_size:
            raise IndexError(i)
        return self.data[h[i]: h[i + 1]]

    def __iter__(self):
        for i in range(self.batch_size):
            yield self[i]

    def dim(self) -> int:
        return 3

    @property
    def shape(self):
        return (self.batch_size, None, self.data.shape[-1])

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def device(self):
        return self.data.device

    def to(self, *a, **k) -> "RaggedTokens":
        """Like Tensor.to for the packed rows; a device move takes ``offsets`` / ``index_down`` along (their integer dtypes
        are kept)."""
        data = self.data.to(*a, **k)
        offsets, down = self.offsets, self.index_down
        if data.device != self.data.device:
            offsets = offsets.to(data.device)
            down = None if down is None else down.to(data.device)
        r = RaggedTokens(data, offsets, down)
        r._host = self._host
        return r

    def with_data(self, data: torch.Tensor) -> "RaggedTokens":
        r = RaggedTokens(data, self.offsets, self.index_down)
        r._host = self._host
        return r

    def to_padded(self, pad_value: float = 0.0):
        """(B, K_max, C) tensor + (B, K_max) bool mask — the layout the reference's detokenizer consumes
        (detokenizer.py:101-120)."""
        c = self.counts
        kmax = max(c) if c else 0
        out = self.data.new_full((self.batch_size, kmax, self.data.shape[-1]), pad_value)
        mask = torch.zeros(self.batch_size, kmax, dtype=torch.bool, device=self.data.device)
        for i, n in enumerate(c):
            out[i, :n] = self[i]
            mask[i, :n] = True
        return out, mask

    def __repr__(self):
        return f"RaggedTokens(B={self.batch_size}, C={self.data.shape[-1]}, dtype={self.data.dtype}, device={self.data.device})"

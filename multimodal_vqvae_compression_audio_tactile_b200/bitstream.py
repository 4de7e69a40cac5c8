"""Code indices <-> bytes: the payload a sender would put on the wire for ``ProposedEval.decode_indices``.

The reference never serialises its codes -- it reports the bitrate analytically,
``est_kbps = tokens_per_sec * rvq_books * log2(rvq_embed) / 1000`` (Training/compare_dacvsproposal_5.py:372-373).
This is that payload made real: ``ceil(log2 K)`` bits per index, indices in [B, books, Tl] order, little-endian bit
order inside a byte, zero-padded to a whole byte.  Host-side numpy: a 75-token frame is < 100 bytes."""
from __future__ import annotations

import math

import numpy as np
import torch


def bits_per_index(K: int) -> int:
    if K < 1:
        raise ValueError("bits_per_index: empty codebook")
    return max(1, (int(K) - 1).bit_length())


def estimated_kbps(books: int, K: int, tokens_per_sec: float = 75.0) -> float:
    """The reference's analytic bitrate (:372-373); equals the packed payload rate when K is a power of two."""
    return tokens_per_sec * books * math.log2(K) / 1000.0


def packed_bytes(shape, K: int) -> int:
    n = int(np.prod(shape))
    return (n * bits_per_index(K) + 7) // 8


def pack_indices(idx: torch.Tensor, K: int) -> bytes:
    """idx: integer tensor (any shape, values in [0, K)) -> bytes."""
    if idx.dtype.is_floating_point:
        raise ValueError("pack_indices: integer tensor expected")
    a = idx.detach().cpu().numpy().astype(np.int64).ravel()
    if a.size and (a.min() < 0 or a.max() >= K):
        raise ValueError(f"pack_indices: index outside [0, {K})")
    nb = bits_per_index(K)
    bits = ((a[:, None] >> np.arange(nb, dtype=np.int64)) & 1).astype(np.uint8).ravel()
    return np.packbits(bits, bitorder="little").tobytes()


def unpack_indices(payload: bytes, shape, K: int, device=None) -> torch.Tensor:
    """Inverse of pack_indices -> int32 tensor of ``shape`` (on ``device`` if given)."""
    n = int(np.prod(shape))
    nb = bits_per_index(K)
    if len(payload) != (n * nb + 7) // 8:
        raise ValueError(f"unpack_indices: {len(payload)} bytes for {n} indices of {nb} bits")
    bits = np.unpackbits(np.frombuffer(payload, dtype=np.uint8), bitorder="little")[: n * nb].reshape(n, nb)
    a = (bits.astype(np.int64) << np.arange(nb, dtype=np.int64)).sum(axis=1)
    if a.size and a.max() >= K:
        raise ValueError(f"unpack_indices: payload holds an index >= {K}")
    t = torch.from_numpy(a.astype(np.int32)).reshape(tuple(shape))
    return t.to(device) if device is not None else t

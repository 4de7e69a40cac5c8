"""Functional entry points that take caller tensors (no packed weights)."""
from __future__ import annotations

import torch

from . import _lib as L
from .engine import Emitter, Engine

_engines = {}


def _engine(device: torch.device) -> Engine:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    eng = _engines.get(key)
    if eng is None:
        eng = _engines[key] = Engine(torch.device("cuda", key[1]))
    return eng


@torch.no_grad()
def nearest_code(x: torch.Tensor, emb: torch.Tensor, precision: str = "tc") -> torch.Tensor:
    """ResidualVQEMA._nearest_l2 (Evaluation/dac_vcpwq_proposed6_latency.py:417-419):
    argmax_k (x @ emb.T - 0.5 * |emb_k|^2), first maximum wins.  x [N, D], emb [K, D] CUDA tensors
    -> int64 [N].  The [N, K] score matrix is never written to memory."""
    if not (x.is_cuda and emb.is_cuda):
        raise L.B2CError("nearest_code: CUDA tensors only (no CPU fallback)")
    if x.dim() != 2 or emb.dim() != 2 or x.shape[1] != emb.shape[1]:
        raise ValueError(f"nearest_code expects x [N, D] and emb [K, D], got {tuple(x.shape)} and {tuple(emb.shape)}")
    n, d = x.shape
    k = emb.shape[0]
    if k == 0:
        raise ValueError("nearest_code: empty codebook")
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=x.device)
    eng = _engine(x.device)
    prec = L.PRECISIONS[{"tc": "bf16x3"}.get(precision, precision)]
    if prec != L.PREC_F32 and not eng.lib.b2c_nearest_tc_eligible(n, d, k):
        prec = L.PREC_F32      # D not a multiple of 8: the FP32 CUDA kernel (same indices by construction)
    key = ("nearest", n, d, k, prec)
    prog = eng.programs.get(key)
    if prog is None:
        em = Emitter(eng)
        scratch = em.arena.alloc(int(eng.lib.b2c_nearest_scratch_bytes(n, d, k, prec)))
        ii = em.new(n)
        em.nearest(em.ext(1), em.ext(2), scratch, ii, n, d, k, prec)
        em.widen(ii, em.ext(3), n)
        prog = eng.programs[key] = em.finish(3)
    xf = x.detach().float().contiguous()
    ef = emb.detach().float().contiguous()
    out = torch.empty(n, dtype=torch.int64, device=x.device)
    eng.run(prog, [xf.data_ptr(), ef.data_ptr(), out.data_ptr()])
    return out

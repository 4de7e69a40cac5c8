"""ORACLE test-case catalogue (test infrastructure): seeded model builders and
synthetic inputs shared by oracle/make_golden.py, tests/, smoke() and bench.py's
CPU-baseline leg.  Everything is a pure function of integer seeds so the GPU box
can rebuild the exact tensors the goldens were made from.

Recipes follow SURVEY.md section 8(d): SEED = 7 for module construction
(Training/compare_dacvsproposal_3.py:50,79-80), inputs U(-1,1) with
torch.Generator().manual_seed(123), 1-s frames at 24 kHz
(Evaluation/dac_vcpwq_proposed6_latency.py:498-499).
"""
from __future__ import annotations

import math

import torch

from . import dac_arch
from .proposed import AR_CHUNK_TOK, nearest_code

MODEL_SEED = 7
INPUT_SEED = 123
CALIB_SEED = 321

# name -> case.  books/K follow the reference sweep grid
# (Training/compare_dacvsproposal_5.py:86-88) and compare_dacvsproposal_3.py:58-63.
CODEC_CASES = {
    # config 1: the compare_dacvsproposal_3.py shape (10 books x 128 codes), B = 2
    "c3_b10k128": dict(books=10, K=128, B=2, T=24000, kind="uniform"),
    # calibrated codebooks so the arg-max is exercised on every book (random init
    # collapses book 0 onto a handful of codes)
    "cal_b8k512": dict(books=8, K=512, B=2, T=24000, kind="uniform", calibrate=True),
    # prefix-truncated RVQ (books_use < books), sine mixture, short ragged frame
    "cal_b4k256_use3_short": dict(books=4, K=256, B=3, T=8000, kind="sines", calibrate=True, books_use=3),
    # the latency script's input: all-zero frame, B = 1
    "zeros_b1k128": dict(books=1, K=128, B=1, T=24000, kind="zeros"),
    # more points of the compare_dacvsproposal_5 sweep grid (BOOKS_LIST x EMBED_LIST, :86-88): the smallest
    # codebook with gaussian frames, and the largest with a books_use prefix (Evaluation/dac_vcpwq_proposed.py stages)
    "cal_b1k128_normal": dict(books=1, K=128, B=1, T=24000, kind="normal", calibrate=True),
    "cal_b10k512_use6": dict(books=10, K=512, B=1, T=24000, kind="uniform", calibrate=True, books_use=6),
    # the two remaining points of config 2's five ((1,128) (4,256) (8,512) (10,128) (10,512)) with ALL
    # books in use on full 1-s frames
    "cal_b4k256": dict(books=4, K=256, B=1, T=24000, kind="normal", calibrate=True),
    "cal_b10k512": dict(books=10, K=512, B=1, T=24000, kind="sines", calibrate=True),
}

# name -> (N, D, K) for ResidualVQEMA._nearest_l2 (config 5 subset the CPU finishes in seconds)
SEARCH_CASES = {
    "n75_d96_k512": (75, 96, 512),
    "n1_d8_k256": (1, 8, 256),
    "n1024_d96_k128": (1024, 96, 128),
    "n4800_d16_k1024": (4800, 16, 1024),
    "n1024_d256_k8192": (1024, 256, 8192),
    "n16384_d64_k2048": (16384, 64, 2048),
    "n333_d32_k4096": (333, 32, 4096),
    "n4800_d128_k256": (4800, 128, 256),
}


def chunk_lengths(tl: int, chunk: int = AR_CHUNK_TOK):
    return [min(tl, s + chunk) - s for s in range(0, tl, chunk)]


def codec_inputs(case, seed: int = INPUT_SEED):
    b, t = case["B"], case["T"]
    g = torch.Generator().manual_seed(seed)
    kind = case.get("kind", "uniform")
    if kind == "zeros":
        return torch.zeros(b, 1, t), torch.zeros(b, 1, t)
    if kind == "uniform":
        a = torch.rand(b, 1, t, generator=g) * 2 - 1
        x = torch.rand(b, 1, t, generator=g) * 2 - 1
        return a, x
    if kind == "normal":
        return 0.1 * torch.randn(b, 1, t, generator=g), 0.1 * torch.randn(b, 1, t, generator=g)
    if kind == "sines":
        n = torch.arange(t, dtype=torch.float32) / 24000.0

        def mix():
            f = 50 + 200 * torch.rand(b, 3, 1, generator=g)
            ph = 2 * math.pi * torch.rand(b, 3, 1, generator=g)
            return (0.5 / 3) * torch.sin(2 * math.pi * f * n + ph).sum(1, keepdim=True)

        return mix(), mix()
    raise ValueError(kind)


def search_inputs(n: int, d: int, k: int, seed: int = 0):
    """x, emb ~ N(0, 1/D): the reference's codebook init (randn/sqrt(dim),
    Evaluation/dac_vcpwq_proposed6_latency.py:413)."""
    g = torch.Generator().manual_seed(seed + 1000003 * d + 7919 * k + n)
    x = torch.randn(n, d, generator=g) / math.sqrt(d)
    emb = torch.randn(k, d, generator=g) / math.sqrt(d)
    return x, emb


def predictor_inputs(b: int = 3, c: int = 1024, t: int = AR_CHUNK_TOK, seed: int = 99):
    g = torch.Generator().manual_seed(seed)
    zt_prev = 0.3 * torch.randn(b, c, t, generator=g)
    za = 2.0 * torch.randn(b, c, t, generator=g)
    return zt_prev, za


def build_backbones(seed: int = MODEL_SEED):
    """build_backbones_for_eval (Evaluation/dac_vcpwq_proposed6_latency.py:527-535)
    with random-init DAC-architecture models instead of the downloaded weights."""
    torch.manual_seed(seed)
    da = dac_arch.DAC().eval()
    dt = dac_arch.DAC().eval()
    return da, dt


@torch.no_grad()
def calibrate_books(model, case):
    """Replace the randn codebooks by rows drawn from the stage-i residuals of a
    seeded calibration batch, so every book has many live codes."""
    cal = dict(case)
    cal["B"] = 2
    a, t = codec_inputs(cal, seed=CALIB_SEED)
    za = model.A_ENC(a)
    qa = model.A_QUANT(za)[0]
    zt = model.T_ENC(t)
    b, c, tl = zt.shape
    rds = []
    for s in range(0, tl, AR_CHUNK_TOK):
        e = min(tl, s + AR_CHUNK_TOK)
        z_pred = model.predict(torch.zeros(b, c, e - s), qa[..., s:e])
        r = zt[..., s:e] - z_pred
        rd = model.proj_down(model.scale.clamp(5e-3, 0.5) * torch.tanh(model.tokennorm(r)))
        rds.append(rd)
    x = torch.cat(rds, dim=-1).permute(0, 2, 1).reshape(-1, rds[0].shape[1])
    g = torch.Generator().manual_seed(CALIB_SEED + 1)
    residual = x.clone()
    for book in model.vq.books:
        k = book.shape[0]
        pick = torch.randint(0, residual.shape[0], (k,), generator=g)
        rows = residual[pick] * (1.0 + 0.05 * torch.randn(k, 1, generator=g))
        rows = rows + 0.02 * residual.std() * torch.randn(rows.shape, generator=g)
        book.data.copy_(rows)
        q = rows[nearest_code(residual, rows)]
        residual = residual - q
    return model


def build_reference_style_model(proposed_cls, case, seed: int = MODEL_SEED):
    """Same construction order as the reference scripts: two DAC models, then
    ProposedEval(A_ENC, A_QUANT, T_ENC, T_DEC, c_lat, books, K)
    (Evaluation/dac_vcpwq_proposed6_latency.py:661-662)."""
    da, dt = build_backbones(seed)
    model = proposed_cls(da.encoder, da.quantizer, dt.encoder, dt.decoder, dac_arch.LATENT_DIM,
                         case["books"], case["K"]).eval()
    if case.get("calibrate"):
        calibrate_books(model, case)
    return model

"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the reference's own layers on the encode -> quantize ->
decode path.  Every function cites the reference lines it follows
(file = Evaluation/dac_vcpwq_proposed6_latency.py unless noted).

Pinned: tests/test_oracle_cpu.py checks this file bit-for-bit against the
reference classes themselves (ast-extracted by oracle/ref_loader.py when
/root/reference is present) and against tests/golden/*.npz, which were produced
by running the reference classes here (oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CODE_DIM = 96        # :336
AR_CHUNK_TOK = 16    # :337


def sinusoid_table(c: int, max_len: int = 8192) -> torch.Tensor:
    """PosEnc1D.__init__ (:340-347): pe[pos, 2i] = sin(pos * w_i), pe[pos, 2i+1] = cos."""
    pe = torch.zeros(max_len, c)
    pos = torch.arange(0, max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, c, 2) * (-math.log(10000.0) / c))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class PosEnc1D(nn.Module):
    def __init__(self, c, max_len=8192):
        super().__init__()
        self.register_buffer("pe", sinusoid_table(c, max_len))

    def forward(self, x):  # :349-351 -- positions are relative to the slice passed in
        n = x.size(-1)
        return x + self.pe[:n, :].T.unsqueeze(0).to(x.dtype)


class TokenNorm(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.ln = nn.LayerNorm(c)

    def forward(self, z):  # :357-360 -- LayerNorm over channels of every token
        return self.ln(z.permute(0, 2, 1)).permute(0, 2, 1)


class CrossPredictor(nn.Module):
    """:362-407.  One pre-LN cross-attention block + FFN; queries come from the
    previous tactile latents, keys/values from the quantised audio latents."""

    def __init__(self, c, heads=8, mlp_mul=2, dropout=0.1):
        super().__init__()
        assert c % heads == 0
        self.pos = PosEnc1D(c)
        self.h = heads
        self.dh = c // heads
        self.ln_q = nn.LayerNorm(c)
        self.ln_kv = nn.LayerNorm(c)
        self.q_proj = nn.Linear(c, c, False)
        self.k_proj = nn.Linear(c, c, False)
        self.v_proj = nn.Linear(c, c, False)
        self.out = nn.Linear(c, c, False)
        self.drop = nn.Dropout(dropout)
        self.ffn = nn.Sequential(nn.LayerNorm(c), nn.Linear(c, mlp_mul * c), nn.GELU(),
                                 nn.Linear(mlp_mul * c, c))

    def _heads(self, x):
        b, t, _ = x.shape
        return x.view(b, t, self.h, self.dh).permute(0, 2, 1, 3)

    def forward(self, zt_prev, za):
        q = self.ln_q(self.pos(zt_prev).permute(0, 2, 1))           # :392,394
        kv = self.ln_kv(self.pos(za).permute(0, 2, 1))              # :393,395
        Q = self._heads(self.q_proj(q))
        K = self._heads(self.k_proj(kv))
        V = self._heads(self.v_proj(kv))
        att = (Q @ K.transpose(-2, -1)) / math.sqrt(self.dh)        # :401
        ctx = att.softmax(dim=-1) @ V                               # :402
        b, h, t, d = ctx.shape
        ctx = ctx.permute(0, 2, 1, 3).contiguous().view(b, t, h * d)
        y = self.out(self.drop(ctx))                                # :404
        y = y + q                                                   # :405 (q is the LayerNorm-ed query)
        y = y + self.ffn(y)                                         # :406
        return y.permute(0, 2, 1)


def nearest_code(x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResidualVQEMA._nearest_l2 (:417-419): argmax_k (x . e_k - 0.5 |e_k|^2),
    first maximum wins."""
    return (x @ emb.t() - 0.5 * (emb * emb).sum(dim=1).unsqueeze(0)).argmax(dim=1)


def nearest_code_scores(x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    return x @ emb.t() - 0.5 * (emb * emb).sum(dim=1).unsqueeze(0)


class ResidualVQEMA(nn.Module):
    """:409-435.  Residual VQ; ``last_indices`` / ``last_margins`` are oracle-side
    additions (the reference throws the indices away)."""

    def __init__(self, dim: int, n_books: int, n_embed: int):
        super().__init__()
        self.books = nn.ParameterList(
            [nn.Parameter(torch.randn(n_embed, dim) / math.sqrt(dim)) for _ in range(n_books)]
        )
        self.last_indices = None
        self.last_margins = None

    def forward(self, z, n_books_use=None):
        if n_books_use is None:
            n_books_use = len(self.books)
        n_books_use = min(n_books_use, len(self.books))
        b, d, t = z.shape
        x = z.permute(0, 2, 1).reshape(b * t, d)
        residual = x
        q_sum = torch.zeros_like(x)
        idxs, margins = [], []
        for cb in list(self.books)[:n_books_use]:
            emb = cb.detach().to(z.dtype).to(z.device)
            sc = nearest_code_scores(residual, emb)
            idx = sc.argmax(dim=1)
            if sc.shape[1] > 1:
                top2 = sc.topk(2, dim=1).values
                margins.append((top2[:, 0] - top2[:, 1]).view(b, t))
            else:
                margins.append(torch.full((b, t), float("inf")))
            idxs.append(idx.view(b, t))
            q = F.embedding(idx, emb)
            q_sum = q_sum + (q - residual).detach() + residual      # :433 -- keep this op order
            residual = residual - q                                 # :434
        self.last_indices = torch.stack(idxs, dim=1) if idxs else torch.zeros(b, 0, t, dtype=torch.long)
        self.last_margins = torch.stack(margins, dim=1) if margins else torch.zeros(b, 0, t)
        return q_sum.view(b, t, d).permute(0, 2, 1).contiguous()


def ema_step(books, z_tokens: torch.Tensor, decay: float) -> list:
    """ResidualVQEMA.ema_step (Training/compare_dacvsproposal_3.py:264-276), in place on ``books`` (a list of [K, D]
    tensors or parameters).  Every book sees the SAME tokens X (no residual update between books).  Returns the
    per-book nearest-code indices and score margins (oracle-side additions for the parity tests)."""
    b, d, t = z_tokens.shape
    x = z_tokens.permute(0, 2, 1).reshape(b * t, d)
    info = []
    with torch.no_grad():
        for cb in books:
            emb = cb.data
            sc = x.to(emb) @ emb.t() - 0.5 * (emb * emb).sum(dim=1).unsqueeze(0)       # :269
            idx = sc.argmax(dim=1)
            top2 = sc.topk(2, dim=1).values
            k = emb.size(0)
            counts = torch.bincount(idx, minlength=k).float().unsqueeze(1)            # :271
            sums = torch.zeros_like(emb)
            sums.index_add_(0, idx, x.to(emb))                                        # :272
            mask = counts.squeeze(1) > 0
            means = torch.zeros_like(emb)
            means[mask] = sums[mask] / (counts[mask] + 1e-9)                          # :274
            emb[mask] = decay * emb[mask] + (1.0 - decay) * means[mask]              # :275
            info.append(dict(idx=idx, margin=top2[:, 0] - top2[:, 1], counts=counts.squeeze(1).long()))
    return info


class ProposedEval(nn.Module):
    """:437-487 (training twin: Training/compare_dacvsproposal_3.py:300-340)."""

    def __init__(self, A_ENC, A_QUANT, T_ENC, T_DEC, c_lat, rvq_books, rvq_embed):
        super().__init__()
        self.A_ENC, self.A_QUANT, self.T_ENC, self.T_DEC = A_ENC, A_QUANT, T_ENC, T_DEC
        for m in (A_ENC, A_QUANT, T_ENC, T_DEC):
            for p in m.parameters():
                p.requires_grad_(False)
        self.predict = CrossPredictor(c=c_lat)
        self.tokennorm = TokenNorm(c_lat)
        self.scale = nn.Parameter(torch.tensor(0.08))
        self.proj_down = nn.Conv1d(c_lat, CODE_DIM, 1)
        self.proj_up = nn.Conv1d(CODE_DIM, c_lat, 1)
        self.vq = ResidualVQEMA(dim=CODE_DIM, n_books=rvq_books, n_embed=rvq_embed)

    # ---- the reference's own schedule: 5 sequential chunks (:451-478) ----
    @torch.no_grad()
    def encode_latents(self, a_1T, t_1T, books_use=None, trace=None):
        za = self.A_ENC(a_1T)
        quant = self.A_QUANT(za)
        qa = quant[0]
        zt = self.T_ENC(t_1T)
        b, c, tl = zt.shape
        z_run = torch.zeros_like(zt)
        idx_all, mar_all, rd_all, zp_all = [], [], [], []
        for s in range(0, tl, AR_CHUNK_TOK):
            e = min(tl, s + AR_CHUNK_TOK)
            zt_prev = torch.zeros(b, c, e - s, device=zt.device, dtype=zt.dtype)
            if s == 0:
                zt_prev[..., 1:] = z_run[..., s:e - 1]
            else:
                zt_prev[...] = z_run[..., s - 1:e - 1]
            z_pred = self.predict(zt_prev, qa[..., s:e])
            r = zt[..., s:e] - z_pred.detach()
            rn = torch.tanh(self.tokennorm(r))
            rd = self.proj_down(self.scale.clamp(5e-3, 0.5) * rn)
            qd = self.vq(rd, n_books_use=books_use)
            z_run[..., s:e] = self.proj_up(qd) + z_pred
            idx_all.append(self.vq.last_indices)
            mar_all.append(self.vq.last_margins)
            rd_all.append(rd)
            zp_all.append(z_pred)
        if trace is not None:
            trace.update(
                za=za, qa=qa, zt=zt, z_run=z_run,
                a_codes=quant[1] if len(quant) > 1 else None,
                a_margin=getattr(self.A_QUANT, "last_margins", None),
                idx=torch.cat(idx_all, dim=-1), margin=torch.cat(mar_all, dim=-1),
                rD=torch.cat(rd_all, dim=-1), z_pred=torch.cat(zp_all, dim=-1),
            )
        return z_run

    @torch.no_grad()
    def forward_eval(self, a_1T, t_1T, books_use=None, trace=None):  # :480-487
        z_run = self.encode_latents(a_1T, t_1T, books_use=books_use, trace=trace)
        y = self.T_DEC(z_run)
        if trace is not None:
            trace["y"] = y
        return y

    # ---- receiver: the reference never decodes from indices; this is its loop with the lookup in place of the search ----
    @torch.no_grad()
    def decode_from_indices(self, a_1T, idx, books_use=None):
        """What a receiver holding the audio frame and the code indices [B, books, Tl] computes: the chunk loop of
        encode_latents (:462-477) with qD = sum over books of book[idx] (plain sums in book order) in place of the
        residual search, then T_DEC (:486).  Oracle for ProposedEval.decode_indices of the CUDA path."""
        za = self.A_ENC(a_1T)
        qa = self.A_QUANT(za)[0]
        b, c, tl = qa.shape
        use = idx.shape[1] if books_use is None else min(books_use, idx.shape[1])
        z_run = torch.zeros_like(qa)
        for s in range(0, tl, AR_CHUNK_TOK):
            e = min(tl, s + AR_CHUNK_TOK)
            zt_prev = torch.zeros(b, c, e - s, dtype=qa.dtype)
            if s == 0:
                zt_prev[..., 1:] = z_run[..., s:e - 1]
            else:
                zt_prev[...] = z_run[..., s - 1:e - 1]
            z_pred = self.predict(zt_prev, qa[..., s:e])
            qd = torch.zeros(b, CODE_DIM, e - s, dtype=qa.dtype)
            for k in range(use):
                qd = qd + F.embedding(idx[:, k, s:e], self.vq.books[k].detach()).permute(0, 2, 1)
            z_run[..., s:e] = self.proj_up(qd) + z_pred
        return self.T_DEC(z_run), z_run

    # ---- the two-pass schedule the CUDA path uses ----
    @torch.no_grad()
    def encode_latents_two_pass(self, a_1T, t_1T, books_use=None):
        """Same result as encode_latents: when chunk s is computed, z_run[s:e-1] is
        still zero, so only the first token of chunks 1.. sees a non-zero query
        input, namely z_hat[s-1], which itself is never a chunk start."""
        za = self.A_ENC(a_1T)
        qa = self.A_QUANT(za)[0]
        zt = self.T_ENC(t_1T)
        b, c, tl = zt.shape
        z_run = torch.zeros_like(zt)

        def chunk(s, e, zt_prev):
            z_pred = self.predict(zt_prev, qa[..., s:e])
            r = zt[..., s:e] - z_pred
            rd = self.proj_down(self.scale.clamp(5e-3, 0.5) * torch.tanh(self.tokennorm(r)))
            return self.proj_up(self.vq(rd, n_books_use=books_use)) + z_pred

        bounds = [(s, min(tl, s + AR_CHUNK_TOK)) for s in range(0, tl, AR_CHUNK_TOK)]
        for s, e in bounds:                                   # pass 1: all chunks, zero query input
            z_run[..., s:e] = chunk(s, e, torch.zeros(b, c, e - s, dtype=zt.dtype))
        fixed = z_run.clone()
        for s, e in bounds[1:]:                               # pass 2: chunk heads only
            zt_prev = torch.zeros(b, c, e - s, dtype=zt.dtype)
            zt_prev[..., 0] = z_run[..., s - 1]
            fixed[..., s] = chunk(s, e, zt_prev)[..., 0]
        return fixed


def build_proposed(dac_a, dac_t, rvq_books: int, rvq_embed: int) -> ProposedEval:
    """build_backbones_for_eval (:527-535) + ProposedEval(...) (:661-662)."""
    return ProposedEval(dac_a.encoder, dac_a.quantizer, dac_t.encoder, dac_t.decoder,
                        dac_a.encoder.block[-1].out_channels, rvq_books, rvq_embed)

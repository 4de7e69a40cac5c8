"""Packet-loss-concealment forward of the reference's PLC scripts on the B200 kernels.

``AllPredPLC`` keeps the constructor, attribute names, state-dict keys and ``forward_step`` result of
PLC/PLC1_eval.py:442-520 (uniform packet loss) and PLC/PLC1_low_mid_high_eval.py:416-500 (burst-loss categories):
audio encoder + DAC quantizer, tactile encoder, token mask, ONE full-length CrossPredictor call, ``torch.where``,
tactile decoder -- here one CUDA program per (batch, length), forward only.  The masks are drawn on the host / with
torch's generator exactly as the reference draws them (they are inputs of the path, not part of it); pass
``mask_tokens`` to fix one.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn

from . import _lib as L
from .engine import Emitter, PackedDecoder, PackedEncoder, PackedPredictor, emit_decoder, emit_encoder, emit_predict_full
from .modules import (CrossPredictor, Decoder, Encoder, ResidualVectorQuantize, TokenNorm, _Top, _as_f32, _require_cuda,
                      pe_table)

PACKET_TOK = 2            # PLC/PLC1_eval.py:74
PACKET_LOSS_PROB = 0.5    # :75
CAT_BURST_MS = {"low": (20.0, 120.0), "medium": (120.0, 320.0), "high": (320.0, 1000.0)}   # PLC/PLC1_low_mid_high_eval.py:89-93
CAT_N_BURSTS = {"low": (1, 2), "medium": (1, 3), "high": (1, 4)}                             # :96-100


def make_token_loss_mask(batch_size: int, T_lat: int, packet_tok: int, p_loss: float, device):
    """PLC/PLC1_eval.py:418-440: packets of `packet_tok` tokens, each dropped with probability p_loss -> [B, T_lat] bool."""
    if packet_tok <= 0 or T_lat <= 0:
        return torch.zeros(batch_size, T_lat, dtype=torch.bool, device=device)
    num_packets = max(1, T_lat // packet_tok)
    lost = torch.rand(batch_size, num_packets, device=device) < p_loss
    mask = lost.unsqueeze(-1).expand(batch_size, num_packets, packet_tok).reshape(batch_size, -1)
    if mask.size(1) > T_lat:
        mask = mask[:, :T_lat]
    elif mask.size(1) < T_lat:
        pad = torch.zeros(batch_size, T_lat - mask.size(1), dtype=torch.bool, device=device)
        mask = torch.cat([mask, pad], dim=1)
    return mask


def make_category_token_loss_mask_for_category(category: str, batch_size: int, T_lat: int, tokens_per_sec: float, device,
                                               burst_ms=None, n_bursts=None):
    """PLC/PLC1_low_mid_high_eval.py:371-414: bursts drawn with python's `random` for a fixed loss category."""
    burst_ms = CAT_BURST_MS if burst_ms is None else burst_ms
    n_bursts = CAT_N_BURSTS if n_bursts is None else n_bursts
    if T_lat <= 0:
        return torch.zeros(batch_size, 0, dtype=torch.bool, device=device)
    if category not in burst_ms:
        raise ValueError(f"Unknown category: {category}")
    min_ms, max_ms = burst_ms[category]
    nb_min, nb_max = n_bursts[category]
    mask = torch.zeros(batch_size, T_lat, dtype=torch.bool)
    for b in range(batch_size):
        min_tok = max(1, int(round(min_ms * tokens_per_sec / 1000.0)))
        max_tok = min(max(min_tok, int(round(max_ms * tokens_per_sec / 1000.0))), T_lat)
        for _ in range(random.randint(nb_min, nb_max)):
            L_b = random.randint(min_tok, max_tok)
            if L_b >= T_lat:
                mask[b, :] = True
                break
            s = random.randint(0, max(0, T_lat - L_b))
            mask[b, s:s + L_b] = True
    return mask.to(device)


class AllPredPLC(_Top):
    """PLC/PLC1_eval.py:442-520.  A_ENC / A_QUANT / T_ENC / T_DEC must be this package's modules."""

    #: frames per program launch (bounds the workspace; PLC evaluation runs B = 1 per file)
    micro_batch = 16

    def __init__(self, A_ENC, A_QUANT, T_ENC, T_DEC, c_lat):
        super().__init__()
        self.A_ENC, self.A_QUANT, self.T_ENC, self.T_DEC = A_ENC, A_QUANT, T_ENC, T_DEC
        for m in (A_ENC, A_QUANT, T_ENC, T_DEC):
            for p in m.parameters():
                p.requires_grad_(False)
        self.predict = CrossPredictor(c=c_lat, heads=8, mlp_mul=2, dropout=0.1)
        self.tokennorm = TokenNorm(c_lat)          # in the checkpoint, unused by the forward (:465)
        self.packet_tok, self.p_loss = PACKET_TOK, PACKET_LOSS_PROB
        self._adopt(A_ENC=lambda pk: pk["a_enc"], T_ENC=lambda pk: pk["t_enc"], T_DEC=lambda pk: pk["t_dec"],
                    A_QUANT=lambda pk: pk["a_q"], predict=lambda pk: pk["pp"])

    def _pack(self, eng):
        for name, m, cls in (("A_ENC", self.A_ENC, Encoder), ("A_QUANT", self.A_QUANT, ResidualVectorQuantize),
                             ("T_ENC", self.T_ENC, Encoder), ("T_DEC", self.T_DEC, Decoder)):
            if not isinstance(m, cls):
                raise L.B2CError(f"AllPredPLC.{name} must be this package's {cls.__name__}; there is no PyTorch fallback path")
        return dict(a_enc=PackedEncoder.pack(eng, self.A_ENC), a_q=eng.pack_dac_rvq(list(self.A_QUANT.quantizers)),
                    n_q=self.A_QUANT.n_codebooks, t_enc=PackedEncoder.pack(eng, self.T_ENC),
                    t_dec=PackedDecoder.pack(eng, self.T_DEC), pp=PackedPredictor.pack_predictor(eng, self.predict))

    def program(self, eng, pk, nb, T):
        """ext slots: 1 a [nb,T], 2 t [nb,T], 3 y [nb,Lout], 4 mask bytes [nb,Tl], 5 audio codes i32 [nb,n_q,Tl],
        6 z_filled [nb,C,Tl] (the latents handed to the decoder, channel-major like the reference's tensor)."""
        pe, pd, pt = self._prec("enc"), self._prec("dec"), self._prec("pred")
        key = ("plc", nb, T, pe, pd, pt)
        prog = eng.programs.get(key)
        if prog is not None:
            return prog
        em = Emitter(eng)
        pp = pk["pp"]
        c = pp.c
        za, Tl = emit_encoder(em, pk["a_enc"], em.ext(1), nb, T, pe)
        qa = em.new(nb * Tl * c)
        em.dac_rvq(pk["a_q"], pk["n_q"], za, qa, em.ext(5), nb, Tl)
        em.drop(za)
        zt, Tl2 = emit_encoder(em, pk["t_enc"], em.ext(2), nb, T, pe)
        if Tl2 != Tl:
            raise L.B2CError("AllPredPLC: audio and tactile frames must have the same length")
        z_pred = emit_predict_full(em, pp, zt, em.ext(4), qa, nb, Tl, pe_table(eng, pp, self.predict, Tl), pt)
        em.drop(qa)
        z_fill = em.new(nb * Tl * c)
        em.select_rows(em.ext(4), z_pred, zt, z_fill, nb * Tl, c)     # where(mask, z_pred, zt_in): zt_in == zt off the mask
        em.drop(z_pred, zt)
        em.transpose(z_fill, em.ext(6), nb, Tl, c)
        emit_decoder(em, pk["t_dec"], z_fill, em.ext(3), nb, Tl, pd)
        prog = eng.programs[key] = em.finish(6, Tl=Tl, Lout=pk["t_dec"].out_len(Tl))
        return prog

    @torch.no_grad()
    def forward_step(self, a_1T, tc_1T, category=None, mask_tokens=None):
        """-> {"y_hat", "tgt", "latent_mask"} as the reference; ``category`` selects the burst-loss mask of
        PLC1_low_mid_high_eval.py, otherwise the uniform packet loss of PLC1_eval.py; ``mask_tokens`` [B, T_lat] bool
        overrides both.  ``last_latents`` keeps z_filled [B, C, T_lat]."""
        _require_cuda(a_1T, tc_1T)
        if a_1T.shape != tc_1T.shape or a_1T.dim() != 3 or a_1T.shape[1] != 1:
            raise ValueError(f"expected two [B, 1, T] tensors, got {tuple(a_1T.shape)} and {tuple(tc_1T.shape)}")
        dev = a_1T.device
        eng, pk = self._engine(dev)
        B, _, T = a_1T.shape
        Tl = pk["t_enc"].out_len(T)
        if B == 0 or Tl <= 0:
            raise ValueError(f"empty batch or frame too short (B={B}, T={T})")
        if mask_tokens is None:
            if category is not None:
                mask_tokens = make_category_token_loss_mask_for_category(category, B, Tl, float(Tl), dev)
            else:
                mask_tokens = make_token_loss_mask(B, Tl, self.packet_tok, self.p_loss, dev)
        if tuple(mask_tokens.shape) != (B, Tl):
            raise ValueError(f"mask_tokens must be [B={B}, T_lat={Tl}], got {tuple(mask_tokens.shape)}")
        mk = mask_tokens.to(device=dev, dtype=torch.uint8).contiguous()
        c, n_q, Lout = pk["pp"].c, pk["n_q"], pk["t_dec"].out_len(Tl)
        a, t = _as_f32(a_1T), _as_f32(tc_1T)
        y = torch.empty(B, 1, Lout, device=dev, dtype=torch.float32)
        codes = torch.empty(B, n_q, Tl, device=dev, dtype=torch.int32)
        z = torch.empty(B, c, Tl, device=dev, dtype=torch.float32)
        mb = min(B, self.micro_batch)
        for b0 in range(0, B, mb):
            nb = min(mb, B - b0)
            prog = self.program(eng, pk, nb, T)
            eng.run(prog, [a[b0:].data_ptr(), t[b0:].data_ptr(), y[b0:].data_ptr(), mk[b0:].data_ptr(),
                           codes[b0:].data_ptr(), z[b0:].data_ptr()])
        self.last_latents, self.last_audio_codes = z, codes
        n = min(Lout, T)
        y = torch.nan_to_num(y[..., :n], nan=0.0, posinf=0.0, neginf=0.0)          # finite_or_zero (:99-100)
        tgt = torch.nan_to_num(tc_1T[..., :n], nan=0.0, posinf=0.0, neginf=0.0)
        return {"y_hat": y.to(a_1T.dtype), "tgt": tgt, "latent_mask": mask_tokens.to(dev).bool().unsqueeze(1)}

    forward = forward_step


def build_plc() -> AllPredPLC:
    """build_backbones (PLC/PLC1_eval.py:523-531) + AllPredPLC(...) with random-init weights."""
    from .modules import DAC
    da, dt = DAC(), DAC()
    return AllPredPLC(da.encoder, da.quantizer, dt.encoder, dt.decoder, 1024)

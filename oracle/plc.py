"""ORACLE (test infrastructure, never shipped): CPU fp32 restatement of the packet-loss-concealment forward,
AllPredPLC.forward_step of PLC/PLC1_eval.py:442-520 (PLC1_low_mid_high_eval.py:416-500 differs only in how the token
mask is drawn).  The predictor / TokenNorm classes are the ones of oracle/proposed.py: the PLC scripts carry verbatim
copies (PLC1_eval.py:336-415; ``ffn(y + q) + (y + q)`` is the same sum as ``y = y + q; y = y + ffn(y)``).

Pinned: tests/test_oracle_cpu.py::test_plc_restatement_equals_reference_classes runs the reference's own AllPredPLC
(ast-extracted) on the same weights, inputs and mask -- bit-equal; tests/golden/plc_*.npz were made from it.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .proposed import CrossPredictor, TokenNorm

PACKET_TOK = 2            # PLC/PLC1_eval.py:74
PACKET_LOSS_PROB = 0.5    # :75


def make_token_loss_mask(batch_size, t_lat, packet_tok, p_loss, device="cpu", generator=None):
    """:418-440 -- packets of `packet_tok` tokens dropped with probability p_loss; -> [B, T_lat] bool."""
    if packet_tok <= 0 or t_lat <= 0:
        return torch.zeros(batch_size, t_lat, dtype=torch.bool, device=device)
    num_packets = max(1, t_lat // packet_tok)
    lost = torch.rand(batch_size, num_packets, device=device, generator=generator) < p_loss
    mask = lost.unsqueeze(-1).expand(batch_size, num_packets, packet_tok).reshape(batch_size, -1)
    if mask.size(1) > t_lat:
        mask = mask[:, :t_lat]
    elif mask.size(1) < t_lat:
        mask = torch.cat([mask, torch.zeros(batch_size, t_lat - mask.size(1), dtype=torch.bool, device=device)], dim=1)
    return mask


class AllPredPLC(nn.Module):
    def __init__(self, A_ENC, A_QUANT, T_ENC, T_DEC, c_lat):
        super().__init__()
        self.A_ENC, self.A_QUANT, self.T_ENC, self.T_DEC = A_ENC, A_QUANT, T_ENC, T_DEC
        for m in (A_ENC, A_QUANT, T_ENC, T_DEC):
            for p in m.parameters():
                p.requires_grad_(False)
        self.predict = CrossPredictor(c=c_lat, heads=8, mlp_mul=2, dropout=0.1)
        self.tokennorm = TokenNorm(c_lat)

    @torch.no_grad()
    def forward_step(self, a_1T, tc_1T, mask_tokens=None, trace=None):
        tw = tc_1T.shape[-1]
        za = self.A_ENC(a_1T)                                   # :479
        qa = self.A_QUANT(za)[0]                                # :480
        zt_full = self.T_ENC(tc_1T)                             # :483
        b, c, t_lat = zt_full.shape
        if mask_tokens is None:
            mask_tokens = make_token_loss_mask(b, t_lat, PACKET_TOK, PACKET_LOSS_PROB, zt_full.device)   # :487-493
        m = mask_tokens.unsqueeze(1)
        zt_in = zt_full * (~m)                                  # :497
        z_pred = self.predict(zt_in, qa)                        # :500 -- ONE call over all T_lat tokens
        z_filled = torch.where(m, z_pred, zt_in)                # :503
        y_hat = self.T_DEC(z_filled)                            # :506
        n = min(y_hat.shape[-1], tc_1T.shape[-1], tw)
        if trace is not None:
            trace.update(za=za, qa=qa, zt=zt_full, z_pred=z_pred, z_filled=z_filled)
        return {"y_hat": torch.nan_to_num(y_hat[..., :n], nan=0.0, posinf=0.0, neginf=0.0),
                "tgt": torch.nan_to_num(tc_1T[..., :n], nan=0.0, posinf=0.0, neginf=0.0), "latent_mask": m}


PLC_CASES = {
    # a 2-s file at B = 1 (the evaluation scripts run per file, :584-612): 150 tokens, one full-length attention
    "b1_t48000": dict(B=1, T=48000, kind="uniform", mask_seed=11),
    # ragged length (T_lat = 70, not a multiple of the 64-key tile), batch 2, sine mixtures
    "b2_t22400": dict(B=2, T=22400, kind="sines", mask_seed=12),
}


def build_plc_model(cls=None, seed: int = 7):
    """build_backbones (PLC/PLC1_eval.py:523-531, random-init DAC architecture) + AllPredPLC(...)."""
    from . import cases, dac_arch
    da, dt = cases.build_backbones(seed)
    cls = AllPredPLC if cls is None else cls
    return cls(da.encoder, da.quantizer, dt.encoder, dt.decoder, dac_arch.LATENT_DIM).eval()


def plc_mask(case, t_lat):
    g = torch.Generator().manual_seed(case["mask_seed"])
    return make_token_loss_mask(case["B"], t_lat, PACKET_TOK, PACKET_LOSS_PROB, generator=g)

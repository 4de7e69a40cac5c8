"""ORACLE (test infrastructure): the training-mode forward of the proposed codec with autograd,
AllPredAR.forward_step (Training/compare_dacvsproposal_3.py:300-340), restated as a function over the oracle's
ProposedEval (same attribute names and parameters as AllPredAR, :284-298), plus the scalar objective the gradient
parity tests differentiate.  Pinned: tests/test_oracle_cpu.py runs the reference's own AllPredAR (ast-extracted) on
the same weights and compares outputs and gradients; tests/golden/train_step.npz holds the reference's values.
Only tests/ may import this."""
from __future__ import annotations

import torch

from oracle import cases, proposed

AR_CHUNK_TOK = 16   # :63

TRAIN_CASE = dict(books=3, K=128, B=2, T=8000, kind="uniform")      # 25 latent tokens: two AR chunks


def finite_or_zero(x):   # :87-88
    return torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)


def forward_step(net, a_1T, tc_1T, keep=None):
    """net: oracle ProposedEval.  Returns the reference's dict; ``keep`` (dict) receives z_run with retain_grad."""
    B, _, Tw = tc_1T.shape
    za = net.A_ENC(a_1T)
    qa, *_ = net.A_QUANT(za)
    zt_teacher = net.T_ENC(tc_1T)
    B, C, Tlat = zt_teacher.shape
    z_run = torch.zeros_like(zt_teacher)
    rD_all = []
    for s in range(0, Tlat, AR_CHUNK_TOK):
        e = min(Tlat, s + AR_CHUNK_TOK)
        zt_prev = torch.zeros(B, C, e - s, dtype=zt_teacher.dtype)
        if s == 0:
            zt_prev[..., 1:] = z_run[..., s:e - 1]
        else:
            zt_prev[...] = z_run[..., s - 1:e - 1]
        qa_chunk = qa[..., s:e]
        z_pred = net.predict(zt_prev, qa_chunk)
        r = zt_teacher[..., s:e] - z_pred.detach()
        rN = torch.tanh(net.tokennorm(r))
        scale = net.scale.clamp(5e-3, 0.5)
        rD = net.proj_down(scale * rN)
        qD = vq_train(net.vq, rD)
        z_hat = z_pred + net.proj_up(qD)
        z_run[..., s:e] = z_hat
        rD_all.append(rD.detach())
    if keep is not None:
        z_run.retain_grad()
        keep["z_run"] = z_run
    y_hat = net.T_DEC(z_run)
    T = min(y_hat.shape[-1], tc_1T.shape[-1], Tw)
    return {"y_hat": finite_or_zero(y_hat[..., :T]), "tgt": finite_or_zero(tc_1T[..., :T]), "z_teacher": zt_teacher,
            "r_tokens": torch.cat(rD_all, dim=-1) if rD_all else None}


def vq_train(vq, z):
    """ResidualVQEMA.forward of the training script (:253-262): all books, straight-through."""
    B, D, T = z.shape
    x = z.permute(0, 2, 1).reshape(B * T, D)
    residual, q_sum = x, torch.zeros_like(x)
    for cb in vq.books:
        emb = cb.detach().to(z.dtype)
        idx = proposed.nearest_code(residual, emb)
        q = torch.nn.functional.embedding(idx, emb)
        q_sum = q_sum + (q - residual).detach() + residual
        residual = residual - q
    return q_sum.view(B, T, D).permute(0, 2, 1).contiguous()


def objective(out):
    """A fixed scalar of the step's output for gradient parity: the waveform L1 term of the training loss (safe_l1,
    :208-209) plus a seeded random projection of y_hat (so that every output sample carries gradient)."""
    y, tgt = out["y_hat"], out["tgt"]
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(11)).to(y.device)
    return torch.nn.functional.l1_loss(y, tgt) + (w * y).mean()


GRAD_KEYS = ("scale", "proj_up.bias", "proj_up.weight", "proj_down.weight", "proj_down.bias", "tokennorm.ln.weight",
             "predict.out.weight", "predict.q_proj.weight", "predict.ffn.3.bias", "predict.ln_kv.weight")


def run_case(net, case=TRAIN_CASE):
    """-> (loss, {param name: grad}, g_z_run, out) for the oracle model `net` in eval mode (dropout inactive)."""
    a, t = cases.codec_inputs(case)
    net.eval()
    for p in net.parameters():
        p.grad = None
    keep = {}
    out = forward_step(net, a, t, keep)
    loss = objective(out)
    loss.backward()
    named = dict(net.named_parameters())
    return float(loss), {k: named[k].grad.clone() for k in GRAD_KEYS}, keep["z_run"].grad.clone(), out

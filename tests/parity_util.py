"""Shared by the GPU parity tests: what counts as agreement between code indices of the CUDA path and of the oracle.

The rule (BASELINE.json north_star: "indices bit-exact except at documented fp near-ties"): along the residual
stages of ONE token the FIRST differing index must sit at an oracle top-1 / top-2 score margin below the stated
tolerance -- after such a flip the token's later stages quantise a different residual and are not comparable.  A
flipped token also changes what its neighbours see: the DAC codes feed the predictor's keys / values of the whole
16-token chunk, and a chunk's last reconstructed latent is the query input of the next chunk's first token
(Evaluation/dac_vcpwq_proposed6_latency.py:462-470), so those tokens are reported as *dependent* and not compared.
Nothing else is waived."""
import numpy as np
import torch

CHUNK = 16   # AR_CHUNK_TOK (:337)


def stage_flips(got, gold, margin, tie):
    """got / gold [B, S, T] integer, margin [B, S, T] (oracle top-1 - top-2).
    -> (ok, n_tokens_flipped, worst_margin, flipped [B, T] bool): ok iff every token's first differing stage is a
    near-tie."""
    got, gold = got.long(), gold.long()
    bad = got != gold
    flipped = bad.any(dim=1)
    if not bad.any():
        return True, 0, 0.0, flipped
    first = (bad.int().cumsum(dim=1) == 1) & bad
    m = margin[first]
    return bool((m < tie).all()), int(flipped.sum()), float(m.max()), flipped


def dependent_tokens(code_flipped, own_flipped, chunk=CHUNK):
    """Tokens whose own-RVQ indices legitimately depend on a token that flipped: the whole chunk of a flipped DAC
    code (keys / values), and the first token of the chunk that follows any flipped token at a chunk end."""
    B, T = code_flipped.shape
    dep = torch.zeros(B, T, dtype=torch.bool)
    for s in range(0, T, chunk):
        e = min(T, s + chunk)
        hit = code_flipped[:, s:e].any(dim=1)
        dep[:, s:e] |= hit[:, None]
        if s > 0:
            dep[:, s] |= (own_flipped[:, s - 1] | dep[:, s - 1])
    return dep


def check_against_oracle(idx, codes, tr, tie_own, tie_code):
    """idx [B, books, T], codes [B, n_q, T] from the CUDA path; tr = the oracle's trace of the same inputs.
    Returns a dict: exact (no flip anywhere), n_code / n_own flipped tokens, and asserts the near-tie rule."""
    ok_c, n_c, worst_c, code_f = stage_flips(codes, tr["a_codes"], tr["a_margin"], tie_code)
    assert ok_c, f"{n_c} tokens with a DAC-code flip that is not a near-tie (worst oracle margin {worst_c:.3e} >= {tie_code})"
    if idx.shape[1] == 0:
        return dict(exact=n_c == 0, n_code=n_c, n_own=0, n_dependent=0)
    _, _, _, own_f_all = stage_flips(idx, tr["idx"], tr["margin"], tie_own)
    dep = dependent_tokens(code_f, own_f_all)
    keep = ~dep
    sel = keep[:, None, :].expand_as(idx)
    got, gold, mar = idx.long().clone(), tr["idx"].long(), tr["margin"]
    got[~sel] = gold[~sel]                      # dependent tokens: not compared
    ok_o, n_o, worst_o, _ = stage_flips(got, gold, mar, tie_own)
    assert ok_o, f"{n_o} tokens with an index flip that is not a near-tie (worst oracle margin {worst_o:.3e} >= {tie_own})"
    return dict(exact=(n_c == 0 and n_o == 0 and not dep.any()), n_code=n_c, n_own=n_o, n_dependent=int(dep.sum()))


def psnr(y, ref):
    mse = float(((y - ref) ** 2).mean())
    peak = float(ref.abs().max()) or 1.0
    return 10 * np.log10(peak * peak / max(mse, 1e-30))

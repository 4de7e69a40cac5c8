"""ORACLE helper (test infrastructure).  Loads the reference's OWN model classes from
/root/reference without importing the script: the scripts ``import dac`` /
``soundfile`` / ``matplotlib`` (absent here) and create directories under
/home/student at import time (Evaluation/dac_vcpwq_proposed6_latency.py:64,77).
Only the top-level definitions named below are compiled, from the file where they
lie.  /root/reference does not exist on the GPU box: callers must check
``available()`` and nothing under ``-m gpu`` may depend on it.
"""
from __future__ import annotations

import ast
import math
import os

REF_ROOT = "/root/reference"
REF_SCRIPT = os.path.join(REF_ROOT, "Evaluation", "dac_vcpwq_proposed6_latency.py")
WANTED = ("CODE_DIM", "AR_CHUNK_TOK", "PosEnc1D", "TokenNorm", "CrossPredictor",
          "ResidualVQEMA", "ProposedEval")


def available() -> bool:
    return os.path.isfile(REF_SCRIPT)


PLC_SCRIPT = os.path.join(REF_ROOT, "PLC", "PLC1_eval.py")
PLC_WANTED = ("PACKET_TOK", "PACKET_LOSS_PROB", "finite_or_zero", "PosEnc1D", "TokenNorm", "CrossPredictor",
              "make_token_loss_mask", "AllPredPLC")
TRAIN_SCRIPT = os.path.join(REF_ROOT, "Training", "compare_dacvsproposal_3.py")
TRAIN_WANTED = ("CODE_DIM", "RVQ_N_BOOKS", "RVQ_EMBED", "EMA_DECAY", "AR_CHUNK_TOK", "ResidualVQEMA")
TRAIN_STEP_WANTED = TRAIN_WANTED + ("finite_or_zero", "PosEnc1D", "TokenNorm", "CrossPredictor", "AllPredAR")


def reference_train_model(oracle_net, books: int, K: int):
    """The reference's AllPredAR (Training/compare_dacvsproposal_3.py:284-340) on the oracle model's backbones and
    parameters (AllPredAR and ProposedEval name their layers alike)."""
    ns = load_reference_classes(TRAIN_SCRIPT, TRAIN_STEP_WANTED)
    ns["RVQ_N_BOOKS"], ns["RVQ_EMBED"] = books, K
    ref = ns["AllPredAR"](oracle_net.A_ENC, oracle_net.A_QUANT, oracle_net.T_ENC, oracle_net.T_DEC, 1024)
    ref.load_state_dict(oracle_net.state_dict(), strict=True)
    return ref.eval()


METRIC_SCRIPT = os.path.join(REF_ROOT, "Evaluation", "compare_dacvsproposal_5_eval.py")
METRIC_WANTED = ("EVAL_SR", "ORIG_3K", "ALIGN_MAX_SHIFT_SAMPLES", "_MEL_CACHE", "resample_f32", "_mel_mag", "stsim_batch",
                 "psnr_batch", "align_pair_24k", "psnr_3k_aligned_batch")


def load_reference_metrics() -> dict:
    """The reference's own metric functions (Evaluation/compare_dacvsproposal_5_eval.py:91-97, :139-223); they call
    torchaudio (Resample, MelScale), which this container has."""
    import torchaudio
    return load_reference_classes(METRIC_SCRIPT, METRIC_WANTED, extra_ns={"torchaudio": torchaudio})


def load_reference_classes(path: str = REF_SCRIPT, wanted=None, extra_ns=None) -> dict:
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    WANTED = globals()["WANTED"] if wanted is None else wanted
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"math": math, "torch": torch, "nn": nn, "F": F}
    ns.update(extra_ns or {})
    for node in tree.body:
        name = None
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
            name = node.name
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
        if name in WANTED:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    missing = [w for w in WANTED if w not in ns]
    if missing:
        raise RuntimeError(f"reference definitions not found: {missing}")
    return ns


class IndexSpy:
    """Wraps the reference's ResidualVQEMA._nearest_l2 so the indices it discards
    (:431-435) are recorded, in call order: chunk-major, book-minor."""

    def __init__(self, ns):
        self.ns = ns
        self.calls = []
        self._orig = ns["ResidualVQEMA"].__dict__["_nearest_l2"]

    def __enter__(self):
        spy = self

        def nearest(x, emb):
            sc = x @ emb.t() - 0.5 * (emb * emb).sum(dim=1).unsqueeze(0)
            idx = sc.argmax(dim=1)
            spy.calls.append(idx.clone())
            return idx

        self.ns["ResidualVQEMA"]._nearest_l2 = staticmethod(nearest)
        return self

    def __exit__(self, *exc):
        self.ns["ResidualVQEMA"]._nearest_l2 = self._orig
        return False

    def indices(self, batch: int, books: int, chunk_lens):
        """-> [B, books, sum(chunk_lens)] int64"""
        import torch
        out = []
        it = iter(self.calls)
        for n in chunk_lens:
            per_book = [next(it).view(batch, n) for _ in range(books)]
            out.append(torch.stack(per_book, dim=1))
        return torch.cat(out, dim=-1)

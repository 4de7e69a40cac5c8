"""ORACLE golden-vector generator (test infrastructure).

Run HERE (the build container, where /root/reference is mounted):

    python -m oracle.make_golden

It executes the REFERENCE's own classes (ast-extracted from
Evaluation/dac_vcpwq_proposed6_latency.py by oracle/ref_loader.py) on top of the
DAC-architecture restatement (oracle/dac_arch.py; the third-party package is
absent) and writes small fixtures to tests/golden/.  The fixtures travel to the
GPU box; /root/reference does not.

Recipe: module construction under torch.manual_seed(7)
(reference SEED, Training/compare_dacvsproposal_3.py:50), inputs U(-1,1) from
torch.Generator().manual_seed(123).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import dac_arch, ref_loader  # noqa: E402
from oracle.cases import (  # noqa: E402
    CODEC_CASES, SEARCH_CASES, build_reference_style_model, codec_inputs, search_inputs,
    chunk_lengths, predictor_inputs,
)

OUT = os.path.join(ROOT, "tests", "golden")


def weights_fingerprint(model) -> np.ndarray:
    """A few float64 sums over the state dict, to detect RNG drift between the box
    that made the goldens and the box that rebuilds the model from the seed."""
    sd = model.state_dict()
    keys = sorted(sd.keys())
    tot = np.array([float(sd[k].double().sum()) for k in keys if sd[k].dtype.is_floating_point])
    return np.array([tot.sum(), np.abs(tot).sum(), float(len(keys))])


def main():
    if not ref_loader.available():
        raise SystemExit("reference not mounted; goldens can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load_reference_classes()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- full-path cases: reference ProposedEval.forward_eval ----
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, case in CODEC_CASES.items():
        if only and name not in only:
            continue
        model = build_reference_style_model(ns["ProposedEval"], case)
        a, t = codec_inputs(case)
        tl = a.shape[-1] // dac_arch.HOP
        with ref_loader.IndexSpy(ns) as spy, torch.no_grad():
            z_run = model.encode_latents(a, t, books_use=case.get("books_use"))
            y = model.T_DEC(z_run)
            za = model.A_ENC(a)
            qa, a_codes, *_ = model.A_QUANT(za)
            zt = model.T_ENC(t)
        books_used = min(case.get("books_use") or case["books"], case["books"])
        idx = spy.indices(a.shape[0], books_used, chunk_lengths(tl))
        np.savez_compressed(
            os.path.join(OUT, f"codec_{name}.npz"),
            fingerprint=weights_fingerprint(model),
            idx=idx.numpy().astype(np.int16),
            a_codes=a_codes.numpy().astype(np.int16),
            y=y.numpy(), z_run=z_run.numpy().astype(np.float32),
            za_sample=za[:, ::64, :].numpy(), zt_sample=zt[:, ::64, :].numpy(),
            qa_sample=qa[:, ::64, :].numpy(),
        )
        print("wrote", name, "y", tuple(y.shape), "idx", tuple(idx.shape))

    if only and not ({"plc", "ema", "metrics", "train"} & set(only)):
        return
    # ---- training step with autograd: the reference's AllPredAR.forward_step + backward ----
    if not only or "train" in only:
        from oracle import training as otr
        from oracle import proposed as oproposed
        case = otr.TRAIN_CASE
        net = build_reference_style_model(oproposed.ProposedEval, case)
        ref = ref_loader.reference_train_model(net, case["books"], case["K"])
        a, t = codec_inputs(case)
        holder = {}
        hook = ref.T_DEC.register_forward_pre_hook(lambda m, inp: (inp[0].retain_grad(), holder.__setitem__("z", inp[0]))[0] and None)
        out = ref.forward_step(a, t)
        hook.remove()
        loss = otr.objective(out)
        loss.backward()
        named = dict(ref.named_parameters())
        blobs = dict(loss=np.array(float(loss)), y_hat=out["y_hat"].detach().numpy(), g_z_run=holder["z"].grad.numpy(),
                     fingerprint=weights_fingerprint(net))
        for k in otr.GRAD_KEYS:
            blobs["grad_" + k] = named[k].grad.numpy()[..., :32] if named[k].grad.dim() == 2 else named[k].grad.numpy()
        np.savez_compressed(os.path.join(OUT, "train_step.npz"), **blobs)
        print("wrote train_step", float(loss), {k: v.shape for k, v in blobs.items()})
    # ---- evaluation metrics: the reference's own functions (Evaluation/compare_dacvsproposal_5_eval.py) ----
    if not only or "metrics" in only:
        from oracle import metrics as om
        mns = ref_loader.load_reference_metrics()
        blobs = {}
        for kind, B in (("shifted", 8), ("plain", 3)):
            ref, est, lags = om.metric_inputs(B=B, kind=kind)
            blobs[f"{kind}_lags"] = np.array(lags)
            blobs[f"{kind}_stsim"] = np.array(mns["stsim_batch"](ref, est))
            blobs[f"{kind}_psnr"] = np.array(mns["psnr_batch"](ref, est))
            blobs[f"{kind}_psnr3k"] = np.array(mns["psnr_3k_aligned_batch"](ref, est))
            blobs[f"{kind}_shift"] = np.array([mns["align_pair_24k"](ref[b:b + 1], est[b:b + 1])[2] for b in range(B)])
            blobs[f"{kind}_ref3k"] = mns["resample_f32"](ref, 24000, 3000).numpy()
            blobs[f"{kind}_ref16k_head"] = mns["resample_f32"](ref[:1], 24000, 16000).numpy()[..., :512]
        np.savez_compressed(os.path.join(OUT, "metrics.npz"), **blobs)
        print("wrote metrics", {k: v.shape for k, v in blobs.items()})
    # ---- packet-loss concealment: the reference's AllPredPLC (PLC/PLC1_eval.py) with a seeded token mask ----
    if not only or "plc" in only:
        from oracle import plc
        pns = ref_loader.load_reference_classes(ref_loader.PLC_SCRIPT, ref_loader.PLC_WANTED)
        model = plc.build_plc_model(pns["AllPredPLC"])
        for name, case in plc.PLC_CASES.items():
            a, t = codec_inputs(case)
            tl = a.shape[-1] // dac_arch.HOP
            mask = plc.plc_mask(case, tl)
            # the reference draws the mask inside forward_step (torch.rand on the default generator): hand it ours
            pns["make_token_loss_mask"] = lambda batch_size, T_lat, packet_tok, p_loss, device, _m=mask: _m
            with torch.no_grad():
                out = model.forward_step(a, t)
            assert torch.equal(out["latent_mask"][:, 0], mask)
            np.savez_compressed(os.path.join(OUT, f"plc_{name}.npz"), y_hat=out["y_hat"].numpy(), mask=mask.numpy(),
                                fingerprint=weights_fingerprint(model))
            print("wrote plc", name, tuple(out["y_hat"].shape), int(mask.sum()), "of", mask.numel(), "tokens masked")
    # ---- ResidualVQEMA.ema_step of Training/compare_dacvsproposal_3.py ----
    if not only or "ema" in only:
        tns = ref_loader.load_reference_classes(ref_loader.TRAIN_SCRIPT, ref_loader.TRAIN_WANTED)
        torch.manual_seed(7)
        vq = tns["ResidualVQEMA"](dim=96, n_books=4, n_embed=128, decay=0.99)
        z = 0.3 * torch.randn(3, 96, 75, generator=torch.Generator().manual_seed(123))
        blobs = dict(n_books=4, decay=0.99, z_tokens=z.numpy())
        for i, b in enumerate(vq.books):
            blobs[f"book{i}_before"] = b.detach().clone().numpy()
        vq.ema_step(z)
        for i, b in enumerate(vq.books):
            blobs[f"book{i}_after"] = b.detach().clone().numpy()
        np.savez_compressed(os.path.join(OUT, "ema_step.npz"), **blobs)
        print("wrote ema_step")
    if only:
        return
    # ---- module-level: reference CrossPredictor + ResidualVQEMA on seeded tensors ----
    torch.manual_seed(7)
    pred = ns["CrossPredictor"](c=1024).eval()
    zp, za = predictor_inputs()
    with torch.no_grad():
        out = pred(zp, za)
    np.savez_compressed(os.path.join(OUT, "predictor.npz"), out=out.numpy())
    print("wrote predictor", tuple(out.shape))

    # ---- _nearest_l2 on the stress shapes ----
    blobs = {}
    for name, (n, d, k) in SEARCH_CASES.items():
        x, emb = search_inputs(n, d, k)
        with torch.no_grad():
            idx = ns["ResidualVQEMA"]._nearest_l2(x, emb)
            sc = x @ emb.t() - 0.5 * (emb * emb).sum(dim=1).unsqueeze(0)
            top2 = sc.topk(2, dim=1).values
        blobs[f"{name}_idx"] = idx.numpy().astype(np.int16 if k <= 32767 else np.int32)
        blobs[f"{name}_margin"] = (top2[:, 0] - top2[:, 1]).numpy()
    np.savez_compressed(os.path.join(OUT, "nearest.npz"), **blobs)
    print("wrote nearest", list(SEARCH_CASES))


if __name__ == "__main__":
    main()

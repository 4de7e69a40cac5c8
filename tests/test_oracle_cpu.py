"""CPU suite, part 1: the oracle is pinned against (a) the reference's own classes when
/root/reference is mounted, (b) the committed golden vectors (made by running the reference
classes, oracle/make_golden.py), (c) the facts the reference's result JSON pins about the
third-party backbone."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, dac_arch, proposed, ref_loader

SMALL = dict(books=3, K=128, B=1, T=4800, kind="uniform")


def test_backbone_pinned_facts():
    # eval_all_vs_dac24_vcpwq_rawPSNR_latency.json:11-12 -> tps = 75, bins = 1024; 3.5_eval.py:75 -> n_q >= 32
    d = dac_arch.DAC()
    n = sum(p.numel() for p in d.parameters())
    assert abs(n / 1e6 - 74.7) < 0.1
    assert d.quantizer.n_codebooks >= 32 and d.quantizer.codebook_size == 1024
    with torch.no_grad():
        z = d.encoder(torch.zeros(1, 1, 24000))
        assert tuple(z.shape) == (1, 1024, 75)
        y = d.decoder(z[..., :5])
    assert y.shape[-1] == 5 * 320 - 8


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_restatement_equals_reference_classes():
    ns = ref_loader.load_reference_classes()
    ref = cases.build_reference_style_model(ns["ProposedEval"], SMALL)
    mine = cases.build_reference_style_model(proposed.ProposedEval, SMALL)
    assert [k for k in ref.state_dict()] == [k for k in mine.state_dict()]
    for k, v in ref.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k
    a, t = cases.codec_inputs(SMALL)
    with ref_loader.IndexSpy(ns) as spy:
        y_ref = ref.forward_eval(a, t, None)
    tr = {}
    y = mine.forward_eval(a, t, None, trace=tr)
    assert torch.equal(y, y_ref)
    assert torch.equal(tr["idx"], spy.indices(1, 3, cases.chunk_lengths(15)))
    # books_use prefix
    assert torch.equal(mine.forward_eval(a, t, 2), ref.forward_eval(a, t, 2))
    # module level
    zp, zk = cases.predictor_inputs()
    assert torch.equal(mine.predict(zp, zk), ref.predict(zp, zk))


def test_two_pass_schedule_is_exact():
    m = cases.build_reference_style_model(proposed.ProposedEval, dict(books=2, K=128, B=2, T=12800, kind="uniform"))
    with torch.no_grad():
        for p in m.predict.parameters():   # amplify the predictor so a wrong schedule is visible
            p.mul_(3.0)
    a, t = cases.codec_inputs(dict(B=2, T=12800, kind="uniform"))
    z1 = m.encode_latents(a, t)
    z2 = m.encode_latents_two_pass(a, t)
    assert torch.equal(z1, z2)


def _near_tie_ok(idx, gold, margin, tol):
    """index mismatches are only tolerated where the oracle's own top-1/top-2 margin is tiny, or
    downstream (later book, same token) of such a flip."""
    bad = idx != gold
    if not bad.any():
        return True
    first = bad.int().cumsum(dim=1) == 1
    first &= bad
    return bool((margin[first] < tol).all())


def test_oracle_reproduces_codec_golden(golden_dir, oracle_models):
    name = "cal_b4k256_use3_short"
    case = cases.CODEC_CASES[name]
    g = np.load(os.path.join(golden_dir, f"codec_{name}.npz"))
    m = oracle_models(name)
    a, t = cases.codec_inputs(case)
    tr = {}
    y = m.forward_eval(a, t, case["books_use"], trace=tr)
    assert _near_tie_ok(tr["idx"], torch.from_numpy(g["idx"].astype(np.int64)), tr["margin"], 1e-5)
    if torch.equal(tr["idx"], torch.from_numpy(g["idx"].astype(np.int64))):
        assert np.abs(y.numpy() - g["y"]).max() < 2e-5
    assert tuple(y.shape) == g["y"].shape


def test_nearest_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nearest.npz"))
    for name, (n, d, k) in cases.SEARCH_CASES.items():
        if n * k > 2 ** 23:
            continue
        x, emb = cases.search_inputs(n, d, k)
        idx = proposed.nearest_code(x, emb).numpy()
        bad = idx != g[f"{name}_idx"]
        assert (g[f"{name}_margin"][bad] < 1e-5).all(), name
        # cross-check against the L2 definition (the reference never calls cdist)
        if n <= 1024:
            alt = torch.cdist(x, emb).argmin(1).numpy()
            assert ((alt != idx).mean()) < 0.01


def test_predictor_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "predictor.npz"))
    torch.manual_seed(7)
    pred = proposed.CrossPredictor(c=1024).eval()
    zp, zk = cases.predictor_inputs()
    with torch.no_grad():
        out = pred(zp, zk)
    assert np.abs(out.numpy() - g["out"]).max() < 1e-4


def test_backbone_restatement_against_the_real_dac_package():
    """Closes the 'parity unpinned' gap of the third-party backbone the day a box has descript-audio-codec: the
    restatement must have dac.DAC's state-dict keys and shapes, and produce its outputs from the same weights."""
    dac = pytest.importorskip("dac", reason="descript-audio-codec is not installed (no network in the build image)")
    torch.manual_seed(0)
    real = dac.DAC(encoder_dim=64, encoder_rates=[2, 4, 5, 8], decoder_dim=1536, decoder_rates=[8, 5, 4, 2],
                   n_codebooks=32, codebook_size=1024, codebook_dim=8, sample_rate=24000).eval()
    mine = dac_arch.DAC().eval()
    sd_real, sd_mine = real.state_dict(), mine.state_dict()
    assert sorted(sd_real) == sorted(sd_mine)
    for k, v in sd_real.items():
        assert tuple(v.shape) == tuple(sd_mine[k].shape), k
    mine.load_state_dict(sd_real)
    x = torch.rand(1, 1, 24000, generator=torch.Generator().manual_seed(5)) * 2 - 1
    with torch.no_grad():
        z_r, z_m = real.encoder(x), mine.encoder(x)
        assert torch.allclose(z_r, z_m, atol=1e-5), float((z_r - z_m).abs().max())
        q_r, q_m = real.quantizer(z_r), mine.quantizer(z_r)
        assert torch.equal(q_r[1], q_m[1])                      # codes
        assert torch.allclose(q_r[0], q_m[0], atol=1e-5)
        y_r, y_m = real.decoder(q_r[0]), mine.decoder(q_r[0])
        assert y_r.shape == y_m.shape
        assert torch.allclose(y_r, y_m, atol=1e-5), float((y_r - y_m).abs().max())


def test_near_tie_rule_of_the_parity_suite():
    """tests/parity_util.py: the first differing stage of a token decides; later stages and dependent tokens do not."""
    from parity_util import check_against_oracle, dependent_tokens, stage_flips
    gold = torch.zeros(1, 3, 40, dtype=torch.long)
    margin = torch.full((1, 3, 40), 0.1)
    got = gold.clone()
    assert stage_flips(got, gold, margin, 1e-5)[:2] == (True, 0)
    got[0, 1, 7] = 5; got[0, 2, 7] = 9            # first flip at stage 1 ...
    assert stage_flips(got, gold, margin, 1e-5)[0] is False
    margin[0, 1, 7] = 1e-7                        # ... is a near-tie: the cascade at stage 2 is not judged
    ok, n, worst, fl = stage_flips(got, gold, margin, 1e-5)
    assert ok and n == 1 and worst < 1e-6 and bool(fl[0, 7])
    # a flipped token at the end of a chunk makes the next chunk's head dependent; a flipped DAC code its whole chunk
    own = torch.zeros(1, 40, dtype=torch.bool); own[0, 15] = True
    code = torch.zeros(1, 40, dtype=torch.bool); code[0, 35] = True
    dep = dependent_tokens(code, own)
    assert dep[0].nonzero().flatten().tolist() == [16] + list(range(32, 40))
    tr = dict(a_codes=torch.zeros(1, 2, 40, dtype=torch.long), a_margin=torch.full((1, 2, 40), 0.1), idx=gold, margin=margin)
    res = check_against_oracle(got, torch.zeros(1, 2, 40, dtype=torch.long), tr, 1e-5, 1e-5)
    assert res == dict(exact=False, n_code=0, n_own=1, n_dependent=0)
    bad = got.clone(); bad[0, 0, 20] = 3          # a real mismatch is never waived
    with pytest.raises(AssertionError):
        check_against_oracle(bad, torch.zeros(1, 2, 40, dtype=torch.long), tr, 1e-5, 1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_plc_restatement_equals_reference_classes():
    """oracle/plc.py against the reference's own AllPredPLC (PLC/PLC1_eval.py, ast-extracted): same weights, same
    frame, same token mask (drawn by the reference's make_token_loss_mask under a seed) -> bit-equal output."""
    from oracle import plc
    ns = ref_loader.load_reference_classes(ref_loader.PLC_SCRIPT, ref_loader.PLC_WANTED)
    ns["DEVICE"] = "cpu"
    ref = plc.build_plc_model(ns["AllPredPLC"])
    mine = plc.build_plc_model()
    assert list(ref.state_dict()) == list(mine.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k
    case = dict(B=1, T=9600, kind="uniform")
    a, t = cases.codec_inputs(case)
    torch.manual_seed(5)
    with torch.no_grad():
        out_ref = ref.forward_step(a, t)
    mask = out_ref["latent_mask"][:, 0]
    assert mask.any() and not mask.all()
    torch.manual_seed(5)
    assert torch.equal(plc.make_token_loss_mask(1, 30, plc.PACKET_TOK, plc.PACKET_LOSS_PROB), mask)
    out = mine.forward_step(a, t, mask_tokens=mask)
    assert torch.equal(out["y_hat"], out_ref["y_hat"]) and torch.equal(out["tgt"], out_ref["tgt"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_ema_step_restatement_equals_reference_class():
    """oracle.proposed.ema_step against ResidualVQEMA.ema_step of Training/compare_dacvsproposal_3.py:264-276."""
    ns = ref_loader.load_reference_classes(ref_loader.TRAIN_SCRIPT, ref_loader.TRAIN_WANTED)
    torch.manual_seed(3)
    vq = ns["ResidualVQEMA"](dim=96, n_books=3, n_embed=64, decay=0.99)
    books = [b.detach().clone() for b in vq.books]
    z = 0.3 * torch.randn(2, 96, 40, generator=torch.Generator().manual_seed(4))
    vq.ema_step(z)
    info = proposed.ema_step(books, z, 0.99)
    for b_ref, b in zip(vq.books, books):
        assert torch.equal(b_ref.data, b)
    assert int(info[0]["counts"].sum()) == 80


def test_plc_and_ema_goldens():
    """The committed fixtures (made by the reference's classes, oracle/make_golden.py) against the restatements."""
    from oracle import plc
    model = plc.build_plc_model()
    for name, case in plc.PLC_CASES.items():
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"plc_{name}.npz"))
        a, t = cases.codec_inputs(case)
        mask = torch.from_numpy(g["mask"])
        assert torch.equal(plc.plc_mask(case, mask.shape[1]), mask)
        out = model.forward_step(a, t, mask_tokens=mask)
        assert torch.equal(out["y_hat"], torch.from_numpy(g["y_hat"])), name
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ema_step.npz"))
    books = [torch.from_numpy(g[f"book{i}_before"]).clone() for i in range(int(g["n_books"]))]
    proposed.ema_step(books, torch.from_numpy(g["z_tokens"]), float(g["decay"]))
    for i, b in enumerate(books):
        assert torch.equal(b, torch.from_numpy(g[f"book{i}_after"])), i


# ---------------------------------------------------------------------------------------------
# evaluation metrics: oracle/metrics.py against the reference's own functions and their goldens
# ---------------------------------------------------------------------------------------------
def _have_torchaudio():
    try:
        import torchaudio  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.skipif(not (ref_loader.available() and _have_torchaudio()), reason="reference tree / torchaudio not here")
def test_metric_restatement_equals_reference_functions():
    """Bit-equal: stsim, psnr, aligned 3 kHz PSNR, best shift, both resamplers (torchaudio's tables restated)."""
    import warnings
    from oracle import metrics as om
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mns = ref_loader.load_reference_metrics()
        ref, est, _ = om.metric_inputs(B=3, T=12000, seed=5)
        assert mns["stsim_batch"](ref, est) == om.stsim_batch(ref, est)
        assert mns["psnr_batch"](ref, est) == om.psnr_batch(ref, est)
        assert mns["psnr_3k_aligned_batch"](ref, est) == om.psnr_3k_aligned_batch(ref, est)
        for b in range(3):
            ra, ea, s = mns["align_pair_24k"](ref[b:b + 1], est[b:b + 1])
            rb, eb, s2 = om.align_pair_24k(ref[b:b + 1], est[b:b + 1])
            assert s == s2 and torch.equal(ra, rb) and torch.equal(ea, eb)
        for sr in (3000, 16000, 44100):
            assert torch.equal(mns["resample_f32"](ref, 24000, sr), om.resample_f32(ref, 24000, sr))


def test_metric_goldens_and_product_tables(golden_dir):
    """tests/golden/metrics.npz was written by the reference's functions; the restatement reproduces it, and the
    constant tables the CUDA path uploads (filter bank of the resampler, mel filters) are the oracle's bit for bit."""
    from oracle import metrics as om
    from multimodal_vqvae_compression_audio_tactile_b200 import metrics as pm
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for kind, B in (("shifted", 8), ("plain", 3)):
        ref, est, lags = om.metric_inputs(B=B, kind=kind)
        assert list(g[f"{kind}_lags"]) == lags
        np.testing.assert_array_equal(np.array(om.stsim_batch(ref, est)), g[f"{kind}_stsim"])
        np.testing.assert_array_equal(np.array(om.psnr_batch(ref, est)), g[f"{kind}_psnr"])
        np.testing.assert_array_equal(np.array(om.psnr_3k_aligned_batch(ref, est)), g[f"{kind}_psnr3k"])
        assert list(g[f"{kind}_shift"]) == lags          # the synthetic lag is what the search finds
        np.testing.assert_array_equal(om.resample_f32(ref, 24000, 3000).numpy(), g[f"{kind}_ref3k"])
    for a, b in ((24000, 3000), (24000, 16000), (3000, 24000), (24000, 44100)):
        ko, wo, oo, no = om.sinc_resample_kernel(a, b)
        kp, wp, op_, np_ = pm.sinc_resample_kernel(a, b)
        assert (wo, oo, no) == (wp, op_, np_) and torch.equal(ko[:, 0], kp)
    assert torch.equal(om.mel_filterbank(257, 0.0, 12000.0, 64, 24000), pm.mel_filterbank())


# ---------------------------------------------------------------------------------------------
# training step with autograd
# ---------------------------------------------------------------------------------------------
def _train_grad(g, k):
    return g[..., :32] if g.dim() == 2 else g


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_training_step_restatement_equals_reference_class():
    """oracle/training.py against the reference's own AllPredAR.forward_step + backward: bit-equal output, loss and
    parameter gradients."""
    import warnings
    from oracle import training as otr
    case = otr.TRAIN_CASE
    net = cases.build_reference_style_model(proposed.ProposedEval, case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_loader.reference_train_model(net, case["books"], case["K"])
        a, t = cases.codec_inputs(case)
        out_r = ref.forward_step(a, t)
        loss_r = otr.objective(out_r)
        loss_r.backward()
        g_ref = {k: dict(ref.named_parameters())[k].grad.clone() for k in otr.GRAD_KEYS}
        loss_o, g_or, _, out_o = otr.run_case(net, case)
    assert float(loss_r) == loss_o and torch.equal(out_r["y_hat"], out_o["y_hat"])
    assert torch.equal(out_r["r_tokens"], out_o["r_tokens"])
    for k in otr.GRAD_KEYS:
        assert torch.equal(g_ref[k], g_or[k]), k


def test_training_step_golden(golden_dir):
    from oracle import training as otr
    g = np.load(os.path.join(golden_dir, "train_step.npz"))
    net = cases.build_reference_style_model(proposed.ProposedEval, otr.TRAIN_CASE)
    loss, grads, g_z, out = otr.run_case(net, otr.TRAIN_CASE)
    assert loss == float(g["loss"])
    np.testing.assert_array_equal(out["y_hat"].detach().numpy(), g["y_hat"])
    np.testing.assert_array_equal(g_z.numpy(), g["g_z_run"])
    for k in otr.GRAD_KEYS:
        np.testing.assert_array_equal(_train_grad(grads[k], k).numpy(), g["grad_" + k])

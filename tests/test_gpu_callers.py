"""GPU parity of the callers either side of the codec path, through the C-ABI library:
the packet-loss-concealment forward against golden vectors made by the reference's AllPredPLC, ResidualVQEMA.ema_step
against the golden made by the reference's class, and the receiver (indices -> reconstruction) against the oracle's
receiver loop."""
import os

import numpy as np
import pytest
import torch

import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from oracle import cases, plc as oplc, proposed
from parity_util import psnr, stage_flips

pytestmark = pytest.mark.gpu

# plan f32: FP32 FFMA everywhere; plan tc: bf16x3 encoders / predictor linears, single-pass bf16 decoder.  The PLC path
# has no quantizer between the predictor and the decoder, so everything is a floating-point tolerance: latents
# |z| ~ 0.1..3, waveform |y| <= ~0.15.
Z_TOL = {"f32": 2e-4, "tc": 1.5e-3}
Y_TOL = {"f32": 3e-5, "tc": 2e-3}
PSNR_MIN = {"f32": 80.0, "tc": 45.0}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def plc_oracle():
    return oplc.build_plc_model()


@pytest.mark.parametrize("plan", ["f32", "tc"])
@pytest.mark.parametrize("name", list(oplc.PLC_CASES))
def test_plc_forward_against_reference_golden(name, plan, dev, golden_dir, plc_oracle):
    case = oplc.PLC_CASES[name]
    g = np.load(os.path.join(golden_dir, f"plc_{name}.npz"))
    net = pkg.build_plc()
    net.load_state_dict(plc_oracle.state_dict())
    for m in (net, net.A_ENC, net.T_ENC, net.T_DEC, net.A_QUANT, net.predict):
        m.precision = plan
    a, t = cases.codec_inputs(case)
    mask = torch.from_numpy(g["mask"])
    out = net.forward_step(a.to(dev), t.to(dev), mask_tokens=mask.to(dev))
    y, y_ref = out["y_hat"].cpu(), torch.from_numpy(g["y_hat"])
    assert tuple(y.shape) == tuple(y_ref.shape)
    assert torch.equal(out["latent_mask"].cpu()[:, 0], mask)
    assert torch.equal(out["tgt"].cpu(), t[..., :y.shape[-1]])
    # latents handed to the decoder: unmasked tokens are the tactile encoder's, masked ones the predictor's
    tr = {}
    plc_oracle.forward_step(a, t, mask_tokens=mask, trace=tr)
    z = net.last_latents.cpu()
    m3 = mask[:, None, :].expand_as(z)
    assert float((z - tr["z_filled"])[m3].abs().max()) < Z_TOL[plan]          # predicted tokens (full-length attention)
    assert float((z - tr["z_filled"])[~m3].abs().max()) < Z_TOL[plan]         # pass-through tokens
    # the DAC codes of the audio frame feed the keys / values: a flip there must be a documented near-tie
    ok, n, worst, _ = stage_flips(net.last_audio_codes.cpu(), plc_oracle.A_QUANT(tr["za"])[1],
                                  plc_oracle.A_QUANT.last_margins, 2e-4 if plan == "tc" else 2e-5)
    assert ok, (n, worst)
    if n == 0:
        assert float((y - y_ref).abs().max()) < Y_TOL[plan]
        assert psnr(y, y_ref) > PSNR_MIN[plan]


def test_plc_masks_and_edges(dev, plc_oracle):
    net = pkg.build_plc()
    net.load_state_dict(plc_oracle.state_dict())
    a, t = cases.codec_inputs(dict(B=2, T=9600, kind="uniform"))
    a, t = a.to(dev), t.to(dev)
    none = torch.zeros(2, 30, dtype=torch.bool, device=dev)
    out0 = net.forward_step(a, t, mask_tokens=none)
    # nothing lost: the latents are the tactile encoder's own and the predictor's output is not used at all
    assert torch.equal(net.last_latents, net.T_ENC(t))
    allm = ~none
    out1 = net.forward_step(a, t, mask_tokens=allm)
    assert not torch.equal(out0["y_hat"], out1["y_hat"])
    # everything lost: the result cannot depend on the tactile frame
    out2 = net.forward_step(a, torch.zeros_like(t), mask_tokens=allm)
    assert torch.equal(out1["y_hat"], out2["y_hat"])
    # drawn masks: the reference's two policies
    torch.manual_seed(0)
    o = net.forward_step(a, t)
    assert o["latent_mask"].shape == (2, 1, 30) and o["latent_mask"].dtype == torch.bool
    o = net.forward_step(a, t, category="medium")
    assert bool(o["latent_mask"].any())
    with pytest.raises(ValueError):
        net.forward_step(a, t, mask_tokens=none[:, :-1])
    with pytest.raises(ValueError):
        net.forward_step(a, t, category="extreme")


@pytest.mark.parametrize("T", [17, 64, 75, 200, 1031])
def test_full_length_predictor_matches_oracle(T, dev):
    """CrossPredictor.forward beyond one chunk: key tiles of 64 with ragged tails, online softmax."""
    torch.manual_seed(7)
    ref = proposed.CrossPredictor(c=1024).eval()
    net = pkg.CrossPredictor(c=1024)
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(T)
    zp = 0.3 * torch.randn(2, 1024, T, generator=g)
    za = 2.0 * torch.randn(2, 1024, T, generator=g)
    with torch.no_grad():
        want = ref(zp, za)
    for plan, tol in (("f32", 5e-5), ("tc", 2e-4)):
        net.precision = plan
        got = net(zp.to(dev), za.to(dev)).cpu()
        assert float((got - want).abs().max()) < tol, (plan, T)


def test_ema_step_against_reference_golden(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "ema_step.npz"))
    n_books = int(g["n_books"])
    vq = pkg.ResidualVQEMA(dim=96, n_books=n_books, n_embed=128).to(dev)
    with torch.no_grad():
        for i, b in enumerate(vq.books):
            b.copy_(torch.from_numpy(g[f"book{i}_before"]))
    vq.decay = float(g["decay"])
    z = torch.from_numpy(g["z_tokens"]).to(dev)
    q_before = vq(z).clone()
    vq.ema_step(z)
    # oracle margins tell whether the GPU's nearest codes could differ anywhere (they select which rows move)
    books = [torch.from_numpy(g[f"book{i}_before"]).clone() for i in range(n_books)]
    info = proposed.ema_step(books, torch.from_numpy(g["z_tokens"]), float(g["decay"]))
    for i, b in enumerate(vq.books):
        after = torch.from_numpy(g[f"book{i}_after"])
        assert torch.equal(vq.last_ema_counts[i].cpu().long(), info[i]["counts"]) or float(info[i]["margin"].min()) < 1e-5
        if torch.equal(vq.last_ema_counts[i].cpu().long(), info[i]["counts"]):
            assert torch.equal(b.detach().cpu(), after), f"book {i}: the update must be bit-equal to the reference's"
        else:
            assert float((b.detach().cpu() - after).abs().max()) < 1e-3
    # the packed copy the CUDA programs read was refreshed: the quantiser now uses the moved codes
    q_after = vq(z)
    assert not torch.equal(q_before, q_after)
    ref_vq = proposed.ResidualVQEMA(96, n_books, 128)
    with torch.no_grad():
        for p, b in zip(ref_vq.books, vq.books):
            p.copy_(b.detach().cpu())
    want = ref_vq(z.cpu())
    ok, n, _, _ = stage_flips(vq.last_indices.cpu(), ref_vq.last_indices, ref_vq.last_margins, 1e-5)
    assert ok
    if n == 0:
        assert torch.equal(q_after.cpu(), want)


@pytest.mark.parametrize("plan", ["f32", "tc"])
def test_receiver_against_oracle_receiver(plan, dev, oracle_models):
    """decode_indices (audio frame + code indices -> reconstruction) against the oracle's receiver loop fed the SAME
    indices: the tolerance is the decoder's; the latents are compared too."""
    name = "cal_b4k256_use3_short"
    case = cases.CODEC_CASES[name]
    ref = oracle_models(name)
    net = pkg.build_proposed(case["books"], case["K"])
    net.load_state_dict(ref.state_dict())
    for m in (net, net.A_ENC, net.T_ENC, net.T_DEC, net.A_QUANT, net.predict, net.vq):
        m.precision = plan
    a, t = cases.codec_inputs(case)
    tr = {}
    ref.forward_eval(a, t, case["books_use"], trace=tr)
    idx = tr["idx"]                                            # the oracle sender's indices
    y_ref, z_ref = ref.decode_from_indices(a, idx, case["books_use"])
    y = net.decode_indices(a.to(dev), idx.to(dev), case["books_use"]).cpu()
    ok, n, worst, _ = stage_flips(net.last_audio_codes.cpu(), tr["a_codes"], tr["a_margin"], 2e-4 if plan == "tc" else 2e-5)
    assert ok, (n, worst)
    if n == 0:
        assert float((y - y_ref).abs().max()) < Y_TOL[plan]
        assert psnr(y, y_ref) > PSNR_MIN[plan]
    # and the receiver's reconstruction is the sender's up to the rounding of q_sum + (q - r) + r  vs  q_sum + q
    assert psnr(y_ref, tr["y"]) > 80.0


@pytest.mark.parametrize("D,K,books", [(96, 512, 8), (96, 128, 10), (96, 256, 4), (64, 64, 3), (128, 512, 2), (32, 192, 5)])
def test_tcgen05_residual_vq_equals_fp32_kernels(D, K, books, dev):
    """rvq_tc_kernel (all books in one launch, scores in TMEM, arg-max + gather + residual update in the epilogue)
    against the FP32 CUDA-core kernels: same indices and the same q_sum bits, including rows whose top codes are
    closer than the tensor-core error bound (duplicated / nearly duplicated codewords force the exact re-score)."""
    g = torch.Generator().manual_seed(D * 1000 + K + books)
    vq = pkg.ResidualVQEMA(dim=D, n_books=books, n_embed=K).to(dev)
    with torch.no_grad():
        for i, b in enumerate(vq.books):
            e = torch.randn(K, D, generator=g) / D ** 0.5 * (0.7 ** i)
            e[K // 2:K // 2 + 8] = e[:8]                                   # exact duplicates: the FIRST must win
            e[K // 2 + 8:K // 2 + 16] = e[8:16] * (1 + 1e-6)               # near-duplicates: inside the error bound
            b.copy_(e.to(dev))
    for B, T in ((1, 1), (1, 75), (3, 100), (64, 75)):
        z = (torch.randn(B, D, T, generator=g) / D ** 0.5).to(dev)
        z[:, :, 0] = vq.books[0][3].detach()[None, :]                      # a token sitting exactly on a codeword
        out = {}
        for plan in ("f32", "tc"):
            vq.precision = plan
            q, idx = vq(z, return_indices=True)
            out[plan] = (q.clone(), idx.clone())
        assert torch.equal(out["f32"][1], out["tc"][1]), (D, K, B, T, int((out["f32"][1] != out["tc"][1]).sum()))
        assert torch.equal(out["f32"][0], out["tc"][0])
        assert int(out["tc"][1].min()) >= 0 and int(out["tc"][1].max()) < K
        # prefix property with fewer books
        vq.precision = "tc"
        q2, idx2 = vq(z, n_books_use=1, return_indices=True)
        assert torch.equal(idx2[:, 0], out["tc"][1][:, 0])
    # the launch count is what the B = 1 latency path cares about: one launch for all books
    eng, wid = vq._engine(dev)
    progs = [p for k, p in eng.programs._d.items() if k[0] == "vq" and k[-1] != 0]
    assert progs and all(p.info["launches"] <= 4 for p in progs), [p.info["launches"] for p in progs]


def test_small_batch_token_kernels_equal_large_batch_kernels(dev):
    """The one-CTA-per-token kernels of the batch-1 streaming path (rvq_token_f32, dac_rvq_token_f32: up to 296 tokens)
    give the bits of the kernels larger batches select, on the same rows: residual VQ (with duplicated codewords:
    the first must win) and the 32-stage DAC quantizer."""
    g = torch.Generator().manual_seed(5)
    D, K, books = 96, 512, 8
    vq = pkg.ResidualVQEMA(dim=D, n_books=books, n_embed=K).to(dev)
    with torch.no_grad():
        for i, b in enumerate(vq.books):
            e = torch.randn(K, D, generator=g) / D ** 0.5 * (0.7 ** i)
            e[K // 2:K // 2 + 8] = e[:8]
            b.copy_(e.to(dev))
    z = (torch.randn(64, D, 75, generator=g) / D ** 0.5).to(dev)
    z[:, :, 0] = vq.books[0][3].detach()[None, :]
    for plan in ("f32", "tc"):
        vq.precision = plan
        q_big, i_big = vq(z, return_indices=True)
        for nb in (1, 3):
            q_s, i_s = vq(z[:nb], return_indices=True)
            assert torch.equal(i_s, i_big[:nb]) and torch.equal(q_s, q_big[:nb]), (plan, nb)
        assert int(i_big[:, 0, 0].max()) == 3 and int(i_big[:, 0, 0].min()) == 3      # the first of the duplicates
    torch.manual_seed(7)
    aq = pkg.ResidualVectorQuantize().to(dev)
    za = torch.randn(8, 1024, 75, generator=g).to(dev) * 2.0
    zq_big, codes_big, *_ = aq(za)
    for nb in (1, 3):
        zq_s, codes_s, *_ = aq(za[:nb])
        assert torch.equal(codes_s, codes_big[:nb]) and torch.equal(zq_s, zq_big[:nb]), nb
    assert len(torch.unique(codes_big)) > 32

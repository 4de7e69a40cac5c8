"""GPU parity of the evaluation metrics through the C-ABI (libb2c.so, b2c_metric_*) against the CPU
oracle (oracle/metrics.py, pinned bit-equal to the reference's own functions) and the committed goldens the
reference's functions wrote.  Floating-point results: tolerances below; the alignment shift is an integer and must
be equal (the synthetic pairs have one clear correlation peak)."""
import os

import numpy as np
import pytest
import torch

from multimodal_vqvae_compression_audio_tactile_b200 import metrics as pm
from oracle import metrics as om

pytestmark = pytest.mark.gpu

PSNR_TOL_DB = 2e-3      # fp32 sums in a different order; 37 dB values
STSIM_TOL = 2e-5
RESAMPLE_TOL = 2e-6     # |x| <= 0.5, 106 taps


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("kind,B", [("shifted", 8), ("plain", 3)])
def test_metrics_against_golden_and_oracle(dev, golden_dir, kind, B):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    ref, est, lags = om.metric_inputs(B=B, kind=kind)
    r, e = ref.to(dev), est.to(dev)
    np.testing.assert_allclose(pm.psnr_batch(r, e), g[f"{kind}_psnr"], atol=PSNR_TOL_DB, rtol=0)
    np.testing.assert_allclose(pm.stsim_batch(r, e), g[f"{kind}_stsim"], atol=STSIM_TOL, rtol=0)
    val, shift = pm.psnr_3k_aligned_tensor(r, e)
    assert shift.tolist() == lags == list(g[f"{kind}_shift"])
    np.testing.assert_allclose(val.cpu().numpy(), g[f"{kind}_psnr3k"], atol=PSNR_TOL_DB, rtol=0)
    assert pm.psnr_3k_aligned_batch(r, e) == [float(v) for v in val.tolist()]
    y = pm.resample_f32(r, 24000, 3000)
    assert y.shape == (B, 1, 2999)
    np.testing.assert_allclose(y.cpu().numpy(), g[f"{kind}_ref3k"], atol=RESAMPLE_TOL, rtol=0)
    y16 = pm.resample_f32(r[:1], 24000, 16000)
    np.testing.assert_allclose(y16.cpu().numpy()[..., :512], g[f"{kind}_ref16k_head"], atol=RESAMPLE_TOL, rtol=0)


def test_alignment_correlations_and_pair_interface(dev):
    """All 401 correlations against the oracle's loop, the reference's single-pair interface, edge shifts +-200."""
    ref, est, lags = om.metric_inputs(B=5, T=6000, seed=9)
    r, e = ref.to(dev), est.to(dev)
    best, corr = pm.align_batch_24k(r, e)
    for b in range(5):
        c_ref = om.xcorr_all_shifts(ref[b, 0], est[b, 0])
        np.testing.assert_allclose(corr[b].cpu().numpy(), c_ref.numpy(), atol=2e-4 * float(c_ref.abs().max()), rtol=0)
        ra, ea, s = pm.align_pair_24k(r[b:b + 1], e[b:b + 1])
        ro, eo, so = om.align_pair_24k(ref[b:b + 1], est[b:b + 1])
        assert s == so == lags[b] == int(best[b])
        assert torch.equal(ra.cpu(), ro) and torch.equal(ea.cpu(), eo)


@pytest.mark.parametrize("T", [257, 1000, 4800, 23992, 24000])
def test_ragged_lengths_and_ratios(dev, T):
    """Frame counts, reflect padding and the resampler's zero padding at lengths that are not multiples of anything."""
    ref, est, _ = om.metric_inputs(B=2, T=T, seed=T, kind="plain")
    r, e = ref.to(dev), est.to(dev)
    np.testing.assert_allclose(pm.stsim_batch(r, e), om.stsim_batch(ref, est), atol=STSIM_TOL, rtol=0)
    np.testing.assert_allclose(pm.psnr_batch(r, e), om.psnr_batch(ref, est), atol=PSNR_TOL_DB, rtol=0)
    np.testing.assert_allclose(pm.psnr_3k_aligned_batch(r, e), om.psnr_3k_aligned_batch(ref, est), atol=PSNR_TOL_DB, rtol=0)
    for sr in (3000, 16000, 44100, 48000):
        y, yo = pm.resample_f32(r, 24000, sr), om.resample_f32(ref, 24000, sr)
        assert y.shape == yo.shape
        np.testing.assert_allclose(y.cpu().numpy(), yo.numpy(), atol=RESAMPLE_TOL, rtol=0)
    up = pm.resample_f32(pm.resample_f32(r, 24000, 3000), 3000, 24000)      # the DAC baselines' round trip (:367-371)
    uo = om.resample_f32(om.resample_f32(ref, 24000, 3000), 3000, 24000)
    np.testing.assert_allclose(up.cpu().numpy(), uo.numpy(), atol=2 * RESAMPLE_TOL, rtol=0)


def test_degenerate_inputs(dev):
    """Identical signals (PSNR clamps at eps = 1e-12 -> 120 dB, ST-SIM = 1), silence (|X| clamps at 1e-8), errors."""
    ref, _, _ = om.metric_inputs(B=2, T=4800, kind="plain")
    r = ref.to(dev)
    assert pm.psnr_batch(r, r) == om.psnr_batch(ref, ref) == [120.0, 120.0]
    np.testing.assert_allclose(pm.stsim_batch(r, r), [1.0, 1.0], atol=1e-6)
    z = torch.zeros(1, 1, 4800)
    np.testing.assert_allclose(pm.stsim_batch(z.to(dev), r[:1]), om.stsim_batch(z, ref[:1]), atol=STSIM_TOL)
    np.testing.assert_allclose(pm.stsim_batch(z.to(dev), z.to(dev)), om.stsim_batch(z, z), atol=STSIM_TOL)
    assert pm.psnr_3k_aligned_batch(z.to(dev), z.to(dev)) == om.psnr_3k_aligned_batch(z, z)
    with pytest.raises(Exception):
        pm.psnr_batch(ref, ref)                      # CPU tensors: no fallback
    with pytest.raises(ValueError):
        pm.stsim_batch(r, r[..., :100])

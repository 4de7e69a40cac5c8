"""ORACLE (test infrastructure): CPU restatement of the reference's evaluation metrics,
Evaluation/compare_dacvsproposal_5_eval.py -- resample_f32 (:91-97), _mel_mag (:142-163), stsim_batch (:166-177),
psnr_batch (:180-185), align_pair_24k (:188-211), psnr_3k_aligned_batch (:213-223).

The reference calls torchaudio for two constant tables (Resample's sinc_interp_hann filter bank, MelScale's HTK
triangular filters); both are restated here from torchaudio 2.x `functional._get_sinc_resample_kernel` /
`melscale_fbanks` so the oracle runs without torchaudio.  Pinned: tests/test_oracle_cpu.py compares every function
below with the reference's own functions (ast-extracted, executed with torchaudio in the build container) and
tests/golden/metrics.npz holds their outputs.  Only tests/, smoke() and bench.py's cpu legs may import this.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EVAL_SR = 24000                  # :52
ORIG_3K = 3000                   # :53
ALIGN_MAX_SHIFT_SAMPLES = 200    # :69


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio.functional._get_sinc_resample_kernel (sinc_interp_hann, dtype=None: float64 grid, float32 result).
    -> (kernel [new, 1, 2*width + orig] float32, width, orig, new) with the rates divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=None)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.to(torch.float32), width, orig, new


def resample_f32(x, sr_in, sr_out):
    if sr_in == sr_out:
        return x
    x = x.to(torch.float32).contiguous()
    kernel, width, orig, new = sinc_resample_kernel(sr_in, sr_out)
    shape = x.size()
    w = x.view(-1, shape[-1])
    n, length = w.shape
    w = F.pad(w, (width, width + orig))
    r = F.conv1d(w[:, None], kernel, stride=orig)
    r = r.transpose(1, 2).reshape(n, -1)
    target = int(math.ceil(new * length / orig))
    r = r[..., :target]
    return r.view(shape[:-1] + r.shape[-1:])


def mel_filterbank(n_freqs=257, f_min=0.0, f_max=12000.0, n_mels=64, sample_rate=EVAL_SR):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> [n_freqs, n_mels] float32."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(zero, torch.min(down, up))


def _mel_mag(x_1T, sr=EVAL_SR, n_fft=512, hop=128, n_mels=64):
    x = x_1T[:, 0, :] if x_1T.dim() == 3 else x_1T
    x = x.to(torch.float32)
    window = torch.hann_window(n_fft, dtype=torch.float32)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=window, center=True, return_complex=True)
    mag = spec.abs().clamp_min_(1e-8)
    fb = mel_filterbank(n_fft // 2 + 1, 0.0, sr * 0.5, n_mels, sr)
    M = torch.matmul(mag.transpose(-1, -2), fb).transpose(-1, -2)      # MelScale.forward
    return M / M.amax(dim=(1, 2), keepdim=True).clamp_min_(1e-8)


@torch.no_grad()
def stsim_batch(ref_1T, est_1T):
    Mref, Mest = _mel_mag(ref_1T), _mel_mag(est_1T)
    Tf = max(Mref.shape[-1], Mest.shape[-1])
    if Mref.shape[-1] != Tf:
        Mref = F.interpolate(Mref, size=Tf, mode="linear", align_corners=False)
    if Mest.shape[-1] != Tf:
        Mest = F.interpolate(Mest, size=Tf, mode="linear", align_corners=False)
    num = (Mref * Mest).sum(dim=1)
    den = (Mref.norm(dim=1) * Mest.norm(dim=1)).clamp_min(1e-8)
    cos_t = (num / den).clamp(-1, 1)
    val = 0.5 * (cos_t.mean(dim=-1) + 1.0)
    return [float(v.item()) for v in val]


@torch.no_grad()
def psnr_batch(ref_1T, est_1T, eps=1e-12):
    ref, est = ref_1T.to(torch.float32), est_1T.to(torch.float32)
    mse = (ref - est).pow(2).mean(dim=(1, 2)).clamp_min(eps)
    return [float(v.item()) for v in 10.0 * torch.log10(1.0 / mse)]


def xcorr_all_shifts(r, e, max_shift=ALIGN_MAX_SHIFT_SAMPLES):
    """The correlations align_pair_24k's loop compares, in loop order (s = -max_shift .. max_shift)."""
    out = []
    for s in range(-max_shift, max_shift + 1):
        if s < 0:
            rs = r[-s:]; es = e[: rs.numel()]
        elif s > 0:
            rs = r[:-s]; es = e[s: s + rs.numel()]
        else:
            rs = r; es = e[: rs.numel()]
        out.append(torch.sum(rs * es))
    return torch.stack(out)


def align_pair_24k(ref_24, est_24, max_shift=ALIGN_MAX_SHIFT_SAMPLES):
    r = ref_24.squeeze(0).squeeze(0)
    e = est_24.squeeze(0).squeeze(0)
    c = xcorr_all_shifts(r.to(torch.float32), e.to(torch.float32), max_shift)
    best_shift, best_corr = 0, -1e18
    for i, s in enumerate(range(-max_shift, max_shift + 1)):
        if c[i] > best_corr:
            best_corr, best_shift = c[i], s
    s = best_shift
    if s < 0:
        r_a = r[-s:]; e_a = e[: r_a.numel()]
    elif s > 0:
        r_a = r[:-s]; e_a = e[s: s + r_a.numel()]
    else:
        r_a = r; e_a = e[: r.numel()]
    return r_a.unsqueeze(0).unsqueeze(0), e_a.unsqueeze(0).unsqueeze(0), best_shift


@torch.no_grad()
def psnr_3k_aligned_batch(ref_24, est_24):
    vals = []
    for b in range(ref_24.size(0)):
        r_al, e_al, _ = align_pair_24k(ref_24[b:b + 1], est_24[b:b + 1])
        vals += psnr_batch(resample_f32(r_al, EVAL_SR, ORIG_3K), resample_f32(e_al, EVAL_SR, ORIG_3K))
    return vals


def metric_inputs(B=4, T=23992, seed=321, kind="shifted"):
    """Synthetic (reference, estimate) pairs with a known lag: band-limited noise + a delayed, scaled, noisy copy."""
    g = torch.Generator().manual_seed(seed)
    n = T + 512
    base = torch.randn(B, n, generator=g)
    k = torch.hann_window(33, periodic=False)
    k = (k / k.sum())[None, None]
    smooth = F.conv1d(base[:, None], k, padding=16)[:, 0]
    smooth = 0.5 * smooth / smooth.abs().amax(dim=1, keepdim=True)
    lags = [0, 37, -121, 200, -200, 5, 163, -64][:B] if kind == "shifted" else [0] * B
    ref = smooth[:, 256:256 + T].clone()
    est = torch.stack([smooth[b, 256 - lags[b]:256 - lags[b] + T] for b in range(B)])
    est = 0.9 * est + 0.02 * torch.randn(B, T, generator=g)
    return ref[:, None].contiguous(), est[:, None].contiguous(), lags

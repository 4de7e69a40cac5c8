"""Evaluation metrics of the reference's Evaluation scripts on the GPU, same names and
return types as Evaluation/compare_dacvsproposal_5_eval.py:

    resample_f32 (:91-97)   stsim_batch (:166-177)   psnr_batch (:180-185)
    align_pair_24k (:188-211)   psnr_3k_aligned_batch (:213-223)

The reference aligns every frame with a 401-iteration Python loop of torch.sum calls and resamples / scores frame by
frame; here a batch is one or two kernel launches of libb2c.so (csrc/kernels_metrics.cuh).  CUDA tensors only --
there is no CPU path.  The two constant tables torchaudio builds for the reference (the sinc_interp_hann polyphase
filter bank of Resample and the HTK triangular filters of MelScale) are computed once per process on the host, the
way weights are packed, and cached per device.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L

EVAL_SR = 24000                  # :52
ORIG_3K = 3000                   # :53
ALIGN_MAX_SHIFT_SAMPLES = 200    # :69

_tables = {}


def _dev_index(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _rows(x: torch.Tensor, what: str) -> torch.Tensor:
    """[B, 1, T] / [B, T] / [T] -> contiguous fp32 [B, T] on the GPU."""
    if not x.is_cuda:
        raise L.B2CError(f"{what}: CUDA tensors only (no CPU fallback)")
    if x.dim() == 3:
        if x.shape[1] != 1:
            raise ValueError(f"{what}: expected [B, 1, T], got {tuple(x.shape)}")
        x = x[:, 0, :]
    elif x.dim() == 1:
        x = x[None]
    elif x.dim() != 2:
        raise ValueError(f"{what}: expected [B, 1, T] or [B, T], got {tuple(x.shape)}")
    return x.detach().to(torch.float32).contiguous()


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The filter bank torchaudio.transforms.Resample(orig_freq, new_freq) builds (sinc_interp_hann; float64 grid,
    float32 taps): -> (kernel [new, 2*width + orig] fp32 CPU tensor, width, orig, new), rates divided by their gcd."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("resampling rates must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None] / orig
    t = torch.arange(0, -new, -1)[:, None] / new + idx     # torchaudio divides the int64 phases in float32
    t = (t * base).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    kern = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * (base / orig)
    return kern.to(torch.float32).contiguous(), width, orig, new


def mel_filterbank(n_freqs: int = 257, f_min: float = 0.0, f_max: float = 12000.0, n_mels: int = 64,
                   sample_rate: int = EVAL_SR) -> torch.Tensor:
    """The [n_freqs, n_mels] matrix of torchaudio.transforms.MelScale(norm=None, mel_scale='htk') (:155-158)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


def _resample_table(device, sr_in, sr_out):
    key = ("rs", str(device), int(sr_in), int(sr_out))
    tab = _tables.get(key)
    if tab is None:
        kern, width, orig, new = sinc_resample_kernel(sr_in, sr_out)
        tab = _tables[key] = (kern.to(device), width, orig, new)
    return tab


def _mel_table(device, n_mels=64, sr=EVAL_SR, n_fft=512):
    key = ("mel", str(device), n_mels, sr, n_fft)
    tab = _tables.get(key)
    if tab is None:
        fb = mel_filterbank(n_fft // 2 + 1, 0.0, sr * 0.5, n_mels, sr)
        nz = fb != 0                                   # bins [lo, hi) of every band's triangle
        lo = torch.where(nz.any(0), nz.int().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        hi = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).int().argmax(0), torch.zeros(n_mels, dtype=torch.long))
        rng = torch.stack([lo, hi], dim=1).to(torch.int32).contiguous()
        tab = _tables[key] = (fb.to(device), rng.to(device))
    return tab


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


@torch.no_grad()
def resample_f32(x: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
    """:91-97.  x [..., T] -> [..., ceil(T * sr_out / sr_in)] fp32."""
    if sr_in == sr_out:
        return x
    if not x.is_cuda:
        raise L.B2CError("resample_f32: CUDA tensors only (no CPU fallback)")
    lib = L.load()
    shape = x.shape
    w = x.detach().to(torch.float32).reshape(-1, shape[-1]).contiguous()
    kern, width, orig, new = _resample_table(x.device, sr_in, sr_out)
    n, length = w.shape
    lout = int(math.ceil(new * length / orig))
    y = torch.empty(n, lout, device=x.device, dtype=torch.float32)
    if n and lout:
        L.check(lib.b2c_metric_resample(_dev_index(x), _stream(x), _p(w), _p(y), _p(kern), n, length, lout, orig, new,
                                        width), "b2c_metric_resample")
    return y.view(shape[:-1] + (lout,))


@torch.no_grad()
def psnr_tensor(ref_1T, est_1T, eps: float = 1e-12) -> torch.Tensor:
    """psnr_batch as a [B] device tensor (no host synchronisation)."""
    r, e = _rows(ref_1T, "psnr_batch"), _rows(est_1T, "psnr_batch")
    if r.shape != e.shape:
        raise ValueError(f"psnr_batch: shapes differ: {tuple(r.shape)} vs {tuple(e.shape)}")
    out = torch.empty(r.shape[0], device=r.device, dtype=torch.float32)
    if r.shape[0]:
        L.check(L.load().b2c_metric_psnr(_dev_index(r), _stream(r), _p(r), _p(e), _p(out), r.shape[0], r.shape[1], eps),
                "b2c_metric_psnr")
    return out


def psnr_batch(ref_1T, est_1T, eps: float = 1e-12):
    """:180-185.  PSNR (dB), peak 1.0 -> list of B floats."""
    return [float(v) for v in psnr_tensor(ref_1T, est_1T, eps).tolist()]


@torch.no_grad()
def align_batch_24k(ref_24, est_24, max_shift: int = ALIGN_MAX_SHIFT_SAMPLES):
    """The search of align_pair_24k (:193-203) for every frame of a batch in one launch:
    -> (best_shift int32 [B], corr fp32 [B, 2*max_shift + 1]) device tensors."""
    r, e = _rows(ref_24, "align_batch_24k"), _rows(est_24, "align_batch_24k")
    if r.shape != e.shape:
        raise ValueError(f"align: shapes differ: {tuple(r.shape)} vs {tuple(e.shape)}")
    B, T = r.shape
    corr = torch.empty(B, 2 * max_shift + 1, device=r.device, dtype=torch.float32)
    best = torch.empty(B, device=r.device, dtype=torch.int32)
    if B:
        L.check(L.load().b2c_metric_xcorr_align(_dev_index(r), _stream(r), _p(r), _p(e), B, T, max_shift, _p(corr),
                                                _p(best)), "b2c_metric_xcorr_align")
    return best, corr


def align_pair_24k(ref_24, est_24, max_shift: int = ALIGN_MAX_SHIFT_SAMPLES):
    """:188-211.  ref_24, est_24 [1, 1, T] -> (r_aligned [1, 1, n], e_aligned [1, 1, n], best_shift)."""
    best, _ = align_batch_24k(ref_24, est_24, max_shift)
    s = int(best[0].item())
    r = ref_24.squeeze(0).squeeze(0)
    e = est_24.squeeze(0).squeeze(0)
    if s < 0:
        r_a = r[-s:]; e_a = e[: r_a.numel()]
    elif s > 0:
        r_a = r[:-s]; e_a = e[s: s + r_a.numel()]
    else:
        r_a = r; e_a = e[: r.numel()]
    return r_a.unsqueeze(0).unsqueeze(0), e_a.unsqueeze(0).unsqueeze(0), s


@torch.no_grad()
def psnr_3k_aligned_tensor(ref_24, est_24, max_shift: int = ALIGN_MAX_SHIFT_SAMPLES, sr_in: int = EVAL_SR,
                           sr_out: int = ORIG_3K, eps: float = 1e-12):
    """psnr_3k_aligned_batch as device tensors: (psnr fp32 [B], best_shift int32 [B]); two launches for the search,
    one for align + resample + PSNR."""
    r, e = _rows(ref_24, "psnr_3k_aligned_batch"), _rows(est_24, "psnr_3k_aligned_batch")
    best, _ = align_batch_24k(r, e, max_shift)
    B, T = r.shape
    out = torch.empty(B, device=r.device, dtype=torch.float32)
    if B:
        kern, width, orig, new = _resample_table(r.device, sr_in, sr_out)
        L.check(L.load().b2c_metric_psnr_resampled(_dev_index(r), _stream(r), _p(r), _p(e), _p(best), _p(out), _p(kern),
                                                   B, T, orig, new, width, eps), "b2c_metric_psnr_resampled")
    return out, best


def psnr_3k_aligned_batch(ref_24, est_24):
    """:213-223 -> list of B floats."""
    return [float(v) for v in psnr_3k_aligned_tensor(ref_24, est_24)[0].tolist()]


@torch.no_grad()
def stsim_tensor(ref_1T, est_1T, n_mels: int = 64) -> torch.Tensor:
    """stsim_batch as a [B] device tensor."""
    r, e = _rows(ref_1T, "stsim_batch"), _rows(est_1T, "stsim_batch")
    if r.shape != e.shape:
        raise ValueError("stsim_batch: this implementation scores equal-length pairs (the evaluation crops both to "
                         f"min(len) first, :443); got {tuple(r.shape)} vs {tuple(e.shape)}")
    B, T = r.shape
    lib = L.load()
    out = torch.empty(B, device=r.device, dtype=torch.float32)
    if B:
        fb, rng = _mel_table(r.device, n_mels)
        scratch = torch.empty(int(lib.b2c_metric_stsim_scratch_bytes(B, T, n_mels)) // 4, device=r.device,
                              dtype=torch.float32)
        L.check(lib.b2c_metric_stsim(_dev_index(r), _stream(r), _p(r), _p(e), _p(fb), _p(rng), _p(scratch), _p(out), B, T,
                                     n_mels),
                "b2c_metric_stsim")
    return out


def stsim_batch(ref_1T, est_1T):
    """:166-177 -> list of B floats."""
    return [float(v) for v in stsim_tensor(ref_1T, est_1T).tolist()]

#!/usr/bin/env python
"""Evaluation metrics on the GPU against the reference's per-frame loops on this box's host cores.
    python tools/metrics_bench.py [--batch 64] [--out gpurun_out/metrics_bench.json]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodal_vqvae_compression_audio_tactile_b200 import metrics as pm
from oracle import metrics as om

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--cpu-frames", type=int, default=4)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
T = 23992
g = torch.Generator().manual_seed(0)
ref = (torch.rand(a.batch, 1, T, generator=g) - 0.5)
est = torch.roll(ref, 17, dims=-1) * 0.9 + 0.01 * torch.randn(a.batch, 1, T, generator=g)
r, e = ref.to(dev), est.to(dev)


def gpu_ms(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_ms(fn):
    fn()
    t = time.perf_counter()
    fn()
    return (time.perf_counter() - t) * 1e3 / a.cpu_frames * a.batch     # scaled to the GPU batch (per-frame loops)


rc, ec = ref[: a.cpu_frames], est[: a.cpu_frames]
res = dict(batch=a.batch, samples=T, cpu_threads=torch.get_num_threads(), cpu_frames_timed=a.cpu_frames, unit="ms per batch")
for name, gfn, cfn in (
        ("psnr_3k_aligned_batch", lambda: pm.psnr_3k_aligned_tensor(r, e), lambda: om.psnr_3k_aligned_batch(rc, ec)),
        ("stsim_batch", lambda: pm.stsim_tensor(r, e), lambda: om.stsim_batch(rc, ec)),
        ("psnr_batch", lambda: pm.psnr_tensor(r, e), lambda: om.psnr_batch(rc, ec)),
        ("resample_24k_3k", lambda: pm.resample_f32(r, 24000, 3000), lambda: om.resample_f32(rc, 24000, 3000))):
    gm, cm = gpu_ms(gfn), cpu_ms(cfn)
    res[name] = dict(gpu_ms=gm, cpu_oracle_ms_scaled=cm, ratio=cm / gm)
    print(name, res[name], flush=True)
# algorithmic bytes of the alignment search: both signals once; the kernel re-reads est from L1/L2 per shift block
res["xcorr_bytes_algorithmic"] = 2 * a.batch * T * 4
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)

#!/usr/bin/env python
"""On-GPU self-test of the tcgen05 implicit-GEMM conv kernel: every distinct conv / convT / linear layer of
the codec, tensor-core path (bf16x3 and bf16) against the FP32 CUDA-core kernel on the same random input.
Prints one line per layer: max |diff| (raw output and activated output), device time of both kernels and
the achieved TFLOP/s.  A developer aid (not part of the test suite or the product path).

    python tools/tc_selftest.py [--group enc|dec|pred|all] [--batch 2] [--T 24000] [--only SUBSTR]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import multimodal_vqvae_compression_audio_tactile_b200 as pkg  # noqa: E402
from multimodal_vqvae_compression_audio_tactile_b200 import _lib as L  # noqa: E402
from multimodal_vqvae_compression_audio_tactile_b200.engine import Emitter, Engine  # noqa: E402


def layers_of(net, T):
    """-> list of (name, module or (weight, bias), Lin, has_alpha)"""
    out = []
    Lx = T
    for name, m in net.T_ENC.named_modules():
        if isinstance(m, pkg.modules.WNConv1d):
            if m.in_channels > 1:
                out.append(("enc." + name, m, Lx))
            Lx = (Lx + 2 * m.padding - m.dilation * (m.kernel_size - 1) - 1) // m.stride + 1
    for name, m in net.T_DEC.named_modules():
        if isinstance(m, pkg.modules.WNConvTranspose1d):
            out.append(("dec." + name, m, Lx))
            Lx = (Lx - 1) * m.stride - 2 * m.padding + m.kernel_size
        elif isinstance(m, pkg.modules.WNConv1d):
            if m.out_channels > 1:
                out.append(("dec." + name, m, Lx))
            Lx = (Lx + 2 * m.padding - m.dilation * (m.kernel_size - 1) - 1) // m.stride + 1
    return out


def run_case(eng, name, w, B, Lin, Lout, dev, precs, reps=3):
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, Lin, w.cin, generator=g) * 2 - 1).to(dev)
    res = (torch.rand(B, Lout, w.cout, generator=g) * 2 - 1).to(dev)
    alpha = eng.pack_vec(torch.rand(w.cout, generator=g) + 0.5)
    n_in, n_out = B * Lin * w.cin, B * Lout * w.cout
    flops = 2.0 * B * Lout * w.cout * w.cin * (w.k if not w.transposed else 2)
    results = {}
    for prec in ["f32"] + precs:
        pr = L.PRECISIONS[prec]
        em = Emitter(eng)
        if pr != L.PREC_F32 and not em.tc_ok(w, Lin, pr):
            results[prec] = None
            continue
        f = L.FMT_OF_PREC[pr]
        act = em.new(n_out)
        kw = dict(out_raw=em.ext(3), out_act=act, alpha=alpha, prec=pr, x_fmt=L.FMT_F32, act_fmt=f)
        if w.transposed:
            em.convT(w, em.ext(1), B, Lin, **kw)
        else:
            em.conv(w, em.ext(1), B, Lin, res=em.ext(2), **kw)
        if f == L.FMT_F32:
            em.transpose(act, em.ext(4), 1, 1, n_out)  # plain copy
        else:
            em.convert(act, f, em.ext(4), L.FMT_F32, n_out)
        prog = em.finish(4)
        raw = torch.empty(B, Lout, w.cout, device=dev)
        a = torch.empty(B, Lout, w.cout, device=dev)
        ext = [x.data_ptr(), res.data_ptr(), raw.data_ptr(), a.data_ptr()]
        eng.run(prog, ext)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            pf = eng.profile(prog, ext)
            ms = sum(r["ms"] for r in pf if r["kind"].startswith("conv"))
            best = min(best, ms)
        results[prec] = (raw, a, best)
        eng.lib.b2c_prog_destroy(prog.handle)
    r0, a0, t0 = results["f32"]
    line = f"{name:34s} B={B} Lin={Lin:6d} {w.cin:4d}->{w.cout:4d} k={w.k:2d} s={w.stride} d={w.dilation} | f32 {t0:8.3f} ms {flops / t0 / 1e9:7.1f} TF/s"
    ok = True
    for prec in precs:
        if results[prec] is None:
            line += f" | {prec}: not eligible"
            continue
        r, a, t = results[prec]
        e_raw = float((r - r0).abs().max())
        e_act = float((a - a0).abs().max())
        scale = float(r0.abs().max())
        tol = (1e-4 if prec == "bf16x3" else 2e-2) * max(scale, 1.0)
        bad = not (e_raw < tol) or not torch.isfinite(r).all()
        ok = ok and not bad
        line += f" | {prec}: {t:7.3f} ms {flops / t / 1e9:7.1f} TF/s err raw {e_raw:.2e} act {e_act:.2e}{' FAIL' if bad else ''}"
    print(line + f" | |out|max {float(r0.abs().max()):.2f}", flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="all")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--T", type=int, default=24000)
    ap.add_argument("--only", default="")
    ap.add_argument("--precs", default="bf16x3,bf16")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    net = pkg.build_proposed(2, 128)
    eng = Engine(dev)
    precs = args.precs.split(",")
    seen = set()
    all_ok = True
    if args.group in ("enc", "dec", "all"):
        for name, m, Lin in layers_of(net, args.T):
            if args.group != "all" and not name.startswith(args.group):
                continue
            if args.only and args.only not in name:
                continue
            key = (m.in_channels, m.out_channels, m.kernel_size, m.stride, m.dilation, m.transposed, Lin)
            if key in seen:
                continue
            seen.add(key)
            w = eng.pack_wnconv(m)
            if m.transposed:
                Lout = (Lin - 1) * m.stride - 2 * m.padding + m.kernel_size
            else:
                Lout = (Lin + 2 * m.padding - m.dilation * (m.kernel_size - 1) - 1) // m.stride + 1
            all_ok &= run_case(eng, name, w, args.batch, Lin, Lout, dev, precs)
    if args.group in ("pred", "all"):
        pr = net.predict
        N = args.batch * 75
        for name, wt, bs in (("pred.q_proj", pr.q_proj.weight, None), ("pred.ffn1", pr.ffn[1].weight, pr.ffn[1].bias),
                             ("pred.ffn2", pr.ffn[3].weight, pr.ffn[3].bias),
                             ("proj_down", net.proj_down.weight, net.proj_down.bias),
                             ("proj_up", net.proj_up.weight, net.proj_up.bias)):
            if args.only and args.only not in name:
                continue
            w = eng.pack_plain(wt, bs)
            all_ok &= run_case(eng, name, w, 1, N, N, dev, precs)
    print("SELFTEST", "OK" if all_ok else "FAILED", flush=True)
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()

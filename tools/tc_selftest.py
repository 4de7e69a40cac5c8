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


def run_ru_case(eng, name, ru_mod, B, Lx, dev, precs, reps=3):
    """Fused ResidualUnit launch (b2c_prog_ru) against the two FP32 conv launches."""
    from multimodal_vqvae_compression_audio_tactile_b200.engine import _pack_ru
    ru = _pack_ru(eng, ru_mod)
    C_ = ru.c7.cout
    g = torch.Generator().manual_seed(2)
    x_raw = (torch.rand(B, Lx, C_, generator=g) * 2 - 1).to(dev)
    a_next = eng.pack_vec(torch.rand(C_, generator=g) + 0.5)
    n = B * Lx * C_
    flops = 2.0 * n * C_ * 8
    out = {}
    for prec in ["f32"] + precs:
        pr = L.PRECISIONS[prec]
        f = L.FMT_OF_PREC[pr]
        em = Emitter(eng)
        x_act = em.new(n)
        # x_act = snake1(x_raw): produce it with the stem-free path: a k=1 identity is not available, so use the
        # activation of an FP32 "conv" we already trust -- here simply convert x_raw (no snake) into the format;
        # the unit under test is linear in how x_act was produced.
        if f == L.FMT_F32:
            em.transpose(em.ext(1), x_act, 1, 1, n)
        else:
            em.convert(em.ext(1), L.FMT_F32, x_act, f, n)
        if pr == L.PREC_F32:
            h = em.new(n)
            em.conv(ru.c7, x_act, B, Lx, out_act=h, alpha=ru.a2, prec=pr, x_fmt=f, act_fmt=f)
            y_act = em.new(n)
            em.conv(ru.c1, h, B, Lx, res=em.ext(1), out_raw=em.ext(2), out_act=y_act, alpha=a_next, prec=pr, x_fmt=f, act_fmt=f)
            em.transpose(y_act, em.ext(3), 1, 1, n)
        else:
            if eng.lib.b2c_ru_tc_eligible(eng.ctx, ru.c7.wid, ru.c1.wid, pr) != 1:
                out[prec] = None
                continue
            y_act = em.new(n)
            L.check(eng.lib.b2c_prog_ru(em.h, ru.c7.wid, ru.a2, ru.c1.wid, em._r(x_act), em._r(em.ext(1)), em._r(em.ext(2)),
                                        em._r(y_act), a_next, B, Lx, ru.c7.dilation, pr, f), "b2c_prog_ru")
            em.convert(y_act, f, em.ext(3), L.FMT_F32, n)
        prog = em.finish(3)
        raw = torch.empty(B, Lx, C_, device=dev)
        act = torch.empty(B, Lx, C_, device=dev)
        ext = [x_raw.data_ptr(), raw.data_ptr(), act.data_ptr()]
        eng.run(prog, ext)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            pf = eng.profile(prog, ext)
            best = min(best, sum(r["ms"] for r in pf if r["kind"].startswith("conv")))
        out[prec] = (raw, act, best)
        eng.lib.b2c_prog_destroy(prog.handle)
    r0, a0, t0 = out["f32"]
    line = f"{name:30s} B={B} L={Lx:6d} C={C_:4d} d={ru.c7.dilation} | f32 2 launches {t0:7.3f} ms"
    ok = True
    for prec in precs:
        if out[prec] is None:
            line += f" | {prec}: not eligible"
            continue
        r, a, t = out[prec]
        e_raw, e_act = float((r - r0).abs().max()), float((a - a0).abs().max())
        tol = (3e-4 if prec == "bf16x3" else 5e-2) * max(float(r0.abs().max()), 1.0)
        bad = not (e_raw < tol) or not torch.isfinite(r).all() or not torch.isfinite(a).all()
        ok = ok and not bad
        line += f" | {prec} fused: {t:7.3f} ms {flops / t / 1e9:7.1f} TF/s err raw {e_raw:.2e} act {e_act:.2e}{' FAIL' if bad else ''}"
    print(line + f" | |y|max {float(r0.abs().max()):.2f}", flush=True)
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
    if args.group in ("ru", "all"):
        T = args.T
        enc, dec = net.T_ENC.block, net.T_DEC.model
        units = [("ru.enc1.d1", enc[1].block[0], T), ("ru.enc1.d3", enc[1].block[1], T), ("ru.enc1.d9", enc[1].block[2], T),
                 ("ru.enc2.d1", enc[2].block[0], T // 2), ("ru.enc2.d9", enc[2].block[2], T // 2),
                 ("ru.dec3.d1", dec[3].block[2], 11996 * T // 24000), ("ru.dec3.d9", dec[3].block[4], 11996 * T // 24000),
                 ("ru.dec4.d1", dec[4].block[2], 23992 * T // 24000), ("ru.dec4.d3", dec[4].block[3], 23992 * T // 24000),
                 ("ru.dec4.d9", dec[4].block[4], 23992 * T // 24000)]
        for name, mod, Lx in units:
            if args.only and args.only not in name:
                continue
            all_ok &= run_ru_case(eng, name, mod, args.batch, Lx, dev, precs)
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

#!/usr/bin/env python
"""Pipeline timeline of the fused residual-unit kernel (developer aid): runs one unit with B2C_TC_DEBUG bit 8 set,
reads the events CTA 0 recorded for a few steady-state tiles and prints them in clock order.

    B2C_TC_DEBUG=8 python tools/ru_trace.py [--unit enc1|enc2|dec3|dec4] [--prec bf16x3|bf16] [--batch 32]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B2C_TC_DEBUG", "8")

import torch  # noqa: E402

import multimodal_vqvae_compression_audio_tactile_b200 as pkg  # noqa: E402
from multimodal_vqvae_compression_audio_tactile_b200 import _lib as L  # noqa: E402
from multimodal_vqvae_compression_audio_tactile_b200.engine import Emitter, Engine, _pack_ru  # noqa: E402

TAGS = {1: "tma  stage(conv7) acquired", 2: "tma  stage(w1) acquired", 10: "mma  G1 start (t1empty ok)",
        11: "mma  G1 stage full", 12: "mma  G1 issued+commit", 13: "mma  G2 hfull ok", 14: "mma  G2 t2empty ok",
        15: "mma  G2 stage full", 16: "mma  G2 issued+commit", 20: "epi  tile start", 21: "epi  t1full ok",
        22: "epi  hempty ok", 23: "epi  A done (hfull arrive)", 24: "epi  t2full ok", 25: "epi  B done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unit", default="enc1")
    ap.add_argument("--prec", default="bf16x3")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    net = pkg.build_proposed(2, 128)
    eng = Engine(dev)
    enc, dec = net.T_ENC.block, net.T_DEC.model
    mod, Lx = {"enc1": (enc[1].block[0], 24000), "enc2": (enc[2].block[0], 12000), "dec3": (dec[3].block[2], 11996),
               "dec4": (dec[4].block[2], 23992)}[args.unit]
    ru = _pack_ru(eng, mod)
    C_ = ru.c7.cout
    B = args.batch
    n = B * Lx * C_
    pr = L.PRECISIONS[args.prec]
    f = L.FMT_OF_PREC[pr]
    x_raw = (torch.rand(B, Lx, C_) * 2 - 1).to(dev)
    a_next = eng.pack_vec(torch.rand(C_) + 0.5)
    em = Emitter(eng)
    x_act = em.new(n)
    em.convert(em.ext(1), L.FMT_F32, x_act, f, n)
    y_act = em.new(n)
    L.check(eng.lib.b2c_prog_ru(em.h, ru.c7.wid, ru.a2, ru.c1.wid, em._r(x_act), em._r(em.ext(1)), em._r(em.ext(2)),
                                em._r(y_act), a_next, B, Lx, ru.c7.dilation, pr, f), "b2c_prog_ru")
    prog = em.finish(2)
    raw = torch.empty(B, Lx, C_, device=dev)
    ext = [x_raw.data_ptr(), raw.data_ptr()]
    buf = (C.c_uint64 * 8192)()
    eng.lib.b2c_debug_ru_trace(buf, 8192)
    for rep in range(3):
        eng.run(prog, ext)
        torch.cuda.synchronize()
        cnt = eng.lib.b2c_debug_ru_trace(buf, 8192)
    pf = eng.profile(prog, ext)
    print("launch ms:", [round(r["ms"], 4) for r in pf if r["kind"].startswith("conv")])
    ev = sorted(((v & ((1 << 44) - 1)), (v >> 44) & 0xFFF, v >> 56) for v in list(buf)[:cnt] if v)
    cnt = len(ev)
    if not ev:
        print("no trace entries (B2C_TC_DEBUG bit 8 not set, or fewer than 9 tiles per CTA)")
        return
    t0 = ev[0][0]
    prev = {}
    print(f"{cnt} events; clock cycles relative to the first")
    for t, tile, tag in ev:
        role = TAGS.get(tag, str(tag))[:4] + (str(tile & 1) if tag >= 20 else "")
        d = t - prev.get(role, t)
        prev[role] = t
        print(f"{t - t0:8d}  (+{d:6d} in role)  tile {tile:3d}  {TAGS.get(tag, tag)}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Power / clock probe (developer aid): loops ONE launch of a layer for a few seconds while nvidia-smi samples board
power and SM clock, to tell power-capped kernels (clock below max under sw_power_cap) from latency-bound ones.
    python tools/power_probe.py [--batch 64] [--secs 2.5]"""
import argparse, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from multimodal_vqvae_compression_audio_tactile_b200 import _lib as L
from multimodal_vqvae_compression_audio_tactile_b200.engine import Emitter, Engine, _pack_ru

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--secs", type=float, default=2.5)
ap.add_argument("--only", default="")
ap.add_argument("--no-program", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = pkg.build_proposed(8, 512)
eng = Engine(dev)
B = a.batch


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()
    def _rd(self):
        for l in self.p.stdout:
            self.rows.append((time.time(), l.strip().split(", ")))
    def window(self, t0, t1):
        r = [x for t, x in self.rows if t0 <= t <= t1]
        if not r: return None
        pw = sorted(float(x[0]) for x in r); ck = sorted(int(x[1]) for x in r)
        cap = sum(1 for x in r if x[2].lower() == "active")
        return pw[len(pw) // 2], ck[len(ck) // 2], cap, len(r)


S = Sampler()
time.sleep(0.5)


def loop(name, prog, ext, flops=0.0):
    eng.run(prog, ext); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time(); e0.record()
    while time.time() - t0 < a.secs:
        for _ in range(50): eng.run(prog, ext)
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    w = S.window(t0 + 0.8, t1)
    print(f"{name:28s} {ms:7.3f} ms/launch  power {w[0]:6.0f} W  clock {w[1]} MHz  capped {w[2]}/{w[3]}  {flops / ms / 1e9 if flops else 0:7.1f} TF/s", flush=True)
    time.sleep(0.7)


def conv_case(name, m, Lin, prec):
    w = eng.pack_wnconv(m)
    pr = L.PRECISIONS[prec]; f = L.FMT_OF_PREC[pr]
    Lout = (Lin + 2 * m.padding - m.dilation * (m.kernel_size - 1) - 1) // m.stride + 1
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, Lin, w.cin, generator=g) * 2 - 1).to(dev)
    res = (torch.rand(B, Lout, w.cout, generator=g) * 2 - 1).to(dev)
    alpha = eng.pack_vec(torch.rand(w.cout, generator=g) + 0.5)
    em = Emitter(eng)
    xa = em.new(B * Lin * w.cin)
    em.convert(em.ext(1), L.FMT_F32, xa, f, B * Lin * w.cin)
    prog0 = em  # conversion happens once below via a separate program
    act = em.new(B * Lout * w.cout)
    em.conv(w, xa, B, Lin, res=em.ext(2), out_raw=em.ext(3), out_act=act, alpha=alpha, prec=pr, x_fmt=f, act_fmt=f)
    prog = em.finish(3)
    raw = torch.empty(B, Lout, w.cout, device=dev)
    # the looped program includes the (cheap) input conversion; report it but it is < 10 % of the wide layers
    loop(name + " " + prec, prog, [x.data_ptr(), res.data_ptr(), raw.data_ptr()], 2.0 * B * Lout * w.cout * w.cin * w.k)


def ru_case(name, mod, Lx, prec):
    ru = _pack_ru(eng, mod)
    C_ = ru.c7.cout
    pr = L.PRECISIONS[prec]; f = L.FMT_OF_PREC[pr]
    g = torch.Generator().manual_seed(2)
    x_raw = (torch.rand(B, Lx, C_, generator=g) * 2 - 1).to(dev)
    a_next = eng.pack_vec(torch.rand(C_, generator=g) + 0.5)
    n = B * Lx * C_
    # program 1: convert once; program 2: the unit alone, reading the converted planes from an external buffer
    planes = torch.empty(n * 4, dtype=torch.uint8, device=dev)
    em = Emitter(eng)
    em.convert(em.ext(1), L.FMT_F32, em.ext(2), f, n)
    p1 = em.finish(2)
    eng.run(p1, [x_raw.data_ptr(), planes.data_ptr()]); torch.cuda.synchronize()
    em = Emitter(eng)
    y_act = em.new(n)
    L.check(eng.lib.b2c_prog_ru(em.h, ru.c7.wid, ru.a2, ru.c1.wid, em._r(em.ext(1)), em._r(em.ext(2)), em._r(em.ext(3)),
                                em._r(y_act), a_next, B, Lx, ru.c7.dilation, pr, f), "b2c_prog_ru")
    prog = em.finish(3)
    raw = torch.empty(B, Lx, C_, device=dev)
    loop(name + " " + prec, prog, [planes.data_ptr(), x_raw.data_ptr(), raw.data_ptr()], 2.0 * n * C_ * 8)


enc, dec = net.T_ENC.block, net.T_DEC.model
_ru, _cv = ru_case, conv_case
def ru_case(name, *r):
    if a.only and not any(o in name for o in a.only.split(",")): return
    _ru(name, *r)
def conv_case(name, *r):
    if a.only and not any(o in name for o in a.only.split(",")): return
    _cv(name, *r)
ru_case("ru.enc1.d1 C64", enc[1].block[0], 24000, "bf16x3")
ru_case("ru.enc2.d1 C128", enc[2].block[0], 12000, "bf16x3")
conv_case("enc3 k7 C256", enc[3].block[0].block[1], 3000, "bf16x3")
conv_case("enc3 k1 C256", enc[3].block[0].block[3], 3000, "bf16x3")
conv_case("enc4 k7 C512", enc[4].block[0].block[1], 600, "bf16x3")
conv_case("dec1 k7 C768", dec[1].block[2].block[1], 600, "bf16")
conv_case("dec2 k7 C384", dec[2].block[2].block[1], 2999, "bf16")
ru_case("ru.dec3.d1 C192", dec[3].block[2], 11996, "bf16")
ru_case("ru.dec3.d3 C192", dec[3].block[3], 11996, "bf16")
ru_case("ru.dec3.d9 C192", dec[3].block[4], 11996, "bf16")
ru_case("ru.dec4.d1 C96", dec[4].block[2], 23992, "bf16")
ru_case("ru.dec4.d9 C96", dec[4].block[4], 23992, "bf16")
# whole codec program
if a.no_program:
    S.p.terminate(); sys.exit(0)
x = torch.rand(B, 1, 24000, device=dev) * 2 - 1
net.forward_eval(x, x); torch.cuda.synchronize()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 2 * a.secs:
    net.forward_eval(x, x); n += 1
e1.record(); torch.cuda.synchronize(); t1 = time.time()
w = S.window(t0 + 0.8, t1)
print(f"{'forward_eval':28s} {e0.elapsed_time(e1) / n:7.3f} ms/program  power {w[0]:6.0f} W  clock {w[1]} MHz  capped {w[2]}/{w[3]}")
S.p.terminate()

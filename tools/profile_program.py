#!/usr/bin/env python
"""Per-launch CUDA-event profile of one codec program (b2c_prog_profile): by-kind totals and the slowest launches.
    python tools/profile_program.py [--batch 1] [--books 8 --codes 512] [--precision tc] [--out file.json]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--books", type=int, default=8)
ap.add_argument("--codes", type=int, default=512)
ap.add_argument("--precision", default="tc")
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = pkg.build_proposed(a.books, a.codes)
net.precision = a.precision
T = 24000
x = torch.rand(a.batch, 1, T, device=dev) * 2 - 1
net.forward_eval(x, x)
eng, pk = net._engine(dev)
prog = net.program(eng, pk, a.batch, T, a.books)
Tl, Lout = prog.info["Tl"], prog.info["Lout"]
y = torch.empty(a.batch, Lout, device=dev)
i1 = torch.empty(a.batch, a.books, Tl, device=dev, dtype=torch.int32)
i2 = torch.empty(a.batch, 32, Tl, device=dev, dtype=torch.int32)
ext = [x.data_ptr(), x.data_ptr(), y.data_ptr(), i1.data_ptr(), i2.data_ptr(), 0]
eng.profile(prog, ext)
best = None
for _ in range(3):
    pf = eng.profile(prog, ext)
    if best is None or sum(r["ms"] for r in pf) < sum(r["ms"] for r in best):
        best = pf
kinds = {}
for r in best:
    k = kinds.setdefault(r["kind"], [0.0, 0])
    k[0] += r["ms"]; k[1] += 1
print("batch", a.batch, "launches", len(best), "total ms %.3f" % sum(r["ms"] for r in best))
print({k: (round(v[0], 3), v[1]) for k, v in kinds.items()})
for i in sorted(range(len(best)), key=lambda i: -best[i]["ms"])[: a.top]:
    r = best[i]
    print(i, r["kind"], "%.4f ms" % r["ms"], "%.2f GF" % (r["flops"] / 1e9), "%.1f MB" % (r["bytes"] / 1e6))
if a.out:
    json.dump(dict(batch=a.batch, launches=best), open(a.out, "w"), indent=1)

#!/usr/bin/env python
"""Batch-1..4 forward_eval latency (p50 over --reps, CUDA events), eager and CUDA-graph replay, one or two launch queues.
    B2C_LANES=0|1 python tools/lat_quick.py [--reps 200]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=200); ap.add_argument("--batches", default="1,2,4")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(7)
net = pkg.build_proposed(8, 512).to(dev).eval()
def p50(fn, reps):
    for _ in range(5): fn()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); out.append(e0.elapsed_time(e1))
    s = np.sort(out); return float(s[len(s)//2]), float(s[int(0.99*len(s))])
for B in [int(b) for b in a.batches.split(",")]:
    x = torch.zeros(B, 1, 24000, device=dev)
    net.use_cuda_graph = False
    y0 = net.forward_eval(x, x).clone(); i0 = net.last_indices.clone()
    e = p50(lambda: net.forward_eval(x, x), a.reps)
    enc = p50(lambda: net.encode_latents(x, x), a.reps)
    z = net.encode_latents(x, x)
    dec = p50(lambda: net.T_DEC(z), a.reps)
    net.use_cuda_graph = True
    y1 = net.forward_eval(x, x).clone()
    g = p50(lambda: net.forward_eval(x, x), a.reps)
    print(f"lanes={os.environ.get('B2C_LANES','auto')} B={B}: forward p50/p99 {e[0]:.3f}/{e[1]:.3f} ms, encode {enc[0]:.3f}, decode {dec[0]:.3f}, graph {g[0]:.3f}/{g[1]:.3f}; "
          f"graph==eager {bool(torch.equal(y0, y1))}", flush=True)

#!/usr/bin/env python
"""Config 5 of BASELINE.json: codebook-search stress sweep (K x D x N) on one GPU.

For every shape: the tcgen05 search (`precision="tc"`) and the FP32 CUDA-core search (`"f32"`) of
ResidualVQEMA._nearest_l2 (Evaluation/dac_vcpwq_proposed6_latency.py:417-419) must return identical indices;
both are compared with the reference expression `(x @ emb.T - 0.5 * (emb * emb).sum(1)).argmax(1)` evaluated by
PyTorch on the same GPU (fp32, TF32 off) -- differences are counted together with the reference's own top-1/top-2
margin so near-ties can be told from errors.  Times are CUDA-event medians; bytes/FLOPs are SURVEY 8(d)'s
algorithmic figures (2*N*D*K FLOP; 4*(N*D + K*D) + 4*N bytes).

    python tools/search_sweep.py [--quick] [--out gpurun_out/search_sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import multimodal_vqvae_compression_audio_tactile_b200 as pkg  # noqa: E402
from oracle import cases  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    Ks = [256, 1024, 8192] if args.quick else [256, 512, 1024, 2048, 4096, 8192]
    Ds = [8, 96, 256] if args.quick else [8, 16, 32, 64, 96, 128, 256]
    Ns = [75, 4800, 65536] if args.quick else [1, 75, 1024, 4800, 16384, 65536]
    rows, bad = [], 0
    for K in Ks:
        for D in Ds:
            for N in Ns:
                x, emb = cases.search_inputs(N, D, K)
                x, emb = x.to(dev), emb.to(dev)
                i_tc = pkg.nearest_code(x, emb, precision="tc")
                i_f32 = pkg.nearest_code(x, emb, precision="f32")
                sc = x @ emb.t() - 0.5 * (emb * emb).sum(1)
                i_ref = sc.argmax(1)
                top2 = sc.topk(2, dim=1).values if K > 1 else None
                margin = (top2[:, 0] - top2[:, 1]) if top2 is not None else torch.zeros(N, device=dev)
                d_tc = i_tc != i_ref
                worst = float(margin[d_tc].max()) if d_tc.any() else 0.0
                same = bool(torch.equal(i_tc, i_f32))
                t_tc = timed(lambda: pkg.nearest_code(x, emb, precision="tc"))
                t_f32 = timed(lambda: pkg.nearest_code(x, emb, precision="f32"))
                t_ref = timed(lambda: (x @ emb.t() - 0.5 * (emb * emb).sum(1)).argmax(1))
                flops = 2.0 * N * D * K
                byts = 4.0 * (N * D + K * D) + 4.0 * N
                row = dict(K=K, D=D, N=N, ms_tc=t_tc, ms_f32=t_f32, ms_torch_gpu=t_ref, tc_equals_f32=same,
                           diff_vs_torch=int(d_tc.sum()), worst_margin_at_diff=worst,
                           tflops_tc=flops / t_tc / 1e9, gbs_tc=byts / t_tc / 1e6,
                           frac_tensor=3 * flops / t_tc / 1e9 / peaks["bf16_tflops"], frac_hbm=byts / t_tc / 1e6 / peaks["hbm_gbs"])
                rows.append(row)
                ok = same and worst < 1e-5
                bad += 0 if ok else 1
                print(f"K={K:5d} D={D:3d} N={N:6d} tc {t_tc:7.3f} ms  f32 {t_f32:7.3f} ms  torch {t_ref:7.3f} ms | "
                      f"tc==f32 {same} diff_vs_torch {int(d_tc.sum())} (worst margin {worst:.1e}) | "
                      f"{row['tflops_tc']:7.1f} TF/s alg ({row['frac_tensor']*100:4.1f}% tensor x3) {row['gbs_tc']:7.1f} GB/s"
                      f"{'' if ok else '  <-- CHECK'}", flush=True)
    if args.out:
        json.dump(dict(peaks=peaks, rows=rows), open(args.out, "w"), indent=1)
    print("SWEEP", "OK" if bad == 0 else f"{bad} shapes to check")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Config 4 of BASELINE.json: batch-1 streaming latency, the B200 counterpart of measure_proposed_latency
(Evaluation/dac_vcpwq_proposed6_latency.py:489-525): zeros [1,1,24000] x 2, 3 warm-ups, then per repetition
`encode_latents` (both encoders, DAC quantizer, predictor + residual VQ) and `T_DEC(z_run)` timed separately --
here with CUDA events over --reps repetitions, reporting mean (the reference's statistic), p50 and p99.
Reported for eager launches and for CUDA-graph replay (`net.use_cuda_graph`), next to the reference's published
numbers (unknown GPU, fp16 autocast) and the CPU oracle on this box's host cores.

    python tools/latency.py [--reps 1000] [--books 10 --codes 512] [--out gpurun_out/latency.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import multimodal_vqvae_compression_audio_tactile_b200 as pkg  # noqa: E402
from oracle import cases, proposed  # noqa: E402


def stats(ms):
    a = np.sort(np.asarray(ms))
    return dict(mean=float(a.mean()), p50=float(a[len(a) // 2]), p99=float(a[min(len(a) - 1, int(0.99 * len(a)))]),
                min=float(a[0]), n=len(a))


def time_fn(fn, reps):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=1000)
    ap.add_argument("--books", type=int, default=10)
    ap.add_argument("--codes", type=int, default=512)
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    case = dict(books=args.books, K=args.codes)
    ref = cases.build_reference_style_model(proposed.ProposedEval, case)
    net = pkg.build_proposed(args.books, args.codes)
    net.load_state_dict(ref.state_dict())
    a = torch.zeros(1, 1, 24000, device=dev)
    t = torch.zeros(1, 1, 24000, device=dev)
    res = {"config": dict(books=args.books, codes=args.codes, frame_samples=24000, batch=1, precision=net.precision,
                          input="zeros (the latency script's input, :498-499)")}
    for mode in ("eager", "cuda_graph"):
        net.use_cuda_graph = mode == "cuda_graph"
        net.T_DEC.micro_batch = 1
        for _ in range(3):
            z = net.encode_latents(a, t, args.books)
            net.T_DEC(z)
        torch.cuda.synchronize()
        z = net.encode_latents(a, t, args.books).clone()
        enc = time_fn(lambda: net.encode_latents(a, t, args.books), args.reps)
        dec = time_fn(lambda: net.T_DEC(z), args.reps)
        e2e = time_fn(lambda: net.forward_eval(a, t, args.books), args.reps)
        res[mode] = dict(encoding_delay_ms=stats(enc), decoding_delay_ms=stats(dec), forward_eval_ms=stats(e2e))
        print(mode, json.dumps(res[mode]), flush=True)
    # CPU oracle (the reference's classes on the restated backbone), fp32, all host threads
    a_c, t_c = torch.zeros(1, 1, 24000), torch.zeros(1, 1, 24000)
    with torch.no_grad():
        ref.encode_latents(a_c, t_c, args.books)
        ce, cd = [], []
        for _ in range(args.cpu_reps):
            t0 = time.perf_counter(); zc = ref.encode_latents(a_c, t_c, args.books); ce.append((time.perf_counter() - t0) * 1e3)
            t0 = time.perf_counter(); ref.T_DEC(zc); cd.append((time.perf_counter() - t0) * 1e3)
    res["cpu_oracle"] = dict(encoding_delay_ms=stats(ce), decoding_delay_ms=stats(cd), threads=torch.get_num_threads())
    res["reference_published"] = dict(encoding_delay_ms=[12.83, 16.27], decoding_delay_ms=[2.75, 2.86],
                                      note="eval_all_vs_dac24_vcpwq_rawPSNR_latency.json, unknown GPU, fp16 autocast, mean of 10")
    print("cpu", json.dumps(res["cpu_oracle"]), flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- signal-seconds per second of encode -> quantize -> decode (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference's CPU path (oracle port)

A step = one forward_eval over one batch of synthetic 1-s frames (24 kHz) per GPU; frames are
independent, so N GPUs run N independent shards (weak scaling, no collective on the hot path; the
code indices are gathered ONCE, after the timed steps).  One JSON line is printed by rank 0; besides the
contract's keys it carries `latency` (config 4: batch-1 p50/p99), `search` (config 5 corners, rows sharded
over the GPUs), `strong_scaling` (fixed global batch, N > 1), `torch_gpu_baseline`, `train_step` (forward_step +
backward through the CUDA decoder) and `eval_metrics` (alignment / PSNR / ST-SIM kernels) as context.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "signal-sec/sec encode+quantize+decode"
UNIT = "signal-s/s"
WORKLOAD = "configs[1]: ProposedEval (compare_dacvsproposal_5 grid point rvqB8_K512), forward_eval on 1-s 24 kHz frames"
BOOKS, K_CODES, T = 8, 512, 24000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=128,
                    help="frames per program run (64: +7 %% over 32, 128: +1 %% over 64 -- fewer partial waves of tiles on 148 "
                         "SMs; two programs per step keep the host copies of the e2e leg under the other program)")
    ap.add_argument("--precision", default=os.environ.get("B2C_PRECISION", "auto"))
    ap.add_argument("--cpu-sample", type=int, default=8, help="frames in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the untimed legs (latency percentiles, search corners, strong-scaling point, torch GPU baseline)")
    ap.add_argument("--latency-reps", type=int, default=500)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


TRAFFIC_CAPTURES = (("r02_ncu_dram_traffic_conv_mb128.json", 128), ("r02_ncu_dram_traffic_conv_mb64.json", 64),
                    ("r01_ncu_dram_traffic_conv_mb64.json", 64))


def pick_traffic(micro_batch: int, family: str, profiles_dir: str = None):
    """DRAM bytes per launch of the dominant family from the committed ncu capture taken at THIS micro-batch (bytes per
    launch scale with it); (None, None) when there is no capture for it or the dominant family is not the bf16x3 one."""
    d = profiles_dir or os.path.join(ROOT, "profiles")
    if family != "conv_tc_x3":
        return None, None
    for name, mb_file in TRAFFIC_CAPTURES:
        tp = os.path.join(d, name)
        if mb_file == micro_batch and os.path.isfile(tp):
            try:
                return json.load(open(tp))["x3"]["dram_bytes_per_launch"], "profiles/" + name
            except Exception:
                continue
    return None, None


def build_oracle():
    from oracle import cases, proposed
    case = dict(books=BOOKS, K=K_CODES)
    return cases.build_reference_style_model(proposed.ProposedEval, case)


def cpu_time_forward(model, frames: int, reps: int, threads: int):
    import torch
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(123)
    a = torch.rand(frames, 1, T, generator=g) * 2 - 1
    t = torch.rand(frames, 1, T, generator=g) * 2 - 1
    with torch.no_grad():
        model.forward_eval(a, t)          # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            model.forward_eval(a, t)
            ts.append(time.perf_counter() - t0)
    return frames / (sum(ts) / len(ts)), ts


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (its own classes restated in
    oracle/proposed.py on the DAC-architecture backbone of oracle/dac_arch.py; /root/reference and the
    `dac` package do not exist on the GPU box), all host threads, fp32, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    model = build_oracle()
    torch.set_num_threads(threads)
    frames = args.cpu_sample
    g = torch.Generator().manual_seed(123)
    a = torch.rand(frames, 1, T, generator=g) * 2 - 1
    t = torch.rand(frames, 1, T, generator=g) * 2 - 1
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            model.forward_eval(a, t)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model.forward_eval(a, t)
        el = time.perf_counter() - t0
    ms = el / args.steps * 1e3
    val = frames / (ms / 1e3)
    sample = f"{frames} one-second frames per step, fp32, torch {torch.__version__} CPU"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "books": BOOKS, "codes": K_CODES,
                                                        "frames_per_step": frames},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def _pctl(ms):
    v = sorted(ms)
    return dict(p50=v[len(v) // 2], p99=v[min(len(v) - 1, int(0.99 * len(v)))], mean=sum(v) / len(v), n=len(v))


def _event_times(torch, fn, reps):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def latency_leg(torch, net, dev, reps):
    """Config 4 (Evaluation/dac_vcpwq_proposed6_latency.py:489-525): zeros [1,1,24000] x 2, 3 warm-ups, then
    encode_latents and T_DEC(z_run) timed separately per repetition -- CUDA events, p50/p99 over `reps`."""
    a = torch.zeros(1, 1, T, device=dev)
    t = torch.zeros(1, 1, T, device=dev)
    for _ in range(3):
        z = net.encode_latents(a, t)
        net.T_DEC(z)
        net.forward_eval(a, t)
    torch.cuda.synchronize()
    z = net.encode_latents(a, t).clone()
    enc = _pctl(_event_times(torch, lambda: net.encode_latents(a, t), reps))
    dec = _pctl(_event_times(torch, lambda: net.T_DEC(z), reps))
    fwd = _pctl(_event_times(torch, lambda: net.forward_eval(a, t), reps))
    net.use_cuda_graph = True
    try:
        for _ in range(3):
            net.forward_eval(a, t)
        torch.cuda.synchronize()
        fwd_g = _pctl(_event_times(torch, lambda: net.forward_eval(a, t), reps))
    finally:
        net.use_cuda_graph = False
    eng, pk = net._engine(dev)
    prog = net.program(eng, pk, 1, T, BOOKS)
    return {"unit": "ms", "batch": 1, "reps": reps, "input": "zeros [1,1,24000] (the latency script's frame)",
            "encode_p50": enc["p50"], "encode_p99": enc["p99"], "decode_p50": dec["p50"], "decode_p99": dec["p99"],
            "forward_p50": fwd["p50"], "forward_p99": fwd["p99"], "forward_graph_p50": fwd_g["p50"],
            "forward_graph_p99": fwd_g["p99"], "launches_forward": prog.info["launches"],
            "reference_published_ms": {"encode": [12.83, 16.27], "decode": [2.75, 2.86], "hardware": "unknown GPU, fp16 autocast"}}


SEARCH_CORNERS = [  # (N, D, K): config 5 corners incl. the K=8192 / D=96 / N=65536 point of SURVEY 8(d)
    (65536, 96, 8192), (65536, 256, 8192), (65536, 8, 256), (16384, 64, 2048), (4800, 96, 512), (1024, 128, 4096),
    (75, 96, 1024), (65536, 96, 1024)]


def search_leg(torch, dist, pkg, dev, world, rank, peaks):
    """Config 5: nearest-code search corners; with N GPUs every rank searches its shard of the N rows (the codebook is
    replicated), time = max over ranks.  frac = the op's bound (max of the executed-MMA time at the sustained bf16 peak
    and the algorithmic bytes at the measured HBM rate) / measured time."""
    from multimodal_vqvae_compression_audio_tactile_b200 import driver
    rows = []
    for (N, D, K) in SEARCH_CORNERS:
        g = torch.Generator().manual_seed(1000003 * D + 7919 * K + N)
        x = torch.randn(N, D, generator=g) / D ** 0.5
        emb = torch.randn(K, D, generator=g) / D ** 0.5
        lo, hi = driver.shard_bounds(N, world, rank)
        n_loc = hi - lo
        ms = float("nan")
        if n_loc > 0:
            xs, es = x[lo:hi].to(dev), emb.to(dev)
            for _ in range(3):
                pkg.nearest_code(xs, es)
            torch.cuda.synchronize()
            v = sorted(_event_times(torch, lambda: pkg.nearest_code(xs, es), 10))
            ms = v[len(v) // 2]
        mt = torch.tensor([0.0 if ms != ms else ms], device=dev)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        ms = mt.item()
        n_max = (N + world - 1) // world
        flops = 2.0 * n_max * D * K
        byts = 4.0 * (n_max * D + K * D) + 4.0 * n_max
        bound_ms = max(3.0 * flops / (peaks["tf_sus"] * 1e12), byts / (peaks["hbm"] * 1e9)) * 1e3
        rows.append({"N": N, "D": D, "K": K, "rows_per_gpu": n_max, "ms": ms, "frac": bound_ms / ms if ms > 0 else None,
                     "tflops_algorithmic_all_gpus": 2.0 * N * D * K / ms / 1e9 if ms > 0 else None})
    return rows


def torch_gpu_leg(torch, ref, dev, frames=16, reps=3):
    """Untimed-context leg ('the bar for the new kernels is PyTorch-eager cuDNN/cuBLAS on the same B200'):
    the reference's own classes on the restated backbone, moved to the GPU, eager, fp32 (TF32 off) and under bf16
    autocast.  Not the product, not part of `value`."""
    import copy
    out = {"frames_per_step": frames, "unit": UNIT}
    try:
        m = copy.deepcopy(ref).to(dev).eval()
        g = torch.Generator().manual_seed(123)
        a = (torch.rand(frames, 1, T, generator=g) * 2 - 1).to(dev)
        t = (torch.rand(frames, 1, T, generator=g) * 2 - 1).to(dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        for name, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16), ("fp16_autocast", torch.float16)):
            def run():
                with torch.no_grad():
                    if ctx is None:
                        return m.forward_eval(a, t)
                    with torch.autocast("cuda", dtype=ctx):
                        return m.forward_eval(a, t)
            run(); run()
            torch.cuda.synchronize()
            ms = sorted(_event_times(torch, run, reps))[reps // 2]
            out[name] = frames / (ms / 1e3)
        del m
        torch.cuda.empty_cache()
    except Exception as e:   # context only: never fails the bench
        out["error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def train_leg(torch, net, dev, frames=6, reps=5):
    """Context leg: one training-mode forward_step + backward (Training/compare_dacvsproposal_3.py:
    386-409, BATCH = 6, 1-s frames): frozen backbones and residual VQ in libb2c.so, T_DEC forward AND backward-data in
    libb2c.so, the 9 M-parameter trainable layers through autograd.  ms per step, CUDA events."""
    out = {"frames": frames, "unit": "ms per forward_step + backward"}
    try:
        import multimodal_vqvae_compression_audio_tactile_b200 as pkg
        host_net = net
        net = pkg.build_proposed(BOOKS, K_CODES)           # the autograd side needs the parameters on the device
        net.load_state_dict(host_net.state_dict())
        net.precision = host_net.precision
        net = net.to(dev)
        g = torch.Generator().manual_seed(5)
        a = (torch.rand(frames, 1, T, generator=g) * 2 - 1).to(dev)
        t = (torch.rand(frames, 1, T, generator=g) * 2 - 1).to(dev)

        def step():
            for p_ in net.parameters():
                p_.grad = None
            o = net.forward_step(a, t)
            torch.nn.functional.l1_loss(o["y_hat"], o["tgt"]).backward()

        with torch.enable_grad():
            step(); step()
            torch.cuda.synchronize()
            ms = sorted(_event_times(torch, step, reps))
        out["ms"] = ms[len(ms) // 2]
        out["signal_s_per_s"] = frames / (out["ms"] / 1e3)
        with torch.no_grad():
            net.forward_eval(a, t)
            torch.cuda.synchronize()
            out["forward_eval_ms"] = sorted(_event_times(torch, lambda: net.forward_eval(a, t), reps))[reps // 2]
        for p_ in net.parameters():
            p_.grad = None
        torch.cuda.empty_cache()
    except Exception as e:   # context only: never fails the bench
        out["error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def metrics_leg(torch, dev, frames=64, reps=10):
    """Context leg: the evaluation metrics of Evaluation/compare_dacvsproposal_5_eval.py on `frames`
    decoded frames: 401-shift alignment + 24k->3k resample + PSNR (psnr_3k_aligned_batch) and ST-SIM."""
    from multimodal_vqvae_compression_audio_tactile_b200 import metrics as pm
    out = {"frames": frames, "samples": 23992, "unit": "ms per batch"}
    try:
        g = torch.Generator().manual_seed(6)
        r = (torch.rand(frames, 1, 23992, generator=g) - 0.5).to(dev)
        e = (torch.roll(r, 17, dims=-1) * 0.9).contiguous()
        for name, fn in (("psnr_3k_aligned_ms", lambda: pm.psnr_3k_aligned_tensor(r, e)),
                         ("stsim_ms", lambda: pm.stsim_tensor(r, e)), ("psnr_ms", lambda: pm.psnr_tensor(r, e))):
            fn(); fn()
            torch.cuda.synchronize()
            v = sorted(_event_times(torch, fn, reps))
            out[name] = v[len(v) // 2]
        out["best_shift"] = int(pm.psnr_3k_aligned_tensor(r, e)[1][0])
    except Exception as e_:
        out["error"] = f"{type(e_).__name__}: {e_}"[:200]
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import multimodal_vqvae_compression_audio_tactile_b200 as pkg
    from multimodal_vqvae_compression_audio_tactile_b200 import driver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # model: random-init weights of the reference architecture, same seed recipe as the oracle
    ref = build_oracle()
    net = pkg.build_proposed(BOOKS, K_CODES)
    net.load_state_dict(ref.state_dict())
    prec = args.precision
    if prec == "auto":
        prec = pkg.DEFAULT_PRECISION
    net.precision = prec
    net.micro_batch = args.micro_batch

    # weak scaling: every rank owns B frames of a world*B global batch (driver.ShardedCodec: contiguous shards)
    B = args.batch
    codec = driver.ShardedCodec(driver.codec_forward_fn(net))
    g = torch.Generator().manual_seed(123 + rank)
    a_host = (torch.rand(B, 1, T, generator=g) * 2 - 1).pin_memory()
    t_host = (torch.rand(B, 1, T, generator=g) * 2 - 1).pin_memory()
    a_dev, t_dev = a_host.to(dev), t_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.zero_()                      # L2 flush between steps (inside the timed region, ~0.1 ms)
        return codec.run_local(a_dev, t_dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y, idx = step()
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = world * B / (ms_step / 1e3)
    # the path's only exchange: ONE final gather of the code indices (north_star), after the timed steps; timed on its own
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    codec.gather_indices(idx)                      # NCCL warm-up (communicator set-up)
    barrier()
    g0.record()
    idx_all = codec.gather_indices(idx)
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    assert idx_all.shape[0] == world * B

    # e2e: same metric through the host-buffer C-ABI entry (H2D + program + D2H inside the timed region)
    y_host = torch.empty(B, 1, net.out_len(T), dtype=torch.float32).pin_memory()       # caller-owned pinned result buffers,
    idx_host = torch.empty(B, BOOKS, net.latent_len(T), dtype=torch.int32).pin_memory()  # reused every step
    for _ in range(2):
        net.forward_eval_host(a_host, t_host, y_out=y_host, idx_out=idx_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        net.forward_eval_host(a_host, t_host, y_out=y_host, idx_out=idx_host)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) / args.steps * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = world * B / (e2e_ms.item() / 1e3)
    h2d_b, d2h_b = net.last_host_bytes
    clocks = sampler.stop() if rank == 0 else None

    # strong-scaling point: a FIXED global batch (one rank's weak-scaling batch) split over the ranks
    strong = None
    if world > 1 and not args.no_extras:
        lo, hi = driver.shard_bounds(B, world, rank)
        a_s, t_s = a_dev[lo:hi], t_dev[lo:hi]
        for _ in range(3):
            net.forward_eval(a_s, t_s)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            flush.zero_()
            net.forward_eval(a_s, t_s)
        s1.record()
        barrier()
        sm = torch.tensor([s0.elapsed_time(s1)], device=dev)
        dist.all_reduce(sm, op=dist.ReduceOp.MAX)
        strong = {"global_batch": B, "frames_per_gpu": hi - lo, "ms_per_step": sm.item() / args.steps,
                  "value": B / (sm.item() / args.steps / 1e3), "unit": UNIT}

    peaks = load_peaks()
    search = None if args.no_extras else search_leg(torch, dist, pkg, dev, world, rank, peaks)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    latency = None if args.no_extras else latency_leg(torch, net, dev, args.latency_reps)

    # ---- roofline of the dominant kernel (conv implicit GEMM), measured live with CUDA events ----
    eng, pk = net._engine(dev)
    mb = min(B, net.micro_batch)
    prog = net.program(eng, pk, mb, T, BOOKS)
    Tl, Lout = prog.info["Tl"], prog.info["Lout"]
    ybuf = torch.empty(mb, Lout, device=dev)
    ibuf = torch.empty(mb, BOOKS, Tl, device=dev, dtype=torch.int32)
    cbuf = torch.empty(mb, 32, Tl, device=dev, dtype=torch.int32)
    ext = [a_dev.data_ptr(), t_dev.data_ptr(), ybuf.data_ptr(), ibuf.data_ptr(), cbuf.data_ptr(), 0]
    eng.profile(prog, ext)
    prof = eng.profile(prog, ext)
    by_kind = {}
    for r in prof:
        d = by_kind.setdefault(r["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        d["ms"] += r["ms"]; d["flops"] += r["flops"]; d["bytes"] += r["bytes"]; d["launches"] += 1
    tot_ms = sum(d["ms"] for d in by_kind.values())
    dom = max((k for k in by_kind if k.startswith("conv")), key=lambda k: by_kind[k]["ms"])
    dd = by_kind[dom]
    achieved = dd["flops"] / (dd["ms"] / 1e3) / 1e12
    # bf16x3 contractions (everything upstream of the quantizer in plan "tc") issue 3 tensor-core FLOPs per
    # algorithmic FLOP: launches before the decoder's first conv are the x3 ones
    executed = dd["flops"] * (3.0 if dom.endswith("_x3") else 1.0)
    # DRAM bytes per launch of the dominant family, from the committed ncu capture of this same bench command
    # (dram__bytes_read.sum + dram__bytes_write.sum of every launch of one program); only valid for the micro-batch
    # it was captured at
    traffic_alg = dd["bytes"] / dd["launches"]
    traffic, traffic_src = pick_traffic(mb, dom)
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sus"], "traffic": traffic, "traffic_algorithmic": traffic_alg,
                "traffic_source": traffic_src,
                "peak_source": f"{peaks['src']} bf16 sustained (MEASURED_PEAKS.json)",
                "share_of_step": dd["ms"] / tot_ms, "launches_per_program": dd["launches"],
                "avg_launch_ms": dd["ms"] / dd["launches"],
                "flops_per_launch_avg": dd["flops"] / dd["launches"],
                "executed_tflops": executed / (dd["ms"] / 1e3) / 1e12,
                "executed_frac": executed / (dd["ms"] / 1e3) / 1e12 / peaks["tf_sus"],
                "note": "achieved = algorithmic conv FLOPs / CUDA-event time of the conv launches of one program; the "
                        "bf16x3 launches execute 3 MMA FLOPs per algorithmic FLOP (executed_tflops); traffic = ncu DRAM "
                        "bytes per launch averaged over the family's launches of one program, traffic_algorithmic = "
                        "the minimal bytes"}
    if args.profile_out:
        with open(args.profile_out, "w") as fh:
            json.dump({"by_kind": by_kind, "launches": prof, "micro_batch": mb}, fh, indent=1)

    launches_per_step = ((B + mb - 1) // mb) * prog.info["launches"] + 1
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"f32": "f32", "bf16x3": "bf16x3+bf16", "bf16": "bf16", "tc": "bf16x3+bf16"}.get(
            prec if isinstance(prec, str) else "tc", "mixed"),
        "data": "synthetic (random-init weights seed 7, U(-1,1) frames)",
        "config": {"workload": WORKLOAD, "books": BOOKS, "codes": K_CODES, "frames_per_gpu_per_step": B,
                   "micro_batch": mb, "precision": prec, "l2": "256 MiB flush between steps, inside the timed region",
                   "parallelism": f"batch-sharded x{world} (driver.ShardedCodec), no hot-path collective; one index "
                                  f"gather after the timed steps (final_gather_ms)"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b},
        "gpu_launches": launches_per_step * args.steps,
        "final_gather_ms": gather_ms,
        "roofline": roofline,
        "kernel_time_ms_per_program": {k: round(v["ms"], 3) for k, v in by_kind.items()},
    }
    if strong is not None:
        out["strong_scaling"] = strong
    if latency is not None:
        out["latency"] = latency
    if search is not None:
        out["search"] = search
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, ts = cpu_time_forward(ref, args.cpu_sample, 2, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_sample} one-second frames, 1 warm-up + 2 timed forward_eval, "
                                         f"fp32 torch CPU ({sum(ts):.1f} s)"}
    if not args.no_extras:
        out["torch_gpu_baseline"] = torch_gpu_leg(torch, ref, dev)
        out["train_step"] = train_leg(torch, net, dev)
        out["eval_metrics"] = metrics_leg(torch, dev)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

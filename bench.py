#!/usr/bin/env python
"""bench.py -- signal-seconds per second of encode -> quantize -> decode (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference's CPU path (oracle port)

A step = one forward_eval over one batch of synthetic 1-s frames (24 kHz) per GPU; frames are
independent, so N GPUs run N independent shards (weak scaling, no collective on the hot path; the
code indices are all-gathered once per step).  One JSON line is printed by rank 0.
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
    ap.add_argument("--batch", type=int, default=128, help="frames per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=64,
                    help="frames per program run (64: measured +7 %% over 32 -- fewer partial waves of tiles on 148 SMs)")
    ap.add_argument("--precision", default=os.environ.get("B2C_PRECISION", "auto"))
    ap.add_argument("--cpu-sample", type=int, default=8, help="frames in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
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

    B = args.batch
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
        y = net.forward_eval(a_dev, t_dev)
        idx = net.last_indices
        if world > 1:                      # the only exchange: final gather of the code indices
            out = [torch.empty_like(idx) for _ in range(world)]
            dist.all_gather(out, idx)
        return y

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = world * B / (ms_step / 1e3)

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
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (conv implicit GEMM), measured live with CUDA events ----
    peaks = load_peaks()
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
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_full_conv_tc.json")
    if os.path.isfile(tpath):
        try:
            ks = json.load(open(tpath))["kernels"]
            traffic_note = [dict(kernel=k["Kernel Name"][:40], us=float(k["gpu__time_duration.sum"]),
                                 dram_MB=float(k["dram__bytes_read.sum"]) + float(k["dram__bytes_write.sum"]),
                                 tensor_pct=float(k["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]))
                            for k in ks]
        except Exception:
            traffic_note = None
    # DRAM bytes per launch of the dominant family, from the committed ncu capture of this same bench command
    # (profiles/r01_ncu_dram_traffic_conv_mb64.json: dram__bytes_read.sum + dram__bytes_write.sum of every launch of
    # one program); only valid for the micro-batch it was captured at
    traffic_alg = dd["bytes"] / dd["launches"]
    tp = os.path.join(ROOT, "profiles", "r01_ncu_dram_traffic_conv_mb64.json")
    if os.path.isfile(tp) and mb == 64 and dom == "conv_tc_x3":
        try:
            traffic = json.load(open(tp))["x3"]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sus"], "traffic": traffic, "traffic_algorithmic": traffic_alg,
                "peak_source": f"{peaks['src']} bf16 sustained (MEASURED_PEAKS.json)",
                "share_of_step": dd["ms"] / tot_ms, "launches_per_program": dd["launches"],
                "avg_launch_ms": dd["ms"] / dd["launches"],
                "flops_per_launch_avg": dd["flops"] / dd["launches"],
                "executed_tflops": executed / (dd["ms"] / 1e3) / 1e12,
                "executed_frac": executed / (dd["ms"] / 1e3) / 1e12 / peaks["tf_sus"],
                "note": "achieved = algorithmic conv FLOPs / CUDA-event time of the conv launches of one program; the "
                        "bf16x3 launches execute 3 MMA FLOPs per algorithmic FLOP (executed_tflops); traffic = ncu DRAM "
                        "bytes per launch averaged over the family's launches of one program "
                        "(profiles/r01_ncu_dram_traffic_conv_mb64.json), traffic_algorithmic = the minimal bytes",
                "ncu_samples": traffic_note}
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
                   "parallelism": f"batch-sharded x{world}, no hot-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "kernel_time_ms_per_program": {k: round(v["ms"], 3) for k, v in by_kind.items()},
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, ts = cpu_time_forward(ref, args.cpu_sample, 2, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_sample} one-second frames, 1 warm-up + 2 timed forward_eval, "
                                         f"fp32 torch CPU ({sum(ts):.1f} s)"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

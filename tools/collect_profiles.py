#!/usr/bin/env python
"""Copy the outputs of tools/run_artifacts.sh (+ run_multi.sh) from gpurun_out/ into profiles/ under their round names."""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst, R = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles"), sys.argv[1] if len(sys.argv) > 1 else "r01"


def first_json(path):
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)


def put(obj, name):
    json.dump(obj, open(os.path.join(dst, name), "w"))


put(first_json(f"{src}/bench_tc.log"), f"{R}_bench_tc_plan.json")
put(first_json(f"{src}/bench_ref.log"), f"{R}_bench_reference_arm.json")
for n in (2, 4, 8):
    if os.path.isfile(f"{src}/bench_{n}gpu.log"):
        put(first_json(f"{src}/bench_{n}gpu.log"), f"{R}_bench_tc_plan_{n}gpu.json")
        put(first_json(f"{src}/bench_{n}gpu_ref.log"), f"{R}_bench_reference_arm_{n}gpu.json")
for a, b in (("latency.json", "latency_b1.json"), ("prof_tc.json", "launch_profile_tc_plan.json"),
             ("search_sweep.json", "search_sweep_config5.json"), ("launches_bench.csv", "ncu_launch_list_bench_tc.csv")):
    shutil.copy(f"{src}/{a}", f"{dst}/{R}_{b}")
for u in ("enc1_bf16x3", "enc2_bf16x3", "dec3_bf16", "dec4_bf16"):
    if os.path.isfile(f"{src}/trace_{u}.log"):
        shutil.copy(f"{src}/trace_{u}.log", f"{dst}/{R}_ru_pipeline_trace_{u}.txt")
py = sys.executable
subprocess.check_call([py, f"{ROOT}/tools/summarize_ncu.py", "launches", f"{src}/launches_bench.csv", f"{dst}/{R}_ncu_launch_list_bench_tc.json"])
subprocess.check_call([py, f"{ROOT}/tools/summarize_ncu.py", "traffic", f"{src}/traffic_conv.csv", f"{dst}/{R}_ncu_dram_traffic_conv_mb64.json"])
b = first_json(f"{src}/bench_tc.log")
print("value", round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), "launches", b["gpu_launches"], "x3 executed frac",
      round(b["roofline"]["executed_frac"], 3), "kernel ms", b["kernel_time_ms_per_program"])

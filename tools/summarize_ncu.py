#!/usr/bin/env python
"""Turn the ncu CSV logs of tools/run_artifacts.sh into the JSON summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_bench.csv profiles/r01_ncu_launch_list_bench_tc.json
    python tools/summarize_ncu.py traffic  gpurun_out/traffic_conv.csv   profiles/r01_ncu_dram_traffic_conv_mb64.json
"""
import collections
import csv
import json
import re
import sys


def read(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    col = {n: H.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    out = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        rec = out.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        rec[r[col["Metric Name"]]] = v * scale          # durations in us, bytes in MB
    return list(out.values())


def short(name):
    return re.sub(r"^void\s+", "", name).replace("b2c::", "").split("(")[0]


def launches(src, dst):
    L = read(src)
    names = [short(r["name"]) for r in L]
    # whole programs only: from the first stem that follows a head to the last head
    heads = [i for i, n in enumerate(names) if n.startswith("head_k7")]
    start = next(i for i in range(heads[0] + 1, len(names)) if names[i].startswith("stem_k7"))
    sel = L[start:heads[-1] + 1]
    n_prog = sum(1 for r in sel if short(r["name"]).startswith("head_k7"))
    by = collections.OrderedDict()
    for r in sel:
        k = by.setdefault(short(r["name"]), [0, 0.0])
        k[0] += 1
        k[1] += r["gpu__time_duration.sum"]
    tot = sum(v[1] for v in by.values())
    conv = sum(v[1] for k, v in by.items() if k.startswith("conv_"))
    json.dump({
        "command": "ncu --metrics gpu__time_duration.sum --clock-control none -s 810 -c 290 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
        "note": "cold-cache, serialised per-launch times (micro-batch 64); whole programs only: compare SHARES with "
                "bench.py's kernel_time_ms_per_program, not absolutes",
        "programs": n_prog, "launches": len(sel), "launches_per_program": len(sel) / max(n_prog, 1),
        "total_us": round(tot, 1), "conv_share": round(conv / tot, 4),
        "by_kernel": [{"kernel": k, "launches": v[0], "total_us": round(v[1], 1), "share": round(v[1] / tot, 4)}
                      for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])],
    }, open(dst, "w"), indent=1)
    print(dst, "programs", n_prog, "launches", len(sel), "conv share", round(conv / tot, 4))


def traffic(src, dst):
    L = read(src)
    fam = {"x3": [0, 0.0, 0.0], "bf16": [0, 0.0, 0.0]}
    rows = []
    for r in L:
        n = short(r["name"])
        x3 = "<1" in n          # conv_tc_kernel<1, 0>, conv_ru_kernel<1>, conv_tc2_kernel<1>
        f = fam["x3" if x3 else "bf16"]
        b = (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]) * 1e6
        f[0] += 1
        f[1] += r["gpu__time_duration.sum"] / 1e3
        f[2] += b
        rows.append({"name": n, "us": round(r["gpu__time_duration.sum"], 3), "dram_read_MB": round(r["dram__bytes_read.sum"], 3),
                     "dram_write_MB": round(r["dram__bytes_write.sum"], 3)})
    json.dump({
        "command": 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k '
                   'regex:"conv_tc_kernel|conv_ru_kernel|conv_tc2_kernel" -s 492 -c 82 python bench.py --steps 1 --warmup 3 --no-cpu-baseline',
        "note": "one program (micro-batch 64) = 82 conv launches; x3 = bf16x3 family (dominant kernel of bench.py roofline)",
        "x3": {"launches": fam["x3"][0], "time_ms": round(fam["x3"][1], 4), "dram_bytes": fam["x3"][2],
               "dram_bytes_per_launch": fam["x3"][2] / max(fam["x3"][0], 1)},
        "bf16": {"launches": fam["bf16"][0], "time_ms": round(fam["bf16"][1], 4), "dram_bytes": fam["bf16"][2]},
        "launches": rows,
    }, open(dst, "w"), indent=1)
    print(dst, {k: (v[0], round(v[1], 3), round(v[2] / 1e9, 2)) for k, v in fam.items()})


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
    if len(sys.argv) > 4:      # the command / note of THIS capture instead of the round-1 defaults
        d = json.load(open(sys.argv[3]))
        d["command"] = sys.argv[4]
        if len(sys.argv) > 5:
            d["note"] = sys.argv[5]
        json.dump(d, open(sys.argv[3], "w"), indent=1)

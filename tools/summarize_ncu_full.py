#!/usr/bin/env python
"""Key metrics of `ncu --set full` captures (.ncu-rep, read with `ncu -i ... --page raw --csv`) -> one JSON under profiles/.
    python tools/summarize_ncu_full.py out.json name=path.ncu-rep[:what] ..."""
import csv, io, json, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]

out = {}
for arg in sys.argv[2:]:
    name, rest = arg.split("=", 1)
    path, _, what = rest.partition(":")
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            m[k] = {"value": vals[i], "unit": units[i]}
    out[name] = {"what": what, "file": path.split("/")[-1], "metrics": m}
json.dump(out, open(sys.argv[1], "w"), indent=1)
for n, r in out.items():
    g = lambda k: r["metrics"].get(k, {}).get("value", "-")
    print(n, "us", g("gpu__time_duration.sum"), "tensor%", g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
          "dram%", g("dram__throughput.avg.pct_of_peak_sustained_elapsed"), "dramR/W MB", g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
          "issue%", g("sm__inst_issued.avg.pct_of_peak_sustained_elapsed"), "L2->SM", g("l1tex__m_xbar2l1tex_read_bytes.sum"))

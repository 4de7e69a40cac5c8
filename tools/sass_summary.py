#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell paths in libb2c.so (runs without a GPU):
counts of tcgen05 MMAs (UTC*MMA), TMA loads (UTMALDG) / stores (UTMASTG), TMEM loads (LDTM) / stores (STTM) and
the register / shared-memory footprint per kernel, from `cuobjdump -sass -res-usage`.

    python tools/sass_summary.py [--out profiles/r02_sass_summary.json]
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal_vqvae_compression_audio_tactile_b200", "libb2c.so")
PATTERNS = {"UTCMMA": r"\bUTC[A-Z0-9]*MMA\b", "UTMALDG": r"\bUTMALDG\b", "UTMASTG": r"\bUTMASTG\b", "LDTM": r"\bLDTM\b",
            "STTM": r"\bSTTM\b", "UTCBAR": r"\bUTCBAR\b", "SYNCS": r"\bSYNCS\b", "MUFU": r"\bMUFU\b", "FFMA": r"\bFFMA\b",
            "UBLKCP": r"\bUBLKCP\b", "REDUX": r"\bREDUX\b"}


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.split("\n")
        return [o if o else n for o, n in zip(out, names)]
    except Exception:
        return names


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--lib", default=LIB)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", args.lib], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = {k.lower(): int(v) for k, v in re.findall(r"(REG|SHARED|LOCAL|STACK):(\d+)", line)}
            cur = None
    kernels = {}
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in PATTERNS}
            kernels[cur]["instructions"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            kernels[cur]["instructions"] += 1
            for k, pat in PATTERNS.items():
                if re.search(pat, line):
                    kernels[cur][k] += 1
    names = list(kernels)
    pretty = demangle(names)
    rows = []
    for n, pn in zip(names, pretty):
        short = re.sub(r"\(.*", "", pn).replace("b2c::", "").replace("void ", "")
        rows.append(dict(kernel=short, **kernels[n], **usage.get(n, {})))
    rows.sort(key=lambda r: (-r["UTCMMA"], r["kernel"]))
    total = {k: sum(r[k] for r in rows) for k in PATTERNS}
    out = dict(library=os.path.relpath(args.lib, ROOT), arch="sm_100a", totals=total, n_kernels=len(rows), kernels=rows)
    txt = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(txt + "\n")
    print(json.dumps(dict(totals=total, n_kernels=len(rows))))
    for r in rows:
        if r["UTCMMA"] or r["UTMALDG"] or r["UTMASTG"] or r["LDTM"]:
            print(f"  {r['kernel'][:60]:60s} MMA {r['UTCMMA']:4d} TMALD {r['UTMALDG']:3d} TMAST {r['UTMASTG']:3d} LDTM {r['LDTM']:3d} "
                  f"STTM {r['STTM']:3d} regs {r.get('reg', '?')}")


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (captured with --import-source on / -lineinfo).

    python profiles/hotlines.py gpurun_out/prof.ncu-rep <kernel regex> <cubin> [min_pct] [mangled-name regex]

The mangled-name regex selects ONE instantiation in the cubin (e.g. tile_deposit_kernelILi1E); without it the
kernel regex is used, which is ambiguous for templates.

ncu's CSV source page is per SASS instruction; the line table comes from `nvdisasm -g` on the cubin the
report was taken from (cuobjdump -xelf all libpyl_b200.so).  Prints, per source line, the share of stall
samples and of executed warp instructions, plus the dominant stall reasons."""
import collections
import csv
import io
import re
import subprocess
import sys


def line_table(cubin, kernel_re):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, func, take = {}, None, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            take = re.search(kernel_re, m.group(1)) is not None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and take:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, kre, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    col = {n: i for i, n in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    table = line_table(cubin, sys.argv[5] if len(sys.argv) > 5 else kre)
    base = None
    agg = collections.defaultdict(lambda: collections.Counter())
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or not r[0].strip():
            continue
        try:
            addr = int(r[0], 16) if not r[0].isdigit() else int(r[0])
        except ValueError:
            continue
        if base is None:
            base = addr
        key = table.get(addr - base, ("?", 0))
        a = agg[key]
        a["samples"] += float(r[col["# Samples"]] or 0)
        a["inst"] += float(r[col["Instructions Executed"]] or 0)
        a["excess"] += float(r[col["L1 Wavefronts Shared Excessive"]] or 0) if "L1 Wavefronts Shared Excessive" in col else 0.0
        for s in stalls:
            a[s] += float(r[col[s]] or 0)
    ts = sum(a["samples"] for a in agg.values()) or 1
    ti = sum(a["inst"] for a in agg.values()) or 1
    print("kernel /%s/: %d samples, %.0f warp instructions" % (kre, ts, ti))
    print("%-28s %8s %8s %10s  top stalls" % ("file:line", "samples%", "inst%", "smem-excess"))
    for key, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        ps, pi = 100 * a["samples"] / ts, 100 * a["inst"] / ti
        if ps < min_pct and pi < min_pct:
            continue
        top = sorted(((a[s], s) for s in stalls), reverse=True)[:3]
        print("%-28s %8.1f %8.1f %10.0f  %s" % ("%s:%d" % key, ps, pi, a["excess"],
                                                 ", ".join("%s %.0f%%" % (s[6:], 100 * v / max(a["samples"], 1)) for v, s in top if v)))


if __name__ == "__main__":
    main()

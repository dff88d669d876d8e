#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into a small committed markdown summary.

    python profiles/summarize.py gpurun_out/prof_x.ncu-rep profiles/r1_x.md "title" [algorithmic_bytes]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global/L2 latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared memory)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LSU queue (lg_throttle)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# %s" % title, "", "Source: `%s` (ncu --set full --clock-control none, one launch)." % rep.split("/")[-1], ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append("## `%s`" % d.get("Kernel Name", "?")[:110])
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k, label in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append("| %s | %s %s |" % (label, d[k], u.get(k, "")))
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale.get(u["dram__bytes_read.sum"], 1) + wr * scale.get(u["dram__bytes_write.sum"], 1)
            lines.append("| DRAM traffic (read+write) | %.3f GB |" % (tot / 1e9))
            if alg:
                lines.append("| algorithmic bytes | %.3f GB (traffic / algorithmic = %.2f) |" % (alg / 1e9, tot / alg))
        except Exception:
            pass
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()

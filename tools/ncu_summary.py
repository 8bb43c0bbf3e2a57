#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per kernel duration, DRAM/L2 traffic, pipe
utilisation, occupancy, issue rate and top stall reasons.  Usage: tools/ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tex.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"## {d['Kernel Name'][:90]}  (id {d['ID']})")
        for w in WANT:
            if w in d and d[w] != "":
                out.append(f"- {w}: {d[w]} {units[hdr.index(w)]}")
        pipes = [(h, d[h]) for h in hdr if h.startswith("sm__inst_executed_pipe_") and h.endswith(".sum") and d[h] not in ("", "0")]
        out.append("- pipes: " + ", ".join(f"{h[len('sm__inst_executed_pipe_'):-4]}={float(v.replace(',', '')):.3g}" for h, v in pipes))
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] != "":
                stalls.append((float(d[h].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            elif h.startswith("smsp__average_warp_latency_issue_stalled_") and d[h] != "":
                stalls.append((float(d[h].replace(",", "")), h[len("smsp__average_warp_latency_issue_stalled_"):].split(".")[0]))
        stalls.sort(reverse=True)
        out.append("- top stalls (warp cycles per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
        out.append("")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()

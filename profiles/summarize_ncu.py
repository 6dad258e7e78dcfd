#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME [units_per_launch]

writes NAME_metrics.csv (selected raw-page metrics of every captured launch) and NAME_sass_mix.txt
(executed warp instructions per opcode and the share of stall samples they carry, from the source page).
units_per_launch (e.g. subapertures) turns instruction counts into per-unit counts.
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, stem = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    hdr, unit, rows = raw[0], raw[1], raw[2:]
    kn = hdr.index("Kernel Name")
    with open(stem + "_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d:%s" % (i, r[kn].split("(")[0]) for i, r in enumerate(rows)])
        for j, (h, u) in enumerate(zip(hdr, unit)):
            if h in KEEP:
                w.writerow([h, u] + [r[j] for r in rows])
    src = page(rep, "source")
    # one block per kernel: title row, header row, instruction rows
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    with open(stem + "_sass_mix.txt", "w") as f:
        for b in blocks[:1]:
            h = b["hdr"]
            iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
            agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
            for r in b["rows"]:
                tok = r[iS].split()
                if not tok:
                    continue
                op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
                a = agg[op]
                a[0] += 1
                a[1] += float(r[iE] or 0)
                a[2] += float(r[iN] or 0)
            tot = sum(a[1] for a in agg.values()) or 1.0
            ts = sum(a[2] for a in agg.values()) or 1.0
            f.write("kernel: %s\nexecuted warp instructions: %.0f" % (b["name"], tot))
            if units:
                f.write("  (%.1f per unit, %g units per launch)" % (tot / units, units))
            f.write("\n%-12s %7s %14s %8s %9s\n" % ("opcode", "static", "executed" + ("/unit" if units else ""), "share", "samples"))
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
                f.write("%-12s %7d %14.1f %8.3f %9.3f\n" % (k, a[0], a[1] / units if units else a[1], a[1] / tot, a[2] / ts))


if __name__ == "__main__":
    main()

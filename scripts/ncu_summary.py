"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed
under profiles/ (run here, no GPU needed).

  python scripts/ncu_summary.py launches <launches.csv> <out.txt>     # per-kernel time shares
  python scripts/ncu_summary.py full <report.ncu-rep> <out.txt>       # key metrics + hot SASS lines
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid/block (first)\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:60]:60s} {a[0]:8d} {a[1] / 1e3:12.1f} {a[1] / 1e3 / a[0]:10.2f} "
                    f"{a[1] / tot * 100:6.1f}%  {a[2]} {a[3]}\n")
    print(open(out).read())


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none --import-source on  ({rep})"]
    dram = None
    for r in rows[2:]:
        lines.append(f"\n== {r[hdr.index('Kernel Name')][:100]}  (ID {r[0]})")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"  {key:84s} {r[i]:>16s} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= mult.get(units[hdr.index("dram__bytes_read.sum")], 1)
            wr *= mult.get(units[hdr.index("dram__bytes_write.sum")], 1)
            dram = rd + wr
            lines.append(f"  dram bytes per launch (read+write) = {dram:.0f}")
        except Exception:
            pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        h0 = next(i for i, r in enumerate(srows) if r and r[0] == "Address")
        sh, srows = srows[h0], srows[h0:]
        try:
            ci = sh.index("Source")
            si = [i for i, h in enumerate(sh) if h.startswith("Warp Stall Sampling (All")][0]
            ei = sh.index("Instructions Executed")
            body = []
            for r in srows[1:]:
                if r and r[0] == "Address":
                    break  # next kernel in the report
                if len(r) > si and r[si].replace(".", "").isdigit():
                    body.append(r)
            tot = sum(float(r[si]) for r in body) or 1.0
            lines.append("\n== hottest SASS lines by warp-stall samples (first kernel in the report)")
            for r in sorted(body, key=lambda r: -float(r[si]))[:25]:
                lines.append(f"  {float(r[si]) / tot * 100:5.1f}%  exec={r[ei]:>10s}  {r[ci][:90]}")
        except Exception as exc:  # layout differs between ncu versions
            lines.append(f"(source page not parsed: {exc})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    return dram


if __name__ == "__main__":
    mode, src, out = sys.argv[1:4]
    if mode == "launches":
        launches(src, out)
    else:
        d = full(src, out)
        if len(sys.argv) > 4 and d is not None:
            json.dump({"dram_bytes_per_launch": d, "source": out}, open(sys.argv[4], "w"))

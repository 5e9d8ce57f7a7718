#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.
usage: scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("kernel: %s  grid %s block %s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEYS) or "warp_issue_stalled" in h and "pct" not in h and "not_issued" not in h:
                lines.append("  %-86s %s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    try:
        srows = list(csv.reader(io.StringIO(src)))
        # find header row containing 'Source' and a sampling column
        hi = next(i for i, r in enumerate(srows) if "Source" in r and any("Sampl" in c for c in r))
        sh = srows[hi]
        ci_src = sh.index("Source")
        ci_smp = next(i for i, c in enumerate(sh) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
        ci_ins = next((i for i, c in enumerate(sh) if c.startswith("Instructions Executed")), None)
        body = []
        for r in srows[hi + 1:]:
            try:
                body.append((float(r[ci_smp] or 0), r[ci_src][:110], r[ci_ins] if ci_ins is not None else ""))
            except (ValueError, IndexError):
                continue
        tot = sum(b[0] for b in body) or 1.0
        lines.append("")
        lines.append("top SASS/source lines by warp-stall samples (share of %d samples):" % tot)
        for smp, s, ins in sorted(body, reverse=True)[:45]:
            lines.append("  %6.2f%%  inst=%-10s %s" % (100 * smp / tot, ins, s))
    except Exception as e:  # source page layout differs between ncu versions
        lines.append("source page not parsed: %r" % (e,))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()

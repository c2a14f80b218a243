#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md / bench.py quote.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [samples_per_launch]
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    samples = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    src = page(rep, "source")
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for ki, r in enumerate(rows[2:]):
        if ki:
            print()
        lo = starts[ki] if ki < len(starts) else None
        hi = starts[ki + 1] if lo is not None and ki + 1 < len(starts) else len(src)
        one(dict(zip(hdr, r)), dict(zip(hdr, units)), src[lo:hi] if lo is not None else None, samples)


def one(d, u, src, samples):
    keys = [
        "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
    ]
    for k in keys:
        if k in d:
            print(f"{k:75s} {d[k]:>18s} {u.get(k, '')}")
    print("-- warp stalls per issued instruction")
    st = [(float(v), h) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:10]:
        print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
    if not src:
        return
    sh = src[1]
    iS, iE = sh.index("Source"), sh.index("Instructions Executed")
    ops = collections.Counter()
    tot = 0
    for row in src[2:]:
        s = row[iS].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        full = s.split()[0]
        op = full.split(".")[0]
        n = int(row[iE])
        tot += n
        ops[full if op in ("LDS", "STS", "LDG", "STG", "UTMALDG", "SYNCS") else op] += n
    print(f"-- executed warp instructions: {tot}" + (f"  = {tot * 32 / samples:.1f} thread-instructions per input sample" if samples else ""))
    for k, v in ops.most_common(24):
        print(f"   {k:24s} {v:12d} {100.0 * v / tot:5.1f}%" + (f"  {v * 32 / samples:6.2f}/sample" if samples else ""))


if __name__ == "__main__":
    main()

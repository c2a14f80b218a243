#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch) into the handful of numbers DESIGN.md / bench.py quote.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [samples_per_launch] [--record KERNEL:STREAMSxCHUNKS]

--record writes dram__bytes_read.sum + dram__bytes_write.sum of the first launch whose name contains KERNEL into
profiles/ncu_traffic.json together with the current commit (bench.py's roofline.traffic reads it from there).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(v):
    return float(str(v).replace(",", ""))


def record(rows, hdr, units, key):
    here = os.path.dirname(os.path.abspath(__file__))
    kernel = key.split(":")[0]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        if kernel in d.get("Kernel Name", ""):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = sum(num(d[k]) * scale.get(u.get(k, "byte"), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            path = os.path.join(here, "ncu_traffic.json")
            data = json.load(open(path)) if os.path.exists(path) else {}
            commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=here).stdout.strip()
            data[key] = {"dram_bytes_per_launch": tot, "kernel": d["Kernel Name"][:60], "duration_ns_under_ncu": num(d.get("gpu__time_duration.sum", 0)),
                         "commit": commit, "report": os.path.basename(sys.argv[1]), "source": "ncu --set full --clock-control none"}
            json.dump(data, open(path, "w"), indent=1, sort_keys=True)
            print(f"recorded {key}: {tot / 1e9:.4f} GB per launch at commit {commit}")
            return
    print(f"no launch of {kernel} in the report")


def main():
    rep = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    samples = float(args[0]) if args else None
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    if "--record" in sys.argv:
        record(rows, hdr, units, sys.argv[sys.argv.index("--record") + 1])
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
    # The source page's per-line counts are warp-level instruction executions summed over the launch; the launch total comes from
    # smsp__inst_executed.sum (the source page can list a line under several views, so its own sum is not a total)
    try:
        launch_tot = num(d["smsp__inst_executed.sum"])
        per_thread = num(d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 32))
    except Exception:
        launch_tot, per_thread = float(tot), 32.0
    scale = launch_tot / tot if tot else 1.0
    print(f"-- executed warp instructions: {launch_tot:.0f}" +
          (f"  = {launch_tot * per_thread / samples:.1f} thread-instructions per input sample" if samples else ""))
    for k, v in ops.most_common(24):
        print(f"   {k:24s} {v * scale:14.0f} {100.0 * v / tot:5.1f}%" + (f"  {v * scale * per_thread / samples:6.2f}/sample" if samples else ""))


if __name__ == "__main__":
    main()

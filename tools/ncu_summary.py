#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export into the handful of counters the design argues from.

    python tools/ncu_summary.py gpurun_out/prof_raw.csv > profiles/rN_<name>.txt
    python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/rN_launches.txt
"""
import csv
import io
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (top pipe)"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy (IMAD) pipe active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe inst %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe inst %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math-pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not-selected / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-scoreboard / issue"),
]


def main(path):
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "")
        key = (short, r[ix.get("launch__grid_size", 0)] if "launch__grid_size" in ix else "")
        if key in seen:
            continue
        if any("nan" in r[ix[k]] for k, _ in KEYS if k in ix):     # a replay that lost its counters: take the next launch
            continue
        seen[key] = 1
        print("== %s" % short)
        for k, label in KEYS:
            if k in ix:
                print("   %-36s %s %s" % (label, r[ix[k]], units[ix[k]]))
        print()


def launches(path):
    """Per-kernel totals and shares of a `--metrics gpu__time_duration.sum --csv` launch list."""
    txt = open(path).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.reader(io.StringIO(txt)))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = {}
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        unit = r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        a = agg.setdefault((name, r[ix["Grid Size"]], r[ix["Block Size"]]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("# kernel | grid | block | launches | mean us | total us | share")
    for (name, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s | %s | %s | %d | %.1f | %.1f | %.1f%%" % (name, g, b, n, t / n, t, 100 * t / tot))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        main(sys.argv[1])

"""Summarise an ncu --page raw --csv dump (one kernel launch per data row) into the handful of metrics the
roofline discussion needs.  Usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [regex ...]"""
import csv
import re
import sys

DEFAULT = [
    r"^gpu__time_duration\.sum$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^smsp__inst_executed\.sum$", r"^sm__inst_executed\.avg\.per_cycle_elapsed$",
    r"^smsp__issue_active\.avg\.pct", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^launch__registers_per_thread$", r"^launch__occupancy_limit", r"^launch__waves_per_multiprocessor$",
    r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|xu|fp64|lsu|uniform|adu|cbu)\.sum$",
    r"^sm__pipe_(fma|alu|xu|fp64|fmaheavy|fmalite)_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_.*pct_of_peak_sustained_active$",
    r"^smsp__thread_inst_executed_per_inst_executed\.ratio$",
    r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__average_warp_latency_issue_stalled.*ratio$",
    r"^sm__cycles_elapsed\.(avg|max)$", r"^smsp__cycles_active\.avg$", r"^launch__(grid|block)_size$",
    r"^smsp__warps_eligible\.avg\.per_cycle_active$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^smsp__inst_executed_op_shared.*sum$", r"^launch__shared_mem_per_block", r"^sm__maximum_warps_per_active_cycle_pct$",
]


def main():
    pats = [re.compile(p) for p in (sys.argv[1:] or DEFAULT)]
    rows = list(csv.reader(sys.stdin))
    hdr, units, data = rows[0], rows[1], rows[2:]
    strip = lambda h: h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1][:1].isupper() else h  # noqa: E731
    for r in data:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        seen = set()
        for i, h in enumerate(hdr):
            # headers look like "SECTION.Group.metric.name"; keep the metric part
            parts = h.split(".")
            metric = h
            for k in range(len(parts)):
                cand = ".".join(parts[k:])
                if re.match(r"^[a-z0-9_]+__", cand):
                    metric = cand
                    break
            if metric in seen:
                continue
            if any(p.search(metric) for p in pats):
                seen.add(metric)
                print(f"{metric:95s} {units[i]:>14s} {r[i]}")


if __name__ == "__main__":
    main()

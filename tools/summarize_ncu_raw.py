#!/usr/bin/env python3
"""One markdown table from `ncu --page raw --csv` exports (tools/r02_profile.sh): usage summarize_ncu_raw.py name=file.csv ..."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy (IMAD) pipe busy % of elapsed"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma pipe busy % of elapsed"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / 32"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
]


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    cols = []
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        cols.append((name, load(path)))
    print("| metric | " + " | ".join(n for n, _ in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for key, label in KEYS:
        cells = []
        for _, d in cols:
            v, u = d.get(key, ("-", ""))
            cells.append(("%s %s" % (v, u)).strip())
        print("| %s | %s |" % (label, " | ".join(cells)))
    # top stall reasons
    cells = []
    for _, d in cols:
        st = []
        for h, (v, u) in d.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(v.replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        cells.append(", ".join("%s %.2f" % (n, x) for x, n in st[:3] if n != "selected"))
    print("| top stalls (warps / issue) | %s |" % " | ".join(cells))
    for name, d in cols:
        print("\n`%s`: %s" % (name, d.get("Kernel Name", ("?", ""))[0][:150]))


if __name__ == "__main__":
    main()

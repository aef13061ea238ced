#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, on the CPU box) into a small tracked table.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/x.csv
    python profiles/ncu_summary.py /tmp/x.csv > profiles/X_summary.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active thr/inst"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if len(r) > 10]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(c) if c in hdr else None, n) for c, n in COLS]
    print("| # | kernel | " + " | ".join(n for _, n in idx) + " | DRAM GB/s |")
    print("|---|---|" + "---|" * (len(idx) + 1))
    for k, r in enumerate(data):
        name = r[ki].split("(")[0].replace("void ", "").replace("b200cd::", "").replace("<unnamed>::", "")
        cells = []
        vals = {}
        for i, n in idx:
            if i is None:
                cells.append("-")
                continue
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                vals[n] = (f, units[i])
                cells.append(f"{f:.4g} {units[i]}".strip())
            except ValueError:
                cells.append(v)
        gbs = "-"
        try:  # section captures carry the rate instead of the two byte counters: traffic = rate x duration
            i = hdr.index("dram__bytes.sum.per_second")
            rate = float(r[i].replace(",", "")) * {"byte/s": 1, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}[units[i]]
            t, u = vals["time"]
            secs = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[u]
            gbs = f"{rate / 1e9:.0f} ({rate * secs / 1e9:.3f} GB)"
        except Exception:
            pass
        try:
            def to_bytes(x):
                f, u = x
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            def to_s(x):
                f, u = x
                return f * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[u]
            gbs = f"{(to_bytes(vals['dram_rd']) + to_bytes(vals['dram_wr'])) / to_s(vals['time']) / 1e9:.0f}"
        except Exception:
            pass
        print(f"| {k} | {name} | " + " | ".join(cells) + f" | {gbs} |")


if __name__ == "__main__":
    main(sys.argv[1])

"""Summarise the per-kernel `ncu --set full` raw CSV exports (profiles/<dir>/*_raw.csv) as a markdown table."""
import csv, glob, os, sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def load_all(path):
    """one dict per captured launch"""
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:]]


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if u in ("ns", "nsecond"):
        return f"{x / 1e6:.3f} ms"
    if u in ("us", "usecond"):
        return f"{x / 1e3:.3f} ms"
    if u in ("ms", "msecond"):
        return f"{x:.3f} ms"
    if u.lower().startswith("gbyte"):
        return f"{x:.2f} GB"
    if u.lower().startswith("mbyte"):
        return f"{x:.1f} MB"
    if u.lower().startswith("kbyte"):
        return f"{x:.0f} KB"
    if u == "byte":
        return f"{x / 1e6:.1f} MB"
    return f"{x:.1f}" if u == "%" else f"{x:g}"


def main(d):
    print("| kernel | " + " | ".join(k[1] for k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for p in sorted(glob.glob(os.path.join(d, "*_raw.csv"))):
        launches = load_all(p)
        for i, m in enumerate(launches):
            cells = [fmt(*m[k]) if k in m else "-" for k, _ in KEYS]
            tag = os.path.basename(p)[:-8] + (f" #{i}" if len(launches) > 1 else "")
            print(f"| {tag} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

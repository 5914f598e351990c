#!/usr/bin/env python
"""One-line-per-kernel summary of an `ncu --set full` report (read with `ncu -i ... --page raw --csv`).
usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep [--md]"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path, md=False):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for name, short in METRICS:
        for i, h in enumerate(hdr):
            if h == name:
                idx[short] = i
    kn = hdr.index("Kernel Name")
    cols = [s for _, s in METRICS if s in idx]
    if md:
        print("| kernel | " + " | ".join(cols) + " |")
        print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[kn]).replace("setok::<unnamed>::", "").replace("void ", "")[:48]
        vals = []
        for c in cols:
            v, u = r[idx[c]], units[idx[c]]
            vals.append(f"{v} {u}".strip() if c in ("time", "dram_rd", "dram_wr") else v)
        if md:
            print(f"| `{name}` | " + " | ".join(vals) + " |")
        else:
            print(f"{name:48s} " + "  ".join(f"{c}={v}" for c, v in zip(cols, vals)))


if __name__ == "__main__":
    main(sys.argv[1], "--md" in sys.argv)

#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import subprocess
import sys


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def main(path, md=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val = val / 1e3 if unit == "ns" else val * 1e3 if unit == "ms" else val * 1e6 if unit == "s" else val
        seq.append((name, val))
        agg.setdefault(name, []).append(val)
    tot = sum(v for _, v in seq)
    print(f"{len(seq)} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised under ncu: compare shares)")
    if md:
        print("\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        d = demangle(k)
        d = re.sub(r"\(.*", "", d).replace("setok::(anonymous namespace)::", "").replace("void ", "")
        if md:
            print(f"| `{d[:70]}` | {len(v)} | {sum(v) / 1e3:.3f} | {sum(v) / tot:.3f} | {sum(v) / len(v):.1f} |")
        else:
            print(f"{d[:70]:70s} n={len(v):4d} sum={sum(v) / 1e3:9.3f} ms share={sum(v) / tot:6.3f} avg={sum(v) / len(v):9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], "--md" in sys.argv)

#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into the two files kept under profiles/:
`<out>_summary.csv` (per kernel: launches, total us, average us, share of all listed launches) and
`<out>_iteration.csv` (the launches of one PCG iteration of the elasticity solve, in order: everything between the
last two `k_cg_spmv<3,...>` launches but one).

usage: tools/summarize_launches.py gpurun_out/<list>.csv profiles/<prefix>"""
import collections
import csv
import sys


def short(name: str) -> str:
    name = name.replace("void ", "").replace("cb::", "")  # ncu prints the namespace only for some builds
    cut = name.find("(")
    return name if cut < 0 else name[:cut]


def main():
    src, out = sys.argv[1], sys.argv[2]
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        launches.append((short(row["Kernel Name"]), us, row.get("Grid Size", ""), row.get("Block Size", "")))
    agg = collections.OrderedDict()
    total = sum(t for _, t, _, _ in launches)
    for name, us, _, _ in launches:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    with open(out + "_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "own", "launches", "total_us", "avg_us", "share_of_listed_time"])
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, int(name.startswith("k_")), n, f"{t:.1f}", f"{t / n:.2f}", f"{t / total:.4f}"])
        w.writerow(["TOTAL", "", len(launches), f"{total:.1f}", "", "1.0"])
    idx = [i for i, l in enumerate(launches) if l[0].startswith("k_cg_spmv<3")]
    if len(idx) >= 3:
        a, b = idx[-3], idx[-2]
        with open(out + "_iteration.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["position", "kernel", "us", "grid", "block"])
            for k, (name, us, g, bl) in enumerate(launches[a:b]):
                w.writerow([k, name, f"{us:.2f}", g, bl])
            w.writerow(["", "ITERATION TOTAL", f"{sum(l[1] for l in launches[a:b]):.2f}", "", ""])
    own = sum(n for name, (n, t) in agg.items() if name.startswith("k_"))
    print(f"{len(launches)} launches ({own} own), {total / 1e3:.2f} ms listed -> {out}_summary.csv, {out}_iteration.csv")


if __name__ == "__main__":
    main()

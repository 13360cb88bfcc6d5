#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
usage: summarize_launches.py launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("stts::", "").replace("void ", "")


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), short(r["Kernel Name"]), ns))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)} (skipped first {skip}), total kernel time {total/1e6:.3f} ms (ncu: serialised, cold cache)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |")


if __name__ == "__main__":
    main()

"""Kernel time shares from an ncu launch list (--metrics gpu__time_duration.sum --csv).

usage: python tools/ncu_launch_shares.py gpurun_out/launches.csv "<note>" > profiles/<name>.txt"""
import csv
import re
import sys


def main(path, note=""):
    rows = [l for l in open(path) if l.startswith('"')]
    per = {}
    n = 0
    tot = 0.0
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        us = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        e = per.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += us
        n += 1
        tot += us
    print(f"launches {n} total us {tot:.1f} ({note})")
    for name, (k, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"  {us:8.1f} us {100 * us / tot:5.1f}% n={k:4d} avg {us / k:8.1f} {name}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

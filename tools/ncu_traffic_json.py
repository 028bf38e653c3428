"""Summarise an ncu CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) per kernel:
launches, DRAM bytes per launch, microseconds per launch under ncu.

usage: python tools/ncu_traffic_json.py gpurun_out/conv_traffic.csv > profiles/r01_ncu_conv_traffic.json"""
import csv
import json
import re
import sys


def main(path):
    rows = [l for l in open(path) if l.startswith('"')]
    per = {}
    for r in csv.DictReader(rows):
        name = re.sub(r"\(CUtensorMap_st.*$", "", r["Kernel Name"])      # drop the argument list
        name = re.sub(r"^void ", "", name)
        e = per.setdefault(name, {}).setdefault(r["ID"], {})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            e["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        elif m.startswith("dram__bytes"):
            e["bytes"] = e.get("bytes", 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    out = {}
    for name, launches in per.items():
        n = len(launches)
        out[name] = {"launches": n, "dram_bytes_per_launch": sum(l.get("bytes", 0.0) for l in launches.values()) / n,
                     "us_per_launch_under_ncu": sum(l.get("us", 0.0) for l in launches.values()) / n}
    json.dump(out, sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main(sys.argv[1])

"""Per-kernel stall summary of an ncu report (source page): top stall reasons, hottest SASS lines, and the samples
grouped by CUDA source line when the report carries -lineinfo.  usage: python tools/ncu_stalls.py REPORT [kernel index]"""
import csv, subprocess, sys, collections

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None:
        cur["rows"].append(row)
for bi, b in enumerate(blocks):
    if which is not None and bi != which:
        continue
    ix = {h: i for i, h in enumerate(b["hdr"])}
    stall = [h for h in b["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in b["rows"])
    print(f"== kernel {bi}: {b['name'][:60]}  samples {tot}  instructions {len(b['rows'])}")
    agg = {c: sum(int(r[ix[c]]) for r in b["rows"]) for c in stall}
    print("   " + "  ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(b["rows"], key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        st = sorted(((c[6:], int(r[ix[c]])) for c in stall if int(r[ix[c]]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"   {r[ix['# Samples']]:>6} x{r[ix['Instructions Executed']]:>9}  {r[ix['Source']].strip()[:64]:64s} {st}")

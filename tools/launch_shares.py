"""Kernel shares from an ncu launch list: python tools/launch_shares.py gpurun_out/launches.csv  (read here, no GPU)"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1], newline="")) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    agg[r[ki].split("(")[0][:100]].append(v / 1e3 if r[ui] in ("ns", "nsecond") else v)
tot = sum(sum(v) for v in agg.values())
print("kernel share of device time under ncu (cold, serialised; compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:100s} n={len(v):3d} avg_us={sum(v)/len(v):9.1f} share={100*sum(v)/tot:5.1f}%")

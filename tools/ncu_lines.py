"""Per-source-line and per-region instruction / stall-sample shares of one kernel from an ncu report (source page, cuda+sass)."""
import csv, collections, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = collections.Counter(); samp = collections.Counter(); src = {}; ops = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Function Name'): continue
    if r[0] != '' and len(r) > 7:
        line = (cur, int(r[0])); src[line] = r[1]
        try: per[line] += int(r[7]); samp[line] += int(r[6])
        except ValueError: pass
    elif len(r) > 7:
        try:
            t = r[3].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op.split('.')[0]] += int(r[7])
        except (ValueError, IndexError): pass
tot = sum(per.values()) or 1; tots = sum(samp.values()) or 1
print("instructions", tot, "samples", tots)
print("by samples:")
for (f, l), n in samp.most_common(top):
    print(f"  {f}:{l:5d} samp {n / tots * 100:5.1f}%  inst {per[(f, l)] / tot * 100:5.1f}%  {src[(f, l)][:120]}")
print("by instructions:")
for (f, l), n in per.most_common(top // 2):
    print(f"  {f}:{l:5d} inst {n / tot * 100:5.1f}%  samp {samp[(f, l)] / tots * 100:5.1f}%  {src[(f, l)][:120]}")
print("ops:", [(k, round(v / sum(ops.values()) * 100, 1)) for k, v in ops.most_common(16)])
if len(sys.argv) > 3:
    # regions: "name:lo-hi,name:lo-hi" over relay2.cuh line numbers
    print("regions:")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        n = sum(v for (f, l), v in per.items() if f == 'relay2.cuh' and lo <= l <= hi); sm = sum(v for (f, l), v in samp.items() if f == 'relay2.cuh' and lo <= l <= hi)
        print(f"  {name:28s} inst {n / tot * 100:5.1f}%  samples {sm / tots * 100:5.1f}%")
    oth = collections.Counter(); oths = collections.Counter()
    for (f, l), v in per.items():
        if f != 'relay2.cuh': oth[f] += v; oths[f] += samp[(f, l)]
    print("  other files:", {k: (round(v / tot * 100, 1), round(oths[k] / tots * 100, 1)) for k, v in oth.items() if v / tot > 0.002})

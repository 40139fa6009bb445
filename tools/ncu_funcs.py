"""Instruction / stall-sample shares per function of relay2.cuh (function starts read from the current source) for one ncu report."""
import csv, collections, subprocess, sys, re, bisect
rep = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "llmapigateway_b200/csrc/relay2.cuh"
base = srcfile.split("/")[-1]
starts = []
for i, t in enumerate(open(srcfile), 1):
    m = re.match(r'^(R2_DEV_NOINLINE|R2_DEV|R2_MEM|R2_GLOBAL|template|static inline|    R2_MEM)\b.*?\b(\w+)\s*\(', t)
    if m and not t.lstrip().startswith("//"):
        starts.append((i, m.group(2)))
    elif re.match(r'^k_\w+\(', t):
        starts.append((i, t.split("(")[0]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; per = collections.Counter(); samp = collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Line No', 'Function Name'): continue
    if r[0] != '' and len(r) > 7:
        try: per[(cur, int(r[0]))] += int(r[7]); samp[(cur, int(r[0]))] += int(r[6])
        except ValueError: pass
tot = sum(per.values()); tots = sum(samp.values())
ls = [s[0] for s in starts]
reg = collections.Counter(); regs = collections.Counter()
for (f, l), v in per.items():
    if f != base: name = '[' + f + ']'
    else:
        i = bisect.bisect_right(ls, l) - 1
        name = starts[i][1] if i >= 0 else '?'
    reg[name] += v; regs[name] += samp[(f, l)]
print("instructions", tot, "samples", tots)
for k, v in reg.most_common(30):
    print(f"{k:28s} inst {v / tot * 100:5.1f}% ({v / 1e6:6.2f} M)  samp {regs[k] / tots * 100:5.1f}%")

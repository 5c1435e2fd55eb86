#!/usr/bin/env python
"""Phase view of an ncu report of the person kernel: the SASS stream is cut at every CTA barrier / mbarrier wait and at
the back edge of every large loop, and for each segment the executed warp-instructions, the stall samples and the opcode
mix are printed (the source view attributes inlined functions to their own files, which hides the phase).
usage: python tools/ncu_phases.py report.ncu-rep [kernel-index]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernels, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
k = kernels[which]
h = k["hdr"]
iA, iS, iN, iX = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
ins = [(int(r[iA], 16), r[iS].strip(), int(r[iN] or 0), int(r[iX] or 0)) for r in k["rows"]]
base = ins[0][0]
tot_s = sum(i[2] for i in ins) or 1
tot_x = sum(i[3] for i in ins) or 1
print(k["name"][:80], " samples", tot_s, " warp-inst", tot_x)
# back edges of large loops
addr_index = {a: n for n, (a, _, _, _) in enumerate(ins)}
cuts = set()
for n, (a, t, s, x) in enumerate(ins):
    if re.search(r"\bBAR\.SYNC|SYNCS\.PHASECHK|WARPSYNC\.ALL", t):
        cuts.add(n + 1)
    m = re.search(r"BRA\S*\s+(?:\S+\s+)?0x([0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = base + int(m.group(1), 16) if int(m.group(1), 16) < 0x100000 else int(m.group(1), 16)
        if tgt in addr_index and tgt < a and n - addr_index[tgt] >= 60:
            cuts.add(addr_index[tgt]); cuts.add(n + 1)
cuts = sorted(c for c in cuts if 0 < c < len(ins))
segs, prev = [], 0
for c in cuts + [len(ins)]:
    if c > prev: segs.append((prev, c))
    prev = c
def op(t):
    p = t.split()
    o = p[1] if p[0].startswith("@") else p[0]
    o = o.split(".")[0] + (".WIDE" if "WIDE" in o else "")
    return o
print(f"{'addr':>7} {'len':>5} {'inst%':>6} {'samp%':>6}  mix")
for a, b in segs:
    s = sum(i[2] for i in ins[a:b]); x = sum(i[3] for i in ins[a:b])
    if s / tot_s < 0.004 and x / tot_x < 0.004: continue
    mix = {}
    for _, t, _, xx in ins[a:b]:
        mix[op(t)] = mix.get(op(t), 0) + xx
    top = ", ".join(f"{kk} {100*v/max(x,1):.0f}" for kk, v in sorted(mix.items(), key=lambda z: -z[1])[:7])
    print(f"{ins[a][0]-base:#7x} {b-a:5d} {100*x/tot_x:6.1f} {100*s/tot_s:6.1f}  {top}")

#!/usr/bin/env python
"""SASS opcode histogram per kernel of the shipped library (evidence for the Blackwell-native instruction mix:
UBLKCP / UBLKPF = 1-D TMA bulk copy / L2 prefetch, SYNCS = mbarrier, FFMA2 / FMUL2 / FADD2 = packed FP32 pipe).
usage: python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "extendedrtirtmodeling.jl_b200", "liberirt_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "UBLKPF", "UBLKRED", "SYNCS", "FFMA2", "FMUL2", "FADD2", "IMAD.WIDE", "MUFU", "UTMALDG", "UTMASTG", "UTCMMA", "HMMA", "RED", "ATOMS", "ATOMG",
         "ACQBULK", "LDS", "STS", "BAR"]
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({os.path.getsize(lib)} bytes)")
print("# columns: kernel, total instructions, then counts of the watched opcode families (prefix match)\n")
tot_watch = collections.Counter()
for blk in txt.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0].strip()
    ops = collections.Counter()
    n = 0
    for line in blk.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "IMAD.WIDE" and op.startswith("IMAD.WIDE")):
                ops[w] += 1
                break
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"erirt::", "", dem)[:110]
    tot_watch.update(ops)
    print(f"{dem}\n    {n} instr: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
print("\n# whole library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot_watch.items(), key=lambda kv: -kv[1])))
print("# no UTMALDG / UTCMMA / HMMA: the path has no dense contraction (tensor cores unused by design); TMA is the 1-D bulk form (UBLKCP).")

#!/usr/bin/env python
"""Static view of a kernel's SASS: every loop (backward branch) with its instruction count and opcode mix.
usage: python tools/sass_loops.py <lib.so> <kernel-name-substring> [min_len]"""
import re, subprocess, sys
lib, name = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 20
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
for b in blocks[1:]:
    fn = b.split("\n", 1)[0]
    if name not in fn:
        continue
    ins = []
    for l in b.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print("==", fn, len(ins), "instructions")
    def op(t):
        parts = t.split()
        o = parts[1] if parts[0].startswith("@") else parts[0]
        return o
    for i, (a, t) in enumerate(ins):
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < a and tgt in addr and i - addr[tgt] + 1 >= minlen:
                    body = ins[addr[tgt]:i + 1]
                    ops = {}
                    for _, tt in body:
                        o = op(tt).split(".")[0]
                        if o == "IMAD" and "WIDE" in op(tt): o = "IMAD.WIDE"
                        if o == "MUFU": o = op(tt)
                        ops[o] = ops.get(o, 0) + 1
                    top = sorted(ops.items(), key=lambda x: -x[1])
                    print(f"{tgt:#x}-{a:#x} len {len(body)}: " + ", ".join(f"{k} {v}" for k, v in top[:22]))

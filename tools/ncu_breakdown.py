#!/usr/bin/env python
"""Summarise an ncu report of the person kernel: key raw metrics, stall mix, instructions per source line group.
usage: python tools/ncu_breakdown.py gpurun_out/prof_person_X.ncu-rep [top_lines]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "launch__grid_size",
        "sm__inst_executed_pipe_xu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    print("--- kernel", r[hdr.index("Kernel Name")][:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]} {units[i]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
stall = collections.Counter()
tot = tots = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix_inst, ix_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        scols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0] != "" and hdr:
        try:
            line = int(r[0])
        except ValueError:
            continue
        off = len(r) - len(hdr)
        try:
            inst, samp = int(r[ix_inst + off]), int(r[ix_samp + off])
        except ValueError:
            continue
        a = agg[(cur, line)]
        a[0] += inst; a[1] += samp; a[2] = ",".join(r[1:2 + off])[:80]
        tot += inst; tots += samp
        for i, h in scols:
            try:
                stall[h] += int(r[i + off])
            except ValueError:
                pass
print(f"total warp instructions {tot:.3e}  samples {tots}")
s = sum(stall.values())
print("stalls:", ", ".join(f"{k[6:]} {100*v/s:.1f}%" for k, v in stall.most_common(9)))
byfile = collections.Counter()
for (f, l), v in agg.items():
    byfile[f] += v[0]
print("by file:", ", ".join(f"{f} {100*v/tot:.1f}%" for f, v in byfile.most_common(8)))
for (f, l), (inst, samp, txt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{f}:{l:4d} inst {100*inst/tot:5.1f}% samp {100*samp/tots:5.1f}%  {txt}")

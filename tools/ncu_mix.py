#!/usr/bin/env python
"""Executed-instruction mix by opcode (and by contiguous SASS region) of the first kernel in an .ncu-rep: tools/ncu_mix.py rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
IE = ix["Instructions Executed"]; SRC = ix["Source"]; S = ix["Warp Stall Sampling (All Samples)"]
tot = sum(int(r[IE] or 0) for r in body)
mix = collections.Counter(); smp = collections.Counter()
for r in body:
    t = r[SRC].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    mix[op] += int(r[IE] or 0); smp[op] += int(r[S] or 0)
print("total warp-inst executed:", tot)
stot = sum(smp.values())
for op, c in mix.most_common(40):
    print("%-10s %12d %6.2f%%   samples %6.2f%%" % (op, c, 100.0 * c / tot, 100.0 * smp[op] / max(stot, 1)))
fp64 = sum(c for op, c in mix.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print("FP64-pipe ops (DFMA DMUL DADD DSETP): %.2f%%" % (100.0 * sum(c for op, c in mix.items() if op in ("DFMA", "DMUL", "DADD", "DSETP")) / tot))
# regions: split the listing where the execution count changes by more than 2x
print("-- regions (start index, instructions, executions per instruction, share of executed) --")
start = 0
def flush(a, b):
    n = b - a; ex = sum(int(r[IE] or 0) for r in body[a:b])
    if ex * 100.0 / tot > 1.0: print("  [%5d..%5d) %5d instr  %12.0f exec/instr  %6.2f%%" % (a, b, n, ex / max(n, 1), 100.0 * ex / tot))
prev = None
for i, r in enumerate(body):
    c = int(r[IE] or 0)
    if prev is not None and (c > 1.3 * prev + 10 or c < prev / 1.3 - 10):
        flush(start, i); start = i
    prev = c
flush(start, len(body))

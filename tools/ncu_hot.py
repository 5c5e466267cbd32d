#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an .ncu-rep (needs --import-source on): tools/ncu_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
S = ix["Warp Stall Sampling (All Samples)"]; IE = ix["Instructions Executed"]; SRC = ix["Source"]
tot = sum(int(r[S] or 0) for r in body)
print("instructions:", len(body), "total samples:", tot, "warp-inst executed:", sum(int(r[IE] or 0) for r in body))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for pos, r in sorted(enumerate(body), key=lambda x: -int(x[1][S] or 0))[:N]:
    top = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %6.2f%% %9s  %-60s %s" % (pos, 100.0 * int(r[S] or 0) / max(tot, 1), r[IE], r[SRC].strip()[:60], " ".join("%s:%d" % (c, v) for v, c in top)))

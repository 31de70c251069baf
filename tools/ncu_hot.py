#!/usr/bin/env python
"""Hottest SASS lines (by warp-stall samples) of one kernel of an .ncu-rep.
   python tools/ncu_hot.py rep.ncu-rep <kernel-index-1-based> [N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot, ts = sum(int(r[ia]) for r in data), sum(int(r[isamp]) for r in data)
print("kernel:", rows[0][1][:100] if rows and len(rows[0]) > 1 else "")
print("total warp-inst", tot, "samples", ts, "sass lines", len(data))
top = sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:n]
for i, r in sorted(top):
    print(f"{i:5d} {r[isrc].strip()[:76]:76s} inst {int(r[ia]):>10d} samp {int(r[isamp]):>6d} ({100*int(r[isamp])/ts:4.1f}%)")

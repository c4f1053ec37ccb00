#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur, hdr = None, None
inst, samp, src, stall = collections.Counter(), collections.Counter(), {}, collections.Counter()
line_stall = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        scols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and len(r) > ie and r[0] != "":
        try:
            key = (cur, int(r[0]))
            inst[key] += int(r[ie])
            samp[key] += int(r[isamp])
            src[key] = r[1]
            for i, h in scols:
                stall[h] += int(r[i])
                line_stall[key][h[6:]] += int(r[i])
        except ValueError:
            pass
ti, ts, tst = sum(inst.values()), sum(samp.values()), sum(stall.values())
print(f"total warp-instructions {ti}  samples {ts}")
print("stalls: " + "  ".join(f"{k[6:]}={v / tst:.3f}" for k, v in stall.most_common(9)))
for key, c in inst.most_common(top):
    top = ",".join(f"{k}:{v}" for k, v in line_stall[key].most_common(2) if v)
    print(f"{key[0]:16s}:{key[1]:4d} inst={c / ti:6.3f} samp={samp[key] / ts:6.3f} [{top}]  {src[key].strip()[:95]}")

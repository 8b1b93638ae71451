#!/usr/bin/env python
"""Warp-stall breakdown of an .ncu-rep captured with --set full --import-source on (read here, no GPU needed):
per kernel block, the share of each stall reason and the SASS rows that collect the most samples.
usage: tools/ncu_stalls.py report.ncu-rep [block-index ...] [top=N]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
blocks = [int(a) for a in sys.argv[2:] if a.isdigit()]
top = next((int(a[4:]) for a in sys.argv[2:] if a.startswith("top=")), 18)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for bi, s in enumerate(starts):
    if blocks and bi not in blocks:
        continue
    e = starts[bi + 1] if bi + 1 < len(starts) else len(rows)
    hdr, data = rows[s + 1], rows[s + 2:e]
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for i, h in stall:
            tot[h] += int(r[i] or 0)
    T = sum(tot.values()) or 1
    print(f"== block {bi}: {rows[s][1]}  ({len(data)} SASS rows, {T} samples)")
    print("   " + "  ".join(f"{h[6:]} {100 * v / T:.1f}%" for h, v in tot.most_common(8)))
    si = hdr.index("# Samples")
    S = sum(int(r[si] or 0) for r in data) or 1
    for r in sorted(data, key=lambda r: -int(r[si] or 0))[:top]:
        reasons = sorted([(int(r[i] or 0), h[6:]) for i, h in stall], reverse=True)[:2]
        print(f"   {100 * int(r[si]) / S:5.1f}%  {r[1].strip()[:58]:58s} {reasons}")

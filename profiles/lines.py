#!/usr/bin/env python
"""Per-source-line share of stall samples / executed instructions / shared-memory wavefronts for one kernel.
    python profiles/lines.py gpurun_out/x.ncu-rep k3_resample [min_pct]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
iW, iWI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
agg = {}
for r in rows:
    if len(r) > iW and r[0].isdigit() and r[2] == "-":
        try:
            v = (int(r[iS]), int(r[iI]), int(r[iW]), int(r[iWI]))
        except ValueError:
            continue
        a = agg.get(int(r[0]), (0, 0, 0, 0, r[1]))
        agg[int(r[0])] = (a[0] + v[0], a[1] + v[1], a[2] + v[2], a[3] + v[3], r[1])
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"# {kern}: {ts} samples, {ti} warp instructions")
for ln in sorted(agg):
    s, i, w, wi, src = agg[ln]
    if s > ts * thr / 100 or i > ti * thr / 100:
        print(f"{ln:5d} samp {100*s/ts:5.1f}% inst {100*i/ti:5.1f}% smem_wf {w:>9d} (ideal {wi:>9d})  {src.strip()[:90]}")

#!/usr/bin/env python
"""Hot CUDA source lines of one kernel of an ncu report (needs -lineinfo and --import-source on).
   python tools/ncu_hot_lines.py <report.ncu-rep> <kernel id> [top N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    try: lines.append((fname, int(r[0]), r[1].strip(), int(r[ii]), int(r[ti]), int(r[si])))
    except ValueError: pass
tot = sum(l[3] for l in lines); ts = sum(l[5] for l in lines)
print(func, "| warp instructions", tot, "| samples", ts)
for l in sorted(lines, key=lambda l: -l[5])[:top]:
    print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s" % (100.0 * l[5] / max(ts, 1), 100.0 * l[3] / max(tot, 1), l[4] / max(l[3], 1), l[0], l[1], l[2][:100]))

#!/usr/bin/env python3
"""Per-source-line summary of an ncu report's source page: python tools/ncu_lines.py report.ncu-rep [top]
(ncu -i report --page source --csv --print-source sass,cuda must work, i.e. the kernel was built with -lineinfo)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows, fname, hdr = [], None, None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try: return int(d.get(k, "0"))
            except ValueError: return 0
        rows.append((fname, int(r[0]), r[1].strip()[:110], num("# Samples"), num("Instructions Executed"), num("stall_wait"), num("stall_short_sb"), num("stall_long_sb"), num("L2 Theoretical Sectors Local")))
ts = sum(x[3] for x in rows) or 1; ti = sum(x[4] for x in rows) or 1
print(f"total samples {ts}, total warp instructions {ti}")
print("by samples:")
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{x[0]}:{x[1]:5d} smp {100*x[3]/ts:5.1f}% inst {100*x[4]/ti:5.1f}% wait {x[5]:6d} ssb {x[6]:6d} lsb {x[7]:6d} loc {x[8]:8d} | {x[2]}")

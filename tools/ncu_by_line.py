#!/usr/bin/env python3
"""Executed warp instructions per CUDA source line of an .ncu-rep captured with --import-source on (kernels built with -lineinfo).
usage: tools/ncu_by_line.py rep [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fpath, hdr, lines, total = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iW, iWi = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        continue
    if hdr is None or len(r) <= iI:
        continue
    if r[0].isdigit():                                    # a CUDA source line: the tool already sums its SASS
        def num(v):
            try:
                return int(v)
            except ValueError:
                return 0
        n = num(r[iI])
        lines.append((n, fpath, int(r[0]), r[1].strip()[:110], num(r[iW]), num(r[iWi])))
        total += n
lines.sort(reverse=True)
print(f"total warp instructions attributed to source lines: {total}")
for n, f, ln, src, wf, wfi in lines[:top]:
    print(f"{n:12d} {100 * n / total:5.1f}%  wf {wf:10d}/{wfi:10d}  {f}:{ln:<4d} {src}")

#!/usr/bin/env python3
"""SASS evidence per kernel of csrc/librv_b200.so (runs here, no GPU): counts of the mnemonics that prove the Blackwell-native
paths -- UTMALDG (TMA tensor load), UTMAPF (TMA L2 prefetch), SYNCS (mbarrier), VIMNMX / VIMNMX3 (packed u16x2 min/max), HFMA2 /
HADD2 (compare-exchanges on the FMA pipe), IDP (byte dot products for the luminance), ATOMS (shared-memory histogram atomics),
SHFL (warp scans of the LUT pass), LDS / STS, plus registers and the kernel-source hash the library was built from.
usage: tools/sass_markers.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvb200  # noqa: E402

MARK = ["UTMALDG", "UTMAPF", "SYNCS", "VIMNMX", "VIMNMX3", "HFMA2", "HADD2", "HMNMX2", "IDP", "ATOMS", "ATOMG", "REDG", "SHFL", "LDS", "STS",
        "LDG", "STG", "FFMA", "FMUL", "FADD", "IMAD", "PRMT", "BAR", "MATCH"]


def main():
    so = rvb200.library_path()
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    print(f"# {os.path.relpath(so, ROOT)}  kernel source hash {rvb200.kernel_source_hash()}  ({rvb200._native.load_library().rv_version().decode()})", file=out)
    print("# cuobjdump -sass | per-kernel mnemonic counts (static instructions, not executed counts)", file=out)
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for (name, cnt), pretty in zip(kernels.items(), demangle):
        pretty = re.sub(r"\(.*", "", pretty).replace("void rv::", "")
        r = regs.get(name, ("?", "?"))
        marks = " ".join(f"{k}={cnt[k]}" for k in MARK if cnt[k])
        print(f"{pretty:34s} instr={cnt['_total']:6d} regs={r[0]} static_smem={r[1]}  {marks}", file=out)


if __name__ == "__main__":
    main()

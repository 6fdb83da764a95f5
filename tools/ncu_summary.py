#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + SASS instruction mix per phase.
usage: tools/ncu_summary.py gpurun_out/prof_chain_TAG.ncu-rep [out.txt] [--json profiles/final_k_chain.json]

--json also writes the per-launch constants bench.py reports (DRAM bytes, executed warp instructions, duration) together with
the hash of the kernel sources they were measured on (rvb200.kernel_source_hash()) and the hash of the measured kernel's SASS
(rvb200.kernel_sass_hash()); bench.py and tests/test_host_logic.py refuse constants when neither the sources nor the kernel's
machine code in the built library are the ones the capture ran on."""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max']


def ncu(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def write_json(path, rep, hdr, d):
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rvb200

    def num(name):
        return float(d[hdr.index(name)].replace(',', ''))
    unit = {h: u for h, u in zip(hdr, UNITS)}

    def to_bytes(name):
        u = unit[name].lower()
        return num(name) * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]

    def to_us(name):
        u = unit[name].lower()
        return num(name) * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}[u]
    kname = d[hdr.index('Kernel Name')]
    if SASS:                                        # per-kernel SASS hashes recorded on the GPU box when the capture was taken
        import re
        norm = lambda n: re.sub(r"^void\s+", "", n).split("(")[0].replace(" ", "")
        sass = {norm(k): v for k, v in json.load(open(SASS[0])).items()}[norm(kname)]
    else:
        sass = rvb200.kernel_sass_hash(kname)
    j = {"kernel": kname, "capture": os.path.basename(rep), "source_hash": HASH[0] if HASH else rvb200.kernel_source_hash(),
         "sass_hash": sass,
         "grid": d[hdr.index('Grid Size')] if 'Grid Size' in hdr else None,
         "dram_bytes_read": to_bytes('dram__bytes_read.sum'), "dram_bytes_write": to_bytes('dram__bytes_write.sum'),
         "warp_instructions": num('smsp__inst_executed.sum'), "duration_us_under_ncu": to_us('gpu__time_duration.sum'),
         "issue_active_pct": num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
         "pipe_alu_pct": num('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
         "pipe_fma_pct": num('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
         "shared_wavefronts": num('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
         "shared_bank_conflicts": num('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
         "note": "one launch over 64 x 1080p frames; ncu --set full --clock-control none (cold cache, serialised)"}
    with open(path, 'w') as fh:
        json.dump(j, fh, indent=1)
        fh.write("\n")


UNITS = []
HASH = []
SASS = []


def main():
    args = list(sys.argv[1:])
    jpath = None
    if '--hash-file' in args:                       # hash recorded on the GPU box when the capture was taken
        i = args.index('--hash-file')
        HASH[:] = [open(args[i + 1]).read().strip()]
        del args[i:i + 2]
    if '--sass-file' in args:                       # {kernel: SASS hash} of the library the capture ran on (rvb200.kernel_sass_hashes())
        i = args.index('--sass-file')
        SASS[:] = [args[i + 1]]
        del args[i:i + 2]
    if '--json' in args:
        i = args.index('--json')
        jpath = args[i + 1]
        del args[i:i + 2]
    rep = args[0]
    out = open(args[1], 'w') if len(args) > 1 else sys.stdout
    rows = ncu(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    d = data[0]
    UNITS[:] = units
    if jpath:
        write_json(jpath, rep, hdr, d)
    print(f"# {rep}\nkernel: {d[hdr.index('Kernel Name')]}", file=out)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:72s} {d[i]:>16s} {units[i]}", file=out)
    stalls = [(hdr[i], float(d[i])) for i in range(len(hdr)) if hdr[i].startswith('smsp__average_warps_issue_stalled') and hdr[i].endswith('_per_issue_active.ratio') and d[i]]
    if not stalls:
        stalls = [(hdr[i], float(d[i])) for i in range(len(hdr)) if 'issue_stalled' in hdr[i] and hdr[i].endswith('.pct') and d[i]]
    print("\nwarp stall reasons (top):", file=out)
    for k, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
        print(f"  {k:90s} {v:10.3f}", file=out)
    rows = ncu(rep, 'source')
    hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if not hidx:
        return
    h = rows[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(rows)
    body = rows[hidx[0] + 1:end]
    iS, iI, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    tot, byop, samp, acc, phases = 0, collections.Counter(), collections.Counter(), 0, []
    ph_inst, ph_samp = [0], [0]          # per phase (code between two BAR.SYNCs): instructions and warp-state samples
    for r in body:
        if len(r) <= iI or not r[iI]:
            continue
        n = int(r[iI]); tot += n
        toks = r[iS].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        byop[op] += n; samp[op] += int(r[iSm] or 0)
        acc += n
        ph_inst[-1] += n; ph_samp[-1] += int(r[iSm] or 0)
        if op == 'BAR':
            phases.append(acc)
            ph_inst.append(0); ph_samp.append(0)
    print(f"\nwarp instructions executed: {tot}", file=out)
    print("cumulative share at each BAR.SYNC: " + ", ".join(f"{100 * x / tot:.1f}%" for x in phases), file=out)
    ts = max(sum(ph_samp), 1)
    print("per phase, share of instructions / share of sampled warp time (a phase whose time share exceeds its instruction "
          "share is waiting, not issuing): " + ", ".join(f"{100 * a / tot:.1f}% / {100 * b / ts:.1f}%" for a, b in zip(ph_inst, ph_samp)), file=out)
    for op, n in byop.most_common(22):
        print(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%   stall samples {samp[op]}", file=out)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (run here, no GPU needed):
   python scripts/summarize_ncu.py gpurun_out/<tag>_prof.ncu-rep profiles/<tag>_ncu_summary.txt
Writes the key raw metrics of the fused kernel, its SASS opcode mix and the shared-memory
instructions with the most excess wavefronts (bank conflicts); also profiles/latest_traffic.json
(dram bytes per launch, read by bench.py's `roofline.traffic`)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
lines = ["ncu --set full --clock-control none summary of %s" % os.path.basename(rep), ""]
kn = hdr.index('Kernel Name')
units = rows[1]
traffic = None      # dram bytes per launch of the DOMINANT kernel (longest gpu__time_duration)
best_dur = -1.0
for r in rows[2:]:
    lines.append("kernel: %s" % r[kn])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("  %-62s %s %s" % (k, r[i], units[i]))
    try:
        rd, wr = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('dram__bytes_write.sum')])
        mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[hdr.index('dram__bytes_read.sum')]]
        lines.append("  dram bytes read+write per launch: %.1f MB" % ((rd + wr) * mult / 1e6))
        dur = float(r[hdr.index('gpu__time_duration.sum')])
        if dur > best_dur:
            best_dur, traffic, traffic_kernel = dur, (rd + wr) * mult, r[kn]
    except Exception:
        pass
    lines.append("")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(srows):
    if 'Source' in r and 'Instructions Executed' in r:
        h = i
        break
if h is not None:
    sh = srows[h]
    iS, iI = sh.index('Source'), sh.index('Instructions Executed')
    iW, iWi = sh.index('L1 Wavefronts Shared'), sh.index('L1 Wavefronts Shared Ideal')
    ops, tot, sm = collections.Counter(), 0, []
    for r in srows[h + 1:]:
        try:
            n = int(r[iI])
        except Exception:
            continue
        s = r[iS].strip()
        m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_.]+)', s)
        op = m.group(2).split('.')[0] if m else s[:10]
        ops[op] += n
        tot += n
        w, wi = int(r[iW] or 0), int(r[iWi] or 0)
        if w:
            sm.append((w - wi, w, wi, n, s))
    lines.append("SASS opcode mix (warp-level instructions executed, all captured launches): total %d" % tot)
    for op, n in ops.most_common(16):
        lines.append("  %-8s %12d  %5.1f%%" % (op, n, 100. * n / max(tot, 1)))
    lines.append("")
    lines.append("shared-memory wavefronts: %d, ideal %d" % (sum(x[1] for x in sm), sum(x[2] for x in sm)))
    lines.append("top shared-memory instructions by excess wavefronts (excess, wavefronts, ideal, executed, SASS):")
    sm.sort(reverse=True)
    for x in sm[:12]:
        lines.append("  %10d %10d %10d %9d  %s" % x)
open(out, 'w').write("\n".join(lines) + "\n")
if traffic:
    with open(os.path.join(os.path.dirname(out), 'latest_traffic.json'), 'w') as f:
        import hashlib
        h = hashlib.sha256()
        csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-soxr_b200", "csrc")
        for name in ("fdr_stage.cuh", "fft_regs.cuh", "fdr_stage.cu"):   # what bench.py hashes: run this right after the capture
            with open(os.path.join(csrc, name), "rb") as g:
                h.update(g.read())
        json.dump({"dram_bytes_per_launch": traffic, "kernel": traffic_kernel[:80], "source": os.path.basename(out),
                   "capture": os.path.basename(rep), "kernel_source_sha": h.hexdigest()[:16]}, f)
print("\n".join(lines[:40]))

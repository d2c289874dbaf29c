#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu --set full capture: python scripts/sass_hot.py rep [kernel-index]
Prints the SASS listing of the first (or given) kernel with executed counts and stall samples, to a file."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
hdr = rows[1]
ia, isrc, ist, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
lines = []
tot_s = tot_e = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Kernel Name' or r[0] == 'Address':
        if r and r[0] == 'Kernel Name': break
        continue
    s, e = int(r[ist] or 0), int(r[iex] or 0)
    tot_s += s; tot_e += e
    lines.append((r[ia], r[isrc].strip(), s, e))
with open(out, 'w') as f:
    f.write("total stall samples %d, instructions executed %d\n" % (tot_s, tot_e))
    for i, (a, src, s, e) in enumerate(lines):
        f.write("%5d %6d %9d  %s\n" % (i, s, e, src))
print(len(lines), tot_s, tot_e)

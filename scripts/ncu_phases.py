#!/usr/bin/env python
"""Per-phase view of an ncu --set full capture: SASS rows of the first k_fdr kernel, segmented at BAR.SYNC,
with executed instructions, stall samples by reason and shared/global wavefronts per segment.
usage: scripts/ncu_phases.py gpurun_out/<tag>_prof.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else 'k_fdr<'
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i = 0
while i < len(rows) and not (rows[i] and rows[i][0] == 'Kernel Name' and pat in rows[i][1]): i += 1
hd = rows[i + 1]; body = []
for r in rows[i + 2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) == len(hd): body.append(r)
col = {n: k for k, n in enumerate(hd)}
stalls = [n for n in hd if n.startswith('stall_') and 'Not Issued' not in n]
seg = []; cur = None
def new(): return dict(n=0, inst=0, samp=0, sh=0, ops={}, **{s: 0 for s in stalls})
cur = new()
for r in body:
    s = r[col['Source']].strip(); op = s.split()[0] if not s.startswith('@') else s.split()[1]
    op = op.split('.')[0]
    ex = int(r[col['Instructions Executed']] or 0)
    cur['n'] += 1; cur['inst'] += ex; cur['samp'] += int(r[col['# Samples']] or 0)
    cur['sh'] += int(r[col['L1 Wavefronts Shared']] or 0)
    cur['ops'][op] = cur['ops'].get(op, 0) + ex
    for st in stalls: cur[st] += int(r[col[st]] or 0)
    if op == 'BAR': seg.append(cur); cur = new()
seg.append(cur)
tot = sum(s['samp'] for s in seg); toti = sum(s['inst'] for s in seg)
print("segments between BAR.SYNC (static order); samples total %d, warp instructions %d" % (tot, toti))
for k, s in enumerate(seg):
    if s['samp'] < tot * 0.003 and s['inst'] < toti * 0.003: continue
    top = sorted(((s[st], st[6:]) for st in stalls), reverse=True)[:6]
    fp = sum(v for o, v in s['ops'].items() if o in ('FADD2', 'FFMA2', 'FMUL2', 'FADD', 'FMUL', 'FFMA'))
    ld = sum(v for o, v in s['ops'].items() if o in ('LDS', 'STS', 'LDG', 'STG'))
    print("seg %2d: %4d sass, inst %5.1f%% (fp %4.1f%% mem %4.1f%%), samples %5.1f%%, sh-wavefronts %5.1f%% | %s" % (
        k, s['n'], 100.0 * s['inst'] / toti, 100.0 * fp / toti, 100.0 * ld / toti, 100.0 * s['samp'] / tot,
        100.0 * s['sh'] / max(1, sum(x['sh'] for x in seg)), ' '.join('%s %.1f' % (n, 100.0 * v / tot) for v, n in top)))

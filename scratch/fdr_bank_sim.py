# Bank-conflict check for the fdr kernel's exchange layouts (8-byte elements: a half-warp of 16 lanes
# must hit 16 distinct 8-byte bank pairs).  Layouts: E1[S1*j + g], g = rc*m' + m;  E2[S2*m + t2], t2 = j + ra*q2.
import itertools, sys
def conflicts(addr_lists):
    worst = 1
    for lanes in addr_lists:
        for h in range(0, len(lanes), 16):
            banks = {}
            for a in lanes[h:h+16]:
                if a is None: continue
                banks.setdefault(a % 16, set()).add(a)
            if banks: worst = max(worst, max(len(v) for v in banks.values()))
    return worst
def p2map(mode, tau, ra, rc):
    if mode == 'jfast':
        j, m = tau % ra, tau // ra
    elif mode == 'row16':      # j = tau>>4, m = tau&15
        j, m = tau >> 4, tau & 15
    elif mode == 'mfast':
        m, j = tau % rc, tau // rc
    elif mode == 'row16m':     # m = tau>>4, j = tau&15
        m, j = tau >> 4, tau & 15
    if j >= ra or m >= rc: return None
    return j, m
def check(ra, rb, rc, T, mode, S1, S2):
    G1 = rb*rc; T3 = ra*rb
    res = {}
    ins = []
    for u in range((G1 + T - 1)//T):
        for j in range(ra):
            ins.append([S1*j + (tau + T*u) if tau + T*u < G1 else None for tau in range(T)])
    res['p1-E1'] = conflicts(ins)
    nt2 = 16*ra if mode == 'row16' else 16*rc if mode == 'row16m' else ra*rc
    ins = []
    for mp in range(rb):
        lanes = []
        for tau in range(nt2):
            jm = p2map(mode, tau, ra, rc)
            lanes.append(None if jm is None else S1*jm[0] + rc*mp + jm[1])
        ins.append(lanes)
    res['p2-E1'] = conflicts(ins)
    ins = []
    for q2 in range(rb):
        lanes = []
        for tau in range(nt2):
            jm = p2map(mode, tau, ra, rc)
            lanes.append(None if jm is None else S2*jm[1] + jm[0] + ra*q2)
        ins.append(lanes)
    res['p2-E2'] = conflicts(ins)
    ins = []
    for m in range(rc):
        ins.append([S2*m + t for t in range(T3)])
    res['p3-E2'] = conflicts(ins)
    return res, nt2
if __name__ == '__main__':
    for (ra, rb, rc, T) in [(8,16,20,160), (12,14,14,196), (16,16,20,320), (16,14,21,336), (14,12,14,196), (8,16,16,128), (10,16,16,256)]:
        G1 = rb*rc; T3 = ra*rb
        found = []
        for mode in ['jfast', 'mfast', 'row16', 'row16m']:
            if mode == 'row16' and rc > 16: continue
            if mode == 'row16m' and ra > 16: continue
            for S1 in range(G1, G1+17):
                ok = False
                for S2 in range(T3, T3+17):
                    r, nt2 = check(ra, rb, rc, T, mode, S1, S2)
                    if max(r.values()) == 1:
                        found.append((mode, S1, S2, nt2)); ok = True; break
                if ok: break
        print((ra, rb, rc, T), 'G1', G1, 'T3', T3, found)

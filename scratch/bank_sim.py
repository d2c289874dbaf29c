# Shared-memory bank-conflict simulator for the FFT pass access patterns.
import itertools, numpy as np

def wavefronts(idx_per_lane, elem_bytes, phys):
    """idx_per_lane: 32 element indices (one warp, one instruction). returns total wavefronts."""
    lanes_per_phase = {4:32, 8:16, 16:8}[elem_bytes]
    total = 0
    for ph in range(0, 32, lanes_per_phase):
        lanes = idx_per_lane[ph:ph+lanes_per_phase]
        words = elem_bytes//4
        bank_addrs = {}
        for i in lanes:
            a = phys(i)*words
            for w in range(words):
                bank_addrs.setdefault((a+w) % 32, set()).add((a+w)//32)
        total += max(len(v) for v in bank_addrs.values())
    return total, 32//lanes_per_phase

def patterns(n, R1):
    T = n//16; R2 = 16
    pats = {}
    # F1 store: thread t -> butterflies g = t + T*u, u < 16/R1 ; store p = k1 + R1*(g//16) + T*(g%16)
    for u in range(max(1, 16//R1)):
        for k1 in range(R1):
            for w in range(max(1, T//32)):
                lanes = [ (k1 + R1*((t + T*u)//16) + T*((t + T*u) % 16)) for t in range(32*w, 32*w+32) if t < T]
                if len(lanes) == 32: pats[('F1st',u,k1,w)] = lanes
    # F2 load/store: butterflies (i3, k1): 16*R1 of them, per thread 16*R1/T = 1 ; thread t -> k1 = t % R1, i3 = t // R1 ; elems k1 + R1*m + T*i3
    for m in range(16):
        for w in range(max(1, T//32)):
            lanes = [ ((t % R1) + R1*m + T*(t//R1)) for t in range(32*w, 32*w+32) if t < T]
            if len(lanes) == 32: pats[('F2',m,w)] = lanes
    # F3 load: thread j<T: j + T*m
    for m in range(16):
        for w in range(max(1,T//32)):
            lanes = [ t + T*m for t in range(32*w, 32*w+32) if t < T]
            if len(lanes) == 32: pats[('F3',m,w)] = lanes
    return pats

def score(n, R1, phys, eb):
    tot = 0; ideal = 0; worst = 0
    for k, lanes in patterns(n, R1).items():
        w, base = wavefronts(lanes, eb, phys)
        tot += w; ideal += base; worst = max(worst, w/base)
    return tot/ideal, worst

for eb in (8, 16):
  for n, R1 in [(512,2),(1024,4),(2048,8),(4096,16)]:
    T = n//16
    best = None
    cands = []
    for a in range(2,9):
        for b in [None]+list(range(a+1,12)):
            def phys(p, a=a, b=b): return p + (p>>a) + ((p>>b) if b is not None else 0)
            s = score(n, R1, phys, eb)
            cands.append((s[0], s[1], a, b))
    cands.sort(key=lambda c:(c[0],c[1],c[2],-1 if c[3] is None else c[3]))
    print(eb, n, R1, cands[:4])

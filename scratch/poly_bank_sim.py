# Bank-conflict simulation of the sliding-window polyphase stage.
import numpy as np, itertools
def wavefronts(word_addrs, width_words):
    # word_addrs: per-lane starting word address; width 1 (LDS.32) -> 32 lanes per wavefront, 2 (LDS.64) -> 16 lanes
    lanes_per = 32 // width_words
    tot = 0
    for ph in range(0, 32, lanes_per):
        banks = {}
        for a in word_addrs[ph:ph+lanes_per]:
            for w in range(width_words):
                banks.setdefault((a+w) % 32, set()).add((a+w)//32)
        tot += max(len(v) for v in banks.values())
    return tot
def sim(pL, pM, nc, R, L, us, nrow, at0=0, pre=9, n_warps=200, E=1, B=None):
    rho = pM/pL
    if B is None: B = int(rho)
    NT = nc + E
    wf_s = wf_c = ideal_s = ideal_c = 0
    rng = np.random.default_rng(0)
    for w in range(n_warps):
        n0 = int(rng.integers(0, 100000))
        lanes_n0 = [n0 + R*l for l in range(32)]
        at = [at0 + n*pM for n in lanes_n0]
        rel = [a//pL - pre + 40 for a in at]     # window start (U index), offset to be positive
        rem = [[(at0 + (n+i)*pM) % pL for n in lanes_n0] for i in range(R)]
        wi = [[(at0 + (n+i)*pM)//pL for n in lanes_n0] for i in range(R)]
        # sample loads: register m -> U idx rel+m
        nsamp = (R-1)*B + NT
        for m in range(nsamp):
            addrs = []
            for l in range(32):
                idx = rel[l] + m
                if L == 2: a = ((idx & 1)*us + (idx >> 1))*2
                else: a = idx*2
                addrs.append(a)
            wf_s += wavefronts(addrs, 2); ideal_s += 2
        for i in range(R):
            e = [wi[i][l] - wi[0][l] - i*B for l in range(32)]
            assert min(e) >= 0 and max(e) <= E, (min(e), max(e))
            for k in range(NT):
                addrs = [rem[i][l]*nrow + (E - e[l]) + k for l in range(32)]
                wf_c += wavefronts(addrs, 1); ideal_c += 1
    outs = n_warps*32*R
    return wf_s/ideal_s, wf_c/ideal_c, (wf_s+wf_c)/outs
for (pL,pM,nc,pre,name) in [(147,320,20,9,'48k->44.1k VHQ'),(147,320,14,6,'48k->44.1k HQ')]:
    for R in (4,5,6,7,8):
        E = int(np.ceil((R-1)*(pM/pL - int(pM/pL))))
        best=None
        for us in (1758,1760,1762,1764,1766,1768,1770,1772):
            for nrow in (nc+E+1|1, (nc+E+2)|1, nc+E+4|1):
                r = sim(pL,pM,nc,R,2,us,nrow,pre=pre,E=E)
                if best is None or r[2] < best[0][2]: best=(r,us,nrow)
        print(name,'R',R,'E',E,'best: samp x%.2f coef x%.2f  wavefronts/output %.2f  us %d nrow %d'%(best[0]+(best[1],best[2])))

print('---- interleaved U layout (L=2: a = idx*2)')
def sim2(pL, pM, nc, R, nrow, pre, E, B, interleaved=True, us=1762, n_warps=150, L=2):
    NT = nc + E
    wf_s = wf_c = ideal_s = ideal_c = 0
    rng = np.random.default_rng(0)
    for w in range(n_warps):
        n0 = int(rng.integers(0, 100000))
        lanes_n0 = [n0 + R*l for l in range(32)]
        wi = [[((n+i)*pM)//pL for n in lanes_n0] for i in range(R)]
        rem = [[((n+i)*pM) % pL for n in lanes_n0] for i in range(R)]
        rel = [wi[0][l] + 40 for l in range(32)]
        nsamp = (R-1)*B + NT
        for m in range(nsamp):
            addrs = []
            for l in range(32):
                idx = rel[l] + m
                a = idx*2 if (interleaved or L == 1) else ((idx & 1)*us + (idx >> 1))*2
                addrs.append(a)
            wf_s += wavefronts(addrs, 2); ideal_s += 2
        for i in range(R):
            e = [wi[i][l] - wi[0][l] - i*B for l in range(32)]
            if min(e) < 0 or max(e) > E: return None
            for k in range(NT):
                addrs = [rem[i][l]*nrow + (E - e[l]) + k for l in range(32)]
                wf_c += wavefronts(addrs, 1); ideal_c += 1
    outs = n_warps*32*R
    return wf_s/ideal_s, wf_c/ideal_c, (wf_s+wf_c)/outs, (ideal_s+ideal_c)/outs
cases = [(147,320,20,'48k->44.1k VHQ'),(147,320,14,'48k->44.1k HQ'),(160,441,12,'44.1k->16k HQ'),(80,147,20,'44.1->48 VHQ (160/147 x2)'),
         (320,441,14,'44.1k->32k HQ'),(441,640,14,'32k->44.1k HQ?')]
for (pL,pM,nc,name) in cases:
    rho = pM/pL; B = int(rho)
    res=[]
    for R in (3,4,5,6,7,8):
        E = int(np.ceil((R-1)*(rho-B)))
        for nrow in range(nc+E, nc+E+12):
            r = sim2(pL,pM,nc,R,nrow,0,E,B)
            if r: res.append((r[2],R,E,nrow,r))
    res.sort()
    for x in res[:3]: print(name,'rho %.3f B %d'%(rho,B),'R',x[1],'E',x[2],'nrow',x[3],'samp x%.2f coef x%.2f wf/out %.2f (ideal %.2f)'%x[4])

print('---- LDS.128 aligned sample loads (window start rounded down to even; E+1)')
def wavefronts_w(word_addrs, width_words):
    lanes_per = 32 // width_words
    tot = 0
    for ph in range(0, 32, lanes_per):
        banks = {}
        for a in word_addrs[ph:ph+lanes_per]:
            for w in range(width_words):
                banks.setdefault((a+w) % 32, set()).add((a+w)//32)
        tot += max(len(v) for v in banks.values())
    return tot
def sim3(pL, pM, nc, R, nrow, E, B, n_warps=150):
    NT = nc + E
    wf_s = wf_c = 0
    rng = np.random.default_rng(0)
    for w in range(n_warps):
        n0 = int(rng.integers(0, 100000))
        lanes_n0 = [n0 + R*l for l in range(32)]
        wi = [[((n+i)*pM)//pL for n in lanes_n0] for i in range(R)]
        rem = [[((n+i)*pM) % pL for n in lanes_n0] for i in range(R)]
        rel = [(wi[0][l] + 40) & ~1 for l in range(32)]
        odd = [(wi[0][l] + 40) & 1 for l in range(32)]
        nsamp = (R-1)*B + NT
        for m in range(0, nsamp, 2):
            addrs = [(rel[l] + m)*2 for l in range(32)]
            wf_s += wavefronts_w(addrs, 4)
        for i in range(R):
            e = [wi[i][l] - wi[0][l] - i*B + odd[l] for l in range(32)]
            if min(e) < 0 or max(e) > E: return None
            for k in range(NT):
                addrs = [rem[i][l]*nrow + (E - e[l]) + k for l in range(32)]
                wf_c += wavefronts_w(addrs, 1)
    outs = n_warps*32*R
    return wf_s/outs, wf_c/outs, (wf_s+wf_c)/outs
for (pL,pM,nc,name) in cases[:2]:
    rho = pM/pL; B = int(rho)
    res=[]
    for R in (4,5,6,7,8,9):
        E = int(np.ceil((R-1)*(rho-B))) + 1
        for nrow in range(nc+E, nc+E+12):
            r = sim3(pL,pM,nc,R,nrow,E,B)
            if r: res.append((r[2],R,E,nrow,r))
    res.sort()
    for x in res[:4]: print(name,'R',x[1],'E',x[2],'nrow',x[3],'samp wf/out %.2f coef wf/out %.2f total %.2f'%x[4])

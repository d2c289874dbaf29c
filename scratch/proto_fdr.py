# Prototype (numpy, double) of the single-pass frequency-domain rational resampler ("fdr"):
# FFT(Nin) -> remap bins x H -> IFFT(Nout), Nin/Nout = M_tot/L_tot, checked against the oracle.
import sys, numpy as np, math
sys.path.insert(0,'/root/repo/oracle')
import soxr_oracle as O

def zero_phase_response(g, step, F, P):
    """sum_i g[i] cos(2 pi F (i - c) step / P), c = (len-1)/2 (g symmetric), for integer array F."""
    c = (len(g)-1)/2.0
    i = np.arange(len(g)) - c
    # chunk to limit memory
    out = np.zeros(len(F))
    for a in range(0, len(F), 256):
        ph = 2*np.pi*np.outer(F[a:a+256], i)*step/P
        out[a:a+256] = np.cos(ph) @ g
    return out

def fdr_setup(P, a):
    st = [s for s in P.stages]
    assert all(s.kind == 'rational' for s in st)
    # composite: y[n] = sum_k gc[Cc + Mt n - Lt k] x[k] on the grid of rate in*Lt
    Lt = 1; Mt = 1
    for s in st: Lt *= s.L; Mt *= s.M
    g = math.gcd(Lt, Mt)
    # (keep un-reduced Lt, Mt for tick arithmetic; reduced ratio for block sizes)
    lr, mr = Lt//g, Mt//g
    Nin, Nout = mr*a, lr*a
    # ticks: stage i output sample j sits at tick j * tick_i where tick_i = prod_{s<=i} M_s * prod_{s>i} L_s
    # composite centre offset D0: y[n] = s(Mt n + D0); derive via stage-by-stage zero-phase delays
    # stage: out[j] = sum_k g[C + M j - L k] in[k] = (in_up * gz)(M j + C - c) on its own fine grid (rate in_stage*L)
    # fine grid tick of this stage in global ticks: tk = prod_{later stages} L
    D0 = 0.0; ticks_per_fine = []
    nst = len(st)
    for i, s in enumerate(st):
        later = 1
        for s2 in st[i+1:]: later *= s2.L
        c = (len(s.g)-1)/2.0
        D0 += (s.C - c) * later
        ticks_per_fine.append(later)
    return dict(Lt=Lt, Mt=Mt, Nin=Nin, Nout=Nout, D0=D0, ticks=ticks_per_fine, st=st, lr=lr, mr=mr)

def fdr_H(S, delta_ticks):
    Nin, Nout, Lt = S['Nin'], S['Nout'], S['Lt']
    Pp = Lt * Nin                     # block period in ticks
    F = np.arange(Nout); F = np.where(F < Nout/2, F, F - Nout).astype(float)
    Gz = np.ones(Nout)
    for s, tk in zip(S['st'], S['ticks']):
        Gz *= zero_phase_response(np.asarray(s.g, float), tk, F, Pp)
    # y = (1/P)*sum_F X[F] Gz[F] e^{2 pi i F (tau)/P}; DFT of zero-stuffed x equals X (no 1/L): amplitude: s = x_hi * gz
    H = Gz * np.exp(2j*np.pi*F*(delta_ticks + S['D0'])/Pp) / Pp
    if Nout % 2 == 0: H[Nout//2] = 0
    return H, F

def fdr_run(x, S, a_ovl, n_out_total):
    Nin, Nout, Lt, Mt, lr, mr = S['Nin'], S['Nout'], S['Lt'], S['Mt'], S['lr'], S['mr']
    a = Nin // mr
    adv_in, adv_out = mr*(a - a_ovl), lr*(a - a_ovl)
    m_in = mr*a_ovl//2; m_out = lr*a_ovl//2              # margins (left)
    # block b: inputs [adv_in*b - m_in, +Nin), outputs block-local o <-> n = adv_out*b - m_out + o ; owns o in [m_out, m_out+adv_out)
    # delta = Mt*nb - Lt*kb = Mt*(-m_out) - Lt*(-m_in) (constant)
    delta = Mt*(-m_out) - Lt*(-m_in)
    H, F = fdr_H(S, delta)
    Fin = (F.astype(int)) % Nin
    n_in = len(x)
    y = np.zeros(n_out_total, complex)
    nb = -(-n_out_total // adv_out)
    for b in range(nb):
        kb = adv_in*b - m_in
        idx = kb + np.arange(Nin)
        ok = (idx >= 0) & (idx < n_in)
        blk = np.zeros(Nin, complex); blk[ok] = x[idx[ok]]
        X = np.fft.fft(blk)
        Y = X[Fin] * H
        yb = np.fft.ifft(Y) * Nout
        n0 = adv_out*b
        n1 = min(n0 + adv_out, n_out_total)
        y[n0:n1] = yb[m_out:m_out + (n1-n0)]
    return y

rng = np.random.default_rng(0)
for (ir, orr, q, nin, a, a_ovl) in [(48000,44100,O.VHQ,20000,32,2),(48000,44100,O.HQ,20000,32,2),(44100,48000,O.VHQ,12000,32,2),
                                     (44100,16000,O.HQ,30000,16,2),(96000,48000,O.HQ,20000,2048,256),(44100,32000,O.HQ,20000,16,2)]:
    P = O.make_plan(ir, orr, q)
    kinds = [(s.sub, s.L, s.M, len(s.g)) for s in P.stages]
    x = rng.standard_normal(nin) + 1j*rng.standard_normal(nin)
    n_out = O.num_out(nin, P.io_ratio)
    S = fdr_setup(P, a)
    y = fdr_run(x, S, a_ovl, n_out)
    ref = O.resample_f64(np.stack([x.real, x.imag], 1), ir, orr, q)
    err = np.abs(y.real - ref[:,0]) + np.abs(y.imag - ref[:,1])
    bad = np.argwhere(err > 1e-7).ravel()
    print(ir, orr, q, kinds, 'Nin/Nout', S['Nin'], S['Nout'], 'rms %.2e'%np.sqrt(np.mean(err[30:]**2)), 'max(after first 30) %.2e'%err[30:].max(),
          'first outputs err', np.round(err[:6],9), 'bad idx', bad[:8], len(bad))

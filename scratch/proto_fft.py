# Prototype (numpy) of the fused FFT-stage kernel's exact dataflow, checked against the oracle.
import sys, numpy as np
sys.path.insert(0,'/root/repo/oracle')
import soxr_oracle as O

def dit_forward(x, R1):
    """3-pass DIT FFT, radices (R1,16,16); input natural x[i] read as the kernel does."""
    n = len(x); T = n//16; R2 = 16
    buf = np.zeros(n, complex)
    # F1: butterfly g loads x[g + (n/R1) m1]; stores g*R1+q1
    for g in range(n//R1):
        v = x[g + (n//R1)*np.arange(R1)]
        buf[np.arange(R1) + R1*(g//16) + T*(g%16)] = np.fft.fft(v)
    # F2: S=R1: groups G<16, j<R1: elements G*R1*R2 + j + R1*m ; twiddle W_{R1R2}^{j m}
    S = R1
    for G in range(n//(R1*R2)):
        for j in range(R1):
            idx = G*R1*R2 + j + R1*np.arange(R2)
            v = buf[idx]*np.exp(-2j*np.pi*j*np.arange(R2)/(R1*R2))
            buf[idx] = np.fft.fft(v)
    # F3: S=T: j<T: elements j + T m ; twiddle W_n^{j m}
    Z = np.zeros(n, complex)
    for j in range(T):
        idx = j + T*np.arange(16)
        v = buf[idx]*np.exp(-2j*np.pi*j*np.arange(16)/n)
        Z[idx] = np.fft.fft(v)      # natural order k = j + T q
    return Z

def dif_inverse(Y, R1):
    """mirror: natural in, passes S=T (16), S=R1 (16), S=1 (R1); returns natural y (unnormalised)."""
    n = len(Y); T = n//16; R2 = 16
    buf = np.zeros(n, complex)
    for j in range(T):
        idx = j + T*np.arange(16)
        v = np.fft.ifft(Y[idx])*16
        buf[idx] = v*np.exp(2j*np.pi*j*np.arange(16)/n)
    for G in range(n//(R1*R2)):
        for j in range(R1):
            idx = G*R1*R2 + j + R1*np.arange(R2)
            v = np.fft.ifft(buf[idx])*R2
            buf[idx] = v*np.exp(2j*np.pi*j*np.arange(R2)/(R1*R2))
    y = np.zeros(n, complex)
    for g in range(n//R1):
        v = np.fft.ifft(buf[np.arange(R1) + R1*(g//16) + T*(g%16)])*R1
        y[g + (n//R1)*np.arange(R1)] = v
    return y

rng = np.random.default_rng(0)
for n,R1 in [(1024,4),(2048,8),(4096,16),(512,2)]:
    x = rng.standard_normal(n)+1j*rng.standard_normal(n)
    Z = dit_forward(x,R1)
    print(n, np.abs(Z-np.fft.fft(x)).max(), np.abs(dif_inverse(Z,R1)/n - x).max())

# ---- full fused pipeline for UP2 + poly, vs oracle
def fused_up2_poly(xL, xR, P, n_out):
    st1, st2 = P.stages
    N, taps, L = st1.N, st1.num_taps, st1.L
    assert L == 2 and st1.M == 1
    n = N//L; R1 = n//256
    V = N - taps + 1; Vin = V//L; pre1 = st1.preload
    g = st1.g
    hc = np.zeros(N); hc[(np.arange(taps)+N-taps+1) % N] = g
    E = [np.fft.fft(hc[r::2])/n for r in range(2)]
    nc, Lp, Mp, at0, pre2 = st2.nc, st2.L, st2.M, st2.at0, st2.preload
    T2 = np.zeros((Lp, nc))
    for ph in range(Lp):
        for t in range(nc):
            gi = (nc-1-t)*Lp + ph - 1
            T2[ph,t] = st2.g[gi] if gi >= 0 else 0.0
    H = nc-1
    z = xL + 1j*xR
    n_in = len(z)
    def X(k):  # zero-extended input
        k = np.asarray(k); ok = (k>=0)&(k<n_in)
        out = np.zeros(k.shape, complex); out[ok] = z[k[ok]]; return out
    out = np.zeros(n_out, complex)
    d = lambda nn: (at0 + nn*Mp)//Lp
    def first_n(b):   # first output owned by block b
        q = b*V + pre2 - nc + 1
        if q <= 0: return 0
        return max(0, -((-(q*Lp - at0))//Mp))
    b_hi = (d(n_out-1) - pre2 + nc - 1)//V
    for b in range(0, b_hi+1):
        k0 = b*Vin - pre1
        blk = X(k0 + np.arange(n))
        Z = dit_forward(blk, R1)
        a = dif_inverse(Z*E[0], R1); bb = dif_inverse(Z*E[1], R1)
        U = np.zeros(H + V, complex)
        U[H+0:H+V:2] = a[:V//2]; U[H+1:H+V:2] = bb[:V//2]
        # halo: u[j] for j in [bV-H, bV) direct form
        for hh in range(H):
            j = b*V - H + hh
            if j < 0: continue
            top = st1.C + j
            klo = max(0, -(-(top-(taps-1))//L)); khi = top//L
            ks = np.arange(klo, khi+1)
            U[hh] = np.sum(g[top - L*ks]*X(ks))
        n0, n1 = first_n(b), min(first_n(b+1), n_out)
        for nn in range(n0, n1):
            at = at0 + nn*Mp
            w0 = at//Lp - pre2 - (b*V - H)
            out[nn] = np.dot(T2[at % Lp], U[w0:w0+nc])
    return out

for (ir, orr, q, nin) in [(48000,44100,O.HQ,5000),(48000,44100,O.VHQ,9000),(32000,44100,O.HQ,3000),(44100,48000,O.VHQ,3000)]:
    P = O.make_plan(ir, orr, q)
    x = rng.standard_normal((nin,2))
    n_out = O.num_out(nin, P.io_ratio)
    y = fused_up2_poly(x[:,0], x[:,1], P, n_out)
    ref = O.resample_f64(x, ir, orr, q)
    print(ir,orr,q, np.abs(y.real-ref[:,0]).max(), np.abs(y.imag-ref[:,1]).max())

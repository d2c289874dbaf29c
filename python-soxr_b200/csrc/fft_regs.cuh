// fft_regs.cuh -- small in-register DFTs (radix 2/4/8/16) on complex pairs.
//
// These replace the radix-2/4 butterfly passes of PFFFT (libsoxr/src/pffft.c:
// 496-558,644-806) as the arithmetic core of the overlap-save stage; the larger
// radices keep each sample in registers for 3-4 butterfly levels between
// shared-memory exchanges.  Outputs are in natural order.  INV selects the
// conjugate (inverse, unnormalised) transform.
//
// float2 arithmetic uses Blackwell's packed f32x2 instructions (FADD2 / FMUL2 /
// FFMA2 on sm_100a): same FLOP rate as scalar code but half the issue slots, and
// the operand swizzles (.F32 broadcast, .LO_HI swap, .NP sign pattern) make a
// complex multiply two instructions and a multiplication by -i free.  Everything
// is phrased through  rot(a) = -i*a = (a.y, -a.x)  because that is the sign
// pattern the hardware swizzle provides.
#pragma once
#include <cuda_runtime.h>

namespace b2 {

template <typename R> struct V2;
template <> struct V2<float> { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define B2_PACKED_F32 1
#else
#define B2_PACKED_F32 0
#endif

// ---- double2 (and host) : scalar -----------------------------------------------
__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 rot(double2 a) { return make_double2(a.y, -a.x); }   // -i * a
// a + s * b, s real
__host__ __device__ __forceinline__ double2 axpy(double s, double2 b, double2 a) {
  return make_double2(fma(s, b.x, a.x), fma(s, b.y, a.y));
}
__host__ __device__ __forceinline__ double2 scale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }

// ---- float2 : packed on the device -----------------------------------------------
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#if B2_PACKED_F32
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#if B2_PACKED_F32
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
__host__ __device__ __forceinline__ float2 rot(float2 a) { return make_float2(a.y, -a.x); }
__host__ __device__ __forceinline__ float2 axpy(float s, float2 b, float2 a) {
#if B2_PACKED_F32
  return __ffma2_rn(b, make_float2(s, s), a);   // swizzled pair (rot) in the A slot, broadcast scalar in B: ptxas folds
                                                  // LO_HI.NP only on A (else it materialises the negation: FADD + MOV)
#else
  return make_float2(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y));
#endif
}
__host__ __device__ __forceinline__ float2 scale(float s, float2 a) {
#if B2_PACKED_F32
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(s * a.x, s * a.y);
#endif
}

// a * b            = b.x * a - b.y * rot(a)
template <typename C> __host__ __device__ __forceinline__ C cmul(C a, C b) { return axpy(-b.y, rot(a), scale(b.x, a)); }
// a * conj(b)      = b.x * a + b.y * rot(a)
template <typename C> __host__ __device__ __forceinline__ C cmulc(C a, C b) { return axpy(b.y, rot(a), scale(b.x, a)); }
// a * (c - i s) forward, a * (c + i s) inverse  (c, s compile-time constants)
template <bool INV, typename C, typename R> __host__ __device__ __forceinline__ C mul_w(C a, R c, R s) {
  return axpy(INV ? -s : s, rot(a), scale(c, a));
}
// (a + w b, a - w b) with w = -i forward, +i inverse
template <bool INV, typename C> __host__ __device__ __forceinline__ void addsub_rot(C a, C b, C &p, C &m) {
  const C t = rot(b);
  if (INV) { p = csub(a, t); m = cadd(a, t); } else { p = cadd(a, t); m = csub(a, t); }
}

template <bool INV, typename C> __host__ __device__ __forceinline__ void dft2(C &a, C &b) {
  const C t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

template <bool INV, typename C> __host__ __device__ __forceinline__ void dft4(C &a, C &b, C &c, C &d) {
  const C t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  a = cadd(t0, t2);
  c = csub(t0, t2);
  addsub_rot<INV>(t1, t3, b, d);
}

template <int R> struct Dft;

template <> struct Dft<2> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[2]) { dft2<INV>(v[0], v[1]); }
};
template <> struct Dft<4> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[4]) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};

// 8 = 4 (over m1) x 2 (over m0), m = 2 m1 + m0
template <> struct Dft<8> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[8]) {
    typedef decltype(v[0].x) R;
    const R h = R(0.70710678118654752440);
    dft4<INV>(v[0], v[2], v[4], v[6]);   // t[0][qb]
    dft4<INV>(v[1], v[3], v[5], v[7]);   // t[1][qb]
    // X[qb + 4 qa] = t[0][qb] +- W8^qb t[1][qb]
    v[3] = mul_w<INV>(v[3], h, h);
    v[7] = mul_w<INV>(v[7], -h, h);
    C x0 = cadd(v[0], v[1]), x4 = csub(v[0], v[1]);
    C x1 = cadd(v[2], v[3]), x5 = csub(v[2], v[3]);
    C x2, x6;
    addsub_rot<INV>(v[4], v[5], x2, x6);   // W8^2 = -i
    C x3 = cadd(v[6], v[7]), x7 = csub(v[6], v[7]);
    v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
  }
};

// 16 = 4 (over m1) x 4 (over m0), m = 4 m1 + m0
template <> struct Dft<16> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[16]) {
    typedef decltype(v[0].x) R;
    const R c1 = R(0.92387953251128675613), s1 = R(0.38268343236508977173), h = R(0.70710678118654752440);
    // step 1: for each m0, DFT4 over m1 of x[4 m1 + m0]  -> t[m0][qb] stored at v[4 qb + m0]
    dft4<INV>(v[0], v[4], v[8], v[12]);
    dft4<INV>(v[1], v[5], v[9], v[13]);
    dft4<INV>(v[2], v[6], v[10], v[14]);
    dft4<INV>(v[3], v[7], v[11], v[15]);
    // step 2: t[m0][qb] *= W16^(m0*qb)   (the W16^4 = -i factor of v[10] is folded into step 3)
    v[5] = mul_w<INV>(v[5], c1, s1);     // m0=1,qb=1 : W^1
    v[6] = mul_w<INV>(v[6], h, h);       // m0=2,qb=1 : W^2
    v[7] = mul_w<INV>(v[7], s1, c1);     // m0=3,qb=1 : W^3
    v[9] = mul_w<INV>(v[9], h, h);       // m0=1,qb=2 : W^2
    v[11] = mul_w<INV>(v[11], -h, h);    // m0=3,qb=2 : W^6
    v[13] = mul_w<INV>(v[13], s1, c1);   // m0=1,qb=3 : W^3
    v[14] = mul_w<INV>(v[14], -h, h);    // m0=2,qb=3 : W^6
    v[15] = mul_w<INV>(v[15], -c1, -s1); // m0=3,qb=3 : W^9 = -W^1
    // step 3: for each qb, DFT4 over m0 -> X[qb + 4 qa] (in place at v[4 qb + qa])
    dft4<INV>(v[0], v[1], v[2], v[3]);
    dft4<INV>(v[4], v[5], v[6], v[7]);
    {   // qb = 2 with c = W16^4 * v[10] = rot(v[10]) forward, -rot(v[10]) inverse
      C &a = v[8], &b = v[9], &c = v[10], &d = v[11];
      C t0, t1;
      addsub_rot<INV>(a, c, t0, t1);     // a + w c, a - w c
      const C t2 = cadd(b, d), t3 = csub(b, d);
      a = cadd(t0, t2);
      c = csub(t0, t2);
      addsub_rot<INV>(t1, t3, b, d);
    }
    dft4<INV>(v[12], v[13], v[14], v[15]);
    // transpose v[4 qb + qa] -> X[qb + 4 qa]
    C t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[2]; v[2] = v[8]; v[8] = t;
    t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;
    t = v[7]; v[7] = v[13]; v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
  }
};

// ---- odd radices (3, 5, 7) and prime-factor composites (12, 14, 20) -------------------
// Needed by the single-pass frequency-domain resampler (fdr_stage.cuh), whose block sizes are
// 160a and 147a points.  PFFFT has radix-3/5 passes too but libsoxr compiles them out
// (pffft.c:320,424,559,808); here they are written from the DFT definition, with the
// symmetric-pair factorisation  X_m, X_{N-m} = a_m -+ i b_m  (forward).
template <> struct Dft<3> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[3]) {
    typedef decltype(v[0].x) R;
    const R s = R(0.86602540378443864676);
    const C t = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const C a = axpy(R(-0.5), t, v[0]);
    v[0] = cadd(v[0], t);
    v[1] = axpy(INV ? -s : s, rot(d), a);
    v[2] = axpy(INV ? s : -s, rot(d), a);
  }
};

template <> struct Dft<5> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[5]) {
    typedef decltype(v[0].x) R;
    const R c1 = R(0.30901699437494742410), c2 = R(-0.80901699437494742410);
    const R s1 = R(0.95105651629515357212), s2 = R(0.58778525229247312917);
    const C t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), d1 = csub(v[1], v[4]), d2 = csub(v[2], v[3]);
    const C a1 = axpy(c2, t2, axpy(c1, t1, v[0])), a2 = axpy(c1, t2, axpy(c2, t1, v[0]));
    C b1 = axpy(s2, d2, scale(s1, d1)), b2 = axpy(-s1, d2, scale(s2, d1));
    b1 = rot(b1);
    b2 = rot(b2);
    v[0] = cadd(v[0], cadd(t1, t2));
    if (INV) { v[1] = csub(a1, b1); v[4] = cadd(a1, b1); v[2] = csub(a2, b2); v[3] = cadd(a2, b2); }
    else { v[1] = cadd(a1, b1); v[4] = csub(a1, b1); v[2] = cadd(a2, b2); v[3] = csub(a2, b2); }
  }
};

template <> struct Dft<7> {
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[7]) {
    typedef decltype(v[0].x) R;
    const R c1 = R(0.62348980185873353053), c2 = R(-0.22252093395631440429), c3 = R(-0.90096886790241912624);
    const R s1 = R(0.78183148246802980871), s2 = R(0.97492791218182360702), s3 = R(0.43388373911755812048);
    const C t1 = cadd(v[1], v[6]), t2 = cadd(v[2], v[5]), t3 = cadd(v[3], v[4]);
    const C d1 = csub(v[1], v[6]), d2 = csub(v[2], v[5]), d3 = csub(v[3], v[4]);
    const C a1 = axpy(c3, t3, axpy(c2, t2, axpy(c1, t1, v[0])));
    const C a2 = axpy(c1, t3, axpy(c3, t2, axpy(c2, t1, v[0])));
    const C a3 = axpy(c2, t3, axpy(c1, t2, axpy(c3, t1, v[0])));
    const C b1 = rot(axpy(s3, d3, axpy(s2, d2, scale(s1, d1))));
    const C b2 = rot(axpy(-s1, d3, axpy(-s3, d2, scale(s2, d1))));
    const C b3 = rot(axpy(s2, d3, axpy(-s1, d2, scale(s3, d1))));
    v[0] = cadd(cadd(v[0], t1), cadd(t2, t3));
    if (INV) {
      v[1] = csub(a1, b1); v[6] = cadd(a1, b1); v[2] = csub(a2, b2); v[5] = cadd(a2, b2);
      v[3] = csub(a3, b3); v[4] = cadd(a3, b3);
    } else {
      v[1] = cadd(a1, b1); v[6] = csub(a1, b1); v[2] = cadd(a2, b2); v[5] = csub(a2, b2);
      v[3] = cadd(a3, b3); v[4] = csub(a3, b3);
    }
  }
};

// Good-Thomas prime-factor DFT for coprime N1, N2: no twiddles between the two sub-transforms,
// n = (N2 n1 + N1 n2) mod N in, k = CRT(k1, k2) out; all index maps are compile-time.
constexpr int pfa_inv_mod(int a, int m) {   // a^-1 mod m (m small)
  int r = 1;
  for (int i = 1; i < m; ++i)
    if ((a * i) % m == 1) r = i;
  return r;
}
template <int N1, int N2> struct PfaDft {
  static constexpr int N = N1 * N2;
  static constexpr int E1 = N2 * pfa_inv_mod(N2 % N1, N1), E2 = N1 * pfa_inv_mod(N1 % N2, N2);
  template <bool INV, typename C> static __host__ __device__ __forceinline__ void run(C (&v)[N1 * N2]) {
    C a[N2][N1];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2)
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) a[n2][n1] = v[(N2 * n1 + N1 * n2) % N];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) Dft<N1>::template run<INV>(a[n2]);
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
      C b[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) b[n2] = a[n2][k1];
      Dft<N2>::template run<INV>(b);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) v[(k1 * E1 + k2 * E2) % N] = b[k2];
    }
  }
};
template <> struct Dft<6> : PfaDft<2, 3> {};
template <> struct Dft<10> : PfaDft<2, 5> {};
template <> struct Dft<12> : PfaDft<4, 3> {};
template <> struct Dft<14> : PfaDft<2, 7> {};
template <> struct Dft<15> : PfaDft<3, 5> {};
template <> struct Dft<20> : PfaDft<4, 5> {};
template <> struct Dft<21> : PfaDft<3, 7> {};
template <> struct Dft<28> : PfaDft<4, 7> {};

}  // namespace b2

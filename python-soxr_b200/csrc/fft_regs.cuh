// fft_regs.cuh -- small in-register DFTs (radix 2/4/8/16) on complex pairs.
//
// These replace the radix-2/4 butterfly passes of PFFFT (libsoxr/src/pffft.c:
// 496-558,644-806) as the arithmetic core of the overlap-save stage; the larger
// radices keep each sample in registers for 3-4 butterfly levels between
// shared-memory exchanges.  Outputs are in natural order.  INV selects the
// conjugate (inverse, unnormalised) transform.
#pragma once
#include <cuda_runtime.h>

namespace b2 {

template <typename R> struct V2;
template <> struct V2<float> { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };

template <typename C> __host__ __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __host__ __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// a * b
template <typename C> __host__ __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
// a * conj(b)
template <typename C> __host__ __device__ __forceinline__ C cmulc(C a, C b) {
  C r;
  r.x = a.x * b.x + a.y * b.y;
  r.y = a.y * b.x - a.x * b.y;
  return r;
}
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename C> __host__ __device__ __forceinline__ C mul_mi(C a) {
  C r;
  if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; }
  return r;
}
// multiply by the constant (c - i s) forward, (c + i s) inverse
template <bool INV, typename C, typename R> __host__ __device__ __forceinline__ C mul_w(C a, R c, R s) {
  C r;
  if (INV) { r.x = a.x * c - a.y * s; r.y = a.y * c + a.x * s; }
  else { r.x = a.x * c + a.y * s; r.y = a.y * c - a.x * s; }
  return r;
}

template <bool INV, typename C> __host__ __device__ __forceinline__ void dft2(C &a, C &b) {
  const C t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

template <bool INV, typename C> __host__ __device__ __forceinline__ void dft4(C &a, C &b, C &c, C &d) {
  const C t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = mul_mi<INV>(csub(b, d));
  a = cadd(t0, t2);
  c = csub(t0, t2);
  b = cadd(t1, t3);
  d = csub(t1, t3);
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
    // t[1][qb] *= W8^qb
    v[3] = mul_w<INV>(v[3], h, h);
    v[5] = mul_mi<INV>(v[5]);
    v[7] = mul_w<INV>(v[7], -h, h);
    // X[qb + 4 qa] = t[0][qb] +- t[1][qb]
    C x0 = cadd(v[0], v[1]), x4 = csub(v[0], v[1]);
    C x1 = cadd(v[2], v[3]), x5 = csub(v[2], v[3]);
    C x2 = cadd(v[4], v[5]), x6 = csub(v[4], v[5]);
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
    // step 2: t[m0][qb] *= W16^(m0*qb)
    v[5] = mul_w<INV>(v[5], c1, s1);     // m0=1,qb=1 : W^1
    v[6] = mul_w<INV>(v[6], h, h);       // m0=2,qb=1 : W^2
    v[7] = mul_w<INV>(v[7], s1, c1);     // m0=3,qb=1 : W^3
    v[9] = mul_w<INV>(v[9], h, h);       // m0=1,qb=2 : W^2
    v[10] = mul_mi<INV>(v[10]);          // m0=2,qb=2 : W^4
    v[11] = mul_w<INV>(v[11], -h, h);    // m0=3,qb=2 : W^6
    v[13] = mul_w<INV>(v[13], s1, c1);   // m0=1,qb=3 : W^3
    v[14] = mul_w<INV>(v[14], -h, h);    // m0=2,qb=3 : W^6
    v[15] = mul_w<INV>(v[15], -c1, -s1); // m0=3,qb=3 : W^9 = -W^1
    // step 3: for each qb, DFT4 over m0 -> X[qb + 4 qa] (in place at v[4 qb + qa])
    dft4<INV>(v[0], v[1], v[2], v[3]);
    dft4<INV>(v[4], v[5], v[6], v[7]);
    dft4<INV>(v[8], v[9], v[10], v[11]);
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

}  // namespace b2

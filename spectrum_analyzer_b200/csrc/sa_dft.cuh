// sa_dft.cuh -- register-resident forward DFTs of size 2/4/8/16 on float2 arrays.
//
// These are the butterflies of the Stockham passes in the fused kernels.  They replace the
// arithmetic the reference delegates to microfft::real (reference src/fft.rs:37,89-106): same
// transform (forward, unnormalised, e^{-2*pi*i*nk/R}), different factorisation (radix-4/8/16
// instead of microfft's radix-2), so results agree to f32 rounding, not bit for bit.
// All outputs are in natural order.
//
// Blackwell-specific: complex values are kept as (re,im) register pairs and every complex
// add/sub/scale is ONE packed instruction (add/mul/fma .rn.f32x2 -> SASS FADD2/FMUL2/FFMA2,
// sm_100+).  ptxas folds the "swap halves and negate one" of a multiplication by -i, and the
// scalar broadcast of a twiddle component, into operand modifiers (.F32x2.LO_HI.NP, .F32), so a
// multiplication by -i is free and a general complex multiply is FMUL2 + FFMA2.
#pragma once
#include <cuda_runtime.h>

namespace sa {

// ---- packed f32x2 primitives (two IEEE binary32 operations per instruction) -----------------
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 bcast2(float s) { return make_float2(s, s); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }
// a * (-i) = (a.y, -a.x): a swizzle, folded into the consumer's operand modifiers
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }
// (a.x + i a.y) * (w.x + i w.y) = w.x * a + w.y * (-a.y, a.x)
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  const float2 t = mul2(bcast2(w.y), make_float2(a.y, a.x));   // (w.y a.y, w.y a.x)
  return fma2(bcast2(w.x), a, make_float2(-t.x, t.y));
}

#define SA_SQRT1_2 0.70710678118654752440f
#define SA_COS_PI_8 0.92387953251128675613f
#define SA_SIN_PI_8 0.38268343236508977173f

// a * (c - i s) for compile-time constants c, s:  c*a + s*(a.y, -a.x)
__device__ __forceinline__ float2 cmul_const(float2 a, float c, float s) {
  return fma2(bcast2(s), make_float2(a.y, -a.x), mul2(bcast2(c), a));
}
// a * W8^1 = a * (1 - i)/sqrt(2) = ((a.x + a.y), (a.y - a.x)) / sqrt(2)
__device__ __forceinline__ float2 cmul_w8_1(float2 a) {
  return mul2(add2(a, make_float2(a.y, -a.x)), bcast2(SA_SQRT1_2));
}
// a * W8^3 = a * (-1 - i)/sqrt(2) = ((a.y - a.x), -(a.x + a.y)) / sqrt(2)
__device__ __forceinline__ float2 cmul_w8_3(float2 a) {
  return mul2(sub2(make_float2(a.y, -a.x), a), bcast2(SA_SQRT1_2));
}

__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
  const float2 t = a;
  a = add2(t, b);
  b = sub2(t, b);
}

__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 t0 = add2(a0, a2), t1 = sub2(a0, a2);
  const float2 t2 = add2(a1, a3), t3 = cmul_mi(sub2(a1, a3));
  a0 = add2(t0, t2);
  a2 = sub2(t0, t2);
  a1 = add2(t1, t3);
  a3 = sub2(t1, t3);
}

__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  o1 = cmul_w8_1(o1);
  o2 = cmul_mi(o2);
  o3 = cmul_w8_3(o3);
  v[0] = add2(e0, o0); v[4] = sub2(e0, o0);
  v[1] = add2(e1, o1); v[5] = sub2(e1, o1);
  v[2] = add2(e2, o2); v[6] = sub2(e2, o2);
  v[3] = add2(e3, o3); v[7] = sub2(e3, o3);
}

// 16 = 4 x 4: n = 4*n1 + n2, k = k1 + 4*k2.
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  // column DFTs over n1 for each n2: v[n2 + 4*k1] = Y[n2][k1]
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
  // twiddles W16^(n2*k1);  W16^m = (cos(m pi/8), -sin(m pi/8))
  v[1 + 4] = cmul_const(v[1 + 4], SA_COS_PI_8, SA_SIN_PI_8);      // W16^1
  v[1 + 8] = cmul_w8_1(v[1 + 8]);                                 // W16^2 = W8^1
  v[1 + 12] = cmul_const(v[1 + 12], SA_SIN_PI_8, SA_COS_PI_8);    // W16^3
  v[2 + 4] = cmul_w8_1(v[2 + 4]);                                 // W16^2
  v[2 + 8] = cmul_mi(v[2 + 8]);                                   // W16^4 = -i
  v[2 + 12] = cmul_w8_3(v[2 + 12]);                               // W16^6 = W8^3
  v[3 + 4] = cmul_const(v[3 + 4], SA_SIN_PI_8, SA_COS_PI_8);      // W16^3
  v[3 + 8] = cmul_w8_3(v[3 + 8]);                                 // W16^6
  v[3 + 12] = cmul_const(v[3 + 12], -SA_COS_PI_8, -SA_SIN_PI_8);  // W16^9 = (-cos, +sin)(pi/8)
  // row DFTs over n2 for each k1: result Z[k2] = X[k1 + 4*k2] lands in v[k2 + 4*k1]
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // transpose 4x4 back to natural order (register renaming)
  float2 t;
#define SA_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t
  SA_SWAP(1, 4); SA_SWAP(2, 8); SA_SWAP(3, 12); SA_SWAP(6, 9); SA_SWAP(7, 13); SA_SWAP(11, 14);
#undef SA_SWAP
}

template <int R>
__device__ __forceinline__ void dft(float2 (&v)[R]) {
  static_assert(R == 2 || R == 4 || R == 8 || R == 16, "unsupported radix");
  if constexpr (R == 2) dft2(v[0], v[1]);
  else if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 8) dft8(v);
  else dft16(v);
}

}  // namespace sa

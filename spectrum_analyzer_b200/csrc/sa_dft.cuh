// sa_dft.cuh -- register-resident forward DFTs of size 2/4/8/16 on float2 arrays.
//
// These are the butterflies of the Stockham passes in sa_spectrum_kernel.cuh.  They
// replace the arithmetic the reference delegates to microfft::real (reference
// src/fft.rs:37,89-106): same transform (forward, unnormalised, e^{-2*pi*i*nk/R}),
// different factorisation (radix-4/8/16 instead of microfft's radix-2), so results
// agree to f32 rounding, not bit for bit.  All outputs are in natural order.
#pragma once
#include <cuda_runtime.h>

namespace sa {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// (a.x + i a.y) * (b.x + i b.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}
// a * (-i)
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

#define SA_SQRT1_2 0.70710678118654752440f
#define SA_COS_PI_8 0.92387953251128675613f
#define SA_SIN_PI_8 0.38268343236508977173f

__device__ __forceinline__ void dft2(float2 &a, float2 &b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
  const float2 t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

// a * W8^1 = a * (1 - i)/sqrt(2);  a * W8^3 = a * (-1 - i)/sqrt(2)
__device__ __forceinline__ float2 cmul_w8_1(float2 a) {
  return make_float2((a.x + a.y) * SA_SQRT1_2, (a.y - a.x) * SA_SQRT1_2);
}
__device__ __forceinline__ float2 cmul_w8_3(float2 a) {
  return make_float2((a.y - a.x) * SA_SQRT1_2, -(a.x + a.y) * SA_SQRT1_2);
}

__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  o1 = cmul_w8_1(o1);
  o2 = cmul_mi(o2);
  o3 = cmul_w8_3(o3);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// 16 = 4 x 4: n = 4*n1 + n2, k = k1 + 4*k2.
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  // column DFTs over n1 for each n2: v[n2 + 4*k1] = Y[n2][k1]
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
  // twiddles W16^(n2*k1)
  const float2 w1 = make_float2(SA_COS_PI_8, -SA_SIN_PI_8);   // W16^1
  const float2 w3 = make_float2(SA_SIN_PI_8, -SA_COS_PI_8);   // W16^3
  const float2 w9 = make_float2(-SA_COS_PI_8, SA_SIN_PI_8);   // W16^9
  v[1 + 4] = cmul(v[1 + 4], w1);          // n2=1,k1=1
  v[1 + 8] = cmul_w8_1(v[1 + 8]);         // n2=1,k1=2 : W16^2 = W8^1
  v[1 + 12] = cmul(v[1 + 12], w3);        // n2=1,k1=3
  v[2 + 4] = cmul_w8_1(v[2 + 4]);         // n2=2,k1=1 : W16^2
  v[2 + 8] = cmul_mi(v[2 + 8]);           // n2=2,k1=2 : W16^4 = -i
  v[2 + 12] = cmul_w8_3(v[2 + 12]);       // n2=2,k1=3 : W16^6 = W8^3
  v[3 + 4] = cmul(v[3 + 4], w3);          // n2=3,k1=1
  v[3 + 8] = cmul_w8_3(v[3 + 8]);         // n2=3,k1=2 : W16^6
  v[3 + 12] = cmul(v[3 + 12], w9);        // n2=3,k1=3 : W16^9
  // row DFTs over n2 for each k1: result Z[k2] = X[k1 + 4*k2] lands in v[k2 + 4*k1]
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // transpose 4x4 back to natural order
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

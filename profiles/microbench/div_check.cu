// Verification: is v/d == two Markstein FMA correction steps on RN(1/d) for every (v, d) the zero-to-one
// scaling can meet (0 <= v <= d, d normal, d's significand not all ones)?  Compares against IEEE __fdiv_rn
// on the GPU over pseudo-random and structured pairs and counts mismatches.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o div_check div_check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float div2(float v, float d, float rd) {
  float q = __fmul_rn(v, rd);
  float r = __fmaf_rn(-q, d, v);
  q = __fmaf_rn(r, rd, q);
  r = __fmaf_rn(-q, d, v);
  return __fmaf_rn(r, rd, q);
}
__device__ __forceinline__ float div1(float v, float d, float rd) {
  const float q = __fmul_rn(v, rd);
  const float r = __fmaf_rn(-q, d, v);
  return __fmaf_rn(r, rd, q);
}
__device__ __forceinline__ uint32_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

__global__ void check(uint64_t seed, int iters, unsigned long long *bad1, unsigned long long *bad2, unsigned long long *n) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  unsigned long long b1 = 0, b2 = 0, cnt = 0;
  for (int it = 0; it < iters; it++) {
    const uint64_t k = seed + tid * 0x9e3779b97f4a7c15ull + (uint64_t)it * 0x632be59bd9b4e019ull;
    // divisor: random significand, exponent in a range spectra meet (2^-40 .. 2^40)
    const uint32_t dm = mix(k) & 0x7fffffu;
    if (dm == 0x7fffffu) continue;   // all-ones significand: the kernel falls back to IEEE division there
    const uint32_t de = 87u + (mix(k ^ 0x1234567ull) % 80u);
    const float d = __uint_as_float((de << 23) | dm);
    const float rd = __frcp_rn(d);
    // dividends: (a) random below d, down to 2^-30 d; (b) values near d*i/2^j (ties of the quotient)
    const uint32_t r2 = mix(k ^ 0xabcdef01ull);
    float v;
    if (it & 1) {
      const uint32_t ve = de - (r2 % 30u);
      v = __uint_as_float((ve << 23) | (mix(k ^ 0x77ull) & 0x7fffffu));
      if (v > d) v = d;
    } else {
      const float t = __uint_as_float((127u << 23) | (r2 & 0x7fffffu)) - 1.0f;   // [0,1)
      v = __fmul_rn(d, t);
      v = __uint_as_float(__float_as_uint(v) + ((r2 >> 29) & 3u) - 1u);            // neighbours of d*t
      if (!(v >= 0.0f) || v > d) v = d * 0.5f;
    }
    if (v != 0.0f && v < 5.4e-20f) continue;   // the kernel's magnitudes below this are flushed to zero
    const float want = __fdiv_rn(v, d);
    b1 += (__float_as_uint(div1(v, d, rd)) != __float_as_uint(want));
    b2 += (__float_as_uint(div2(v, d, rd)) != __float_as_uint(want));
    cnt++;
  }
  atomicAdd(bad1, b1); atomicAdd(bad2, b2); atomicAdd(n, cnt);
}

int main() {
  unsigned long long *d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
  for (int rep = 0; rep < 8; rep++) check<<<148 * 16, 256>>>(0x5A17ull + rep * 977ull, 4096, d, d + 1, d + 2);
  unsigned long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("pairs %llu  mismatches: one correction step %llu, two steps %llu  (%s)\n", h[2], h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Microbenchmark: can tensor memory (TMEM) serve as a per-thread read-only table store, and does
// tcgen05.ld traffic contend with the shared-memory (LSU) pipe?   One CTA of 512 threads per SM.
//   mode 0: LDS.64 only (16 per iteration per thread)        -> shared-memory wavefront rate
//   mode 1: tcgen05.ld.32x32b.x32 only (1 per iteration)     -> TMEM read rate
//   mode 2: both per iteration                               -> overlap
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                 "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                 "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, uint32_t *out, long long *cycles) {
  __shared__ uint32_t s_taddr;
  __shared__ __align__(16) float2 s_tab[16 * 128];
  const int tid = threadIdx.x, warp = tid >> 5, j = tid & 127;
  for (int i = tid; i < 16 * 128; i += 512) s_tab[i] = make_float2((float)i, (float)(i * 3));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&s_taddr)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tbase = s_taddr;
  // lane quadrant of this warp, columns [0,32) shared by the 4 warps with the same (warp & 3)
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16);
  if (warp < 4) {
    uint32_t v[32];
#pragma unroll
    for (int k = 0; k < 32; k++) v[k] = (uint32_t)(j * 100 + k);
    tmem_st32(taddr, v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");

  uint32_t acc = 0;
  float facc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int s = 0; s < 16; s++) {
        const float2 w = s_tab[((j + it) & 127) + s * 128];
        facc += w.x * w.y;
      }
    }
    if (MODE == 1 || MODE == 2) {
      uint32_t r[32];
      tmem_ld32(taddr, r);
#pragma unroll
      for (int k = 0; k < 32; k++) acc ^= r[k] + (uint32_t)it;
    }
  }
  const long long t1 = clock64();
  // check: every thread must have read its own lane's values
  uint32_t r[32];
  tmem_ld32(taddr, r);
  uint32_t bad = 0;
#pragma unroll
  for (int k = 0; k < 32; k++) bad |= (r[k] != (uint32_t)(j * 100 + k));
  out[blockIdx.x * 512 + tid] = acc ^ __float_as_uint(facc);
  if (bad) atomicAdd((unsigned long long *)&cycles[1], 1ull);
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "n"(128));
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t *out; long long *cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * 512 * sms));
  CK(cudaMalloc(&cyc, 16));
  const int iters = 20000;
  for (int mode = 0; mode < 3; mode++) {
    CK(cudaMemset(cyc, 0, 16));
    for (int rep = 0; rep < 2; rep++) {
      if (mode == 0) bench<0><<<sms, 512>>>(iters, out, cyc);
      if (mode == 1) bench<1><<<sms, 512>>>(iters, out, cyc);
      if (mode == 2) bench<2><<<sms, 512>>>(iters, out, cyc);
      CK(cudaDeviceSynchronize());
    }
    long long h[2];
    CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    const double c = (double)h[0] / iters;
    // per iteration per SM: LDS: 16 warps * 16 LDS.64 * 256 B = 64 KiB; LDTM: 16 warps * 32 regs * 128 B = 64 KiB
    printf("mode %d: %.1f cycles/iter/SM  bad=%lld", mode, c, h[1] / 2);
    if (mode == 0) printf("  LDS  %.1f B/clk/SM\n", 65536.0 / c);
    if (mode == 1) printf("  LDTM %.1f B/clk/SM\n", 65536.0 / c);
    if (mode == 2) printf("  LDS+LDTM %.1f B/clk/SM total (sum of the two alone would be the no-contention case)\n", 131072.0 / c);
  }
  return 0;
}

// sa_tmem.cuh -- tensor memory (TMEM) as a per-thread read-only table store.
//
// The fused spectrum kernel is bound by the LSU/shared-memory data pipe (profiles/README.md), and a
// quarter of that traffic is per-thread constants that never change from frame to frame: the window
// multipliers of the 32 samples a thread owns, its pass twiddles and its recombination twiddles.
// Blackwell's tensor memory is 128 lanes x 512 columns x 32 bit per SM with its own datapath:
// with the 32x32b shape, thread t of a warp reads / writes lane 32*(warp%4)+t, consecutive columns
// into consecutive registers.  That is exactly a per-thread private array, and reading it does not
// touch the shared-memory pipe (profiles/microbench/tmem_ld.cu: LDS alone 128 B/clk/SM, LDTM alone
// 185 B/clk/SM, both together 245 B/clk/SM).  No tensor-core instruction is involved.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace sa {

// One warp allocates NCOLS columns (power of two >= 32); the base address lands in *smem_slot.
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
               :: "r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(base), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// Address of column `col` in this warp's lane quadrant.
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int warp, int col) {
  return base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)col;
}

#define SA_R16(a) a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]

// 16 / 32 consecutive columns of the calling thread's lane -> registers (waits for the data).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               "tcgen05.wait::ld.sync.aligned;\n"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                 "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
               "tcgen05.wait::ld.sync.aligned;\n"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
                 "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
                 "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
                 "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
               : "r"(taddr));
}
// registers -> 16 consecutive columns of the calling thread's lane (waits for completion).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
               "tcgen05.wait::st.sync.aligned;\n"
               :: "r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
                  "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15])
               : "memory");
}


// ---- bulk asynchronous copy global -> shared (TMA, 1-D) completing on an mbarrier ---------------------
// Used to stage the next frame's samples: unlike ld.global into registers, the copy occupies no
// register scoreboard, so the statistics phase that runs meanwhile never waits for it (ptxas shares its
// six scoreboards between ld.global and the LDS / SHFL / ATOMS of that phase; measured: "prefetching"
// into registers was slower than not prefetching at all, profiles/README.md).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// one arrival + `bytes` expected from the copy engine
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one arrival (release at CTA scope: the arriving thread's earlier shared-memory accesses, and those of the
// threads it synchronised with before, happen before whatever a waiter does after its wait)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
// shared-memory counter with acquire-release semantics at CTA scope: the thread that sees the last count may act on
// what the earlier arrivals (and the threads they synchronised with) did before arriving
__device__ __forceinline__ uint32_t atoms_add_acq_rel(uint32_t *addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;\n" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}
// Orders earlier generic-proxy accesses to shared memory before later async-proxy accesses (bulk copies) to it.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// The same for waits that are expected to take a while (role hand-over): the hint lets the hardware keep the warp
// suspended instead of re-issuing the probe, which would compete with the warp that is being waited for.
#ifndef SA_MBAR_HINT_NS
#define SA_MBAR_HINT_NS 2000
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
  if (SA_MBAR_HINT_NS == 0) { mbar_wait(bar, parity); return; }
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SA_MBAR_HINT_NS) : "memory");
  } while (!done);
}

// 4 keys of one 16-byte chunk: is any of them inside the candidate range [c_lo, c_lo + c_w]?  9 instructions.
__device__ __forceinline__ uint32_t chunk_hit(const uint4 k, uint32_t c_lo, uint32_t c_w) {
  uint32_t hit;
  asm("{\n.reg .pred p, q;\n.reg .u32 d0, d1, d2, d3;\n"
      "sub.u32 d0, %1, %5;\n sub.u32 d1, %2, %5;\n sub.u32 d2, %3, %5;\n sub.u32 d3, %4, %5;\n"
      "setp.le.u32 p, d0, %6;\n setp.le.u32 q, d1, %6;\n setp.le.or.u32 p, d2, %6, p;\n setp.le.or.u32 q, d3, %6, q;\n"
      "or.pred p, p, q;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(hit) : "r"(k.x), "r"(k.y), "r"(k.z), "r"(k.w), "r"(c_lo), "r"(c_w));
  return hit;
}
// shared-memory atomic as a single instruction (the compiler's warp-aggregated expansion of atomicAdd(.., 1) costs
// more than the whole test it would sit in; at most a few lanes are active where this is used)
__device__ __forceinline__ uint32_t atoms_add(uint32_t *addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;\n" : "=r"(old) : "r"(smem_u32(addr)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_inc(uint32_t *addr) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;\n" : "=r"(old) : "r"(smem_u32(addr)) : "memory");
  return old;
}


}  // namespace sa

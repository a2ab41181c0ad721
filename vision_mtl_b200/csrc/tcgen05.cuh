// Thin inline-PTX layer over the Blackwell tensor-core path (sm_100a): tcgen05.mma with
// kind::tf32 operands in 128B-swizzled shared memory, fp32 accumulators in TMEM, mbarrier
// completion.  Bit layouts follow the UMMA descriptor definitions of the PTX ISA
// (shared-memory matrix descriptor, instruction descriptor for .kind::tf32).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vmtl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// try_wait parks the thread until the phase completes or the hint expires; without a hint it returns
// after a few tens of ns and the retry loops of idle warps eat the issue slots of the working ones.
constexpr uint32_t kMbarSuspendNs = 2000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    // a tile's MMAs take microseconds; a wait this long means a lost arrival -> fail loudly
    if (!done && ++spins > (1u << 22)) __trap();
  } while (!done);
}

// Busy-polling wait for the one thread whose latency is the kernel's clock (the MMA issuer): no parking,
// so the hand-over costs a shared-memory round trip instead of a wake-up.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 28)) __trap();
  } while (!done);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// barrier among a sub-group of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- proxies / fences -------------------------------------------------------------------
// generic-proxy st.shared -> async-proxy (tensor core) reads
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one full warp executes these) ----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---- descriptors ------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B.  K-major operand: rows of 128 bytes
// (32 tf32), 8-row groups 1024 bytes apart (SBO).  `addr` must lie in a 1024B-aligned tile;
// advancing K inside the 128-byte row = adding bytes to the start address.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// MN-major tf32 operands have ONE legal shared-memory layout: SWIZZLE_128B_BASE32B.  Rows (one per
// K index) of 128 bytes = 32 consecutive M/N elements, 4-row groups of 512 bytes, the 32-byte
// chunk index XORed with (row % 4).  LBO = byte stride between 32-element atoms along M/N,
// SBO = byte stride between 4-row groups along K (512 when rows are dense).
__device__ __forceinline__ uint64_t smem_desc_mn_tf32(uint32_t addr, uint32_t lbo_bytes,
                                                      uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of 16-byte chunk `c` (0..7) of K-row `row` inside a [rows x 128B] BASE32B tile
__device__ __forceinline__ uint32_t sw128b32_off(int row, int c) {
  return (uint32_t)(row * 128 + ((((c >> 1) ^ (row & 3)) << 5) | ((c & 1) << 4)));
}

// Instruction descriptor for kind::tf32, fp32 accumulate, M x N tile.
// a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4)                    // D format  = F32
         | (2u << 7)                  // A format  = TF32
         | (2u << 10)                 // B format  = TF32
         | ((uint32_t)a_mn << 15)     // A major
         | ((uint32_t)b_mn << 16)     // B major
         | ((uint32_t)(N >> 3) << 17) // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 with BF16 operands, fp32 accumulate (K = 16 per instruction)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// x = hi + mid + O(2^-17 |x|) with hi, mid bf16: 8 floats -> 8 packed hi and 8 packed mid (element 2q in the
// low half of word q)
__device__ __forceinline__ void bf16_split8(const float* x, uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float a = x[2 * q], b = x[2 * q + 1];
    uint32_t hw;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hw) : "f"(b), "f"(a));  // upper half <- b, lower half <- a
    const float ra = a - __uint_as_float(hw << 16), rb = b - __uint_as_float(hw & 0xffff0000u);
    uint32_t mw;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mw) : "f"(rb), "f"(ra));
    h[q] = hw;
    m[q] = mw;
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem].  A occupies TMEM lanes 0..127 (row i = lane i), one 32-bit
// column per K element; the 8 columns of a K-step start at `a_tmem`.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- TMA (cp.async.bulk.tensor) ------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
// 2-D tiled load: box (x = inner coordinate in elements, y = row) -> smem, completes `bar`'s tx count
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// 2-D tiled store smem -> global (rows/cols outside the tensor are clipped); bulk async-group
__device__ __forceinline__ void tma_store_2d(const void* tmap, int x, int y, uint32_t smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(x), "r"(y)
               : "memory");
}
// same, but the tile is ADDED to global memory (fp32 reduction performed at L2)
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, int x, int y, uint32_t smem_src) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed stores have finished READING shared memory (the staging buffer may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- registers -> TMEM: thread (lane l of warp w) writes 16 consecutive columns of lane 32*(w%4)+l
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM -> registers: warp w reads lanes 32*(w%4)..+31, one row per thread ----------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// un-waited forms: issue several loads, then ONE tmem_wait_ld() before the registers are used
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties 4 loaded registers to a point after the wait (volatile asms keep their order), so the compiler
// cannot schedule a use of them above tmem_wait_ld()
__device__ __forceinline__ void tmem_pin4(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]));
}

// ---- tf32 hi/lo split -----------------------------------------------------------------------
// hi = round-to-nearest tf32(a) (exactly representable, so the tensor core's own fp32->tf32
// conversion cannot change it); lo = a - hi is exact in fp32 and has <= 13 significant bits.
__device__ __forceinline__ float tf32_hi(float a) {
  // == cvt.rna.tf32.f32 (nearest, ties away from zero) for finite inputs, as two integer ops: the cvt
  // expands to ~5 SASS instructions (inf/nan handling) and the converter warps run it per element.
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}

// byte offset of 16-byte chunk `c` (0..7) of row `row` inside a [rows x 128B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int c) {
  return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
}

}  // namespace tc
}  // namespace vmtl

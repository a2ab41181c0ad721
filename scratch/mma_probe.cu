// Probe: tcgen05.mma kind::tf32 issue cost / completion time / TMEM layout for a few shapes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision_mtl_b200/csrc -I include scratch/mma_probe.cu -o scratch/mma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tcgen05.cuh"
using namespace vmtl::tc;

__device__ __forceinline__ void mma12_block(uint32_t tmem, uint64_t aH, uint64_t aL, uint64_t bH, uint64_t bL,
                                            uint32_t idesc, uint32_t first) {
  asm volatile(
      "{\n\t.reg .pred p0, p1;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "setp.ne.b32 p0, %6, 0;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
      "mov.b64 ah, %1; mov.b64 al, %2; mov.b64 bh, %3; mov.b64 bl, %4;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %5, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %5, p1;\n\t"
      "add.s64 ah, ah, 2; add.s64 al, al, 2; add.s64 bh, bh, 2; add.s64 bl, bl, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %5, p1;\n\t"
      "add.s64 ah, ah, 2; add.s64 al, al, 2; add.s64 bh, bh, 2; add.s64 bl, bl, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %5, p1;\n\t"
      "add.s64 ah, ah, 2; add.s64 al, al, 2; add.s64 bh, bh, 2; add.s64 bl, bl, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %5, p1;\n\t}"
      ::"r"(tmem), "l"(aH), "l"(aL), "l"(bH), "l"(bL), "r"(idesc), "r"(first)
      : "memory");
}

template <int M, int N>
__global__ void __launch_bounds__(128, 1) probe_fused(long long* times, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = idesc_tf32(M, N, 0, 0);
  long long t_issue = 0, t_done = 0;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
#pragma unroll
      for (int atom = 0; atom < 4; ++atom)
        mma12_block(tmem, smem_desc_sw128(a, 16, 1024), smem_desc_sw128(a + 4096, 16, 1024),
                    smem_desc_sw128(b, 16, 1024), smem_desc_sw128(b + 4096, 16, 1024), idesc, atom > 0);
      mma_commit(smem_u32(&bar));
      const long long t1 = clock64();
      mbar_wait(smem_u32(&bar), rep & 1);
      const long long t2 = clock64();
      t_issue += t1 - t0;
      t_done += t2 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { times[0] = t_issue / reps; times[1] = t_done / reps; }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N>
void run_fused(const char* name) {
  long long* times; cudaMalloc(&times, 16);
  auto k = probe_fused<M, N>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  k<<<1, 128, 60 * 1024>>>(times, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, times, 16, cudaMemcpyDeviceToHost);
  printf("%-28s err=%d  issue %lld cyc (%.1f/MMA)  done %lld cyc (%.1f/MMA)\n", name, (int)e, h[0], h[0] / 48.0, h[1], h[1] / 48.0);
}

template <int M, int N, int NMMA>
__global__ void __launch_bounds__(128, 1) probe(float* out, long long* times, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // [128 rows][128B] one atom (K = 32 floats), zero padded
  uint8_t* sB = smem + 16384;         // [256 rows][128B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  // A[r][0] = r+1 ; B[n][0] = n+1  (element k=0 lives in chunk 0 -> physical chunk (0 ^ (r&7)))
  for (int r = threadIdx.x; r < 128; r += blockDim.x) *(float*)(sA + sw128_off(r, 0)) = (float)(r + 1);
  for (int n = threadIdx.x; n < 256; n += blockDim.x) *(float*)(sB + sw128_off(n, 0)) = (float)(n + 1);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = idesc_tf32(M, N, 0, 0);
  long long t_issue = 0, t_done = 0;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      const uint32_t a = smem_u32(sA), b = smem_u32(sB);
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        const uint32_t ko = (i & 3) * 32;
        mma_tf32(tmem, smem_desc_sw128(a + ko, 16, 1024), smem_desc_sw128(b + ko, 16, 1024), idesc, i > 0);
      }
      mma_commit(smem_u32(&bar));
      const long long t1 = clock64();
      mbar_wait(smem_u32(&bar), rep & 1);
      const long long t2 = clock64();
      t_issue += t1 - t0;
      t_done += t2 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { times[0] = t_issue / reps; times[1] = t_done / reps; }
  // dump TMEM: thread (warp w, lane l) reads lane 32*w + l, columns 0..15 and N-16..N-1
  tc_fence_after_sync();
  float v[16];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  tmem_ld16(tmem + ((uint32_t)(w * 32) << 16), v);
  for (int j = 0; j < 16; ++j) out[(w * 32 + l) * 32 + j] = v[j];
  tmem_ld16(tmem + ((uint32_t)(w * 32) << 16) + (N >= 32 ? N - 16 : 0), v);
  for (int j = 0; j < 16; ++j) out[(w * 32 + l) * 32 + 16 + j] = v[j];
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int M, int N, int NMMA>
void run(const char* name, bool dump) {
  float* out; long long* times;
  cudaMalloc(&out, 128 * 32 * 4); cudaMalloc(&times, 16);
  cudaMemset(out, 0, 128 * 32 * 4);
  auto k = probe<M, N, NMMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  k<<<1, 128, 60 * 1024>>>(out, times, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; std::vector<float> o(128 * 32);
  cudaMemcpy(h, times, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  printf("%-28s err=%d  issue %lld cyc (%.1f/MMA)  done %lld cyc (%.1f/MMA)\n", name, (int)e, h[0], (double)h[0] / NMMA,
         h[1], (double)h[1] / NMMA);
  if (dump) {  // D[r][n] = NMMA/4 * (r+1)(n+1) (only k-step 0 has data; i&3==0 quarter of the MMAs)
    const float scale = (float)((NMMA + 3) / 4);
    for (int lane : {0, 1, 15, 16, 31, 32, 33, 47, 48, 63, 64, 65, 96, 127}) {
      printf("  lane %3d: col0 -> row %6.1f  col1/col0 = %.2f   col[N-1]/col0 = %.1f\n", lane, o[lane * 32] / scale,
             o[lane * 32] != 0 ? o[lane * 32 + 1] / o[lane * 32] : 0.f, o[lane * 32] != 0 ? o[lane * 32 + 31] / o[lane * 32] : 0.f);
    }
  }
}

int main() {
  run_fused<128, 32>("fused-asm M128 N32 x48");
  run_fused<128, 64>("fused-asm M128 N64 x48");
  run<128, 32, 48>("M128 N32 x48", false);
  run<128, 32, 16>("M128 N32 x16", false);
  run<128, 64, 48>("M128 N64 x48", false);
  run<128, 128, 48>("M128 N128 x48", false);
  run<128, 256, 48>("M128 N256 x48", false);
  run<64, 256, 48>("M64 N256 x48", false);
  run<64, 64, 48>("M64 N64 x48", false);
  return 0;
}

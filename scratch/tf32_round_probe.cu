// Does tcgen05.mma kind::tf32 TRUNCATE or ROUND the fp32 operand bits it is given?
// D = A * B with A[r][0] = test value, B[n][0] = 1, everything else 0 (SS form, K-major SW128, M=128, N=32).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision_mtl_b200/csrc -o scratch/tf32_round_probe scratch/tf32_round_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include "tcgen05.cuh"
using namespace vmtl::tc;

__global__ void __launch_bounds__(128, 1) probe(const float* vals, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;           // [128 rows][128 B]
  uint8_t* sB = smem + 16384;   // [32 rows][128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  *(float*)(sA + sw128_off(threadIdx.x, 0)) = vals[threadIdx.x];
  if (threadIdx.x < 32) *(float*)(sB + sw128_off(threadIdx.x, 0)) = 1.0f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 32);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mma_tf32(tmem, smem_desc_sw128(smem_u32(sA), 16, 1024), smem_desc_sw128(smem_u32(sB), 16, 1024),
             idesc_tf32(128, 32, 0, 0), 0);
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  __syncthreads();
  tc_fence_after_sync();
  float v[16];
  tmem_ld16(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  out[threadIdx.x] = v[0];
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> h(128, 0.f);
  // mantissa patterns below the tf32 cut (13 dropped bits): 0x0fff (just under half), 0x1000 (half), 0x1001, 0x1fff
  const uint32_t pats[8] = {0x0000, 0x0fff, 0x1000, 0x1001, 0x1800, 0x1fff, 0x3000, 0x2fff};
  for (int i = 0; i < 8; ++i) { uint32_t b = 0x3f800000u | pats[i]; memcpy(&h[i], &b, 4); b |= 0x80000000u; memcpy(&h[8 + i], &b, 4); }
  float *dv, *dout;
  cudaMalloc(&dv, 512); cudaMalloc(&dout, 512);
  cudaMemcpy(dv, h.data(), 512, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  probe<<<1, 128, 40 * 1024>>>(dv, dout);
  printf("err=%d\n", (int)cudaDeviceSynchronize());
  std::vector<float> o(128);
  cudaMemcpy(o.data(), dout, 512, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 16; ++i) {
    uint32_t bi, bo; memcpy(&bi, &h[i], 4); memcpy(&bo, &o[i], 4);
    printf("in 0x%08x -> out 0x%08x  (%s)\n", bi, bo, (bo == (bi & 0xffffe000u)) ? "truncated" : ((bo == ((bi + 0x1000u) & 0xffffe000u)) ? "rounded (rna)" : "other"));
  }
  return 0;
}

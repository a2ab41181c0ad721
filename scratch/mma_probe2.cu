// Probe 2: A operand from TMEM (tcgen05.st + tcgen05.mma [d],[a],bdesc), layout + timing.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tcgen05.cuh"
using namespace vmtl::tc;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int N, int NMMA>
__global__ void __launch_bounds__(128, 1) probe(float* out, long long* times, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;  // [256 rows][128B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  // B[n][k] = (n+1) * (k==0 ? 1 : (k==5 ? 100 : 0))
  for (int n = threadIdx.x; n < 256; n += blockDim.x) {
    *(float*)(sB + sw128_off(n, 0)) = (float)(n + 1);
    *(float*)(sB + sw128_off(n, 1) + 4) = 100.f * (float)(n + 1);  // k = 5
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int row = w * 32 + l;
  // A region at columns 256.. : A[row][k]: k=0 -> row+1 ; k=5 -> 0.001*(row+1) ; others 0  (16 columns = 2 k-steps)
  float a[16];
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  a[0] = (float)(row + 1);
  a[5] = 0.001f * (float)(row + 1);
  const uint32_t a_col = 256;
  tmem_st16(tmem + ((uint32_t)(w * 32) << 16) + a_col, a);
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
  long long t_issue = 0, t_done = 0;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after_sync();
      const long long t0 = clock64();
      const uint32_t b = smem_u32(sB);
#pragma unroll
      for (int i = 0; i < NMMA; ++i)
        mma_tf32_ts(tmem, tmem + a_col, smem_desc_sw128(b, 16, 1024), idesc, i > 0);  // k-step 0 each time
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
  tc_fence_after_sync();
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(w * 32) << 16), v);
  for (int j = 0; j < 16; ++j) out[row * 16 + j] = v[j];
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NMMA>
void run(const char* name) {
  float* out; long long* times;
  cudaMalloc(&out, 128 * 16 * 4); cudaMalloc(&times, 16);
  auto k = probe<N, NMMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  k<<<1, 128, 40 * 1024>>>(out, times, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; std::vector<float> o(128 * 16);
  cudaMemcpy(h, times, 16, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), out, 128 * 16 * 4, cudaMemcpyDeviceToHost);
  printf("%-24s err=%d(%s) issue %lld cyc (%.1f/MMA) done %lld cyc (%.1f/MMA)\n", name, (int)e, cudaGetErrorString(e), h[0],
         (double)h[0] / NMMA, h[1], (double)h[1] / NMMA);
  // expected D[r][n] = NMMA * (r+1)(n+1) * (1 + 0.001*100) = NMMA*(r+1)(n+1)*1.1 if k=5 pairs with k=5
  for (int r : {0, 1, 31, 32, 64, 127})
    printf("   row %3d: D[r][0]/NMMA = %9.3f (expect %.3f)  D[r][1]/D[r][0] = %.2f\n", r, o[r * 16] / NMMA, 1.1 * (r + 1),
           o[r * 16] != 0 ? o[r * 16 + 1] / o[r * 16] : 0.f);
}

int main() {
  run<32, 48>("TS M128 N32 x48");
  run<64, 48>("TS M128 N64 x48");
  run<128, 48>("TS M128 N128 x48");
  return 0;
}

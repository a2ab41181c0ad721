// Probe 3: tcgen05.mma kind::tf32 throughput on ALL SMs with concurrent TMEM traffic.
//   warps 0-3  : tcgen05.st loops (converter stand-in)   if n_st > 0 (warps 0..n_st-1)
//   warps 4-7  : tcgen05.ld loops (epilogue stand-in)    if n_ld > 0
//   warp 8     : MMA issuer, TS (A in TMEM) or SS (A in smem)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vision_mtl_b200/csrc -o scratch/mma_probe3 scratch/mma_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tcgen05.cuh"
using namespace vmtl::tc;

template <int WAITK>
__device__ __forceinline__ void wait_k(uint32_t bar, uint32_t parity) {
  if (WAITK == 0) { mbar_wait(bar, parity); return; }
  uint32_t done;
  do {
    if (WAITK == 1)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

template <int N, bool TS, int WAITK = 0, int MODE = 0, int MM = 128>
__global__ void __launch_bounds__(288, 1) probe(long long* times, int nmma, int n_st, int n_ld, int st_gap, int group) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i & 63);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  long long cnt = 0;
  if (w < 4) {
    if (w < n_st) {
      float a[16];
      for (int j = 0; j < 16; ++j) a[j] = 0.01f * (l + j);
      const uint32_t base = tmem + ((uint32_t)(w * 32) << 16) + 384;  // columns 384..447: never read by the MMAs
      while (!stop) {
        tmem_st16(base, a); tmem_st16(base + 16, a); tmem_st16(base + 32, a); tmem_st16(base + 48, a);
        tmem_wait_st();
        ++cnt;
        if (st_gap) __nanosleep(st_gap);
      }
      if (l == 0) times[8 + blockIdx.x * 16 + w] = cnt;
    }
  } else if (w < 8) {
    if (w - 4 < n_ld) {
      float v[16], acc = 0.f;
      const uint32_t base = tmem + ((uint32_t)((w & 3) * 32) << 16) + 448;  // columns 448..511
      while (!stop) {
        tmem_ld16(base, v); acc += v[0];
        tmem_ld16(base + 16, v); acc += v[3];
        ++cnt;
        if (st_gap) __nanosleep(st_gap);
      }
      if (l == 0) times[8 + blockIdx.x * 16 + w] = cnt + (acc == 123.f);
    }
  } else if (l == 0) {
    constexpr uint32_t idesc = MODE == 2 ? idesc_tf32(MM, N, 1, 1) : (MODE == 3 ? idesc_tf32(MM, N, 0, 1) : idesc_tf32(MM, N, 0, 0));
    const uint32_t b = smem_u32(smem);          // B: [256 rows][128 B]
    const uint32_t a_s = smem_u32(smem + 65536);  // A (SS)
    tc_fence_after_sync();
    const long long t0 = clock64();
    int ph = 0;
    for (int i = 0; i < nmma; i += group) {
      for (int j = 0; j < group; ++j) {
        const uint32_t ks = (uint32_t)(j & 3);
        const uint64_t db = smem_desc_sw128(b + ks * 32, 16, 1024);
        if (MODE == 2)
          mma_tf32(tmem + (uint32_t)((i / group) & 1) * 128, smem_desc_mn_tf32(a_s + ks * 1024, 16384, 512),
                   smem_desc_mn_tf32(b + ks * 1024, 16384, 512), idesc, j > 0);
        else if (MODE == 3)
          mma_tf32_ts(tmem + (uint32_t)((i / group) & 1) * 128, tmem + 256 + ks * 8, smem_desc_mn_tf32(b + ks * 1024, 16384, 512), idesc, j > 0);
        else if (TS)
          mma_tf32_ts(tmem + (uint32_t)((i / group) & 1) * 128, tmem + 256 + ks * 8, db, idesc, j > 0);
        else
          mma_tf32(tmem + (uint32_t)((i / group) & 1) * 128, smem_desc_sw128(a_s + ks * 32, 16, 1024), db, idesc, j > 0);
      }
      if (i) { wait_k<WAITK>(smem_u32(&bar), ph); ph ^= 1; }  // previous group done (double-buffered accumulator)
      mma_commit(smem_u32(&bar));
    }
    wait_k<WAITK>(smem_u32(&bar), ph);
    const long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) times[0] = t1 - t0;
    times[2048 + blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, int WAITK = 0, int MODE = 0, int MM = 128>
void run(const char* name, int n_st, int n_ld, int gap, int group, int grid = 148) {
  long long* times;
  cudaMalloc(&times, 8192 * 8);
  cudaMemset(times, 0, 8192 * 8);
  auto k = probe<N, TS, WAITK, MODE, MM>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  const int nmma = 8192;
  k<<<grid, 288, 140 * 1024>>>(times, nmma, n_st, n_ld, gap, group);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(8192);
  cudaMemcpy(h.data(), times, 8192 * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[2048 + i] / grid;
  long long st = 0, ld = 0;
  for (int w = 0; w < 4; ++w) { st += h[8 + w]; ld += h[8 + 4 + w]; }
  printf("%-30s grid %3d st_warps %d ld_warps %d gap %4d group %2d: %7.1f cyc/MMA  (err %d) st iters/warp %lld ld iters/warp %lld\n",
         name, grid, n_st, n_ld, gap, group, mean / nmma, (int)e, n_st ? st / n_st : 0, n_ld ? ld / n_ld : 0);
  cudaFree(times);
}

int main() {
  run<64, true, 2, 0>("TS Kmaj N64 g32", 0, 0, 0, 32);
  run<64, false, 2, 1>("SS Kmaj N64 g32", 0, 0, 0, 32);
  run<64, false, 2, 2>("SS MN/MN M128 N64 g32", 0, 0, 0, 32);
  run<64, false, 2, 2, 64>("SS MN/MN M64 N64 g32", 0, 0, 0, 32);
  run<32, false, 2, 2>("SS MN/MN M128 N32 g32", 0, 0, 0, 32);
  run<128, false, 2, 2>("SS MN/MN M128 N128 g32", 0, 0, 0, 32);
  run<64, true, 2, 3>("TS MN-B N64 g32", 0, 0, 0, 32);
  run<128, true, 2, 3>("TS MN-B N128 g32", 0, 0, 0, 32);
  run<32, true, 2, 3>("TS MN-B N32 g32", 0, 0, 0, 32);
  run<64, true, 2, 3>("TS MN-B N64 g32 +st+ld gap", 4, 4, 500, 32);
  run<64, false, 2, 2>("SS MN/MN M128 N64 g32 +st+ld gap", 4, 4, 500, 32);
  run<64, true, 2, 0>("TS Kmaj N64 g32 +st+ld gap", 4, 4, 500, 32);
  return 0;
}

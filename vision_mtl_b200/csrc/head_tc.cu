// Segmentation head (1x1 conv, 32 -> C <= 32) fused with softmax / argmax / cross-entropy / confusion,
// projection on tensor cores: TMA -> smem -> TMEM (tf32 hi/lo) -> tcgen05 -> per-pixel epilogue.
//
// Replaces nn.Conv2d(32,C,1) + softmax + argmax + CrossEntropyLoss + the torchmetrics statistics
// (mtan_model.py:367-376,401-404; lit_module.py:31,123,137-138,109-111).
//
// Why tensor cores here: the CUDA-core version (head_loss.cu) spends ~1450 instructions per pixel
// (608 FMA + 152 shared loads + accurate exp/log) and sits at 20 % of the HBM roofline.  A 128-pixel
// tile is a [128 x 32] x [32 x 32] contraction = 8 tcgen05.mma (3xTF32 with the stacked W operand:
// A_hi @ [W_hi;W_lo]^T and A_lo @ W_hi^T per K-step) -- ~360 tensor-pipe cycles against the ~400
// cycles the tile's 17 KB take at the roofline -- and tcgen05.ld hands every epilogue thread exactly
// one pixel's logits, which is the shape softmax / argmax / CE want.
//
// Roles (27 warps): 0-7 two converter groups (one pixel row per thread: smem -> hi/lo -> TMEM; alternate
// tiles), 8-23 four epilogue groups (group g owns accumulator buffer g = tiles g, g+4, ...), 24 TMA
// producer, 25-26 two MMA issuers (alternate tiles).  Every role is a serial chain of ~1-1.6k cycles per
// tile (mbarrier round trips, 8 blocking MMA issues of ~60 cycles, a ~300-instruction epilogue), against
// ~750 cycles a tile's bytes take at the HBM roofline -- hence several instances of each role in flight.
// Feature tile = ONE 128B-swizzled atom of 16 KB; 8 stages -> 128 KB in flight per SM.
#include <cuda.h>
#include <math.h>

#include "head_internal.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace vmtl {

using namespace tc;

constexpr int kHtThreads = 27 * 32;
constexpr int kHtTile = 128;
constexpr int kHtStages = 8;
constexpr int kHtStageBytes = kHtTile * 128;  // 16 KB

struct HtSmem {
  static constexpr int kW = kHtStages * kHtStageBytes;  // stacked W operand: [64 rows x 128 B]
  static constexpr int kMisc = kW + 64 * 128;
  static constexpr int kConf = kMisc + 512;             // uint32 [32*32]
  static constexpr int kBytes = kConf + 32 * 32 * 4 + 1024;
};

__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

#ifdef VMTL_HT_PROF
__device__ long long g_ht_prof[148][16];
#define PROF_T0(v) const long long v = clock64()
#define PROF_ACC(acc, v) acc += clock64() - v
#define PROF_DECL(...) long long __VA_ARGS__
#define PROF_OUT(slot, val) if (lane == 0) g_ht_prof[blockIdx.x][slot] = (val)
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PROF_STAMP(slot) if (threadIdx.x == 0) g_ht_prof[blockIdx.x][slot] = gtimer()
#else
#define PROF_T0(v)
#define PROF_ACC(acc, v)
#define PROF_DECL(...)
#define PROF_OUT(slot, val)
#define PROF_STAMP(slot)
#endif

template <int CPAD>
__global__ void __launch_bounds__(kHtThreads, 1)
    head_ce_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_f, const float* __restrict__ W,
                          const float* __restrict__ bias, const int64_t* __restrict__ target, int64_t P, int C,
                          int64_t ignore_index, double* __restrict__ partial, uint8_t* __restrict__ pred,
                          unsigned long long* __restrict__ conf) {
  using L = HtSmem;
  constexpr int S = kHtStages;
  extern __shared__ uint8_t smem_raw[];
  PROF_STAMP(13);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sW = smem + L::kW;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);   // full[8] empty[8] aready[4] amma[4] dfull[4] dfree[4]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 288);
  float* s_bias = reinterpret_cast<float*>(smem + L::kMisc + 320);  // [32]
  unsigned int* s_conf = reinterpret_cast<unsigned int*>(smem + L::kConf);
  __shared__ double s_part[16][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_empty = [&](int s) { return bar0 + 64u + 8u * (uint32_t)s; };
  auto bar_aready = [&](int b) { return bar0 + 128u + 8u * (uint32_t)b; };
  auto bar_dfull = [&](int b) { return bar0 + 192u + 8u * (uint32_t)b; };
  auto bar_dfree = [&](int b) { return bar0 + 224u + 8u * (uint32_t)b; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_aready(i), 128);
      mbar_init(bar_dfull(i), 1);
      mbar_init(bar_dfree(i), 128);
    }
    fence_mbar_init();
  }
  if (warp == 25) tmem_alloc(smem_u32(s_tmem), 512);
  if (warp == 24 && lane == 0) tma_prefetch_desc(&tmap_f);
  for (int i = threadIdx.x; i < 32; i += kHtThreads) s_bias[i] = i < C ? bias[i] : 0.f;
  for (int i = threadIdx.x; i < 32 * 32; i += kHtThreads) s_conf[i] = 0u;
  // stacked W operand: row n (< 32) = W_hi[n], row 32 + n = W_lo[n]; classes >= C are zero rows
  for (int e = threadIdx.x; e < 32 * 32; e += kHtThreads) {
    const int n = e >> 5, k = e & 31;
    const float w = n < C ? W[n * 32 + k] : 0.f;
    const float hi = tf32_hi(w);
    *reinterpret_cast<float*>(sW + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + ((k & 3) << 2)) = hi;
    *reinterpret_cast<float*>(sW + (32 + n) * 128 + (((k >> 2) ^ ((32 + n) & 7)) << 4) + ((k & 3) << 2)) = w - hi;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_d0 = tmem_base + 256;
  PROF_STAMP(14);

  const int64_t ntiles = (P + kHtTile - 1) / kHtTile;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  double loss_acc = 0.0, n_acc = 0.0;

  if (warp < 8) {
    // ------------------------------------------------------------------ converters (thread = pixel row;
    // two groups of 4 warps take alternate tiles)
    const int row = (warp & 3) * 32 + lane;
    PROF_DECL(w_full = 0, w_amma = 0, w_st = 0);
    PROF_T0(t_begin);
    for (int64_t it = warp >> 2; it < nitems; it += 2) {
      const int s = (int)(it % S), ab = (int)(it & 3);
      PROF_T0(t0);
      mbar_wait(bar_full(s), (uint32_t)((it / S) & 1));
      PROF_ACC(w_full, t0);
      const uint8_t* st = smem + s * kHtStageBytes;
      float4 c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = *reinterpret_cast<const float4*>(st + sw128_off(row, j));
      mbar_arrive(bar_empty(s));
      if (it >= 4) {
        PROF_T0(t1);
        mbar_wait(bar_dfull(ab), (uint32_t)(((it >> 2) - 1) & 1));  // tile it-4's MMAs have read this A buffer
        PROF_ACC(w_amma, t1);
        tc_fence_after_sync();
      }
      const uint32_t ta = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)(ab * 64);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = c[g * 4 + j];
          hi[4 * j] = tf32_hi(a.x); hi[4 * j + 1] = tf32_hi(a.y); hi[4 * j + 2] = tf32_hi(a.z); hi[4 * j + 3] = tf32_hi(a.w);
          lo[4 * j] = a.x - hi[4 * j]; lo[4 * j + 1] = a.y - hi[4 * j + 1];
          lo[4 * j + 2] = a.z - hi[4 * j + 2]; lo[4 * j + 3] = a.w - hi[4 * j + 3];
        }
        tmem_st16(ta + g * 16, hi);
        tmem_st16(ta + 32 + g * 16, lo);
      }
      PROF_T0(t2);
      tmem_wait_st();
      PROF_ACC(w_st, t2);
      tc_fence_before_sync();
      mbar_arrive(bar_aready(ab));
    }
#ifdef VMTL_HT_PROF
    if (warp == 0) {
      PROF_OUT(0, clock64() - t_begin);
      PROF_OUT(1, w_full);
      PROF_OUT(2, w_amma);
      PROF_OUT(3, w_st);
    }
#endif
  } else if (warp < 24) {
    // ------------------------------------------------------------------ epilogue groups (tile it -> group it % 4)
    const int g = (warp - 8) >> 2, quad = warp & 3;
    const uint32_t taddr = tmem_d0 + (((uint32_t)quad * 32) << 16) + (uint32_t)(g * 64);
    PROF_DECL(w_dfull = 0);
    PROF_T0(t_begin);
    auto pixel_of = [&](int64_t it) { return (blockIdx.x + it * gridDim.x) * kHtTile + quad * 32 + lane; };
    // the label is fetched one tile ahead: its DRAM latency would otherwise sit in this group's serial chain
    int64_t t_next = (g < nitems && pixel_of(g) < P) ? __ldg(target + pixel_of(g)) : ignore_index;
    for (int64_t it = g; it < nitems; it += 4) {
      const int64_t p = pixel_of(it);
      const int64_t t = t_next;
      t_next = (it + 4 < nitems && pixel_of(it + 4) < P) ? __ldg(target + pixel_of(it + 4)) : ignore_index;
      PROF_T0(t0);
      mbar_wait(bar_dfull(g), (uint32_t)((it >> 2) & 1));
      PROF_ACC(w_dfull, t0);
      tc_fence_after_sync();
      float l[CPAD], l2[CPAD];
      tmem_ld16_nowait(taddr, l);            // A_hi W_hi + A_lo W_hi
      tmem_ld16_nowait(taddr + 32, l2);      // A_hi W_lo
      if (CPAD == 20) {
        tmem_ld4_nowait(taddr + 16, l + 16);
        tmem_ld4_nowait(taddr + 48, l2 + 16);
      } else if (CPAD == 32) {
        tmem_ld16_nowait(taddr + 16, l + 16);
        tmem_ld16_nowait(taddr + 48, l2 + 16);
      }
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < CPAD; c += 4) {
        tmem_pin4(l + c);
        tmem_pin4(l2 + c);
      }
      tc_fence_before_sync();
      mbar_arrive(bar_dfree(g));
#pragma unroll
      for (int c = 0; c < CPAD; ++c) l[c] = l[c] + l2[c] + s_bias[c];
      // classes >= C exist only in the last few columns of a CPAD bucket (16: C<=16, 20: 17..20, 32: 21..32)
      constexpr int CMIN = CPAD == 16 ? 1 : (CPAD == 20 ? 17 : 21);
#pragma unroll
      for (int c = CMIN; c < CPAD; ++c) l[c] = c < C ? l[c] : -INFINITY;
      float m = l[0];
      int arg = 0;
#pragma unroll
      for (int c = 1; c < CPAD; ++c)
        if (l[c] > m) {
          m = l[c];
          arg = c;
        }
      const float mneg = -m * kLog2e;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) sum += fast_ex2(fmaf(l[c], kLog2e, mneg));
      // l[target]: two-level select (3 per group of 4 + 2 per group) instead of a compare per class
      const int ti = (int)t;
      const bool q1 = (ti & 3) == 1, q2 = (ti & 3) == 2, q3 = (ti & 3) == 3;
      float lt = 0.f;
#pragma unroll
      for (int gq = 0; gq < CPAD / 4; ++gq) {
        float v = l[4 * gq];
        v = q1 ? l[4 * gq + 1] : v;
        v = q2 ? l[4 * gq + 2] : v;
        v = q3 ? l[4 * gq + 3] : v;
        lt = (ti >> 2) == gq ? v : lt;
      }
      if (p < P) {
        if (pred) pred[p] = (uint8_t)arg;
        if (t != ignore_index && (uint64_t)t < (uint64_t)C) {
          loss_acc += (double)(fmaf(fast_lg2(sum), kLn2, m) - lt);
          n_acc += 1.0;
          if (conf) red_shared_add(smem_u32(s_conf) + 4u * (uint32_t)(ti * C + arg), 1u);
        }
      }
    }
#ifdef VMTL_HT_PROF
    if (warp == 8) {
      PROF_OUT(4, clock64() - t_begin);
      PROF_OUT(5, w_dfull);
    }
#endif
  } else if (warp == 24) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      PROF_DECL(w_empty = 0);
      PROF_T0(t_begin);
      for (int64_t it = 0; it < nitems; ++it) {
        const int s = (int)(it % S);
        PROF_T0(t0);
        if (it >= S) mbar_wait(bar_empty(s), (uint32_t)(((it / S) - 1) & 1));
        PROF_ACC(w_empty, t0);
        mbar_expect_tx(bar_full(s), (uint32_t)kHtStageBytes);
        tma_load_2d(smem_u32(smem + s * kHtStageBytes), &tmap_f, 0, (int)((blockIdx.x + it * gridDim.x) * kHtTile),
                    bar_full(s));
      }
      PROF_OUT(6, clock64() - t_begin);
      PROF_OUT(7, w_empty);
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_wide = idesc_tf32(kHtTile, 64, 0, 0);
    constexpr uint32_t idesc_n = idesc_tf32(kHtTile, 32, 0, 0);
    const uint32_t bW = smem_u32(sW);
    PROF_DECL(w_aready = 0, w_dfree = 0, w_issue = 0, w_commit = 0);
    PROF_T0(t_begin);
    for (int64_t it = warp - 25; it < nitems; it += 2) {
      const int ab = (int)(it & 3);
      PROF_T0(t0);
      mbar_wait(bar_aready(ab), (uint32_t)((it >> 2) & 1));
      PROF_ACC(w_aready, t0);
      PROF_T0(t1);
      if (it >= 4) mbar_wait(bar_dfree(ab), (uint32_t)(((it >> 2) - 1) & 1));
      PROF_ACC(w_dfree, t1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_d0 + (uint32_t)(ab * 64);
      PROF_T0(t2);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a_hi = tmem_base + (uint32_t)(ab * 64 + ks * 8);
        const uint64_t dB = smem_desc_sw128(bW + ks * 32, 16, 1024);
        mma_tf32_ts(d_tmem, a_hi, dB, idesc_wide, ks != 0);
        mma_tf32_ts(d_tmem, a_hi + 32, dB, idesc_n, 1);
      }
      PROF_ACC(w_issue, t2);
      PROF_T0(t3);
      mma_commit(bar_dfull(ab));
      PROF_ACC(w_commit, t3);
    }
    PROF_OUT(11, w_issue);
    PROF_OUT(12, w_commit);
    PROF_OUT(8, clock64() - t_begin);
    PROF_OUT(9, w_aready);
    PROF_OUT(10, w_dfree);
  }
  tc_fence_before_sync();
  __syncthreads();
  PROF_STAMP(15);
  if (warp == 25) tmem_dealloc(tmem_base, 512);
  if (conf)
    for (int i = threadIdx.x; i < C * C; i += kHtThreads) {
      const unsigned int v = s_conf[i];
      if (v) atomicAdd(conf + i, (unsigned long long)v);
    }
  // (loss sum, valid count) partial of this CTA: the 16 epilogue warps in fixed order
  loss_acc = warp_sum(loss_acc);
  n_acc = warp_sum(n_acc);
  if (warp >= 8 && warp < 24 && lane == 0) {
    s_part[warp - 8][0] = loss_acc;
    s_part[warp - 8][1] = n_acc;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < 16; ++w) s += s_part[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
#ifdef VMTL_HT_PROF
  __syncthreads();
  if (threadIdx.x == 0) g_ht_prof[blockIdx.x][3] = gtimer();
#endif
}

int head_ce_tc_fwd(const float* feat, const float* W, const float* b, const int64_t* target, int64_t P, int C,
                   int64_t ignore_index, double* partial, int max_blocks, int* grid_out, uint8_t* pred,
                   int64_t* conf, cudaStream_t st) {
  if (C > 32 || P < 1) return VMTL_EUNSUPPORTED;
  CUtensorMap tmap;
  if (!make_tmap_2d(&tmap, feat, P, 32, kHtTile)) return VMTL_ECUDA;
  const int64_t ntiles = (P + kHtTile - 1) / kHtTile;
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  if (grid > max_blocks) grid = max_blocks;
  *grid_out = grid;
  unsigned long long* cf = reinterpret_cast<unsigned long long*>(conf);
#define VMTL_HT(CP)                                                                                          \
  do {                                                                                                       \
    if (cudaFuncSetAttribute(head_ce_tc_fwd_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                             HtSmem::kBytes) != cudaSuccess)                                                 \
      return VMTL_ECUDA;                                                                                     \
    head_ce_tc_fwd_kernel<CP><<<grid, kHtThreads, HtSmem::kBytes, st>>>(tmap, W, b, target, P, C,            \
                                                                         ignore_index, partial, pred, cf);   \
  } while (0)
  if (C <= 16)
    VMTL_HT(16);
  else if (C <= 20)
    VMTL_HT(20);
  else
    VMTL_HT(32);
#undef VMTL_HT
  return launch_status();
}

}  // namespace vmtl

#ifdef VMTL_HT_PROF
extern "C" int vmtl_debug_head_prof(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, vmtl::g_ht_prof, sizeof(long long) * 148 * 16) == cudaSuccess ? 0 : 1;
}
#endif

// Backward of the fused segmentation head + cross-entropy on tensor cores.
//
// Replaces the autograd chain of nn.Conv2d(32,C,1) + CrossEntropyLoss (mtan_model.py:367-376,401-404;
// lit_module.py:31,123): per 128-pixel tile
//   (1) logits = f . W^T + b             [128 x 32] x [32 x C]      tcgen05, A = f hi/lo in TMEM
//   (2) dl = (softmax - onehot) * g/n    one pixel per thread       epilogue-1 warps
//   (3) dfeat = dl . W                   [128 x C] x [C x 32]       tcgen05, A = dl hi/lo in TMEM
//   (4) D3 += [dl_hi;dl_mid]^T . [f_hi | f_mid]   contraction over the tile's pixels in bf16x2 (each operand
//       split into two bf16 = 16 mantissa bits; the sum over >= 1e5 pixels of signed terms keeps the result
//       at ~1e-5 relative), kind::f16 has K = 16, so 8 MMAs per tile; both operands MN-major in shared
//       memory (one 128-byte row per pixel, exactly what a pixel-per-thread writer produces); the
//       accumulator stays in TMEM for all tiles of the CTA.
//   db accumulates in registers of the epilogue-1 threads.
// The CUDA-core version (head_loss.cu) spends ~3500 instructions per pixel on the three contractions
// (1824 FMA) and runs at 20 % of the HBM roofline.
//
// Roles (18 warps): 0-3 converters, 4-11 two epilogue-1 groups (tile parity), 12-15 epilogue-2 (dfeat),
// 16 TMA producer, 17 MMA issuer.  TMEM region R1[it % 3] holds f hi/lo, then the dfeat accumulator of the
// tile; R2[it % 2] holds its logits accumulator, then dl hi/lo.  Three f sets / two dl sets in shared memory:
// the converter of tile it only needs tile it-3 retired, so it runs ahead of the MMA chain
// (logits -> softmax -> dfeat + pixel contraction), whose single issuing thread is the kernel's clock
// (30 tcgen05.mma of ~50 cycles per tile).
#include <cuda.h>
#include <math.h>

#include "head_internal.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace vmtl {

using namespace tc;

constexpr int kHbThreads = 18 * 32;
constexpr int kHbTile = 128;

struct HbSmem {
  static constexpr int kSlot = kHbTile * 128;          // 16 KB: [128 rows x 32 floats]
  static constexpr int kStages = 4;
  // MN-major bf16 operand slots of the pixel contraction, one 128-byte row per pixel = 64 bf16 =
  // [hi(32) | mid(32)]: dl sets 0,1 then f sets 0,1,2.  (The M = 128 A operand spans 2 slots from a dl
  // set: the second is whatever follows and feeds accumulator rows 64..127, which nobody reads.)
  static constexpr int kOps = kStages * kSlot;
  static constexpr int kOut = kOps + 5 * kSlot;        // dfeat staging for the TMA store
  static constexpr int kW1 = kOut + kSlot;             // [W_hi ; W_lo]     rows = class,   K = feature
  static constexpr int kW2 = kW1 + 64 * 128;           // [W^T_hi ; W^T_lo] rows = feature, K = class
  static constexpr int kMisc = kW2 + 64 * 128;
  static constexpr int kBytes = kMisc + 512 + 1024;
};

#ifdef VMTL_HT_PROF
__device__ long long g_hb_prof[148][8];
#define HB_T0(v) const long long v = clock64()
#define HB_ACC(slot, v) hb_prof[slot] += clock64() - v
#else
#define HB_T0(v)
#define HB_ACC(slot, v)
#endif

template <int CPAD>
__global__ void __launch_bounds__(kHbThreads, 1)
    head_ce_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_df,
                          const float* __restrict__ W, const float* __restrict__ bias,
                          const int64_t* __restrict__ target, int64_t P, int C, int64_t ignore_index,
                          const double* __restrict__ fwd_out, const float* __restrict__ gscale, int write_dfeat,
                          float* __restrict__ partial /* [grid][CPAD][33] */) {
  using L = HbSmem;
  constexpr int S = L::kStages;
  constexpr int KS2 = CPAD == 16 ? 2 : (CPAD == 20 ? 3 : 4);  // K-steps (8 classes each) of dfeat = dl . W
  constexpr int CMIN = CPAD == 16 ? 1 : (CPAD == 20 ? 17 : 21);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer
  uint8_t* sW1 = smem + L::kW1;
  uint8_t* sW2 = smem + L::kW2;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 224);
  float* s_bias = reinterpret_cast<float*>(smem + L::kMisc + 256);  // [32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_empty = [&](int s) { return bar0 + 32u + 8u * (uint32_t)s; };
  auto bar_conv = [&](int r) { return bar0 + 64u + 8u * (uint32_t)r; };      // [3] f hi/lo in TMEM + smem
  auto bar_d1full = [&](int b) { return bar0 + 88u + 8u * (uint32_t)b; };    // [2] logits accumulator complete
  auto bar_dl = [&](int b) { return bar0 + 104u + 8u * (uint32_t)b; };        // [2] dl hi/lo in TMEM + smem
  auto bar_d2full = [&](int r) { return bar0 + 120u + 8u * (uint32_t)r; };    // [3] dfeat accumulator complete
  auto bar_d2free = [&](int r) { return bar0 + 144u + 8u * (uint32_t)r; };   // [3] ... and read out
  // pixel contraction of tile t retired: barrier t % 6 (its f set t % 3 and dl set t % 2 are free again)
  auto bar_mma3 = [&](int64_t t) { return bar0 + 168u + 8u * (uint32_t)(t % 6); };
  auto par_mma3 = [&](int64_t t) { return (uint32_t)((t / 6) & 1); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 128);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_d1full(b), 1);
      mbar_init(bar_dl(b), 128);
    }
    for (int r = 0; r < 3; ++r) {
      mbar_init(bar_conv(r), 128);
      mbar_init(bar_d2full(r), 1);
      mbar_init(bar_d2free(r), 128);
    }
    for (int t = 0; t < 6; ++t) mbar_init(bar_mma3(t), 1);
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(smem_u32(s_tmem), 512);
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_f);
    tma_prefetch_desc(&tmap_df);
  }
  for (int i = threadIdx.x; i < 32; i += kHbThreads) s_bias[i] = i < C ? bias[i] : 0.f;
  for (int e = threadIdx.x; e < 32 * 32; e += kHbThreads) {
    const int c = e >> 5, j = e & 31;  // class, feature
    const float w = c < C ? W[c * 32 + j] : 0.f;
    const float hi = tf32_hi(w), lo = w - hi;
    // W1: row = class c, K index = feature j (K-major, SWIZZLE_128B); stacked lo rows at 32 + c
    *reinterpret_cast<float*>(sW1 + c * 128 + (((j >> 2) ^ (c & 7)) << 4) + ((j & 3) << 2)) = hi;
    *reinterpret_cast<float*>(sW1 + (32 + c) * 128 + (((j >> 2) ^ ((32 + c) & 7)) << 4) + ((j & 3) << 2)) = lo;
    // W2: row = feature j, K index = class c
    *reinterpret_cast<float*>(sW2 + j * 128 + (((c >> 2) ^ (j & 7)) << 4) + ((c & 3) << 2)) = hi;
    *reinterpret_cast<float*>(sW2 + (32 + j) * 128 + (((c >> 2) ^ ((32 + j) & 7)) << 4) + ((c & 3) << 2)) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  auto R1 = [&](int r) { return tmem_base + (uint32_t)(r * 64); };        // r = tile % 3
  auto R2 = [&](int b) { return tmem_base + (uint32_t)(192 + b * 64); };  // b = tile % 2
  const uint32_t tmem_d3 = tmem_base + 320;
  auto dl_set = [&](int b) { return smem + L::kOps + b * L::kSlot; };
  auto f_set = [&](int r) { return smem + L::kOps + (2 + r) * L::kSlot; };

  const int64_t ntiles = (P + kHbTile - 1) / kHbTile;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int quad = warp & 3, row = quad * 32 + lane;
  float db_acc[CPAD];
#pragma unroll
  for (int c = 0; c < CPAD; ++c) db_acc[c] = 0.f;

  if (warp < 4) {
    // ------------------------------------------------------------------ converters (thread = pixel row)
    for (int64_t it = 0; it < nitems; ++it) {
      const int s = (int)(it % S), r = (int)(it % 3);
      mbar_wait(bar_full(s), (uint32_t)((it / S) & 1));
      const uint8_t* st = smem + s * L::kSlot;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(st + sw128_off(row, j));
      mbar_arrive(bar_empty(s));
      if (it >= 3) {  // tile it-3: its dfeat accumulator (R1[r]) is read out, its f set is consumed
        mbar_wait(bar_d2free(r), (uint32_t)(((it / 3) - 1) & 1));
        mbar_wait(bar_mma3(it - 3), par_mma3(it - 3));
        tc_fence_after_sync();
      }
      const uint32_t ta = R1(r) + (((uint32_t)quad * 32) << 16);
      uint8_t* fs = f_set(r);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = v[g * 4 + j];
          hi[4 * j] = tf32_hi(a.x); hi[4 * j + 1] = tf32_hi(a.y); hi[4 * j + 2] = tf32_hi(a.z); hi[4 * j + 3] = tf32_hi(a.w);
          lo[4 * j] = a.x - hi[4 * j]; lo[4 * j + 1] = a.y - hi[4 * j + 1];
          lo[4 * j + 2] = a.z - hi[4 * j + 2]; lo[4 * j + 3] = a.w - hi[4 * j + 3];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // features 16g + 8j .. +7 -> bf16 hi (chunk 2g + j) and mid (chunk 4 + 2g + j)
          const float x8[8] = {v[g * 4 + 2 * j].x, v[g * 4 + 2 * j].y, v[g * 4 + 2 * j].z, v[g * 4 + 2 * j].w,
                               v[g * 4 + 2 * j + 1].x, v[g * 4 + 2 * j + 1].y, v[g * 4 + 2 * j + 1].z,
                               v[g * 4 + 2 * j + 1].w};
          uint4 bh, bm;
          bf16_split8(x8, bh, bm);
          *reinterpret_cast<uint4*>(fs + sw128_off(row, 2 * g + j)) = bh;
          *reinterpret_cast<uint4*>(fs + sw128_off(row, 4 + 2 * g + j)) = bm;
        }
        tmem_st16(ta + g * 16, hi);
        tmem_st16(ta + 32 + g * 16, lo);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_conv(r));
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ epilogue-1: softmax -> dl (tile parity g)
    const int g = (warp - 4) >> 2;
    const float scale = gscale[0] / (float)fwd_out[1];
    auto pixel_of = [&](int64_t it) { return (blockIdx.x + it * gridDim.x) * kHbTile + row; };
    int64_t t_next = (g < nitems && pixel_of(g) < P) ? __ldg(target + pixel_of(g)) : ignore_index;
    const uint32_t taddr = R2(g) + (((uint32_t)quad * 32) << 16);
    uint8_t* ds = dl_set(g);
    for (int64_t it = g; it < nitems; it += 2) {
      const uint32_t k2 = (uint32_t)(it >> 1);
      const int64_t p = pixel_of(it);
      const int64_t t = t_next;
      t_next = (it + 2 < nitems && pixel_of(it + 2) < P) ? __ldg(target + pixel_of(it + 2)) : ignore_index;
      mbar_wait(bar_d1full(g), k2 & 1);
      tc_fence_after_sync();
      float l[CPAD], l2[CPAD];
      tmem_ld16_nowait(taddr, l);
      tmem_ld16_nowait(taddr + 32, l2);
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
#pragma unroll
      for (int c = 0; c < CPAD; ++c) l[c] = l[c] + l2[c] + s_bias[c];
#pragma unroll
      for (int c = CMIN; c < CPAD; ++c) l[c] = c < C ? l[c] : -INFINITY;
      float m = l[0];
#pragma unroll
      for (int c = 1; c < CPAD; ++c) m = fmaxf(m, l[c]);
      const float mneg = -m * kLog2e;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        l[c] = fast_ex2(fmaf(l[c], kLog2e, mneg));
        sum += l[c];
      }
      const bool valid = p < P && t != ignore_index && (uint64_t)t < (uint64_t)C;
      const float k = valid ? scale * __fdividef(1.f, sum) : 0.f;
      const int ti = valid ? (int)t : -1;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        l[c] = fmaf(l[c], k, c == ti ? -scale : 0.f);  // (softmax - onehot) * g / n_valid ; 0 for invalid pixels
        db_acc[c] += l[c];
      }
      if (it >= 2) {  // dl set g still feeds tile it-2's pixel contraction
        mbar_wait(bar_mma3(it - 2), par_mma3(it - 2));
        tc_fence_after_sync();
      }
      // dl: tf32 hi/lo to TMEM (A operand of dfeat = dl . W, columns = classes) and bf16 hi/mid to smem
      // (MN-major row of this pixel for the contraction over pixels)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float d16[16], hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c = h * 16 + e;
          d16[e] = c < CPAD ? l[c < CPAD ? c : 0] : 0.f;
          hi[e] = tf32_hi(d16[e]);
          lo[e] = d16[e] - hi[e];
        }
        tmem_st16(taddr + h * 16, hi);
        tmem_st16(taddr + 32 + h * 16, lo);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 bh, bm;
          bf16_split8(d16 + 8 * j, bh, bm);
          *reinterpret_cast<uint4*>(ds + sw128_off(row, 2 * h + j)) = bh;
          *reinterpret_cast<uint4*>(ds + sw128_off(row, 4 + 2 * h + j)) = bm;
        }
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_dl(g));
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------ epilogue-2: dfeat rows -> TMA store
    uint8_t* stage = smem + L::kOut;
    const bool leader = threadIdx.x == 12 * 32;
    for (int64_t it = 0; it < nitems; ++it) {
      const int r = (int)(it % 3);
      mbar_wait(bar_d2full(r), (uint32_t)((it / 3) & 1));
      tc_fence_after_sync();
      const uint32_t taddr = R1(r) + (((uint32_t)quad * 32) << 16);
      float a[32], a2[32];
      tmem_ld16_nowait(taddr, a);
      tmem_ld16_nowait(taddr + 16, a + 16);
      tmem_ld16_nowait(taddr + 32, a2);
      tmem_ld16_nowait(taddr + 48, a2 + 16);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        tmem_pin4(a + c);
        tmem_pin4(a2 + c);
      }
      tc_fence_before_sync();
      mbar_arrive(bar_d2free(r));
      if (write_dfeat) {
        if (leader) tma_store_wait_read();  // the previous tile's store has read the staging buffer
        named_barrier_sync(2, 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stage + sw128_off(row, j)) =
              make_float4(a[4 * j] + a2[4 * j], a[4 * j + 1] + a2[4 * j + 1], a[4 * j + 2] + a2[4 * j + 2],
                          a[4 * j + 3] + a2[4 * j + 3]);
        fence_proxy_async_smem();
        named_barrier_sync(2, 128);
        if (leader) {
          tma_store_2d(&tmap_df, 0, (int)((blockIdx.x + it * gridDim.x) * kHbTile), smem_u32(stage));
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  } else if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int64_t it = 0; it < nitems; ++it) {
        const int s = (int)(it % S);
        if (it >= S) mbar_wait(bar_empty(s), (uint32_t)(((it / S) - 1) & 1));
        mbar_expect_tx(bar_full(s), (uint32_t)L::kSlot);
        tma_load_2d(smem_u32(smem + s * L::kSlot), &tmap_f, 0, (int)((blockIdx.x + it * gridDim.x) * kHbTile),
                    bar_full(s));
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer (logits run one tile ahead)
    constexpr uint32_t idesc_wide = idesc_tf32(kHbTile, 64, 0, 0);
    constexpr uint32_t idesc_n = idesc_tf32(kHbTile, 32, 0, 0);
    constexpr uint32_t idesc_px = idesc_bf16(128, 64, 1, 1);  // bf16, both operands MN-major (rows = pixels)
    const uint32_t bW1 = smem_u32(sW1), bW2 = smem_u32(sW2);
#ifdef VMTL_HT_PROF
    long long hb_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long hb_begin = clock64();
#endif
    for (int64_t i = 0; i <= nitems; ++i) {
      if (i < nitems) {
        const int b = (int)(i & 1), r = (int)(i % 3);
        HB_T0(t0);
        mbar_wait(bar_conv(r), (uint32_t)((i / 3) & 1));
        HB_ACC(0, t0);
        tc_fence_after_sync();
        HB_T0(t1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dB = smem_desc_sw128(bW1 + ks * 32, 16, 1024);
          mma_tf32_ts(R2(b), R1(r) + ks * 8, dB, idesc_wide, ks != 0);
          mma_tf32_ts(R2(b), R1(r) + 32 + ks * 8, dB, idesc_n, 1);
        }
        mma_commit(bar_d1full(b));
        HB_ACC(1, t1);
      }
      if (i >= 1) {
        const int64_t j = i - 1;
        const int b = (int)(j & 1), r = (int)(j % 3);
        HB_T0(t2);
        mbar_wait(bar_dl(b), (uint32_t)((j >> 1) & 1));
        HB_ACC(2, t2);
        tc_fence_after_sync();
        HB_T0(t3);
#pragma unroll
        for (int ks = 0; ks < KS2; ++ks) {
          const uint64_t dB = smem_desc_sw128(bW2 + ks * 32, 16, 1024);
          mma_tf32_ts(R1(r), R2(b) + ks * 8, dB, idesc_wide, ks != 0);
          mma_tf32_ts(R1(r), R2(b) + 32 + ks * 8, dB, idesc_n, 1);
        }
        mma_commit(bar_d2full(r));
        HB_ACC(3, t3);
        HB_T0(t4);
        const uint32_t oA = smem_u32(dl_set(b)), oB = smem_u32(f_set(r));
#pragma unroll
        for (int ks = 0; ks < kHbTile / 16; ++ks)  // 16 pixel rows (two 8-row swizzle groups) per MMA
          mma_bf16(tmem_d3, smem_desc_sw128(oA + ks * 2048, L::kSlot, 1024),
                   smem_desc_sw128(oB + ks * 2048, L::kSlot, 1024), idesc_px, (j | ks) != 0);
        mma_commit(bar_mma3(j));
        HB_ACC(4, t4);
      }
    }
#ifdef VMTL_HT_PROF
    hb_prof[5] = clock64() - hb_begin;
    for (int q = 0; q < 8; ++q) g_hb_prof[blockIdx.x][q] = hb_prof[q];
#endif
  }
  // ---------------------------------------------------------------------- drain dW / db partials
  if (nitems > 0 && warp == 17 && lane == 0) {
    const int64_t jl = nitems - 1;
    mbar_wait(bar_mma3(jl), par_mma3(jl));
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  float* s_red = reinterpret_cast<float*>(smem);  // [2][32][32] dW halves, then [8][CPAD] db
  float* s_db = s_red + 2 * 32 * 32;
  if (nitems > 0 && (warp == 4 || warp == 5)) {
    // D3 rows 0..31 = dl_hi^T [f_hi | f_lo], rows 32..63 = dl_lo^T [f_hi | f_lo]; thread = class row
    const uint32_t taddr = tmem_d3 + (((uint32_t)(warp - 4) * 32) << 16);
    float a[32], a2[32];
    tmem_ld16_nowait(taddr, a);
    tmem_ld16_nowait(taddr + 16, a + 16);
    tmem_ld16_nowait(taddr + 32, a2);
    tmem_ld16_nowait(taddr + 48, a2 + 16);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      tmem_pin4(a + c);
      tmem_pin4(a2 + c);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) s_red[((warp - 4) * 32 + lane) * 32 + j] = a[j] + a2[j];
  }
  if (warp >= 4 && warp < 12) {
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      const float r = warp_sum(db_acc[c]);
      if (lane == 0) s_db[(warp - 4) * CPAD + c] = r;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, 512);
  float* out = partial + (int64_t)blockIdx.x * CPAD * 33;
  for (int e = threadIdx.x; e < CPAD * 33; e += kHbThreads) {
    const int c = e / 33, j = e - c * 33;
    float r = 0.f;
    if (nitems > 0) {
      if (j < 32) {
        r = s_red[c * 32 + j] + s_red[(32 + c) * 32 + j];
      } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) r += s_db[w * CPAD + c];
      }
    }
    out[e] = r;
  }
}

int head_ce_tc_bwd(const float* feat, const float* W, const float* b, const int64_t* target, int64_t P, int C,
                   int64_t ignore_index, const double* fwd_out, const float* gscale, float* dfeat, float* partial,
                   int max_blocks, int* grid_out, cudaStream_t st) {
  if (C > 32 || P < 1) return VMTL_EUNSUPPORTED;
  CUtensorMap tmap_f, tmap_df;
  if (!make_tmap_2d(&tmap_f, feat, P, 32, kHbTile) || !make_tmap_2d(&tmap_df, dfeat ? dfeat : feat, P, 32, kHbTile))
    return VMTL_ECUDA;
  const int64_t ntiles = (P + kHbTile - 1) / kHbTile;
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  if (grid > max_blocks) grid = max_blocks;
  *grid_out = grid;
#define VMTL_HB(CP)                                                                                           \
  do {                                                                                                        \
    if (cudaFuncSetAttribute(head_ce_tc_bwd_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                             HbSmem::kBytes) != cudaSuccess)                                                  \
      return VMTL_ECUDA;                                                                                      \
    head_ce_tc_bwd_kernel<CP><<<grid, kHbThreads, HbSmem::kBytes, st>>>(                                      \
        tmap_f, tmap_df, W, b, target, P, C, ignore_index, fwd_out, gscale, dfeat != nullptr, partial);       \
  } while (0)
  if (C <= 16)
    VMTL_HB(16);
  else if (C <= 20)
    VMTL_HB(20);
  else
    VMTL_HB(32);
#undef VMTL_HB
  return launch_status();
}

}  // namespace vmtl

#ifdef VMTL_HT_PROF
extern "C" int vmtl_debug_head_bwd_prof(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, vmtl::g_hb_prof, sizeof(long long) * 148 * 8) == cudaSuccess ? 0 : 1;
}
#endif

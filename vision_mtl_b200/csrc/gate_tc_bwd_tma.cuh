// MTAN gate backward on tensor cores, TMA generation (K = 128, N = 32 or a multiple of 64).
// Included by gate_tc.cu.  Same machinery as gate_tc_tma.cuh: raw fp32 tiles arrive in shared memory by TMA
// (>= 128 KB per SM in flight), the tf32 hi/lo A operand lives in TENSOR MEMORY.
//
//   pass 1  gate_tc_sdw_tma_kernel : ds, the batch sums of the BatchNorm backward and the two pixel
//           contractions dW is linear in (see the comment at the kernel), one read of (dy, s, z, h);
//   finalize (gate.cu)             : dgamma, dbeta, c1, c2, dbias and dW from the per-CTA partials (fp64);
//   pass 2  gate_tc_dh_tma_kernel  : unit = (128-row tile, 32-column atom a).  TMA brings dy_a, s_a, z_a
//           ([128 x 32] each, 48 KB per unit).  Converter threads (one row each) rebuild
//           dz = gamma*invstd*(du - c1 - zhat*c2) in registers and put its tf32 hi/lo parts into TMEM.
//           MMA: dh[128 x 128] += dz_a[128 x 32] @ W[a*32.., :]  (A from TMEM, B = W^T atoms K-major in smem,
//           3xTF32); dh leaves as TMA bulk stores.  dz never touches HBM.
#pragma once

namespace vmtl {

// ------------------------------------------------------------------------------------------- B1
template <int NA>
struct DhTmaSmem {
  static constexpr int kSlot = kTileM * 128;                 // [128 rows x 32 floats]
  static constexpr int kStage = 3 * kSlot;                   // dy_a, s_a, z_a
  static constexpr int kStages = 2;
  static constexpr int kBhi = kStages * kStage;              // W^T hi: NA atoms of [128 rows(k) x 128 B]
  static constexpr int kBlo = kBhi + NA * kSlot;
  static constexpr int kOut = kBlo + NA * kSlot;             // dh tile staging for the TMA store: 4 atoms
  static constexpr int kMisc = kOut + 4 * kSlot;
  static constexpr int kBytes = kMisc + 256 + 6 * 64 * 4 + 1024;
};

template <int NA, int NCH, bool SPLIT>
__global__ void __launch_bounds__(kTmaThreads, 1)
    gate_tc_dh_tma_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_s,
                          const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_dh,
                          const float* __restrict__ W /* [N,128] */,
                          const float* __restrict__ coefA, const float* __restrict__ coefB,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ c1, const float* __restrict__ c2, int64_t M,
                          float* __restrict__ dh, int n0 /* first gate column of this pass */,
                          int accumulate /* dh += */) {
  using namespace tc;
  using L = DhTmaSmem<NA>;
  constexpr int S = L::kStages;
  constexpr int N = NA * 32;        // gate columns per chunk; the launch covers NCH chunks from column n0
  constexpr int KH = 128;
  constexpr uint32_t kACols = 256;  // two A buffers of 128 columns: atom a -> hi [a*64, +32), lo [a*64+32, +32)
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sBhi = smem + L::kBhi;
  uint8_t* sBlo = smem + L::kBlo;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 192);
  float* s_coef = reinterpret_cast<float*>(smem + L::kMisc + 256);  // [6][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_empty = [&](int s) { return bar0 + 32u + 8u * (uint32_t)s; };
  // one aready barrier per (A buffer, atom): a consumer must observe every phase of a barrier
  // before its producers can complete the next one
  auto bar_aready = [&](int b, int a) { return bar0 + 64u + 8u * (uint32_t)(b * 2 + a); };
  auto bar_amma = [&](int b) { return bar0 + 96u + 8u * (uint32_t)b; };
  auto bar_dfull = [&](int b) { return bar0 + 112u + 8u * (uint32_t)b; };
  auto bar_dfree = [&](int b) { return bar0 + 128u + 8u * (uint32_t)b; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_aready(i, 0), 256);
      mbar_init(bar_aready(i, 1), 256);
      mbar_init(bar_amma(i), 1);
      mbar_init(bar_dfull(i), 1);
      mbar_init(bar_dfree(i), 256);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_s);
    tma_prefetch_desc(&tmap_z);
  }
  // W^T operand of chunk c: element (k, n) = W[n0 + c*N + n][k]; rows k, K-major along n, atom = n / 32
  auto stage_w = [&](int c) {
    const float* Wc = W + ((int64_t)n0 + c * N) * KH;
    constexpr int PER = N * KH / 256;  // elements per converter thread
#pragma unroll
    for (int i0 = 0; i0 < PER; i0 += 8) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __ldg(Wc + threadIdx.x + 256 * (i0 + i));  // 8 loads in flight, then the stores
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = threadIdx.x + 256 * (i0 + i);
        const int n = e / KH, k = e - n * KH;
        const float hi = tf32_hi(w[i]);
        const uint32_t off = (uint32_t)((n >> 5) * L::kSlot) + sw128_off(k, (n & 31) >> 2) + (uint32_t)((n & 3) << 2);
        *reinterpret_cast<float*>(sBhi + off) = hi;
        if (SPLIT) *reinterpret_cast<float*>(sBlo + off) = w[i] - hi;
      }
    }
    fence_proxy_async_smem();
  };
  for (int i = threadIdx.x; i < N; i += kTmaThreads) {
    s_coef[0 * 64 + i] = coefA[n0 + i];
    s_coef[1 * 64 + i] = coefB[n0 + i];
    s_coef[2 * 64 + i] = mean[n0 + i];
    s_coef[3 * 64 + i] = invstd[n0 + i];
    s_coef[4 * 64 + i] = c1[n0 + i];
    s_coef[5 * 64 + i] = c2[n0 + i];
  }
  if (NCH == 1 && warp < 8) stage_w(0);  // several chunks: the converters restage W per (tile, chunk)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_d0 = tmem_base + kACols;

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < 8) {
    // ------------------------------------------------------------------ converters (thread = row)
    const int quad = warp & 3, ch = warp >> 2;  // lane quadrant, 16-column half of the atom
    const int row = quad * 32 + lane;
    int64_t u = 0;   // unit = (tile, chunk, atom): one TMA stage
    int64_t tc = 0;  // (tile, chunk): one TMEM A buffer, one W^T chunk in shared memory
    for (int64_t it = 0; it < nitems; ++it) {
      const int64_t grow = (blockIdx.x + it * gridDim.x) * kTileM + row;
      const bool row_ok = grow < M;
#pragma unroll
      for (int c = 0; c < NCH; ++c, ++tc) {
        const int tb = (int)(tc & 1);
#pragma unroll
        for (int a = 0; a < NA; ++a, ++u) {
          const int s = (int)(u % S);
          mbar_wait(bar_full(s), (uint32_t)((u / S) & 1));
          const uint8_t* st = smem + s * L::kStage;
          float4 vdy[4], vs[4], vz[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = sw128_off(row, ch * 4 + j);
            vdy[j] = *reinterpret_cast<const float4*>(st + o);
            vs[j] = *reinterpret_cast<const float4*>(st + L::kSlot + o);
            vz[j] = *reinterpret_cast<const float4*>(st + 2 * L::kSlot + o);
          }
          float d[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 A, B, mu, rs, k1, k2;
            if (NCH == 1) {
              const int c4 = (a * 32 + ch * 16 + j * 4) >> 2;
              A = reinterpret_cast<const float4*>(s_coef)[0 * 16 + c4];
              B = reinterpret_cast<const float4*>(s_coef)[1 * 16 + c4];
              mu = reinterpret_cast<const float4*>(s_coef)[2 * 16 + c4];
              rs = reinterpret_cast<const float4*>(s_coef)[3 * 16 + c4];
              k1 = reinterpret_cast<const float4*>(s_coef)[4 * 16 + c4];
              k2 = reinterpret_cast<const float4*>(s_coef)[5 * 16 + c4];
            } else {  // several chunks: per-column coefficients straight from global (L1-resident, 6 KB)
              const int col = n0 + c * N + a * 32 + ch * 16 + j * 4;
              A = __ldg(reinterpret_cast<const float4*>(coefA + col));
              B = __ldg(reinterpret_cast<const float4*>(coefB + col));
              mu = __ldg(reinterpret_cast<const float4*>(mean + col));
              rs = __ldg(reinterpret_cast<const float4*>(invstd + col));
              k1 = __ldg(reinterpret_cast<const float4*>(c1 + col));
              k2 = __ldg(reinterpret_cast<const float4*>(c2 + col));
            }
            auto dz1 = [](float g, float sv, float zv, float A_, float B_, float mu_, float r_, float k1_, float k2_) {
              const float act = sigmoidf_fast(fmaf(A_, zv, B_));
              return A_ * (g * sv * act * (1.f - act) - k1_ - (zv - mu_) * r_ * k2_);
            };
            d[4 * j] = dz1(vdy[j].x, vs[j].x, vz[j].x, A.x, B.x, mu.x, rs.x, k1.x, k2.x);
            d[4 * j + 1] = dz1(vdy[j].y, vs[j].y, vz[j].y, A.y, B.y, mu.y, rs.y, k1.y, k2.y);
            d[4 * j + 2] = dz1(vdy[j].z, vs[j].z, vz[j].z, A.z, B.z, mu.z, rs.z, k1.z, k2.z);
            d[4 * j + 3] = dz1(vdy[j].w, vs[j].w, vz[j].w, A.w, B.w, mu.w, rs.w, k1.w, k2.w);
          }
          float hi[16], lo[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            if (!row_ok) d[e] = 0.f;
            hi[e] = tf32_hi(d[e]);
            lo[e] = d[e] - hi[e];
          }
          mbar_arrive(bar_empty(s));  // this thread holds everything it needs from stage s in registers
          if (a == 0) {
            if (NCH > 1) {
              // W^T of this chunk replaces the previous one: the MMAs of (tile, chunk) tc-1 must be done
              // (they retire in order, so tc-2 -- the previous user of this A buffer -- is done too)
              if (tc >= 1) {
                mbar_wait(bar_amma((int)((tc - 1) & 1)), (uint32_t)(((tc - 1) >> 1) & 1));
                tc_fence_after_sync();
              }
              stage_w(c);
            } else if (tc >= 2) {  // the MMAs of tile tc-2 have consumed this A buffer
              mbar_wait(bar_amma(tb), (uint32_t)(((tc >> 1) - 1) & 1));
              tc_fence_after_sync();
            }
          }
          const uint32_t ta = tmem_base + (((uint32_t)quad * 32) << 16) + (uint32_t)(tb * 128 + a * 64 + ch * 16);
          tmem_st16(ta, hi);
          if (SPLIT) tmem_st16(ta + 32, lo);
          tmem_wait_st();
          tc_fence_before_sync();
          mbar_arrive(bar_aready(tb, a));
        }
      }
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------ epilogue: dh rows
    const int ew = warp - 8;
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      mbar_wait(bar_dfull(b), (uint32_t)((it >> 1) & 1));
      tc_fence_after_sync();
      const int64_t grow = (blockIdx.x + it * gridDim.x) * kTileM + (ew & 3) * 32 + lane;
      const int col0 = (ew >> 2) * 64;
      const uint32_t taddr = tmem_d0 + (((uint32_t)(ew & 3) * 32) << 16) + (uint32_t)(b * KH + col0);
      // A row per thread is the wrong shape for global stores (32 x 16 B pieces per request).  The
      // tile goes through a 128B-swizzled staging buffer and leaves as four TMA bulk stores.
      if (threadIdx.x == 8 * 32) tma_store_wait_read();  // previous tile's stores have read the staging
      named_barrier_sync(2, 256);
      const int trow = (ew & 3) * 32 + lane;
#pragma unroll
      for (int j = 0; j < 64; j += 16) {
        float t16[16];
        tmem_ld16(taddr + j, t16);
        uint8_t* atom = smem + L::kOut + ((col0 + j) >> 5) * L::kSlot;
        const int c0 = ((col0 + j) & 31) >> 2;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          *reinterpret_cast<float4*>(atom + sw128_off(trow, c0 + e)) =
              make_float4(t16[4 * e], t16[4 * e + 1], t16[4 * e + 2], t16[4 * e + 3]);
      }
      tc_fence_before_sync();
      mbar_arrive(bar_dfree(b));
      fence_proxy_async_smem();
      named_barrier_sync(2, 256);
      if (threadIdx.x == 8 * 32 && dh) {
        const int row0 = (int)((blockIdx.x + it * gridDim.x) * kTileM);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (accumulate)  // later column passes add their share of dh in L2 (TMA reduce), no read-back
            tma_reduce_add_2d(&tmap_dh, a * 32, row0, smem_u32(smem + L::kOut + a * L::kSlot));
          else
            tma_store_2d(&tmap_dh, a * 32, row0, smem_u32(smem + L::kOut + a * L::kSlot));
        }
        tma_store_commit();
      }
      (void)grow;
    }
    if (threadIdx.x == 8 * 32) tma_store_wait_all();
  } else if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      constexpr int UPT = NCH * NA;  // units (32-column atoms) per tile
      // unit v -> first row of its tile / first gate column of its atom
      auto unit_row0 = [&](int64_t v) { return (int)((blockIdx.x + (v / UPT) * gridDim.x) * kTileM); };
      auto unit_col0 = [&](int64_t v) { return n0 + (int)(v % UPT) * 32; };
      const int64_t nunits = nitems * UPT;
      for (int64_t u = 0; u < nunits; ++u) {
        const int s = (int)(u % S);
        if (u >= S) mbar_wait(bar_empty(s), (uint32_t)(((u / S) - 1) & 1));  // every converter has read unit u-S
        mbar_expect_tx(bar_full(s), (uint32_t)L::kStage);
        const uint32_t dst = smem_u32(smem + s * L::kStage);
        const int row0 = unit_row0(u), col0 = unit_col0(u);
        tma_load_2d(dst, &tmap_dy, col0, row0, bar_full(s));
        tma_load_2d(dst + L::kSlot, &tmap_s, col0, row0, bar_full(s));
        tma_load_2d(dst + 2 * L::kSlot, &tmap_z, col0, row0, bar_full(s));
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = idesc_tf32(kTileM, KH, 0, 0);
    const uint32_t bH = smem_u32(sBhi), bL = smem_u32(sBlo);
    int64_t tc = 0;
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      const uint32_t d_tmem = tmem_d0 + (uint32_t)(b * KH);
      if (it >= 2) mbar_wait(bar_dfree(b), (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
      for (int c = 0; c < NCH; ++c, ++tc) {  // dh of the tile accumulates in TMEM over all column chunks
        const int tb = (int)(tc & 1);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          mbar_wait(bar_aready(tb, a), (uint32_t)((tc >> 1) & 1));
          tc_fence_after_sync();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a_hi = tmem_base + (uint32_t)(tb * 128 + a * 64 + ks * 8);
            const uint64_t dBh = smem_desc_sw128(bH + a * L::kSlot + ks * 32, 16, 1024);
            if (SPLIT) {
              mma_tf32_ts(d_tmem, a_hi + 32, dBh, idesc, (c | a | ks) != 0);
              mma_tf32_ts(d_tmem, a_hi, smem_desc_sw128(bL + a * L::kSlot + ks * 32, 16, 1024), idesc, 1);
              mma_tf32_ts(d_tmem, a_hi, dBh, idesc, 1);
            } else {
              mma_tf32_ts(d_tmem, a_hi, dBh, idesc, (c | a | ks) != 0);
            }
          }
        }
        mma_commit(bar_amma(tb));  // A buffer tb and the W^T chunk are consumed
      }
      mma_commit(bar_dfull(b));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------- pass 1
// Statistics + dW in ONE pass over (dy, s, z, h).  dz = A (du - c1 - zhat c2) needs the batch sums c1 = sum(du)/M,
// c2 = sum(du zhat)/M first -- but dW is LINEAR in them:
//     dW[n,k] = sum_r dz[r,n] h[r,k] = A_n ( P1[n,k] - c1_n hsum[k] - c2_n P2[n,k] ),
//     P1 = du^T h,  P2 = zhat^T h,  hsum = sum_r h[r,:]
// so this kernel accumulates P1, P2, hsum and the column sums (sum du, sum du zhat, sum zhat) while it writes
// ds = dy a, and a finalize kernel combines them once the sums are known.  No dz tensor is ever written, and
// (dy, s, z) are read here once and in the dh pass once: 4M(2K + 7N) bytes per gate backward instead of 4M(2K + 9N).
//
// unit = 64 pixel rows x NA 32-column atoms.  TMA brings the h half ([64 x 128], SW128) and dy, z, s straight into the
// MN-major tf32 operand layout (SWIZZLE_128B_ATOM_32B).  Converter warps: (i) transpose h^T into TMEM as hi/lo
// (lane = hidden channel, column = pixel row), (ii) walk the 16-byte chunks of dy / z / s in place -- elementwise,
// so the swizzle only matters for finding a chunk's columns -- and overwrite them with the B operand
//     [du_hi | zhat_hi | du_lo | zhat_lo]   (4 NA atoms; du_hi <- dy's slot, zhat_hi <- z's, du_lo <- s's)
// while ds leaves for global memory as full 128-byte lines.  MMA per 8 pixel rows:
//     D[:, 0:4N'] += h_hi^T @ [du_hi | zhat_hi | du_lo | zhat_lo] ,  D[:, 0:2N'] += h_lo^T @ [du_hi | zhat_hi]
// accumulating over ALL units of the CTA in TMEM (4 N' = 128 or 256 columns); drained once.
template <int NA>
struct SdwSmem {
  static constexpr int kHalfRows = 64;
  static constexpr int kSlotH = kHalfRows * 128;             // one K-atom of the h half [64 rows x 128 B]
  static constexpr int kStageH = 4 * kSlotH;                 // 32 KB
  static constexpr int kSlotD = kHalfRows * 128;             // one 32-column atom of a [64 x N'] operand, MN-major
  static constexpr int kStageD = 4 * NA * kSlotD;            // atom groups: du_hi, zhat_hi, du_lo, zhat_lo
  static constexpr int kStage = kStageH + kStageD;           // 64 KB (N' = 32) / 96 KB (N' = 64)
  static constexpr int kStages = NA == 1 ? 3 : 2;            // 192 KB in flight either way
  static constexpr int kMisc = kStages * kStage;
  static constexpr int kCoef = 4 * NA * 32 * 4;              // A, B, mean, invstd of this launch's columns
  static constexpr int kBytes = kMisc + 256 + kCoef + 1024;
};

template <int NA, bool SPLIT>
__global__ void __launch_bounds__(kTmaThreads, 1)
    gate_tc_sdw_tma_kernel(const __grid_constant__ CUtensorMap tmap_h /* box [32 x 64], SW128 */,
                           const __grid_constant__ CUtensorMap tmap_dy /* box [32 x 64], SW128_ATOM_32B */,
                           const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_s,
                           const float* __restrict__ h_coef /* nullptr, or [A1 | B1]: the h operand is the hidden
                              layer's pre-activation c and h = max(A1 c + B1, 0) is rebuilt here */,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ mean, const float* __restrict__ invstd, int64_t M,
                           float* __restrict__ ds /* [M, n_total] or nullptr */,
                           float* __restrict__ pw_partial /* [grid][2][N][128]: P1, P2 of the CTA's column chunk */,
                           float* __restrict__ hs_partial /* [grid][128] */,
                           float* __restrict__ col_partial /* [grid][3][N]: sum du, sum du zhat, sum zhat */,
                           int nch, int n_total) {
  using namespace tc;
  using L = SdwSmem<NA>;
  constexpr int S = L::kStages;
  constexpr int N = NA * 32;
  constexpr int KH = 128;
  constexpr int DC = SPLIT ? 4 * N : 2 * N;  // accumulator columns
  constexpr uint32_t kACols = 256;  // A half buffer hb: hi [hb*128, +64), lo [hb*128+64, +64); column = pixel row
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 192);
  float* s_coef = reinterpret_cast<float*>(smem + L::kMisc + 256);  // [4][N]: A, B, mean, invstd
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // n_total = N * nch gate columns: column chunks of N are spread over the CTAs of ONE launch (chunk = blockIdx.x %
  // nch, fixed per CTA); the CTAs of a chunk share its 64-row units.  h is re-read once per chunk -- from L2, the
  // chunk-mates work on the same units at the same time.
  const int chunk = (int)blockIdx.x % nch, cslot = (int)blockIdx.x / nch, cgrid = (int)gridDim.x / nch;
  const int n0 = chunk * N;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_aready = [&](int b) { return bar0 + 64u + 8u * (uint32_t)b; };
  auto bar_umma = [&](int s) { return bar0 + 80u + 8u * (uint32_t)s; };    // MMAs of the unit in stage s done
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_umma(s), 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(bar_aready(i), 512);
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_z);
    tma_prefetch_desc(&tmap_s);
  }
  for (int i = threadIdx.x; i < N; i += kTmaThreads) {
    const float a = gamma[n0 + i] * invstd[n0 + i];
    s_coef[i] = a;
    s_coef[N + i] = beta[n0 + i] - mean[n0 + i] * a;
    s_coef[2 * N + i] = mean[n0 + i];
    s_coef[3 * N + i] = invstd[n0 + i];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_d = tmem_base + kACols;

  const int64_t nhalves = (M + L::kHalfRows - 1) / L::kHalfRows;  // units of 64 pixel rows
  const int64_t nunits = cslot < nhalves ? (nhalves - cslot + cgrid - 1) / cgrid : 0;

  // per-thread sums (converter threads only): hidden channel k over its 32-row half; 4 gate columns per atom
  float hsum_acc = 0.f;
  float acc_du[NA][4], acc_dz[NA][4], acc_z[NA][4];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc_du[a][e] = acc_dz[a][e] = acc_z[a][e] = 0.f;
  // A thread's 16-byte chunk inside an atom: physical chunk (t % 8) of row (t / 8) [+ 32], t = tid % 256; the ATOM_32B
  // swizzle XORs the 32-byte index with (row % 4), which is the same for both rows -> a fixed logical column group
  const int t256 = (int)(threadIdx.x & 255);
  const int cj = t256 & 7, crow = t256 >> 3;
  const int cl = ((((cj >> 1) ^ (crow & 3)) << 1) | (cj & 1));  // logical 16-byte chunk: columns 4*cl .. 4*cl+3 of the atom
  // The elementwise pass over (dy, z, s) is split between the two groups of 8 warps: group B (warps 8-15, which have
  // no other per-unit work) takes chunks [0, kChunksB), group A (warps 0-7, after the h transposition) the rest.
  constexpr int kChunks = NA * 2, kChunksB = NA == 1 ? 2 : 3;

  // (ii) dy, z, s -> ds (global) and the B operand [du_hi | zhat_hi | du_lo | zhat_lo], in place
  auto elementwise = [&](uint8_t* dsm, int64_t row0, int i_begin, int i_end) {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      if (i < i_begin || i >= i_end) continue;
      const int atom = i >> 1;
      const int row = crow + 32 * (i & 1);
      const uint32_t o = (uint32_t)(t256 + 256 * (i & 1)) * 16u;
      const float4 gy = *reinterpret_cast<const float4*>(dsm + (0 * NA + atom) * L::kSlotD + o);
      const float4 zv = *reinterpret_cast<const float4*>(dsm + (1 * NA + atom) * L::kSlotD + o);
      const float4 sv = *reinterpret_cast<const float4*>(dsm + (2 * NA + atom) * L::kSlotD + o);
      const int col = atom * 32 + cl * 4;
      const float4 A = *reinterpret_cast<const float4*>(s_coef + col);
      const float4 B = *reinterpret_cast<const float4*>(s_coef + N + col);
      const float4 mu = *reinterpret_cast<const float4*>(s_coef + 2 * N + col);
      const float4 rs = *reinterpret_cast<const float4*>(s_coef + 3 * N + col);
      const bool valid = row0 + row < M;  // rows past M arrive as zeros: du = ds = 0 there, zhat must be forced
      const float gyv[4] = {gy.x, gy.y, gy.z, gy.w}, zvv[4] = {zv.x, zv.y, zv.z, zv.w}, svv[4] = {sv.x, sv.y, sv.z, sv.w};
      const float Av[4] = {A.x, A.y, A.z, A.w}, Bv[4] = {B.x, B.y, B.z, B.w};
      const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
      float dsv[4], du[4], zh[4], duh[4], zhh[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float act = sigmoidf_fast(fmaf(Av[e], zvv[e], Bv[e]));
        dsv[e] = gyv[e] * act;
        du[e] = gyv[e] * svv[e] * act * (1.f - act);
        zh[e] = valid ? (zvv[e] - muv[e]) * rsv[e] : 0.f;
        acc_du[atom][e] += du[e];
        acc_dz[atom][e] = fmaf(du[e], zh[e], acc_dz[atom][e]);
        acc_z[atom][e] += zh[e];
        duh[e] = tf32_hi(du[e]);
        zhh[e] = tf32_hi(zh[e]);
      }
      if (ds != nullptr && valid)
        stg_stream(reinterpret_cast<float4*>(ds + (row0 + row) * n_total + n0 + col), make_float4(dsv[0], dsv[1], dsv[2], dsv[3]));
      *reinterpret_cast<float4*>(dsm + (0 * NA + atom) * L::kSlotD + o) = make_float4(duh[0], duh[1], duh[2], duh[3]);
      *reinterpret_cast<float4*>(dsm + (1 * NA + atom) * L::kSlotD + o) = make_float4(zhh[0], zhh[1], zhh[2], zhh[3]);
      if (SPLIT) {
        *reinterpret_cast<float4*>(dsm + (2 * NA + atom) * L::kSlotD + o) =
            make_float4(du[0] - duh[0], du[1] - duh[1], du[2] - duh[2], du[3] - duh[3]);
        *reinterpret_cast<float4*>(dsm + (3 * NA + atom) * L::kSlotD + o) =
            make_float4(zh[0] - zhh[0], zh[1] - zhh[1], zh[2] - zhh[2], zh[3] - zhh[3]);
      }
    }
  };

  if (warp < 8) {
    // ------------------------------------------------------------------ converters A: h^T -> TMEM, then their chunks
    const int quad = warp & 3, rh = warp >> 2;  // lane quadrant (k / 32), 32-row half of the unit
    const int k = quad * 32 + lane;
    const bool pre = h_coef != nullptr;
    const float a1 = pre ? h_coef[k] : 1.f, b1 = pre ? h_coef[KH + k] : 0.f;
    for (int64_t u = 0; u < nunits; ++u) {
      const int s = (int)(u % S), hb = (int)(u & 1);
      const int64_t row0 = (cslot + u * cgrid) * L::kHalfRows;
      mbar_wait(bar_full(s), (uint32_t)((u / S) & 1));
      uint8_t* stage = smem + s * L::kStage;
      {  // (i) h^T -> TMEM (thread = hidden channel k)
        const uint8_t* hs = stage + (k >> 5) * L::kSlotH;  // K-atom holding channel k
        float hv[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const int row = rh * 32 + r;
          hv[r] = *reinterpret_cast<const float*>(hs + sw128_off(row, (k & 31) >> 2) + ((k & 3) << 2));
          // folded BatchNorm + ReLU of the hidden layer; rows past M arrive as zeros and must stay zeros
          if (pre) hv[r] = row0 + row < M ? fmaxf(fmaf(a1, hv[r], b1), 0.f) : 0.f;
          hsum_acc += hv[r];
        }
        if (u >= 2) {  // the MMAs of unit u-2 (same A half buffer) are done
          const int64_t up = u - 2;
          mbar_wait(bar_umma((int)(up % S)), (uint32_t)((up / S) & 1));
          tc_fence_after_sync();
        }
        const uint32_t ta = tmem_base + (((uint32_t)quad * 32) << 16) + (uint32_t)(hb * 128 + rh * 32);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float hi[16], lo[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            hi[e] = tf32_hi(hv[g * 16 + e]);
            lo[e] = hv[g * 16 + e] - hi[e];
          }
          tmem_st16(ta + g * 16, hi);
          if (SPLIT) tmem_st16(ta + 64 + g * 16, lo);
        }
      }
      elementwise(stage + L::kStageH, row0, kChunksB, kChunks);
      fence_proxy_async_smem();
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(bar_aready(hb));
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------ converters B: elementwise only
    for (int64_t u = 0; u < nunits; ++u) {
      const int s = (int)(u % S), hb = (int)(u & 1);
      const int64_t row0 = (cslot + u * cgrid) * L::kHalfRows;
      mbar_wait(bar_full(s), (uint32_t)((u / S) & 1));
      elementwise(smem + s * L::kStage + L::kStageH, row0, 0, kChunksB);
      fence_proxy_async_smem();
      mbar_arrive(bar_aready(hb));
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int64_t u = 0; u < nunits; ++u) {
        const int s = (int)(u % S);
        // stage reusable once the MMAs of its previous unit are done (they were issued after every converter
        // had finished reading and rewriting the stage)
        if (u >= S) mbar_wait(bar_umma(s), (uint32_t)(((u / S) - 1) & 1));
        const int row0 = (int)((cslot + u * cgrid) * L::kHalfRows);
        mbar_expect_tx(bar_full(s), (uint32_t)(L::kStageH + 3 * NA * L::kSlotD));
        const uint32_t dst = smem_u32(smem + s * L::kStage);
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_2d(dst + a * L::kSlotH, &tmap_h, a * 32, row0, bar_full(s));
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const uint32_t d0 = dst + L::kStageH + a * L::kSlotD;
          tma_load_2d(d0, &tmap_dy, n0 + a * 32, row0, bar_full(s));
          tma_load_2d(d0 + NA * L::kSlotD, &tmap_z, n0 + a * 32, row0, bar_full(s));
          tma_load_2d(d0 + 2 * NA * L::kSlotD, &tmap_s, n0 + a * 32, row0, bar_full(s));
        }
      }
    }
  } else if (warp == 17 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_wide = idesc_tf32(KH, DC, 0, 1);  // A from TMEM (K-major form), B MN-major
    constexpr uint32_t idesc_n = idesc_tf32(KH, 2 * N, 0, 1);
    for (int64_t u = 0; u < nunits; ++u) {
      const int s = (int)(u % S), hb = (int)(u & 1);
      mbar_wait(bar_aready(hb), (uint32_t)((u >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t bD = smem_u32(smem + s * L::kStage + L::kStageH);
#pragma unroll
      for (int ks = 0; ks < L::kHalfRows / 8; ++ks) {  // 8 pixel rows per MMA = two 4-row groups
        const uint32_t a_hi = tmem_base + (uint32_t)(hb * 128 + ks * 8);
        const uint64_t dB = smem_desc_mn_tf32(bD + ks * 1024, L::kSlotD, 512);
        mma_tf32_ts(tmem_d, a_hi, dB, idesc_wide, (u | ks) != 0);
        if (SPLIT) mma_tf32_ts(tmem_d, a_hi + 64, dB, idesc_n, 1);
      }
      mma_commit(bar_umma(s));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  float* out = pw_partial + (int64_t)blockIdx.x * 2 * N * KH;
  const int64_t q_stride = (int64_t)N * KH;  // P1 -> P2
  if (nunits > 0) {
    if (warp == 17 && lane == 0) {  // wait for the last unit's MMAs
      const int64_t ul = nunits - 1;
      mbar_wait(bar_umma((int)(ul % S)), (uint32_t)((ul / S) & 1));
    }
    __syncthreads();
    tc_fence_after_sync();
    if (warp < 8) {  // thread (k = lane quadrant row) drains [P1 | P2][k][n]; partial layout is [q][n][k]
      const int k = (warp & 3) * 32 + lane;
      const int col0 = (warp >> 2) * (N / 2);
      const uint32_t taddr = tmem_d + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)col0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int j = 0; j < N / 2; j += 16) {
          float t16[16], t2[16];
          tmem_ld16(taddr + q * N + j, t16);
          if (SPLIT) {
            tmem_ld16(taddr + (2 + q) * N + j, t2);
#pragma unroll
            for (int e = 0; e < 16; ++e) t16[e] += t2[e];
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) out[q * q_stride + (int64_t)(col0 + j + e) * KH + k] = t16[e];
        }
      }
    }
  } else {
    for (int e = threadIdx.x; e < 2 * N * KH; e += kTmaThreads)
      out[(e / (N * KH)) * q_stride + (e % (N * KH))] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, kTmemCols);
  // ---- per-CTA column sums and hsum: fixed-order block reductions through shared memory (stage 0 is idle) ----
  float* s_red = reinterpret_cast<float*>(smem);  // [512][NA][12] then [256] hsum
  float* s_hs = s_red + 512 * NA * 12;
  if (warp < 16) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s_red[(threadIdx.x * NA + a) * 12 + e] = acc_du[a][e];
        s_red[(threadIdx.x * NA + a) * 12 + 4 + e] = acc_dz[a][e];
        s_red[(threadIdx.x * NA + a) * 12 + 8 + e] = acc_z[a][e];
      }
    if (warp < 8) s_hs[threadIdx.x] = hsum_acc;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 3 * N; o += kTmaThreads) {
    const int q = o / N, col = o % N;
    const int atom = col >> 5, c = (col & 31) >> 2, e = col & 3;
    float acc = 0.f;
    for (int t = 0; t < 512; ++t) {  // the 64 converter threads (both groups) whose chunk holds this column, in thread order
      const int tj = t & 7, tr = (t >> 3) & 31;
      if (((((tj >> 1) ^ (tr & 3)) << 1) | (tj & 1)) == c) acc += s_red[(t * NA + atom) * 12 + q * 4 + e];
    }
    col_partial[((int64_t)blockIdx.x * 3 + q) * N + col] = acc;
  }
  for (int k = threadIdx.x; k < KH; k += kTmaThreads) {
    // converter thread (quad, rh) held channel k = quad*32 + lane: tid = (rh*4 + quad)*32 + lane
    hs_partial[(int64_t)blockIdx.x * KH + k] = s_hs[k] + s_hs[128 + k];
  }
}

// ------------------------------------------------------------------------------------------- host side
// N = 32 or any multiple of 64.
// pass 1: ONE launch.  N = 32: NA = 1; otherwise 64-column chunks (NA = 2: 256 accumulator columns in TMEM) spread
//         over the CTAs of the launch, grid a multiple of the chunk count.
template <bool SPLIT>
static int launch_sdw_tma(const float* dy, const float* h, const float* h_coef, const float* s, const float* z, const float* gamma,
                          const float* beta, const float* mean, const float* invstd, int64_t M, int N, float* ds,
                          float* pw_partial, float* hs_partial, float* col_partial, int grid, int nch, cudaStream_t st) {
  CUtensorMap t_h, t_dy, t_z, t_s;
  if (!make_tmap_2d_sw(&t_h, h, M, 128, 64, false) || !make_tmap_2d_sw(&t_dy, dy, M, N, 64, true) ||
      !make_tmap_2d_sw(&t_z, z, M, N, 64, true) || !make_tmap_2d_sw(&t_s, s, M, N, 64, true))
    return VMTL_ECUDA;
  if (N == 32) {
    auto k = gate_tc_sdw_tma_kernel<1, SPLIT>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SdwSmem<1>::kBytes) != cudaSuccess)
      return VMTL_ECUDA;
    k<<<grid, kTmaThreads, SdwSmem<1>::kBytes, st>>>(t_h, t_dy, t_z, t_s, h_coef, gamma, beta, mean, invstd, M, ds, pw_partial,
                                                     hs_partial, col_partial, 1, N);
    return launch_status();
  }
  auto k = gate_tc_sdw_tma_kernel<2, SPLIT>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SdwSmem<2>::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  k<<<grid, kTmaThreads, SdwSmem<2>::kBytes, st>>>(t_h, t_dy, t_z, t_s, h_coef, gamma, beta, mean, invstd, M, ds, pw_partial,
                                                   hs_partial, col_partial, nch, N);
  return launch_status();
}

// pass 2: one launch covers up to 256 gate columns (NCH chunks of 64 whose dh contributions accumulate in TMEM, W^T
// restaged per chunk); wider gates take further launches that add into dh with TMA reduce-adds.
template <int NA_DH, int NCH, bool SPLIT>
static int launch_dh_tma(const CUtensorMap& t_dy, const CUtensorMap& t_s, const CUtensorMap& t_z, const CUtensorMap& t_dh,
                         const float* W, const GateWs& ws, int64_t M, float* dh, int n0, int grid, cudaStream_t st) {
  auto k1 = gate_tc_dh_tma_kernel<NA_DH, NCH, SPLIT>;
  if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, DhTmaSmem<NA_DH>::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  k1<<<grid, kTmaThreads, DhTmaSmem<NA_DH>::kBytes, st>>>(t_dy, t_s, t_z, t_dh, W, ws.coefA, ws.coefB, ws.mean, ws.invstd,
                                                           ws.c1, ws.c2, M, dh, n0, n0 != 0);
  return launch_status();
}

template <bool SPLIT>
static int launch_dh_passes(const float* dy, const float* s, const float* z, const float* W, const GateWs& ws, int64_t M,
                            int N, float* dh, int grid, cudaStream_t st) {
  CUtensorMap t_dy, t_s, t_z, t_dh;
  if (!make_tmap_2d_sw(&t_dy, dy, M, N, kTileM, false) || !make_tmap_2d_sw(&t_s, s, M, N, kTileM, false) ||
      !make_tmap_2d_sw(&t_z, z, M, N, kTileM, false) || !make_tmap_2d_sw(&t_dh, dh, M, 128, kTileM, false))
    return VMTL_ECUDA;
  if (N == 32) return launch_dh_tma<1, 1, SPLIT>(t_dy, t_s, t_z, t_dh, W, ws, M, dh, 0, grid, st);
  // One launch over several chunks pays a W^T restage per (tile, chunk); it wins only where a CTA has a
  // single tile anyway (then the alternative is one latency-bound launch per chunk).  Otherwise: one
  // 64-column pass per launch, later passes add into dh.
  const int cols_per_launch = (M + kTileM - 1) / kTileM <= grid ? 256 : 64;
  for (int n0 = 0; n0 < N; n0 += cols_per_launch) {
    const int nch = (N - n0 >= cols_per_launch ? cols_per_launch : N - n0) / 64;
    int rc;
    switch (nch) {
      case 1: rc = launch_dh_tma<2, 1, SPLIT>(t_dy, t_s, t_z, t_dh, W, ws, M, dh, n0, grid, st); break;
      case 2: rc = launch_dh_tma<2, 2, SPLIT>(t_dy, t_s, t_z, t_dh, W, ws, M, dh, n0, grid, st); break;
      case 3: rc = launch_dh_tma<2, 3, SPLIT>(t_dy, t_s, t_z, t_dh, W, ws, M, dh, n0, grid, st); break;
      default: rc = launch_dh_tma<2, 4, SPLIT>(t_dy, t_s, t_z, t_dh, W, ws, M, dh, n0, grid, st); break;
    }
    if (rc != VMTL_OK) return rc;
  }
  return VMTL_OK;
}

}  // namespace vmtl

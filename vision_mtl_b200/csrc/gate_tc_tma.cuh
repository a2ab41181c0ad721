// MTAN gate forward (K = 128, N = 32 or a multiple of 64): TMA -> smem -> TMEM -> tcgen05.
// Included by gate_tc.cu.  TRAIN: z + per-CTA column partials of (sum z, sum z^2).  EVAL: BatchNorm folded
// into (coefA, coefB), y = s * sigmoid(coefA * z + coefB) straight from the accumulator, z never stored.
//
// Why this shape (measured on B200, scratch/mma_probe*.cu and ncu, profiles/):
//  * the contraction is HBM-bound only if >= ~128 KB per SM are in flight; a register-staged
//    pipeline (one 64 KB tile in 256 threads' registers) tops out near 3.4 TB/s.  Here the raw fp32
//    tile lands in shared memory by TMA (cp.async.bulk.tensor, 128B swizzle, 3 x 64 KB stages for
//    N = 32, 2 for N = 64), so up to 192 KB per SM are in flight and no register holds a load;
//  * the tf32 hi/lo split of the A operand is written to TENSOR MEMORY (tcgen05.st, one row per
//    thread) and consumed from there (tcgen05.mma with A in TMEM), which frees the 128 KB of
//    shared memory the split copies used to take -- that is what pays for the TMA stages;
//  * a kind::tf32 MMA with N <= 64 occupies the tensor pipe ~45 cycles regardless of operands, so
//    W_hi and W_lo are stacked into one B operand of 2N rows:  per K-step
//        D[:, 0:2N] (+)= A_hi @ [W_hi ; W_lo]^T ,   D[:, 0:N] += A_lo @ W_hi^T
//    (2 MMAs instead of 3); the epilogue adds the two column halves.
//
// Roles (18 warps): 0-7 converters (smem -> hi/lo -> TMEM; TMEM lane quadrant = warp % 4, K-atom
// of the half = warp / 4), 8-15 epilogue, 16 TMA producer, 17 MMA issuer.
// mbarriers: full[s] (TMA tx), empty[s] (256 converter arrivals), aready[h] (256), amma[h]
// (tcgen05.commit: A half h consumed), dfull[b] (commit), dfree[b] (256 epilogue arrivals).
#pragma once

#include <cuda.h>

namespace vmtl {

constexpr int kTmaThreads = 18 * 32;

template <int NC>
struct TmaSmem {
  static constexpr int kStage = kTileM * 512;              // raw fp32 tile: 4 K-atoms of [128 x 128 B]
  static constexpr int kStages = NC <= 32 ? 3 : 2;
  static constexpr int kAtomB = 2 * NC * 128;              // rows [0,NC) = W_hi, [NC,2NC) = W_lo
  static constexpr int kB = kStages * kStage;
  static constexpr int kMisc = kB + 4 * kAtomB;
  static constexpr int kBytes = kMisc + 256 + 3 * 64 * 4 + 2 * 128 * 4 + 1024;
};

// N = NC * nch.  nch > 1: column chunks of NC are spread over CTAs (chunk = blockIdx.x % nch, fixed per CTA so
// its W operand is staged once); gridDim.x is a multiple of nch.
// PRE: the operand is the PRE-activation c of the hidden layer and the converters fold the BatchNorm + ReLU in front
// of the gate into the conversion, h = max(A1 c + B1, 0) (h_coef = [A1 | B1], [2][128]): h is never materialised.
template <int NC, bool SPLIT, bool EVAL, bool PRE>
__global__ void __launch_bounds__(kTmaThreads, 1)
    gate_tc_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap_h, const float* __restrict__ h_coef,
                           const float* __restrict__ W,
                           const float* __restrict__ bias, int64_t M, int nch,
                           float* __restrict__ out /* TRAIN: z ; EVAL: y */,
                           float* __restrict__ partial /* TRAIN: [gridDim.x][2][N] */,
                           const float* __restrict__ s_in /* EVAL: shared features [M,N] */,
                           const float* __restrict__ coefA, const float* __restrict__ coefB) {
  using namespace tc;
  using L = TmaSmem<NC>;
  constexpr int S = L::kStages;
  const int N = NC * nch;
  constexpr int V = NC / 2;                  // columns per epilogue thread
  constexpr int DC = SPLIT ? 2 * NC : NC;    // accumulator columns per buffer
  constexpr uint32_t kACols = 256;           // TMEM: A halves [h*128, +64) hi, [+64, +128) lo
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sB = smem + L::kB;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 192);
  float* s_bias = reinterpret_cast<float*>(smem + L::kMisc + 256);
  float* s_cA = s_bias + 64;
  float* s_cB = s_cA + 64;
  float* s_hc = s_cB + 64;  // PRE: [A1 | B1] of the hidden layer's BatchNorm, 2 x 128

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % nch;          // this CTA's column chunk (fixed: W staged once)
  const int cslot = blockIdx.x / nch;          // position among the CTAs that share the chunk
  const int cgrid = gridDim.x / nch;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };            // [0,4)
  auto bar_empty = [&](int s) { return bar0 + 32u + 8u * (uint32_t)s; };     // [4,8)
  auto bar_aready = [&](int hh) { return bar0 + 64u + 8u * (uint32_t)hh; };  // [8,10)
  auto bar_amma = [&](int hh) { return bar0 + 80u + 8u * (uint32_t)hh; };    // [10,12)
  auto bar_dfull = [&](int b) { return bar0 + 96u + 8u * (uint32_t)b; };     // [12,14)
  auto bar_dfree = [&](int b) { return bar0 + 112u + 8u * (uint32_t)b; };    // [14,16)

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_aready(i), 256);
      mbar_init(bar_amma(i), 1);
      mbar_init(bar_dfull(i), 1);
      mbar_init(bar_dfree(i), 256);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  if (warp == 16 && lane == 0) tma_prefetch_desc(&tmap_h);
  if (PRE)
    for (int i = threadIdx.x; i < 256; i += kTmaThreads) s_hc[i] = h_coef[i];
  for (int i = threadIdx.x; i < NC; i += kTmaThreads) {
    s_bias[i] = bias[chunk * NC + i];
    if (EVAL) {
      s_cA[i] = coefA[chunk * NC + i];
      s_cB[i] = coefB[chunk * NC + i];
    }
  }
  if (warp < 8) {  // stacked W operand: row n = W_hi[n], row NC + n = W_lo[n]; K-major, SW128
    constexpr int PER = NC * 32 / 256;  // float4 per thread (4 or 8): all loads in flight, then the stores
    float4 wv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i)
      wv[i] = __ldg(reinterpret_cast<const float4*>(W) + (int64_t)chunk * NC * 32 + threadIdx.x + 256 * i);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int q = threadIdx.x + 256 * i;
      const int n = q >> 5, kc = q & 31;
      const float4 w = wv[i];
      const float4 hi = make_float4(tf32_hi(w.x), tf32_hi(w.y), tf32_hi(w.z), tf32_hi(w.w));
      uint8_t* atom = sB + (kc >> 3) * L::kAtomB;
      *reinterpret_cast<float4*>(atom + sw128_off(n, kc & 7)) = hi;
      if (SPLIT)
        *reinterpret_cast<float4*>(atom + sw128_off(NC + n, kc & 7)) =
            make_float4(w.x - hi.x, w.y - hi.y, w.z - hi.z, w.w - hi.w);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_d0 = tmem_base + kACols;

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t nitems = cslot < ntiles ? (ntiles - cslot + cgrid - 1) / cgrid : 0;
  double st_sum = 0.0, st_sq = 0.0;

  if (warp < 8) {
    // ------------------------------------------------------------------ converters
    const int quad = warp & 3, a2 = warp >> 2;  // lane quadrant, K-atom inside the half
    const int row = quad * 32 + lane;
    for (int64_t it = 0; it < nitems; ++it) {
      const int s = (int)(it % S);
      mbar_wait(bar_full(s), (uint32_t)((it / S) & 1));
      const uint8_t* stage = smem + s * L::kStage;
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const uint8_t* atom = stage + (kh * 2 + a2) * (kTileM * 128);
        float4 c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = *reinterpret_cast<const float4*>(atom + sw128_off(row, j));
        if (PRE) {  // h = max(A1 c + B1, 0); all lanes read the same coefficients (shared-memory broadcast)
          const float* hc = s_hc + (kh * 2 + a2) * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 A1 = *reinterpret_cast<const float4*>(hc + 4 * j), B1 = *reinterpret_cast<const float4*>(hc + 128 + 4 * j);
            c[j].x = fmaxf(fmaf(A1.x, c[j].x, B1.x), 0.f);
            c[j].y = fmaxf(fmaf(A1.y, c[j].y, B1.y), 0.f);
            c[j].z = fmaxf(fmaf(A1.z, c[j].z, B1.z), 0.f);
            c[j].w = fmaxf(fmaf(A1.w, c[j].w, B1.w), 0.f);
          }
        }
        if (it > 0) {  // MMAs that read this A half for the previous tile are done
          mbar_wait(bar_amma(kh), (uint32_t)((it - 1) & 1));
          tc_fence_after_sync();
        }
        const uint32_t ta = tmem_base + (((uint32_t)quad * 32) << 16) + (uint32_t)(kh * 128 + a2 * 32);
#pragma unroll
        for (int g = 0; g < 2; ++g) {  // 16 columns (= 4 chunks) per tcgen05.st
          float hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 a = c[g * 4 + j];
            hi[4 * j] = tf32_hi(a.x); hi[4 * j + 1] = tf32_hi(a.y); hi[4 * j + 2] = tf32_hi(a.z); hi[4 * j + 3] = tf32_hi(a.w);
            lo[4 * j] = a.x - hi[4 * j]; lo[4 * j + 1] = a.y - hi[4 * j + 1];
            lo[4 * j + 2] = a.z - hi[4 * j + 2]; lo[4 * j + 3] = a.w - hi[4 * j + 3];
          }
          tmem_st16(ta + g * 16, hi);
          if (SPLIT) tmem_st16(ta + 64 + g * 16, lo);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(bar_aready(kh));
      }
      mbar_arrive(bar_empty(s));  // this thread has read everything it needs from stage s
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 8;
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      const int64_t tile = cslot + it * cgrid;
      const int col0 = (ew >> 2) * V;
      const int64_t grow = tile * kTileM + (ew & 3) * 32 + lane;
      const bool row_ok = grow < M;
      float4 sv[EVAL ? V / 4 : 1];
      if (EVAL && row_ok) {  // the shared features of this row: in flight while the MMAs finish
        const float4* sp = reinterpret_cast<const float4*>(s_in + grow * N + chunk * NC + col0);
#pragma unroll
        for (int j = 0; j < V / 4; ++j) sv[j] = ldg_stream(sp + j);
      }
      mbar_wait(bar_dfull(b), (uint32_t)((it >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t taddr = tmem_d0 + (((uint32_t)(ew & 3) * 32) << 16) + (uint32_t)(b * DC + col0);
      float v[V];
#pragma unroll
      for (int j = 0; j < V; j += 16) {
        float t16[16];
        tmem_ld16(taddr + j, t16);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[j + e] = t16[e] + s_bias[col0 + j + e];
        if (SPLIT) {
          tmem_ld16(taddr + NC + j, t16);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[j + e] += t16[e];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(bar_dfree(b));
      if (EVAL) {
        if (row_ok) {
          float4* yp = reinterpret_cast<float4*>(out + grow * N + chunk * NC + col0);
#pragma unroll
          for (int j = 0; j < V; j += 4) {
            float4 y;
            y.x = sv[j / 4].x * sigmoidf_acc(fmaf(s_cA[col0 + j], v[j], s_cB[col0 + j]));
            y.y = sv[j / 4].y * sigmoidf_acc(fmaf(s_cA[col0 + j + 1], v[j + 1], s_cB[col0 + j + 1]));
            y.z = sv[j / 4].z * sigmoidf_acc(fmaf(s_cA[col0 + j + 2], v[j + 2], s_cB[col0 + j + 2]));
            y.w = sv[j / 4].w * sigmoidf_acc(fmaf(s_cA[col0 + j + 3], v[j + 3], s_cB[col0 + j + 3]));
            stg_stream(yp + j / 4, y);
          }
        }
        continue;
      }
      if (row_ok) {
        float4* zp = reinterpret_cast<float4*>(out + grow * N + chunk * NC + col0);
#pragma unroll
        for (int j = 0; j < V; j += 4) stg_stream(zp + j / 4, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
      float sq[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (!row_ok) v[j] = 0.f;
        sq[j] = v[j] * v[j];
      }
      st_sum += (double)butterfly_colsum<V>(v, lane);
      st_sq += (double)butterfly_colsum<V>(sq, lane);
    }
  } else if (warp == 16) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int64_t it = 0; it < nitems; ++it) {
        const int s = (int)(it % S);
        if (it >= S) mbar_wait(bar_empty(s), (uint32_t)(((it / S) - 1) & 1));
        const int row0 = (int)((cslot + it * cgrid) * kTileM);
        mbar_expect_tx(bar_full(s), (uint32_t)L::kStage);
        const uint32_t dst = smem_u32(smem + s * L::kStage);
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_2d(dst + a * (kTileM * 128), &tmap_h, a * 32, row0, bar_full(s));
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_wide = idesc_tf32(kTileM, DC, 0, 0);
    constexpr uint32_t idesc_n = idesc_tf32(kTileM, NC, 0, 0);
    const uint32_t bW = smem_u32(sB);
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      const uint32_t d_tmem = tmem_d0 + (uint32_t)(b * DC);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(bar_aready(kh), (uint32_t)(it & 1));
        if (kh == 0 && it >= 2) mbar_wait(bar_dfree(b), (uint32_t)(((it >> 1) - 1) & 1));
        tc_fence_after_sync();
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a_hi = tmem_base + (uint32_t)(kh * 128 + a2 * 32 + ks * 8);
            const uint64_t dB = smem_desc_sw128(bW + (kh * 2 + a2) * L::kAtomB + ks * 32, 16, 1024);
            mma_tf32_ts(d_tmem, a_hi, dB, idesc_wide, (kh | a2 | ks) != 0);
            if (SPLIT) mma_tf32_ts(d_tmem, a_hi + 64, dB, idesc_n, 1);
          }
        }
        mma_commit(bar_amma(kh));
        if (kh == 1) mma_commit(bar_dfull(b));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, kTmemCols);
  if (EVAL) return;
  // per-CTA column partials: the four quadrant warps of a column half, summed in fixed order
  double* s_red = reinterpret_cast<double*>(smem);  // [8 epilogue warps][V][2]; stage 0 is idle now
  if (warp >= 8 && warp < 16 && lane < V) {
    s_red[((warp - 8) * V + lane) * 2] = st_sum;
    s_red[((warp - 8) * V + lane) * 2 + 1] = st_sq;
  }
  __syncthreads();
  for (int col = threadIdx.x; col < N; col += kTmaThreads) {
    double a = 0.0, bq = 0.0;
    if (col / NC == chunk) {  // columns of other chunks belong to other CTAs: contribute zeros
      const int cc = col % NC, half = cc / V, l = cc % V;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a += s_red[((half * 4 + q) * V + l) * 2];
        bq += s_red[((half * 4 + q) * V + l) * 2 + 1];
      }
    }
    partial[(int64_t)blockIdx.x * 2 * N + col] = (float)a;
    partial[(int64_t)blockIdx.x * 2 * N + N + col] = (float)bq;
  }
}

template <int NC, bool SPLIT, bool EVAL, bool PRE>
static int launch_fwd_tma1(const CUtensorMap& tmap, const float* h_coef, const float* W, const float* bias, int64_t M,
                           int nch, float* out, float* partial, const float* s_in, const float* coefA,
                           const float* coefB, int grid, cudaStream_t st) {
  using L = TmaSmem<NC>;
  auto kern = gate_tc_fwd_tma_kernel<NC, SPLIT, EVAL, PRE>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  kern<<<grid, kTmaThreads, L::kBytes, st>>>(tmap, h_coef, W, bias, M, nch, out, partial, s_in, coefA, coefB);
  return launch_status();
}

template <int NC, bool SPLIT, bool EVAL>
static int launch_fwd_tma(const float* h, const float* h_coef, const float* W, const float* bias, int64_t M, int nch,
                          float* out, float* partial, const float* s_in, const float* coefA, const float* coefB,
                          int grid, cudaStream_t st) {
  CUtensorMap tmap;
  if (!make_tmap_2d(&tmap, h, M, 128, kTileM)) return VMTL_ECUDA;
  return h_coef ? launch_fwd_tma1<NC, SPLIT, EVAL, true>(tmap, h_coef, W, bias, M, nch, out, partial, s_in, coefA, coefB, grid, st)
                : launch_fwd_tma1<NC, SPLIT, EVAL, false>(tmap, h_coef, W, bias, M, nch, out, partial, s_in, coefA, coefB, grid, st);
}

// grid of the forward: (row tile, column chunk) items over the SMs, a multiple of nch
static int fwd_tma_grid(int64_t M, int nch) {
  const int64_t items = ((M + kTileM - 1) / kTileM) * nch;
  int grid = (int)(items < sm_count() ? items : sm_count());
  grid = grid / nch * nch;
  return grid < nch ? nch : grid;
}

template <bool EVAL>
static int dispatch_fwd_tma(const float* h, const float* h_coef, const float* W, const float* bias, int64_t M, int N,
                            int split3, float* out, float* partial, const float* s_in, const float* coefA,
                            const float* coefB, int grid, cudaStream_t st) {
  const int nch = N <= 64 ? 1 : N / 64;
  if (N == 32)
    return split3 ? launch_fwd_tma<32, true, EVAL>(h, h_coef, W, bias, M, 1, out, partial, s_in, coefA, coefB, grid, st)
                  : launch_fwd_tma<32, false, EVAL>(h, h_coef, W, bias, M, 1, out, partial, s_in, coefA, coefB, grid, st);
  return split3 ? launch_fwd_tma<64, true, EVAL>(h, h_coef, W, bias, M, nch, out, partial, s_in, coefA, coefB, grid, st)
                : launch_fwd_tma<64, false, EVAL>(h, h_coef, W, bias, M, nch, out, partial, s_in, coefA, coefB, grid, st);
}

}  // namespace vmtl

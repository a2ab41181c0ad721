// Warp-specialised tensor-core forward of the MTAN gate (training mode, K = 128, N in {32, 64}).
// Included by gate_tc.cu (uses its tile helpers).
//
// Measured on B200 (scratch/mma_probe*.cu): a tcgen05.mma kind::tf32 with N <= 64 occupies the
// tensor pipe for ~45 cycles whatever its operands (smem or TMEM), and the issuing thread is
// blocked for that time.  With 3 MMAs per K-step (3xTF32) a 128-row tile costs 48 x 45 = 2160
// cycles, more than the 1830 cycles its 80 KB take at the HBM roofline.  Hence:
//   * W_hi and W_lo are stacked into ONE B operand of 2N rows: per K-step
//         D[:, 0:2N] (+)= A_hi @ [W_hi ; W_lo]^T     (N_mma = 2N)
//         D[:, 0:N ]  += A_lo @ W_hi^T               (N_mma = N)
//     i.e. 2 MMAs instead of 3; the epilogue adds the two column halves.
//   * the MMAs are issued by a dedicated warp, so the ~1500 cycles it spends blocked per tile
//     overlap the producers' global loads / hi-lo split / smem stores and the epilogue;
//   * the A tile is produced and consumed in two K-halves (atoms {0,1} and {2,3} are separate smem
//     regions): the split+store of one half overlaps the MMAs of the other, and each half's
//     registers are refilled from global as soon as they are stored -> loads always in flight.
//
// Roles (17 warps): warps 0-7 producers, 8-15 epilogue (TMEM lane quadrant = warp % 4),
// warp 16 MMA issuer.  mbarriers: full[h] (256 producer arrivals), mma[h] (tcgen05.commit: half
// h consumed), dfull[b] (commit: accumulator b complete), dfree[b] (256 epilogue arrivals).
#pragma once

namespace vmtl {

constexpr int kWs2Threads = 17 * 32;

template <int NC>
struct Ws2Smem {
  static constexpr int kAtomA = kTileM * 128;
  static constexpr int kAtomB = 2 * NC * 128;  // rows [0,NC) = W_hi, rows [NC,2NC) = W_lo
  static constexpr int kAhi = 0;
  static constexpr int kAlo = kAhi + 4 * kAtomA;
  static constexpr int kB = kAlo + 4 * kAtomA;
  static constexpr int kMisc = kB + 4 * kAtomB;
  static constexpr int kBytes = kMisc + 128 + 64 * 4 + 1024;
};

// 8 float4 per thread = one K-half ([128 rows x 64 floats]) of a tile; a warp covers 2 rows x 256 B
__device__ __forceinline__ void ws2_load_half(const float* __restrict__ src, int64_t row0, int64_t M, int kh,
                                              float4 (&regs)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = i * 256 + threadIdx.x;
    const int row = q >> 4, kc = kh * 16 + (q & 15);
    const int64_t grow = row0 + row;
    regs[i] = grow < M ? ldg_stream(reinterpret_cast<const float4*>(src) + grow * 32 + kc)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <bool SPLIT>
__device__ __forceinline__ void ws2_store_half(uint8_t* hi_base, uint8_t* lo_base, int kh, const float4 (&regs)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = i * 256 + threadIdx.x;
    const int row = q >> 4, kc = kh * 16 + (q & 15);
    const uint32_t off = (uint32_t)((kc >> 3) * (kTileM * 128)) + tc::sw128_off(row, kc & 7);
    const float4 a = regs[i];
    const float4 hi = make_float4(tc::tf32_hi(a.x), tc::tf32_hi(a.y), tc::tf32_hi(a.z), tc::tf32_hi(a.w));
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    if (SPLIT)
      *reinterpret_cast<float4*>(lo_base + off) = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
  }
}

template <int NC, bool SPLIT>
__global__ void __launch_bounds__(kWs2Threads, 1)
    gate_tc_fwd_ws2_kernel(const float* __restrict__ h, const float* __restrict__ W,
                           const float* __restrict__ bias, int64_t M, float* __restrict__ z_out,
                           float* __restrict__ partial /* [gridDim.x][2][N] */) {
  using namespace tc;
  using L = Ws2Smem<NC>;
  constexpr int N = NC;
  constexpr int V = NC / 2;                 // columns per epilogue thread
  constexpr int DC = SPLIT ? 2 * NC : NC;   // accumulator columns per buffer
  constexpr uint32_t kTmemCols = 2 * DC;    // 64..256, power of two
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sAhi = smem + L::kAhi;
  uint8_t* sAlo = smem + L::kAlo;
  uint8_t* sB = smem + L::kB;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);  // full[2] mma[2] dfull[2] dfree[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 96);
  float* s_bias = reinterpret_cast<float*>(smem + L::kMisc + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int hh) { return bar0 + 8u * (uint32_t)hh; };
  auto bar_mma = [&](int hh) { return bar0 + 16u + 8u * (uint32_t)hh; };
  auto bar_dfull = [&](int b) { return bar0 + 32u + 8u * (uint32_t)b; };
  auto bar_dfree = [&](int b) { return bar0 + 48u + 8u * (uint32_t)b; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_full(i), 256);
      mbar_init(bar_mma(i), 1);
      mbar_init(bar_dfull(i), 1);
      mbar_init(bar_dfree(i), 256);
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kWs2Threads) s_bias[i] = bias[i];
  if (warp < 8) {  // stacked W operand: row n = W_hi[n], row NC + n = W_lo[n]; K-major, SW128
    for (int q = threadIdx.x; q < NC * 32; q += 256) {
      const int n = q >> 5, kc = q & 31;
      const float4 w = __ldg(reinterpret_cast<const float4*>(W) + q);
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

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  double st_sum = 0.0, st_sq = 0.0;

  if (warp < 8) {
    // ------------------------------------------------------------------ producers
    float4 r0[8], r1[8];
    if (nitems > 0) {
      ws2_load_half(h, (int64_t)blockIdx.x * kTileM, M, 0, r0);
      ws2_load_half(h, (int64_t)blockIdx.x * kTileM, M, 1, r1);
    }
    for (int64_t it = 0; it < nitems; ++it) {
      const int64_t next_row0 = (blockIdx.x + (it + 1) * gridDim.x) * kTileM;
      const bool has_next = it + 1 < nitems;
      if (it > 0) mbar_wait(bar_mma(0), (uint32_t)((it - 1) & 1));  // half 0 of the previous tile consumed
      ws2_store_half<SPLIT>(sAhi, sAlo, 0, r0);
      fence_proxy_async_smem();
      mbar_arrive(bar_full(0));
      if (has_next) ws2_load_half(h, next_row0, M, 0, r0);
      if (it > 0) mbar_wait(bar_mma(1), (uint32_t)((it - 1) & 1));
      ws2_store_half<SPLIT>(sAhi, sAlo, 1, r1);
      fence_proxy_async_smem();
      mbar_arrive(bar_full(1));
      if (has_next) ws2_load_half(h, next_row0, M, 1, r1);
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 8;  // lane quadrant ew & 3, column half ew >> 2
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      mbar_wait(bar_dfull(b), (uint32_t)((it >> 1) & 1));
      tc_fence_after_sync();
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int col0 = (ew >> 2) * V;
      const int64_t row = tile * kTileM + (ew & 3) * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (((uint32_t)(ew & 3) * 32) << 16) + (uint32_t)(b * DC + col0);
      float v[V];
#pragma unroll
      for (int j = 0; j < V; j += 16) {
        float t16[16];
        tmem_ld16(taddr + j, t16);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[j + e] = t16[e] + s_bias[col0 + j + e];
        if (SPLIT) {  // + A_hi @ W_lo^T, accumulated in the second column half
          tmem_ld16(taddr + NC + j, t16);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[j + e] += t16[e];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(bar_dfree(b));
      if (row_ok) {
        float4* zp = reinterpret_cast<float4*>(z_out + row * N + col0);
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
  } else if (lane == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc_wide = idesc_tf32(kTileM, DC, 0, 0);
    constexpr uint32_t idesc_n = idesc_tf32(kTileM, NC, 0, 0);
    const uint32_t aH = smem_u32(sAhi), aL = smem_u32(sAlo), bW = smem_u32(sB);
    for (int64_t it = 0; it < nitems; ++it) {
      const int b = (int)(it & 1);
      const uint32_t d_tmem = tmem_base + (uint32_t)(b * DC);
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(bar_full(kh), (uint32_t)(it & 1));
        if (kh == 0 && it >= 2) mbar_wait(bar_dfree(b), (uint32_t)(((it >> 1) - 1) & 1));
        tc_fence_after_sync();
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
          const int atom = kh * 2 + a2;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ao = atom * L::kAtomA + ks * 32, bo = atom * L::kAtomB + ks * 32;
            const uint64_t dB = smem_desc_sw128(bW + bo, 16, 1024);
            mma_tf32(d_tmem, smem_desc_sw128(aH + ao, 16, 1024), dB, idesc_wide, (kh | a2 | ks) != 0);
            if (SPLIT) mma_tf32(d_tmem, smem_desc_sw128(aL + ao, 16, 1024), dB, idesc_n, 1);
          }
        }
        mma_commit(bar_mma(kh));
        if (kh == 1) mma_commit(bar_dfull(b));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, kTmemCols);
  // per-CTA column partials: the four quadrant warps of a column half, summed in fixed order
  double* s_red = reinterpret_cast<double*>(sAhi);  // [8 epilogue warps][V][2]
  if (warp >= 8 && warp < 16 && lane < V) {
    s_red[((warp - 8) * V + lane) * 2] = st_sum;
    s_red[((warp - 8) * V + lane) * 2 + 1] = st_sq;
  }
  __syncthreads();
  for (int col = threadIdx.x; col < N; col += kWs2Threads) {
    const int half = col / V, l = col % V;
    double a = 0.0, bq = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a += s_red[((half * 4 + q) * V + l) * 2];
      bq += s_red[((half * 4 + q) * V + l) * 2 + 1];
    }
    partial[(int64_t)blockIdx.x * 2 * N + col] = (float)a;
    partial[(int64_t)blockIdx.x * 2 * N + N + col] = (float)bq;
  }
}

template <int NC, bool SPLIT>
static int launch_fwd_ws(const float* h, const float* W, const float* bias, int64_t M, float* z, float* partial,
                         int grid, cudaStream_t st) {
  using L = Ws2Smem<NC>;
  auto kern = gate_tc_fwd_ws2_kernel<NC, SPLIT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  kern<<<grid, kWs2Threads, L::kBytes, st>>>(h, W, bias, M, z, partial);
  return launch_status();
}

}  // namespace vmtl

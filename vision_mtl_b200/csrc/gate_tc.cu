// MTAN attention gate -- tensor-core contraction kernels (tcgen05.mma kind::tf32, TMEM).
//
// z[M,N] = h[M,K] @ W[N,K]^T + bias, K = 128 (or 64), fp32-grade accuracy through a tf32
// hi/lo split:  h = h_hi + h_lo, W = W_hi + W_lo,  z ~= h_lo W_hi + h_hi W_lo + h_hi W_hi
// (3 MMAs per K step, fp32 accumulation in TMEM; the dropped h_lo*W_lo term is ~2^-22).
//
// One persistent CTA per SM walks 128-row tiles of h:
//   global -> registers (coalesced 512B rows, prefetched one tile ahead)
//          -> hi/lo split -> 128B-swizzled shared memory (K-major UMMA operand layout)
//          -> tcgen05.mma (one elected thread, 16 K-steps x 3 passes) -> TMEM [128 x NC] fp32
//          -> tcgen05.ld (one row per thread) -> epilogue.
// The accumulator is double-buffered in TMEM so the epilogue of tile i-1 and the global loads
// of tile i+1 overlap the MMAs of tile i.  N > 64 is processed in 64-column chunks (the W
// chunk is re-staged per work item; those sites carry ~10% of the gate bytes).
//
// Epilogues:
//   TRAIN : z = acc + bias -> save_z ; per-column sum / sum of squares via a warp butterfly
//           (31 shuffles per 32 columns), accumulated in fp64 registers across tiles and
//           written once per CTA -> fixed-order finalize (deterministic batch statistics).
//   EVAL  : y = s * sigmoid(A*(acc + bias) + B)  (BN folded), single pass.
#include <math.h>

#include "gate_internal.cuh"
#include "tcgen05.cuh"

namespace vmtl {

using namespace tc;

constexpr int kTcThreads = 256;
constexpr int kTileM = 128;

// shared memory carve-up (bytes), all operand tiles 1024B aligned
template <int KATOMS, int NC>
struct FwdSmem {
  static constexpr int kAtomA = kTileM * 128;          // one K-atom (32 floats) of the A tile
  static constexpr int kAtomB = NC * 128;              // one K-atom of the W chunk
  static constexpr int kAhi = 0;
  static constexpr int kAlo = kAhi + KATOMS * kAtomA;
  static constexpr int kBhi = kAlo + KATOMS * kAtomA;
  static constexpr int kBlo = kBhi + KATOMS * kAtomB;
  static constexpr int kMisc = kBlo + KATOMS * kAtomB;  // mbarrier, tmem address, bias/coef
  static constexpr int kBytes = kMisc + 64 + 3 * 256 * 4 + 1024 /*alignment slack*/;
};

// butterfly transpose-reduce: on return lane l holds the sum over the 32 lanes of v[l % V]
template <int V>
__device__ __forceinline__ float butterfly_colsum(float (&v)[V], int lane) {
#pragma unroll
  for (int s = V / 2; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? v[j] : v[j + s];
      const float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = V; o < 32; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// stage rows [row0, row0+rows) x K floats of a row-major fp32 matrix into hi/lo swizzled tiles
template <int KATOMS, int ROWS>
__device__ __forceinline__ void load_tile_regs(const float* __restrict__ src, int64_t row0, int64_t nrows_total,
                                               float4 (&regs)[ROWS * KATOMS * 8 / kTcThreads]) {
  constexpr int K4 = KATOMS * 8;  // float4 per row
  constexpr int PER = ROWS * K4 / kTcThreads;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int q = it * kTcThreads + threadIdx.x;
    const int row = q / K4, kc = q % K4;
    const int64_t grow = row0 + row;
    regs[it] = grow < nrows_total
                   ? ldg_stream(reinterpret_cast<const float4*>(src) + grow * K4 + kc)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int KATOMS, int ROWS, bool SPLIT>
__device__ __forceinline__ void store_tile_split(uint8_t* hi_base, uint8_t* lo_base, int atom_bytes,
                                                 const float4 (&regs)[ROWS * KATOMS * 8 / kTcThreads]) {
  constexpr int K4 = KATOMS * 8;
  constexpr int PER = ROWS * K4 / kTcThreads;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int q = it * kTcThreads + threadIdx.x;
    const int row = q / K4, kc = q % K4;
    const int atom = kc >> 3, c = kc & 7;
    const uint32_t off = (uint32_t)(atom * atom_bytes) + sw128_off(row, c);
    const float4 a = regs[it];
    float4 hi = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    if (SPLIT)
      *reinterpret_cast<float4*>(lo_base + off) = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
  }
}

// issue the MMAs of one [128 x NC] x K work item (single thread)
template <int KATOMS, int NC, bool SPLIT>
__device__ __forceinline__ void issue_item(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                           uint32_t b_lo, uint32_t bar) {
  constexpr uint32_t idesc = idesc_tf32(kTileM, NC, 0, 0);
  constexpr int kAtomA = kTileM * 128, kAtomB = NC * 128;
  uint32_t acc = 0;
#pragma unroll
  for (int atom = 0; atom < KATOMS; ++atom) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 8 tf32 (32 bytes) per MMA K-step
      const uint32_t ao = atom * kAtomA + ks * 32, bo = atom * kAtomB + ks * 32;
      const uint64_t dAh = smem_desc_sw128(a_hi + ao, 16, 1024);
      const uint64_t dBh = smem_desc_sw128(b_hi + bo, 16, 1024);
      if (SPLIT) {
        const uint64_t dAl = smem_desc_sw128(a_lo + ao, 16, 1024);
        const uint64_t dBl = smem_desc_sw128(b_lo + bo, 16, 1024);
        mma_tf32(tmem_d, dAl, dBh, idesc, acc);
        acc = 1;
        mma_tf32(tmem_d, dAh, dBl, idesc, 1);
      }
      mma_tf32(tmem_d, dAh, dBh, idesc, acc);
      acc = 1;
    }
  }
  mma_commit(bar);
}

template <int KATOMS, int NC, int NCH, bool SPLIT, bool EVAL>
__global__ void __launch_bounds__(kTcThreads, 1)
    gate_tc_fwd_kernel(const float* __restrict__ h, const float* __restrict__ W,
                       const float* __restrict__ bias, const float* __restrict__ s_in,
                       const float* __restrict__ coefA, const float* __restrict__ coefB, int64_t M,
                       float* __restrict__ out /* TRAIN: z ; EVAL: y */,
                       float* __restrict__ partial /* TRAIN: [gridDim.x][2][N] */) {
  using L = FwdSmem<KATOMS, NC>;
  constexpr int N = NC * NCH;
  constexpr int K4 = KATOMS * 8;
  constexpr int V = NC / 2;          // columns per thread in the epilogue
  constexpr uint32_t kTmemCols = 2 * NC;  // 64 or 128: power of two >= 32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sAhi = smem + L::kAhi;
  uint8_t* sAlo = smem + L::kAlo;
  uint8_t* sBhi = smem + L::kBhi;
  uint8_t* sBlo = smem + L::kBlo;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 16);
  float* s_bias = reinterpret_cast<float*>(smem + L::kMisc + 64);
  float* s_cA = s_bias + 256;
  float* s_cB = s_cA + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = smem_u32(s_bar);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kTcThreads) {
    s_bias[i] = bias[i];
    if (EVAL) {
      s_cA[i] = coefA[i];
      s_cB[i] = coefB[i];
    }
  }
  // W chunk staging: rows = output channels of the chunk, K-major, same swizzle as A
  auto stage_w = [&](int chunk) {
    constexpr int PERW = NC * K4 / kTcThreads;
    float4 wr[PERW];
    load_tile_regs<KATOMS, NC>(W + (int64_t)chunk * NC * KATOMS * 32, 0, NC, wr);
    store_tile_split<KATOMS, NC, SPLIT>(sBhi, sBlo, L::kAtomB, wr);
  };
  if (NCH == 1) stage_w(0);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t nitems = my_tiles * NCH;

  constexpr int PERA = kTileM * K4 / kTcThreads;
  float4 areg[PERA];
  if (nitems > 0) load_tile_regs<KATOMS, kTileM>(h, (int64_t)blockIdx.x * kTileM, M, areg);

  double st_sum[NCH], st_sq[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) st_sum[c] = st_sq[c] = 0.0;

  // epilogue of work item `it` (its accumulator is complete)
  auto epilogue = [&](int64_t it) {
    const int64_t tile = blockIdx.x + (it / NCH) * gridDim.x;
    const int chunk = (int)(it % NCH);
    const int col0 = chunk * NC + (warp >> 2) * V;  // first column of this thread
    const int64_t row = tile * kTileM + (warp & 3) * 32 + lane;
    const bool row_ok = row < M;
    const uint32_t taddr = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)((it & 1) * NC + (warp >> 2) * V);
    float v[V];
#pragma unroll
    for (int j = 0; j < V; j += 16) {
      float t16[16];
      tmem_ld16(taddr + j, t16);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[j + e] = t16[e] + s_bias[col0 + j + e];
    }
    if (EVAL) {
      if (row_ok) {
        const float4* sp = reinterpret_cast<const float4*>(s_in + row * N + col0);
        float4* yp = reinterpret_cast<float4*>(out + row * N + col0);
#pragma unroll
        for (int j = 0; j < V; j += 4) {
          const float4 sv = ldg_stream(sp + j / 4);
          float4 y;
          y.x = sv.x * sigmoidf_acc(fmaf(s_cA[col0 + j], v[j], s_cB[col0 + j]));
          y.y = sv.y * sigmoidf_acc(fmaf(s_cA[col0 + j + 1], v[j + 1], s_cB[col0 + j + 1]));
          y.z = sv.z * sigmoidf_acc(fmaf(s_cA[col0 + j + 2], v[j + 2], s_cB[col0 + j + 2]));
          y.w = sv.w * sigmoidf_acc(fmaf(s_cA[col0 + j + 3], v[j + 3], s_cB[col0 + j + 3]));
          stg_stream(yp + j / 4, y);
        }
      }
    } else {
      if (row_ok) {
        float4* zp = reinterpret_cast<float4*>(out + row * N + col0);
#pragma unroll
        for (int j = 0; j < V; j += 4) stg_stream(zp + j / 4, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      }
      float sq[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (!row_ok) v[j] = 0.f;
        sq[j] = v[j] * v[j];
      }
      const float cs = butterfly_colsum<V>(v, lane);
      const float cq = butterfly_colsum<V>(sq, lane);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c == chunk) {
          st_sum[c] += (double)cs;
          st_sq[c] += (double)cq;
        }
    }
  };

  for (int64_t it = 0; it < nitems; ++it) {
    const int chunk = (int)(it % NCH);
    if (it > 0) {
      mbar_wait(bar, (uint32_t)((it - 1) & 1));  // MMAs of item it-1 done: smem operands reusable
      tc_fence_after_sync();
    }
    if (chunk == 0) store_tile_split<KATOMS, kTileM, SPLIT>(sAhi, sAlo, L::kAtomA, areg);
    if (NCH > 1) stage_w(chunk);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after_sync();
      issue_item<KATOMS, NC, SPLIT>(tmem_base + (uint32_t)((it & 1) * NC), smem_u32(sAhi), smem_u32(sAlo),
                                    smem_u32(sBhi), smem_u32(sBlo), bar);
    }
    // prefetch the next A tile while the tensor core works
    if (chunk == NCH - 1 && it + 1 < nitems) {
      const int64_t next_tile = blockIdx.x + ((it + 1) / NCH) * gridDim.x;
      load_tile_regs<KATOMS, kTileM>(h, next_tile * kTileM, M, areg);
    }
    if (it > 0) epilogue(it - 1);
  }
  if (nitems > 0) {
    mbar_wait(bar, (uint32_t)((nitems - 1) & 1));
    tc_fence_after_sync();
    epilogue(nitems - 1);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);

  if (!EVAL) {
    // per-CTA column partials: quadrant warps (same column half) summed in fixed order
    double* s_red = reinterpret_cast<double*>(sAhi);  // [8 warps][NCH][V][2]
    if (lane < V) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        s_red[((warp * NCH + c) * V + lane) * 2] = st_sum[c];
        s_red[((warp * NCH + c) * V + lane) * 2 + 1] = st_sq[c];
      }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < N; col += kTcThreads) {
      const int c = col / NC, half = (col % NC) / V, l = col % V;
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int w = half * 4 + q;
        a += s_red[((w * NCH + c) * V + l) * 2];
        b += s_red[((w * NCH + c) * V + l) * 2 + 1];
      }
      partial[(int64_t)blockIdx.x * 2 * N + col] = (float)a;
      partial[(int64_t)blockIdx.x * 2 * N + N + col] = (float)b;
    }
  }
}

template <int KATOMS, int NC, int NCH, bool SPLIT, bool EVAL>
static int launch_fwd(const float* h, const float* W, const float* bias, const float* s,
                      const float* coefA, const float* coefB, int64_t M, float* out, float* partial,
                      int grid, cudaStream_t st) {
  using L = FwdSmem<KATOMS, NC>;
  auto kern = gate_tc_fwd_kernel<KATOMS, NC, NCH, SPLIT, EVAL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  kern<<<grid, kTcThreads, L::kBytes, st>>>(h, W, bias, s, coefA, coefB, M, out, partial);
  return launch_status();
}

template <bool EVAL>
static int dispatch_fwd(const float* h, const float* W, const float* bias, const float* s,
                        const float* coefA, const float* coefB, int64_t M, int K, int N, int split3,
                        float* out, float* partial, int grid, cudaStream_t st) {
#define VMTL_FWD(KA, NC_, NCH_)                                                                   \
  (split3 ? launch_fwd<KA, NC_, NCH_, true, EVAL>(h, W, bias, s, coefA, coefB, M, out, partial,   \
                                                  grid, st)                                       \
          : launch_fwd<KA, NC_, NCH_, false, EVAL>(h, W, bias, s, coefA, coefB, M, out, partial,  \
                                                   grid, st))
  if (K == 128) {
    switch (N) {
      case 32: return VMTL_FWD(4, 32, 1);
      case 64: return VMTL_FWD(4, 64, 1);
      case 128: return VMTL_FWD(4, 64, 2);
      case 192: return VMTL_FWD(4, 64, 3);
      case 256: return VMTL_FWD(4, 64, 4);
      default: return VMTL_EUNSUPPORTED;
    }
  }
  if (K == 64) {
    switch (N) {
      case 32: return VMTL_FWD(2, 32, 1);
      case 64: return VMTL_FWD(2, 64, 1);
      case 128: return VMTL_FWD(2, 64, 2);
      case 192: return VMTL_FWD(2, 64, 3);
      case 256: return VMTL_FWD(2, 64, 4);
      default: return VMTL_EUNSUPPORTED;
    }
  }
#undef VMTL_FWD
  return VMTL_EUNSUPPORTED;
}

static int tc_grid(int64_t M) {
  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int sms = sm_count();
  return (int)(ntiles < sms ? ntiles : sms);
}

int gate_tc_fwd_gemm(const float* h, const float* W, const float* bias, int64_t M, int K, int N,
                     int split3, float* z_out, float* partial, int partial_rows, int* nparts,
                     cudaStream_t st) {
  const int grid = tc_grid(M);
  if (grid > partial_rows) return VMTL_EWORKSPACE;
  *nparts = grid;
  return dispatch_fwd<false>(h, W, bias, nullptr, nullptr, nullptr, M, K, N, split3, z_out, partial, grid, st);
}

int gate_tc_fwd_eval(const float* h, const float* s, const float* W, const float* bias,
                     const float* coefA, const float* coefB, int64_t M, int K, int N, int split3,
                     float* y, cudaStream_t st) {
  return dispatch_fwd<true>(h, W, bias, s, coefA, coefB, M, K, N, split3, y, nullptr, tc_grid(M), st);
}

int gate_tc_bwd_gemm(const float*, const float*, const float*, const float*, const float*, const GateWs&,
                     const float*, int64_t, int, int, int, float*, float*, int, int*, float*,
                     cudaStream_t) {
  return VMTL_EUNSUPPORTED;  // falls back to the materialised-dz path in gate.cu for now
}

}  // namespace vmtl

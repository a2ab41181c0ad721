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
#include "tma_host.cuh"

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

template <int KATOMS, int ROWS, bool SPLIT, bool MN32 = false>
__device__ __forceinline__ void store_tile_split(uint8_t* hi_base, uint8_t* lo_base, int atom_bytes,
                                                 const float4 (&regs)[ROWS * KATOMS * 8 / kTcThreads]) {
  constexpr int K4 = KATOMS * 8;
  constexpr int PER = ROWS * K4 / kTcThreads;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int q = it * kTcThreads + threadIdx.x;
    const int row = q / K4, kc = q % K4;
    const int atom = kc >> 3, c = kc & 7;
    const uint32_t off = (uint32_t)(atom * atom_bytes) + (MN32 ? sw128b32_off(row, c) : sw128_off(row, c));
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
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
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

}  // namespace vmtl
#include "gate_tc_ws.cuh"   // warp-specialised forward (uses the tile helpers above)
#include "gate_tc_tma.cuh"  // TMA + TMEM-resident A operand forward
#include "gate_tc_bwd_tma.cuh"  // same machinery for dh / dW
namespace vmtl {

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
  int grid = tc_grid(M);
  // VMTL_GATE_FWD = tma (default) | ws | base : kernel generation, for A/B comparisons
  static const int variant = [] {
    const char* e = getenv("VMTL_GATE_FWD");
    if (e && e[0] == 'b') return 0;
    if (e && e[0] == 'w') return 1;
    return 2;
  }();
  if (variant == 2 && K == 128 && (N == 32 || N == 64 || N == 128 || N == 192 || N == 256)) {
    // N > 64: (row tile, 64-column chunk) items; a CTA keeps one chunk so its W operand is staged once
    const int nch = N <= 64 ? 1 : N / 64;
    const int64_t items = ((M + kTileM - 1) / kTileM) * nch;
    grid = (int)(items < sm_count() ? items : sm_count());
    grid = grid / nch * nch;
    if (grid < nch) grid = nch;
    if (grid > partial_rows) return VMTL_EWORKSPACE;
    *nparts = grid;
#define VMTL_TMA(NC_, NCH_)                                                                         \
  (split3 ? launch_fwd_tma<NC_, NCH_, true>(h, W, bias, M, z_out, partial, grid, st)                \
          : launch_fwd_tma<NC_, NCH_, false>(h, W, bias, M, z_out, partial, grid, st))
    switch (N) {
      case 32: return VMTL_TMA(32, 1);
      case 64: return VMTL_TMA(64, 1);
      case 128: return VMTL_TMA(64, 2);
      case 192: return VMTL_TMA(64, 3);
      default: return VMTL_TMA(64, 4);
    }
#undef VMTL_TMA
  }
  if (grid > partial_rows) return VMTL_EWORKSPACE;
  *nparts = grid;
  if (variant >= 1 && K == 128 && (N == 32 || N == 64)) {
    if (N == 32)
      return split3 ? launch_fwd_ws<32, true>(h, W, bias, M, z_out, partial, grid, st)
                    : launch_fwd_ws<32, false>(h, W, bias, M, z_out, partial, grid, st);
    return split3 ? launch_fwd_ws<64, true>(h, W, bias, M, z_out, partial, grid, st)
                  : launch_fwd_ws<64, false>(h, W, bias, M, z_out, partial, grid, st);
  }
  return dispatch_fwd<false>(h, W, bias, nullptr, nullptr, nullptr, M, K, N, split3, z_out, partial, grid, st);
}

int gate_tc_fwd_eval(const float* h, const float* s, const float* W, const float* bias,
                     const float* coefA, const float* coefB, int64_t M, int K, int N, int split3,
                     float* y, cudaStream_t st) {
  return dispatch_fwd<true>(h, W, bias, s, coefA, coefB, M, K, N, split3, y, nullptr, tc_grid(M), st);
}

// =============================================================================================
// Backward, phase B on tensor cores (N in {32, 64}, K = 128).
//
//   B1  gate_tc_dh_kernel : per 128-row tile, dz = gamma*invstd*(du - c1 - zhat*c2) is rebuilt
//       from (dy, s, z) in registers, written once to a [M,N] scratch, split into tf32 hi/lo in
//       swizzled smem and contracted with W^T:  dh[128 x K] = dz[128 x N] @ W[N x K]  (K-major
//       operands, N/8 K-steps x 3 passes, accumulator double-buffered in TMEM).  db = sum dz is
//       accumulated per thread (fixed column group) and reduced once per CTA.
//   B2  gate_tc_dw_kernel : dW^T[K x N] += h^T[K x rows] @ dz[rows x N].  Both operands are read
//       "MN-major" straight from their natural row-major tiles (one 128-byte row of 32 channels
//       per pixel = one K index) in the SWIZZLE_128B_BASE32B layout tf32 requires; the 8 pixel
//       rows of one MMA K-step are two 4-row groups.  The accumulator stays in TMEM across ALL
//       tiles of the CTA and is drained once -> one [N x K] partial per CTA, fixed-order second
//       stage.
// =============================================================================================
template <int NATOMS>
struct DhSmem {
  static constexpr int kAtom = kTileM * 128;                    // [128 rows x 128 B]
  static constexpr int kAhi = 0;                                // dz hi, NATOMS atoms
  static constexpr int kAlo = kAhi + NATOMS * kAtom;
  static constexpr int kBhi = kAlo + NATOMS * kAtom;            // W^T hi: rows = k (128), NATOMS atoms
  static constexpr int kBlo = kBhi + NATOMS * kAtom;
  static constexpr int kMisc = kBlo + NATOMS * kAtom;
  static constexpr int kBytes = kMisc + 64 + 6 * 64 * 4 + 1024;
};

template <int NATOMS, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
    gate_tc_dh_kernel(const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ z,
                      const float* __restrict__ W /* [N,128] */, const float* __restrict__ coefA,
                      const float* __restrict__ coefB, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ c1,
                      const float* __restrict__ c2, int64_t M, float* __restrict__ dz_out,
                      float* __restrict__ dh, float* __restrict__ db_partial /* [grid][N] */) {
  using L = DhSmem<NATOMS>;
  constexpr int N = NATOMS * 32;
  constexpr int KH = 128;                 // hidden width = MMA N
  constexpr int N4 = N / 4;
  constexpr int PER = kTileM * N4 / kTcThreads;  // float4 of dz per thread per tile (4 or 8)
  constexpr uint32_t kTmemCols = 2 * KH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sAhi = smem + L::kAhi;
  uint8_t* sAlo = smem + L::kAlo;
  uint8_t* sBhi = smem + L::kBhi;
  uint8_t* sBlo = smem + L::kBlo;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 16);
  float* s_coef = reinterpret_cast<float*>(smem + L::kMisc + 64);  // [6][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = smem_u32(s_bar);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kTcThreads) {
    s_coef[0 * 64 + i] = coefA[i];
    s_coef[1 * 64 + i] = coefB[i];
    s_coef[2 * 64 + i] = mean[i];
    s_coef[3 * 64 + i] = invstd[i];
    s_coef[4 * 64 + i] = c1[i];
    s_coef[5 * 64 + i] = c2[i];
  }
  // W^T staging: element (k, n) of the B operand = W[n][k]; rows k, K-major along n
  for (int e = threadIdx.x; e < N * KH; e += kTcThreads) {
    const int n = e / KH, k = e - n * KH;  // coalesced read of W
    const float w = W[e];
    const float hi = tf32_hi(w);
    const uint32_t off = (uint32_t)((n >> 5) * L::kAtom) + sw128_off(k, (n & 31) >> 2) + (uint32_t)((n & 3) << 2);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    if (SPLIT) *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // per-thread column group is fixed: q = it*256 + tid, c4 = q % N4 = tid % N4
  const int c4 = threadIdx.x % N4;
  float4 cA, cB, cMu, cRs, cK1, cK2;
  {
    const float4* p = reinterpret_cast<const float4*>(s_coef);
    cA = p[0 * 16 + c4]; cB = p[1 * 16 + c4]; cMu = p[2 * 16 + c4];
    cRs = p[3 * 16 + c4]; cK1 = p[4 * 16 + c4]; cK2 = p[5 * 16 + c4];
  }
  float4 db_acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 rdy[PER], rs[PER], rz[PER];
  auto load_inputs = [&](int64_t tile) {
    load_tile_regs<NATOMS, kTileM>(dy, tile * kTileM, M, rdy);
    load_tile_regs<NATOMS, kTileM>(s, tile * kTileM, M, rs);
    load_tile_regs<NATOMS, kTileM>(z, tile * kTileM, M, rz);
  };
  auto dz1 = [](float g, float sv, float zv, float A, float B, float mu, float r, float k1, float k2) {
    const float a = sigmoidf_acc(fmaf(A, zv, B));
    return A * (g * sv * a * (1.f - a) - k1 - (zv - mu) * r * k2);
  };
  // dz of the tile held in registers -> global scratch + hi/lo smem operand
  auto produce_dz = [&](int64_t tile) {
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int q = it * kTcThreads + threadIdx.x;
      const int row = q / N4, kc = q % N4;
      const int64_t grow = tile * kTileM + row;
      float4 d;
      d.x = dz1(rdy[it].x, rs[it].x, rz[it].x, cA.x, cB.x, cMu.x, cRs.x, cK1.x, cK2.x);
      d.y = dz1(rdy[it].y, rs[it].y, rz[it].y, cA.y, cB.y, cMu.y, cRs.y, cK1.y, cK2.y);
      d.z = dz1(rdy[it].z, rs[it].z, rz[it].z, cA.z, cB.z, cMu.z, cRs.z, cK1.z, cK2.z);
      d.w = dz1(rdy[it].w, rs[it].w, rz[it].w, cA.w, cB.w, cMu.w, cRs.w, cK1.w, cK2.w);
      if (grow >= M) d = make_float4(0.f, 0.f, 0.f, 0.f);
      else stg_stream(reinterpret_cast<float4*>(dz_out) + grow * N4 + kc, d);
      db_acc.x += d.x; db_acc.y += d.y; db_acc.z += d.z; db_acc.w += d.w;
      const uint32_t off = (uint32_t)((kc >> 3) * L::kAtom) + sw128_off(row, kc & 7);
      const float4 hi = make_float4(tf32_hi(d.x), tf32_hi(d.y), tf32_hi(d.z), tf32_hi(d.w));
      *reinterpret_cast<float4*>(sAhi + off) = hi;
      if (SPLIT)
        *reinterpret_cast<float4*>(sAlo + off) = make_float4(d.x - hi.x, d.y - hi.y, d.z - hi.z, d.w - hi.w);
    }
  };
  auto issue = [&](int64_t it) {
    constexpr uint32_t idesc = idesc_tf32(kTileM, KH, 0, 0);
    const uint32_t d_tmem = tmem_base + (uint32_t)((it & 1) * KH);
    const uint32_t aH = smem_u32(sAhi), aL = smem_u32(sAlo), bH = smem_u32(sBhi), bL = smem_u32(sBlo);
    uint32_t acc = 0;
#pragma unroll
    for (int atom = 0; atom < NATOMS; ++atom)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t o = atom * L::kAtom + ks * 32;
        const uint64_t dAh = smem_desc_sw128(aH + o, 16, 1024), dBh = smem_desc_sw128(bH + o, 16, 1024);
        if (SPLIT) {
          mma_tf32(d_tmem, smem_desc_sw128(aL + o, 16, 1024), dBh, idesc, acc);
          acc = 1;
          mma_tf32(d_tmem, dAh, smem_desc_sw128(bL + o, 16, 1024), idesc, 1);
        }
        mma_tf32(d_tmem, dAh, dBh, idesc, acc);
        acc = 1;
      }
    mma_commit(bar);
  };
  auto epilogue = [&](int64_t it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    const int64_t row = tile * kTileM + (warp & 3) * 32 + lane;
    const int col0 = (warp >> 2) * (KH / 2);
    const uint32_t taddr = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)((it & 1) * KH + col0);
#pragma unroll
    for (int j = 0; j < KH / 2; j += 16) {
      float t16[16];
      tmem_ld16(taddr + j, t16);
      if (row < M) {
        float4* o = reinterpret_cast<float4*>(dh + row * KH + col0 + j);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          stg_stream(o + e, make_float4(t16[4 * e], t16[4 * e + 1], t16[4 * e + 2], t16[4 * e + 3]));
      }
    }
  };

  if (nitems > 0) load_inputs(blockIdx.x);
  for (int64_t it = 0; it < nitems; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    if (it > 0) {
      mbar_wait(bar, (uint32_t)((it - 1) & 1));
      tc_fence_after_sync();
    }
    produce_dz(tile);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after_sync();
      issue(it);
    }
    if (it + 1 < nitems) load_inputs(tile + gridDim.x);
    if (it > 0 && dh) epilogue(it - 1);
  }
  if (nitems > 0) {
    mbar_wait(bar, (uint32_t)((nitems - 1) & 1));
    tc_fence_after_sync();
    if (dh) epilogue(nitems - 1);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
  // db partial of this CTA: threads sharing a column group summed in fixed order
  float4* s_red = reinterpret_cast<float4*>(sAhi);
  s_red[threadIdx.x] = db_acc;
  __syncthreads();
  if (threadIdx.x < N4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = threadIdx.x; t < kTcThreads; t += N4) {
      const float4 v = s_red[t];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(db_partial + (int64_t)blockIdx.x * N)[threadIdx.x] = a;
  }
}

template <int NATOMS>
struct DwSmem {
  static constexpr int kAtom = kTileM * 128;
  static constexpr int kHhi = 0;                         // h tile, 4 atoms (k), MN-major A
  static constexpr int kHlo = kHhi + 4 * kAtom;
  static constexpr int kDhi = kHlo + 4 * kAtom;          // dz tile, NATOMS atoms (n), MN-major B
  static constexpr int kDlo = kDhi + NATOMS * kAtom;
  static constexpr int kMisc = kDlo + NATOMS * kAtom;
  static constexpr int kBytes = kMisc + 64 + 1024;
};

template <int NATOMS, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
    gate_tc_dw_kernel(const float* __restrict__ h, const float* __restrict__ dz, int64_t M,
                      float* __restrict__ dw_partial /* [grid][N][128] */) {
  using L = DwSmem<NATOMS>;
  constexpr int N = NATOMS * 32;
  constexpr int KH = 128;
  constexpr uint32_t kTmemCols = N < 32 ? 32 : N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* sHhi = smem + L::kHhi;
  uint8_t* sHlo = smem + L::kHlo;
  uint8_t* sDhi = smem + L::kDhi;
  uint8_t* sDlo = smem + L::kDlo;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kMisc);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::kMisc + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = smem_u32(s_bar);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t nitems = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  float4 rh[kTileM * 32 / kTcThreads];
  float4 rd[kTileM * NATOMS * 8 / kTcThreads];
  if (nitems > 0) {
    load_tile_regs<4, kTileM>(h, (int64_t)blockIdx.x * kTileM, M, rh);
    load_tile_regs<NATOMS, kTileM>(dz, (int64_t)blockIdx.x * kTileM, M, rd);
  }
  for (int64_t it = 0; it < nitems; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    if (it > 0) {
      mbar_wait(bar, (uint32_t)((it - 1) & 1));
      tc_fence_after_sync();
    }
    store_tile_split<4, kTileM, SPLIT, true>(sHhi, sHlo, L::kAtom, rh);
    store_tile_split<NATOMS, kTileM, SPLIT, true>(sDhi, sDlo, L::kAtom, rd);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after_sync();
      // D[k (128) x n (N)] += sum over the tile's 128 pixel rows, 8 rows (= one 1024 B group) per MMA
      constexpr uint32_t idesc = idesc_tf32(KH, N, 1, 1);
      const uint32_t hH = smem_u32(sHhi), hL = smem_u32(sHlo), dH = smem_u32(sDhi), dL = smem_u32(sDlo);
      uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll 1
      for (int ks = 0; ks < kTileM / 8; ++ks) {
        const uint32_t o = ks * 1024;
        // MN-major tf32: LBO = stride between 32-element atoms along M/N, SBO = 4-row group stride
        const uint64_t aH = smem_desc_mn_tf32(hH + o, L::kAtom, 512), bH = smem_desc_mn_tf32(dH + o, L::kAtom, 512);
        if (SPLIT) {
          mma_tf32(tmem_base, smem_desc_mn_tf32(hL + o, L::kAtom, 512), bH, idesc, acc);
          acc = 1;
          mma_tf32(tmem_base, aH, smem_desc_mn_tf32(dL + o, L::kAtom, 512), idesc, 1);
        }
        mma_tf32(tmem_base, aH, bH, idesc, acc);
        acc = 1;
      }
      mma_commit(bar);
    }
    if (it + 1 < nitems) {
      load_tile_regs<4, kTileM>(h, (tile + gridDim.x) * kTileM, M, rh);
      load_tile_regs<NATOMS, kTileM>(dz, (tile + gridDim.x) * kTileM, M, rd);
    }
  }
  float* out = dw_partial + (int64_t)blockIdx.x * N * KH;
  if (nitems > 0) {
    mbar_wait(bar, (uint32_t)((nitems - 1) & 1));
    tc_fence_after_sync();
    // drain: thread (k = TMEM lane) holds dW^T[k][n0..]; partial layout is dW[n][k]
    const int k = (warp & 3) * 32 + lane;
    const int col0 = (warp >> 2) * (N / 2);
    const uint32_t taddr = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)col0;
#pragma unroll
    for (int j = 0; j < N / 2; j += 16) {
      float t16[16];
      tmem_ld16(taddr + j, t16);
#pragma unroll
      for (int e = 0; e < 16; ++e) out[(int64_t)(col0 + j + e) * KH + k] = t16[e];
    }
  } else {
    for (int e = threadIdx.x; e < N * KH; e += kTcThreads) out[e] = 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int NATOMS, bool SPLIT>
static int launch_bwd(const float* dy, const float* h, const float* s, const float* z, const float* W,
                      const GateWs& ws, int64_t M, float* dh, float* dw_partial, float* db_partial,
                      int grid, cudaStream_t st) {
  auto k1 = gate_tc_dh_kernel<NATOMS, SPLIT>;
  auto k2 = gate_tc_dw_kernel<NATOMS, SPLIT>;
  if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, DhSmem<NATOMS>::kBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, DwSmem<NATOMS>::kBytes) != cudaSuccess)
    return VMTL_ECUDA;
  k1<<<grid, kTcThreads, DhSmem<NATOMS>::kBytes, st>>>(dy, s, z, W, ws.coefA, ws.coefB, ws.mean, ws.invstd, ws.c1,
                                                       ws.c2, M, ws.dz, dh, db_partial);
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  k2<<<grid, kTcThreads, DwSmem<NATOMS>::kBytes, st>>>(h, ws.dz, M, dw_partial);
  return launch_status();
}

int gate_tc_bwd_gemm(const float* dy, const float* h, const float* s, const float* z, const float* W,
                     const GateWs& ws, const float* gamma, int64_t M, int K, int N, int split3,
                     float* dh, float* dw_partial, int slots, int* nslots, float* db_partial,
                     cudaStream_t st) {
  (void)gamma;
  if (K != 128 || (N != 32 && N % 64 != 0) || N > 1024 || !ws.dz) return VMTL_EUNSUPPORTED;
  const int grid = tc_grid(M);
  if (grid > slots || grid > ws.partial_rows) return VMTL_EWORKSPACE;
  *nslots = grid;
  static const bool use_tma = [] {  // VMTL_GATE_BWD=base selects the register-staged kernels
    const char* e = getenv("VMTL_GATE_BWD");
    return !(e && e[0] == 'b');
  }();
  if (use_tma || N > 64) {
#define VMTL_BWD_TMA(NDW)                                                                                    \
  (split3 ? launch_bwd_tma<NDW, true>(dy, h, s, z, W, ws, M, N, dh, dw_partial, db_partial, grid, st)         \
          : launch_bwd_tma<NDW, false>(dy, h, s, z, W, ws, M, N, dh, dw_partial, db_partial, grid, st))
    if (N == 32) return VMTL_BWD_TMA(1);
    if (N % 128 == 0) return VMTL_BWD_TMA(4);
    return VMTL_BWD_TMA(2);
#undef VMTL_BWD_TMA
  }
  if (N == 32)
    return split3 ? launch_bwd<1, true>(dy, h, s, z, W, ws, M, dh, dw_partial, db_partial, grid, st)
                  : launch_bwd<1, false>(dy, h, s, z, W, ws, M, dh, dw_partial, db_partial, grid, st);
  return split3 ? launch_bwd<2, true>(dy, h, s, z, W, ws, M, dh, dw_partial, db_partial, grid, st)
                : launch_bwd<2, false>(dy, h, s, z, W, ws, M, dh, dw_partial, db_partial, grid, st);
}

}  // namespace vmtl

// Task heads fused with their losses and metrics.
//
//   head_ce_*      : 1x1 seg head (nn.Conv2d(32,C,1), mtan_model.py:367-376,401-404) fused with
//                    softmax/argmax (lit_module.py:137-138), nn.CrossEntropyLoss
//                    (lit_module.py:31,123) and the confusion matrix behind the torchmetrics
//                    calls (lit_module.py:109-111).  Logits [P,C] never reach HBM.
//   ce_logits_*    : the same loss/argmax/confusion on precomputed logits (3x3 heads of
//                    basic/csnet, basic_model.py:30-41).
//   head_silog_*   : 1x1 depth head + sigmoid (lit_module.py:139) + SILog moments
//                    (losses.py:14-36) + MAE / abs-rel error sums (lit_module.py:68,112).
//
// Algorithmic bytes per pixel (fp32, Cin = 32):
//   head_ce fwd 4*32 + 8 (+1 pred)   bwd 2*4*32 + 8
//   head_silog fwd 4*32 + 4 (+4 pred) bwd 2*4*32 + 4
//   ce_logits fwd 4C + 8 (+1)         bwd 2*4C + 8
// All cross-block reductions are two-stage with a fixed summation order (deterministic);
// the confusion matrix uses integer atomics only.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include "head_internal.cuh"
#include "tcgen05.cuh"
#include "vmtl_common.cuh"

namespace vmtl {

constexpr int kLossThreads = 256;
constexpr int kHeadCin = 32;  // fused 1x1 heads are specialised for the MTAN head width

static int loss_grid(int64_t P, int px_per_block, int blocks_per_sm) {
  int64_t want = (P + px_per_block - 1) / px_per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// Upper bound of blocks any loss kernel launches -> partial buffers are sized from it.
static int loss_max_blocks() { return sm_count() * 8; }

// ---------------------------------------------------------------------------------------
// per-pixel softmax cross-entropy on a register-resident logit vector
// ---------------------------------------------------------------------------------------
// cpad_for() buckets: CPAD 16 <- C in [1,16], 20 <- [17,20], 32 <- [21,32]: classes below kCmin<CPAD> exist
// for every C of the bucket, so only the last few columns need a run-time `c < C` test.
template <int CPAD>
constexpr int kCmin = CPAD == 16 ? 1 : (CPAD == 20 ? 17 : 21);
#define VMTL_HAS_CLASS(c, C) ((c) < kCmin<CPAD> || (c) < (C))

template <int CPAD>
struct PixelCE {
  float lse;    // log-sum-exp of the C logits
  float mneg;   // -max * log2(e)
  float inv_s;  // 1 / sum_c exp(l_c - max)
  int arg;
  // l[c] for c >= C is overwritten with -inf (its probability is then exactly 0)
  __device__ __forceinline__ void run(float (&l)[CPAD], int C) {
#pragma unroll
    for (int c = kCmin<CPAD>; c < CPAD; ++c) l[c] = c < C ? l[c] : -INFINITY;
    float m = l[0];
    arg = 0;
#pragma unroll
    for (int c = 1; c < CPAD; ++c)
      if (l[c] > m) {
        m = l[c];
        arg = c;
      }
    mneg = -m * kLog2e;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPAD; ++c) s += fast_ex2(fmaf(l[c], kLog2e, mneg));
    lse = fmaf(fast_lg2(s), kLn2, m);
    inv_s = __fdividef(1.f, s);
  }
  // softmax probability of a class with logit lc
  __device__ __forceinline__ float prob(float lc) const { return fast_ex2(fmaf(lc, kLog2e, mneg)) * inv_s; }
};

// l[t] without a compare per class: select inside groups of 4 by t % 4, then the group by t / 4
template <int CPAD>
__device__ __forceinline__ float pick(const float (&l)[CPAD], int t) {
  const bool q1 = (t & 3) == 1, q2 = (t & 3) == 2, q3 = (t & 3) == 3;
  float r = 0.f;
#pragma unroll
  for (int g = 0; g < CPAD / 4; ++g) {
    float v = l[4 * g];
    v = q1 ? l[4 * g + 1] : v;
    v = q2 ? l[4 * g + 2] : v;
    v = q3 ? l[4 * g + 3] : v;
    r = (t >> 2) == g ? v : r;
  }
  return r;
}

__device__ __forceinline__ bool target_valid(int64_t t, int C, int64_t ignore_index) {
  return t != ignore_index && (uint64_t)t < (uint64_t)C;
}

// block reduction of (loss, count) to one partial per block; deterministic
__device__ __forceinline__ void block_partial2(double a, double b, double* partial) {
  __shared__ double s_p[2][kLossThreads / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_p[0][warp] = a;
    s_p[1][warp] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) s += s_p[threadIdx.x][w];
    partial[(int64_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kFinThreads)
    ce_finalize(const double* __restrict__ partial, int nblocks, double* __restrict__ out,
                            float* __restrict__ loss) {
  __shared__ double s_v[2];
  const int col = threadIdx.x & 31;
  const double v = block_colsum(partial, nblocks, 2, col < 2 ? col : 0, col < 2);
  if (threadIdx.x < 2) s_v[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double s = s_v[0], n = s_v[1];
    out[0] = s;
    out[1] = n;
    if (loss) loss[0] = (float)(s / n);
  }
}

__device__ __forceinline__ void flush_conf(const unsigned int* s_conf, int CC,
                                           unsigned long long* conf) {
  for (int i = threadIdx.x; i < CC; i += blockDim.x) {
    const unsigned int v = s_conf[i];
    if (v) atomicAdd(conf + i, (unsigned long long)v);
  }
}

// ---------------------------------------------------------------------------------------
// Fused 1x1 head + CE.  One thread per pixel; the [256 x 32] feature tile is staged through
// shared memory with a 16-byte XOR swizzle so that both the coalesced global->shared copy
// and the per-thread row reads are bank-conflict free.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int swz(int row, int chunk) { return row * 8 + (chunk ^ (row & 7)); }

__device__ __forceinline__ void stage_tile32(float4* s_tile, const float4* __restrict__ g, int64_t p0,
                                             int npx) {
  // tile rows [p0, p0+npx) x 8 float4 ; consecutive threads read consecutive 16B
  const int n = npx * 8;
  const float4* src = g + p0 * 8;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += kLossThreads) {
    const float4 v = ldg_stream(src + i);
    s_tile[swz(i >> 3, i & 7)] = v;
  }
}

template <int CPAD>
__device__ __forceinline__ void head_logits(const float4* s_tile, const float* s_w, const float* s_b,
                                            int row, int C, float (&f)[kHeadCin], float (&l)[CPAD]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = s_tile[swz(row, j)];
    f[4 * j] = v.x;
    f[4 * j + 1] = v.y;
    f[4 * j + 2] = v.z;
    f[4 * j + 3] = v.w;
  }
#pragma unroll
  for (int c = 0; c < CPAD; ++c) {
    float acc = 0.f;
    if (c < C) {
      const float4* wr = reinterpret_cast<const float4*>(s_w + c * kHeadCin);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w4 = wr[j];
        acc = fmaf(w4.x, f[4 * j], acc);
        acc = fmaf(w4.y, f[4 * j + 1], acc);
        acc = fmaf(w4.z, f[4 * j + 2], acc);
        acc = fmaf(w4.w, f[4 * j + 3], acc);
      }
      acc += s_b[c];
    }
    l[c] = acc;
  }
}

template <int CPAD>
__global__ void __launch_bounds__(kLossThreads)
    head_ce_fwd_kernel(const float4* __restrict__ feat, const float* __restrict__ W,
                       const float* __restrict__ b, const int64_t* __restrict__ target, int64_t P,
                       int C, int64_t ignore_index, double* __restrict__ partial,
                       uint8_t* __restrict__ pred, unsigned long long* __restrict__ conf) {
  __shared__ __align__(16) float4 s_tile[kLossThreads * 8];
  __shared__ __align__(16) float s_w[CPAD * kHeadCin];
  __shared__ float s_b[CPAD];
  extern __shared__ unsigned int s_conf[];  // [C*C] when conf != nullptr
  for (int i = threadIdx.x; i < C * kHeadCin; i += blockDim.x) s_w[i] = W[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_b[i] = b[i];
  if (conf)
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_conf[i] = 0u;

  double loss_acc = 0.0, n_acc = 0.0;
  const int64_t ntiles = (P + kLossThreads - 1) / kLossThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kLossThreads;
    const int npx = (int)((P - p0) < kLossThreads ? (P - p0) : kLossThreads);
    __syncthreads();  // previous tile fully consumed (and s_w/s_conf initialised)
    stage_tile32(s_tile, feat, p0, npx);
    __syncthreads();
    if ((int)threadIdx.x < npx) {
      float f[kHeadCin], l[CPAD];
      head_logits<CPAD>(s_tile, s_w, s_b, threadIdx.x, C, f, l);
      PixelCE<CPAD> ce;
      ce.run(l, C);
      const int64_t t = target[p0 + threadIdx.x];
      if (pred) pred[p0 + threadIdx.x] = (uint8_t)ce.arg;
      if (target_valid(t, C, ignore_index)) {
        loss_acc += (double)(ce.lse - pick<CPAD>(l, (int)t));
        n_acc += 1.0;
        if (conf) atomicAdd(&s_conf[(int)t * C + ce.arg], 1u);
      }
    }
  }
  __syncthreads();
  if (conf) flush_conf(s_conf, C * C, conf);
  block_partial2(loss_acc, n_acc, partial);
}

// Backward: recompute logits, dl = (softmax - onehot) * g / n_valid, dfeat = dl @ W,
// dW = dl^T @ feat (register-tiled over the tile held in shared memory), db = sum dl.
// Warp-level layout for dW: lane = cg*8 + kg ; kg = float4 group of the 32 input channels,
// cg in [0,4) owns classes cg, cg+4, ... ; each of the 8 warps walks 32 of the 256 pixels.
template <int CPAD>
__global__ void __launch_bounds__(kLossThreads)
    head_ce_bwd_kernel(const float4* __restrict__ feat, const float* __restrict__ W,
                       const float* __restrict__ b, const int64_t* __restrict__ target, int64_t P,
                       int C, int64_t ignore_index, const double* __restrict__ fwd_out,
                       const float* __restrict__ gscale, float4* __restrict__ dfeat,
                       float* __restrict__ partial /* [grid][CPAD*33] */) {
  constexpr int CQ = CPAD / 4;       // classes per lane in the dW pass
  constexpr int DLS = CPAD + 1;      // padded row stride of the dl tile
  __shared__ __align__(16) float4 s_tile[kLossThreads * 8];
  __shared__ __align__(16) float s_w[CPAD * kHeadCin];
  __shared__ float s_b[CPAD];
  extern __shared__ float s_dyn[];  // dl tile [256][DLS]
  float* s_dl = s_dyn;
  for (int i = threadIdx.x; i < CPAD * kHeadCin; i += blockDim.x) s_w[i] = i < C * kHeadCin ? W[i] : 0.f;
  for (int i = threadIdx.x; i < CPAD; i += blockDim.x) s_b[i] = i < C ? b[i] : 0.f;

  const float scale = gscale[0] / (float)fwd_out[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kg = lane & 7, cg = lane >> 3;
  float4 dw_acc[CQ];
#pragma unroll
  for (int i = 0; i < CQ; ++i) dw_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float db_acc[CQ];  // lanes with kg == 0 own the bias gradient of their classes
#pragma unroll
  for (int i = 0; i < CQ; ++i) db_acc[i] = 0.f;

  const int64_t ntiles = (P + kLossThreads - 1) / kLossThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kLossThreads;
    const int npx = (int)((P - p0) < kLossThreads ? (P - p0) : kLossThreads);
    __syncthreads();
    stage_tile32(s_tile, feat, p0, npx);
    __syncthreads();
    float l[CPAD];
    float f[kHeadCin];
    const int row = threadIdx.x;
    bool valid = false;
    if (row < npx) {
      head_logits<CPAD>(s_tile, s_w, s_b, row, C, f, l);
      PixelCE<CPAD> ce;
      ce.run(l, C);
      const int64_t t = target[p0 + row];
      valid = target_valid(t, C, ignore_index);
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float d = 0.f;
        if (valid) d = (ce.prob(l[c]) - (c == (int)t ? 1.f : 0.f)) * scale;
        l[c] = d;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CPAD; ++c) l[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < CPAD; ++c) s_dl[row * DLS + c] = l[c];
    // dfeat = dl @ W  (per pixel, registers), staged back through the swizzled tile
    float df[kHeadCin];
#pragma unroll
    for (int k = 0; k < kHeadCin; ++k) df[k] = 0.f;
    if (valid) {
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        if (c < C) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + c * kHeadCin);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w4 = wr[j];
            df[4 * j] = fmaf(l[c], w4.x, df[4 * j]);
            df[4 * j + 1] = fmaf(l[c], w4.y, df[4 * j + 1]);
            df[4 * j + 2] = fmaf(l[c], w4.z, df[4 * j + 2]);
            df[4 * j + 3] = fmaf(l[c], w4.w, df[4 * j + 3]);
          }
        }
      }
    }
    __syncthreads();  // s_dl complete; everyone has read its feature row into registers
    // dW pass over this tile (reads s_tile = features, s_dl)
    for (int pr = warp; pr < npx; pr += kLossThreads / 32) {
      const float4 fv = s_tile[swz(pr, kg)];
#pragma unroll
      for (int i = 0; i < CQ; ++i) {
        const float d = s_dl[pr * DLS + cg + 4 * i];
        dw_acc[i].x = fmaf(d, fv.x, dw_acc[i].x);
        dw_acc[i].y = fmaf(d, fv.y, dw_acc[i].y);
        dw_acc[i].z = fmaf(d, fv.z, dw_acc[i].z);
        dw_acc[i].w = fmaf(d, fv.w, dw_acc[i].w);
        if (kg == 0) db_acc[i] += d;
      }
    }
    __syncthreads();  // features consumed -> reuse the tile for the dfeat store
    if (dfeat) {
      if (row < npx) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          s_tile[swz(row, j)] = make_float4(df[4 * j], df[4 * j + 1], df[4 * j + 2], df[4 * j + 3]);
      }
      __syncthreads();
      const int n = npx * 8;
      float4* dst = dfeat + p0 * 8;
#pragma unroll 4
      for (int i = threadIdx.x; i < n; i += kLossThreads) stg_stream(dst + i, s_tile[swz(i >> 3, i & 7)]);
    }
  }
  // cross-warp reduction of dW / db : reuse s_dl as [8 warps][CPAD*33]
  __syncthreads();
  float* s_red = s_dyn;
  constexpr int ROW = CPAD * (kHeadCin + 1);
#pragma unroll
  for (int i = 0; i < CQ; ++i) {
    const int c = cg + 4 * i;
    float* dst = s_red + warp * ROW + c * (kHeadCin + 1);
    dst[4 * kg] = dw_acc[i].x;
    dst[4 * kg + 1] = dw_acc[i].y;
    dst[4 * kg + 2] = dw_acc[i].z;
    dst[4 * kg + 3] = dw_acc[i].w;
    if (kg == 0) dst[kHeadCin] = db_acc[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ROW; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += s_red[w * ROW + i];
    partial[(int64_t)blockIdx.x * ROW + i] = s;
  }
}

// dW[c][k] / db[c] = fixed-order fp64 sum of the block partials ([grid][CPAD][33])
__global__ void __launch_bounds__(kFinThreads)
    head_ce_bwd_finalize(const float* __restrict__ partial, int nblocks, int CPAD, int C,
                                     float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ROW = CPAD * (kHeadCin + 1);
  const bool ok = i < C * (kHeadCin + 1);
  const double s = block_colsum(partial, nblocks, (int64_t)ROW, ok ? i : 0, ok);
  if (threadIdx.x >= 32 || !ok) return;
  const int c = i / (kHeadCin + 1), k = i % (kHeadCin + 1);
  if (k < kHeadCin)
    dW[c * kHeadCin + k] = (float)s;
  else
    db[c] = (float)s;
}

// ---------------------------------------------------------------------------------------
// CE on precomputed logits.  NCHW: class planes are read with unit stride across lanes.
// NHWC: the [256 x C] tile is contiguous; it is staged through shared memory (odd row
// stride) so global accesses stay 128-bit and coalesced.
// ---------------------------------------------------------------------------------------
// element offset of (pixel p, class 0) in a [B, C, HW] tensor; 32-bit division whenever the pixel index fits
__device__ __forceinline__ int64_t nchw_offset(int64_t p, int64_t HW, int C) {
  if (p <= 0xffffffffll && HW <= 0xffffffffll) {
    const uint32_t bi = (uint32_t)p / (uint32_t)HW;
    const uint32_t hw = (uint32_t)p - bi * (uint32_t)HW;
    return (int64_t)bi * C * HW + hw;
  }
  const int64_t bi = p / HW;
  return bi * C * HW + (p - bi * HW);
}

template <int CPAD, bool NHWC>
__device__ __forceinline__ void load_logits_tile(const float* __restrict__ logits, int64_t p0, int npx,
                                                 int64_t HW, int C, float* s_l, int CS,
                                                 float (&l)[CPAD]) {
  if (NHWC) {
    const int n = npx * C;
    const float* src = logits + p0 * C;  // 16B aligned: p0 is a multiple of 256
    const int n4 = n >> 2;
    const uint32_t cmagic = 0xffffffffu / (uint32_t)C + 1u;  // idx / C == umulhi(idx, cmagic) for idx < 2^27
    for (int q = threadIdx.x; q < n4; q += kLossThreads) {
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(src) + q);
      int idx = q * 4;
      int r = (int)__umulhi((uint32_t)idx, cmagic), c = idx - r * C;
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s_l[r * CS + c] = vv[e];
        if (++c == C) {
          c = 0;
          ++r;
        }
      }
    }
    for (int idx = n4 * 4 + threadIdx.x; idx < n; idx += kLossThreads) {
      const int r = idx / C, c = idx - r * C;
      s_l[r * CS + c] = src[idx];
    }
    __syncthreads();
    if ((int)threadIdx.x < npx) {
#pragma unroll
      for (int c = 0; c < CPAD; ++c) l[c] = VMTL_HAS_CLASS(c, C) ? s_l[threadIdx.x * CS + c] : 0.f;
    }
  } else {
    if ((int)threadIdx.x < npx) {
      const float* ptr = logits + nchw_offset(p0 + threadIdx.x, HW, C);
#pragma unroll
      for (int c = 0; c < CPAD; ++c, ptr += HW) l[c] = VMTL_HAS_CLASS(c, C) ? __ldg(ptr) : 0.f;
    }
  }
}

template <int CPAD, bool NHWC>
__global__ void __launch_bounds__(kLossThreads)
    ce_logits_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t P,
                         int64_t HW, int C, int64_t ignore_index, double* __restrict__ partial,
                         uint8_t* __restrict__ pred, unsigned long long* __restrict__ conf) {
  extern __shared__ unsigned int s_dynu[];  // [C*C conf][NHWC: 256*CS floats]
  unsigned int* s_conf = s_dynu;
  const int CS = C | 1;
  float* s_l = reinterpret_cast<float*>(s_dynu + (conf ? C * C : 0));
  if (conf)
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_conf[i] = 0u;
  double loss_acc = 0.0, n_acc = 0.0;
  const int64_t ntiles = (P + kLossThreads - 1) / kLossThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kLossThreads;
    const int npx = (int)((P - p0) < kLossThreads ? (P - p0) : kLossThreads);
    __syncthreads();
    float l[CPAD];
    load_logits_tile<CPAD, NHWC>(logits, p0, npx, HW, C, s_l, CS, l);
    if ((int)threadIdx.x < npx) {
      const int64_t t = __ldg(target + p0 + threadIdx.x);  // issued before the softmax math
      PixelCE<CPAD> ce;
      ce.run(l, C);
      if (pred) pred[p0 + threadIdx.x] = (uint8_t)ce.arg;
      if (target_valid(t, C, ignore_index)) {
        loss_acc += (double)(ce.lse - pick<CPAD>(l, (int)t));
        n_acc += 1.0;
        if (conf) atomicAdd(&s_conf[(int)t * C + ce.arg], 1u);
      }
    }
  }
  __syncthreads();
  if (conf) flush_conf(s_conf, C * C, conf);
  block_partial2(loss_acc, n_acc, partial);
}

template <int CPAD, bool NHWC>
__global__ void __launch_bounds__(kLossThreads)
    ce_logits_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t P,
                         int64_t HW, int C, int64_t ignore_index, const double* __restrict__ fwd_out,
                         const float* __restrict__ gscale, float* __restrict__ dlogits) {
  extern __shared__ unsigned int s_dynu[];
  const int CS = C | 1;
  float* s_l = reinterpret_cast<float*>(s_dynu);
  const float scale = gscale[0] / (float)fwd_out[1];
  const int64_t ntiles = (P + kLossThreads - 1) / kLossThreads;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kLossThreads;
    const int npx = (int)((P - p0) < kLossThreads ? (P - p0) : kLossThreads);
    __syncthreads();
    float l[CPAD];
    load_logits_tile<CPAD, NHWC>(logits, p0, npx, HW, C, s_l, CS, l);
    if ((int)threadIdx.x < npx) {
      const int64_t t = __ldg(target + p0 + threadIdx.x);
      PixelCE<CPAD> ce;
      ce.run(l, C);
      const bool valid = target_valid(t, C, ignore_index);
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float d = 0.f;
        if (valid) d = (ce.prob(l[c]) - (c == (int)t ? 1.f : 0.f)) * scale;
        l[c] = d;
      }
      if (NHWC) {
#pragma unroll
        for (int c = 0; c < CPAD; ++c)
          if (VMTL_HAS_CLASS(c, C)) s_l[threadIdx.x * CS + c] = l[c];
      } else {
        float* ptr = dlogits + nchw_offset(p0 + threadIdx.x, HW, C);
#pragma unroll
        for (int c = 0; c < CPAD; ++c, ptr += HW)
          if (VMTL_HAS_CLASS(c, C)) *ptr = l[c];
      }
    }
    if (NHWC) {
      __syncthreads();
      const int n = npx * C;
      float* dst = dlogits + p0 * C;
      const int n4 = n >> 2;
      const uint32_t cmagic = 0xffffffffu / (uint32_t)C + 1u;
      for (int q = threadIdx.x; q < n4; q += kLossThreads) {
        int idx = q * 4;
        int r = (int)__umulhi((uint32_t)idx, cmagic), c = idx - r * C;
        float vv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          vv[e] = s_l[r * CS + c];
          if (++c == C) {
            c = 0;
            ++r;
          }
        }
        stg_stream(reinterpret_cast<float4*>(dst) + q, make_float4(vv[0], vv[1], vv[2], vv[3]));
      }
      for (int idx = n4 * 4 + threadIdx.x; idx < n; idx += kLossThreads) {
        const int r = idx / C, c = idx - r * C;
        dst[idx] = s_l[r * CS + c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// NHWC logits through bulk async copies.  A 256-pixel tile of [P, C] logits is ONE contiguous run of
// 256*C floats: a single cp.async.bulk brings it into shared memory as is (double-buffered, issued one
// tile ahead by thread 0), each thread then reads its own pixel's C floats (stride C words).  No
// per-element staging stores, no index arithmetic; the backward sends dlogits back the same way
// (cp.async.bulk shared -> global).  The ragged last tile (npx < 256) is read / written directly.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}

template <int CPAD, bool BWD>
__global__ void __launch_bounds__(kLossThreads)
    ce_logits_nhwc_bulk_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t P, int C,
                               int64_t ignore_index, double* __restrict__ partial, uint8_t* __restrict__ pred,
                               unsigned long long* __restrict__ conf, const double* __restrict__ fwd_out,
                               const float* __restrict__ gscale, float* __restrict__ dlogits) {
  using namespace tc;
  extern __shared__ __align__(128) uint8_t s_raw[];  // [2 stages][256*C floats] | conf hist (fwd) | 2 mbarriers
  const uint32_t tile_bytes = (uint32_t)kLossThreads * (uint32_t)C * 4u;
  float* s_tile = reinterpret_cast<float*>(s_raw);
  unsigned int* s_conf = reinterpret_cast<unsigned int*>(s_raw + 2 * tile_bytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_raw + 2 * tile_bytes + ((!BWD && conf) ? (size_t)C * C * 4 : 0));
  s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_bar) + 7) & ~(uintptr_t)7);
  const uint32_t bar0 = smem_u32(s_bar);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  if (!BWD && conf)
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_conf[i] = 0u;
  __syncthreads();

  const float scale = BWD ? gscale[0] / (float)fwd_out[1] : 0.f;
  const int64_t nfull = P / kLossThreads;  // full tiles only; the ragged tail is handled below
  const int64_t nt = blockIdx.x < nfull ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_of = [&](int64_t i) { return blockIdx.x + i * gridDim.x; };
  auto issue = [&](int64_t i) {  // thread 0: tile i -> stage i & 1
    const uint32_t b = bar0 + 8u * (uint32_t)(i & 1);
    mbar_expect_tx(b, tile_bytes);
    bulk_load_1d(smem_u32(s_raw) + (uint32_t)(i & 1) * tile_bytes, logits + tile_of(i) * kLossThreads * C, tile_bytes, b);
  };
  double loss_acc = 0.0, n_acc = 0.0;

  auto pixel = [&](float (&l)[CPAD], int64_t p) {  // forward bookkeeping or dl, for one pixel held in l[]
    const int64_t t = __ldg(target + p);
    PixelCE<CPAD> ce;
    ce.run(l, C);
    const bool valid = target_valid(t, C, ignore_index);
    if (!BWD) {
      if (pred) pred[p] = (uint8_t)ce.arg;
      if (valid) {
        loss_acc += (double)(ce.lse - pick<CPAD>(l, (int)t));
        n_acc += 1.0;
        if (conf) atomicAdd(&s_conf[(int)t * C + ce.arg], 1u);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CPAD; ++c) l[c] = valid ? (ce.prob(l[c]) - (c == (int)t ? 1.f : 0.f)) * scale : 0.f;
    }
  };

  if (threadIdx.x == 0 && nt > 0) issue(0);
  for (int64_t i = 0; i < nt; ++i) {
    const int s = (int)(i & 1);
    if (threadIdx.x == 0 && i + 1 < nt) {
      if (BWD) tma_store_wait_read();  // the store of tile i-1 has finished reading stage s^1
      issue(i + 1);
    }
    mbar_wait(bar0 + 8u * (uint32_t)s, (uint32_t)((i >> 1) & 1));
    float* row = s_tile + (size_t)s * (tile_bytes / 4) + (size_t)threadIdx.x * C;
    float l[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) l[c] = VMTL_HAS_CLASS(c, C) ? row[c] : 0.f;
    const int64_t p = tile_of(i) * kLossThreads + threadIdx.x;
    pixel(l, p);
    if (BWD) {
#pragma unroll
      for (int c = 0; c < CPAD; ++c)
        if (VMTL_HAS_CLASS(c, C)) row[c] = l[c];
      fence_proxy_async_smem();
    }
    __syncthreads();  // stage s fully consumed (fwd) / fully rewritten with dl (bwd)
    if (BWD && threadIdx.x == 0) {
      bulk_store_1d(dlogits + tile_of(i) * kLossThreads * C, smem_u32(s_raw) + (uint32_t)s * tile_bytes, tile_bytes);
      tma_store_commit();
    }
  }
  if (BWD && threadIdx.x == 0) tma_store_wait_all();
  // ragged tail: the last P % 256 pixels, one CTA, straight from / to global memory
  const int64_t p_tail = nfull * kLossThreads + threadIdx.x;
  if (blockIdx.x == (unsigned)(nfull % gridDim.x) && p_tail < P) {
    float l[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) l[c] = VMTL_HAS_CLASS(c, C) ? __ldg(logits + p_tail * C + c) : 0.f;
    pixel(l, p_tail);
    if (BWD) {
#pragma unroll
      for (int c = 0; c < CPAD; ++c)
        if (VMTL_HAS_CLASS(c, C)) dlogits[p_tail * C + c] = l[c];
    }
  }
  if (!BWD) {
    __syncthreads();
    if (conf) flush_conf(s_conf, C * C, conf);
    block_partial2(loss_acc, n_acc, partial);
  }
}

// ---------------------------------------------------------------------------------------
// Depth head + sigmoid + SILog moments + error sums.
// LPP lanes cooperate on one pixel (LPP = Cin/4 float4 per pixel); LPP == 1 means the
// input already is the depth logit (basic / csnet).
// ---------------------------------------------------------------------------------------
struct SilogAcc {
  double n, sg, sgg, sabs, srel;
};

__device__ __forceinline__ float sigmoid_fast(float u) { return __fdividef(1.f, 1.f + fast_ex2(-u * kLog2e)); }
// log(p) - log(t) through MUFU lg2, p = sigmoid(z).  Below z = -80 the fast sigmoid flushes to 0 (its exponential
// overflows) while the reference's stays a finite denormal with log p = z to 1e-35: take z there.
__device__ __forceinline__ float log_ratio(float z, float p, float t) {
  const float lt = fast_lg2(t);
  return z < -80.f ? fmaf(-lt, kLn2, z) : (fast_lg2(p) - lt) * kLn2;
}

__device__ __forceinline__ void silog_pixel(float zd, float t, float min_depth, SilogAcc& a, float& p_out) {
  const float p = sigmoid_fast(zd);
  p_out = p;
  const float d = fabsf(p - t);
  a.sabs += (double)d;
  if (t > min_depth) {
    const float g = log_ratio(zd, p, t);
    a.n += 1.0;
    a.sg += (double)g;
    a.sgg += (double)g * (double)g;
    a.srel += (double)__fdividef(d, t);
  }
}

template <int LPP>
__global__ void __launch_bounds__(kLossThreads)
    head_silog_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                          const float* __restrict__ b, const float* __restrict__ target, int64_t P,
                          float min_depth, double* __restrict__ partial, float* __restrict__ pred) {
  SilogAcc a{0.0, 0.0, 0.0, 0.0, 0.0};
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  if (LPP == 1) {
    for (int64_t p = gtid; p < P; p += nthr) {
      float pr;
      silog_pixel(__ldg(feat + p), __ldg(target + p), min_depth, a, pr);
      if (pred) pred[p] = pr;
    }
  } else {
    const int sub = threadIdx.x % LPP;
    const float4 w4 = reinterpret_cast<const float4*>(w)[sub];
    const float bias = b[0];
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    const int64_t ngrp = nthr / LPP;  // pixels in flight per sweep (multiple of 32/LPP)
    const int lane = threadIdx.x & 31;
    // warp-uniform loop bound: every lane runs the same trip count so the shuffles are legal
    int64_t pbase = (gtid - lane) / LPP;
    const int pl = lane / LPP;
    constexpr int U = 4;
    for (; pbase < P; pbase += U * ngrp) {
      float4 v[U];
      float tg[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = pbase + pl + u * ngrp;
        ok[u] = p < P;
        v[u] = ok[u] ? ldg_stream(f4 + p * LPP + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        tg[u] = ok[u] ? __ldg(target + p) : 0.f;
      }
      // after the butterfly every lane of a pixel group holds all U dot products: lane `sub` (< U) takes
      // pixel u = sub, so the sigmoid/log/fp64 part runs once per warp instead of once per unrolled pixel
      float dsel = 0.f, tsel = 0.f;
      bool oksel = false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float d = v[u].x * w4.x + v[u].y * w4.y + v[u].z * w4.z + v[u].w * w4.w;
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (sub == u) {
          dsel = d;
          tsel = tg[u];
          oksel = ok[u];
        }
      }
      if (sub < U && oksel) {
        float pr;
        silog_pixel(dsel + bias, tsel, min_depth, a, pr);
        if (pred) pred[pbase + pl + sub * ngrp] = pr;
      }
    }
  }
  __shared__ double s_p[5][kLossThreads / 32];
  double v5[5] = {a.n, a.sg, a.sgg, a.sabs, a.srel};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double r = warp_sum(v5[i]);
    if (lane == 0) s_p[i][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int wi = 0; wi < kLossThreads / 32; ++wi) s += s_p[threadIdx.x][wi];
    partial[(int64_t)blockIdx.x * 5 + threadIdx.x] = s;
  }
}

// out[0..4] = {n, sum g, sum g^2, sum|p-t|, sum|p-t|/t}, out[7] = P  ->  out[5] = mean g, out[6] = D, scalars
__device__ __forceinline__ void silog_from_moments(double* __restrict__ out, float* __restrict__ scalars) {
  const double n = out[0], sg = out[1], sgg = out[2];
  const double mean = sg / n;
  // unbiased variance, as torch.var (losses.py:35); n <= 1 gives NaN like the reference
  const double var = (sgg - sg * mean) / (n - 1.0);
  const double D = var + 0.15 * mean * mean;
  out[5] = mean;
  out[6] = D;
  if (scalars) {
    scalars[0] = (float)(10.0 * sqrt(D));
    scalars[1] = (float)(out[3] / out[7]);
    scalars[2] = (float)(out[4] / n);
  }
}

__global__ void __launch_bounds__(kFinThreads)
    silog_finalize(const double* __restrict__ partial, int nblocks, int64_t P,
                               double* __restrict__ out, float* __restrict__ scalars) {
  __shared__ double s[5];
  const int col = threadIdx.x & 31;
  const double acc = block_colsum(partial, nblocks, 5, col < 5 ? col : 0, col < 5);
  if (threadIdx.x < 5) s[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = s[0];
    out[1] = s[1];
    out[2] = s[2];
    out[3] = s[3];
    out[4] = s[4];
    out[7] = (double)P;
    silog_from_moments(out, scalars);
  }
}

// global-batch statistics: out[0..4] and out[7] hold all-reduced sums; re-derive mean, D and the scalars
__global__ void silog_refinalize(double* __restrict__ out, float* __restrict__ scalars) {
  if (threadIdx.x == 0) silog_from_moments(out, scalars);
}

// d silog / d zd_i = (5/sqrt(D)) * (2 (g_i - mean)/(n-1) + 0.3 mean / n) * (1 - p_i)   
template <int LPP>
__global__ void __launch_bounds__(kLossThreads)
    head_silog_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                          const float* __restrict__ b, const float* __restrict__ target, int64_t P,
                          float min_depth, const double* __restrict__ fwd_out,
                          const float* __restrict__ gscale, float* __restrict__ dfeat,
                          float* __restrict__ partial /* [grid][4*LPP + 1] */) {
  const double n = fwd_out[0], mean = fwd_out[5], D = fwd_out[6];
  const double k0 = (double)gscale[0] * 5.0 / sqrt(D);
  const float cA = (float)(k0 * 2.0 / (n - 1.0));
  const float cB = (float)(k0 * 0.3 * mean / n);
  const float fmean = (float)mean;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  if (LPP == 1) {
    for (int64_t p = gtid; p < P; p += nthr) {
      const float t = __ldg(target + p);
      float dz = 0.f;
      if (t > min_depth) {
        const float zd = __ldg(feat + p);
        const float pr = sigmoid_fast(zd);
        const float g = log_ratio(zd, pr, t);
        dz = (cA * (g - fmean) + cB) * (1.f - pr);
      }
      dfeat[p] = dz;
    }
    return;
  } else {
    const int sub = threadIdx.x % LPP;
    const float4 w4 = reinterpret_cast<const float4*>(w)[sub];
    const float bias = b[0];
    const float4* f4 = reinterpret_cast<const float4*>(feat);
    float4* df4 = reinterpret_cast<float4*>(dfeat);
    const int64_t ngrp = nthr / LPP;
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    float dbacc = 0.f;
    const int lane_ = threadIdx.x & 31;
    const int pl = lane_ / LPP;
    constexpr int U = 4;  // pixels in flight per thread
    for (int64_t pbase = (gtid - lane_) / LPP; pbase < P; pbase += U * ngrp) {  // warp-uniform bound
      float4 v[U];
      float tg[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = pbase + pl + u * ngrp;
        ok[u] = p < P;
        v[u] = ok[u] ? ldg_stream(f4 + p * LPP + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        tg[u] = ok[u] ? __ldg(target + p) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = pbase + pl + u * ngrp;
        float d = v[u].x * w4.x + v[u].y * w4.y + v[u].z * w4.z + v[u].w * w4.w;
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        float dz = 0.f;
        if (ok[u] && tg[u] > min_depth) {
          const float pr = sigmoid_fast(d + bias);
          const float g = log_ratio(d + bias, pr, tg[u]);
          dz = (cA * (g - fmean) + cB) * (1.f - pr);
        }
        if (dfeat && ok[u])
          stg_stream(df4 + p * LPP + sub, make_float4(dz * w4.x, dz * w4.y, dz * w4.z, dz * w4.w));
        dw.x = fmaf(dz, v[u].x, dw.x);
        dw.y = fmaf(dz, v[u].y, dw.y);
        dw.z = fmaf(dz, v[u].z, dw.z);
        dw.w = fmaf(dz, v[u].w, dw.w);
        if (sub == 0) dbacc += dz;
      }
    }
    // reduce lanes that share `sub` inside the warp, then across warps
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
      dw.x += __shfl_xor_sync(0xffffffffu, dw.x, o);
      dw.y += __shfl_xor_sync(0xffffffffu, dw.y, o);
      dw.z += __shfl_xor_sync(0xffffffffu, dw.z, o);
      dw.w += __shfl_xor_sync(0xffffffffu, dw.w, o);
      dbacc += __shfl_xor_sync(0xffffffffu, dbacc, o);
    }
    constexpr int ROW = 4 * LPP + 1;
    __shared__ float s_r[kLossThreads / 32][ROW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < LPP) {
      s_r[warp][4 * lane] = dw.x;
      s_r[warp][4 * lane + 1] = dw.y;
      s_r[warp][4 * lane + 2] = dw.z;
      s_r[warp][4 * lane + 3] = dw.w;
      if (lane == 0) s_r[warp][4 * LPP] = dbacc;
    }
    __syncthreads();
    if (threadIdx.x < ROW) {
      float s = 0.f;
      for (int wi = 0; wi < kLossThreads / 32; ++wi) s += s_r[wi][threadIdx.x];
      partial[(int64_t)blockIdx.x * ROW + threadIdx.x] = s;
    }
  }
}

__global__ void __launch_bounds__(kFinThreads)
    silog_bwd_finalize(const float* __restrict__ partial, int nblocks, int cin,
                                   float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const double s = block_colsum(partial, nblocks, (int64_t)(cin + 1), i <= cin ? i : 0, i <= cin);
  if (threadIdx.x >= 32 || i > cin) return;
  if (i < cin)
    dw[i] = (float)s;
  else
    db[0] = (float)s;
}

static int cpad_for(int C) { return C <= 16 ? 16 : (C <= 20 ? 20 : (C <= 32 ? 32 : 0)); }

}  // namespace vmtl

using namespace vmtl;

extern "C" size_t vmtl_loss_workspace_bytes(int64_t P) {
  (void)P;
  // largest per-block partial row: head_ce_bwd = 32*33 floats
  return (size_t)loss_max_blocks() * 32 * (kHeadCin + 1) * sizeof(float) + 256;
}

extern "C" int vmtl_head_ce_fwd(const float* feat, const float* W, const float* b, const int64_t* target,
                                int64_t P, int Cin, int C, int64_t ignore_index, double* out,
                                float* loss, uint8_t* pred, int64_t* conf, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!feat || !W || !b || !target || !out || !workspace || P < 0 || C < 1) return VMTL_EINVAL;
  if (Cin != kHeadCin) return VMTL_EUNSUPPORTED;
  const int cpad = cpad_for(C);
  if (!cpad) return VMTL_EUNSUPPORTED;
  if (!aligned16(feat)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  // tensor-core projection (head_tc.cu); inputs smaller than one tile keep the CUDA-core kernel
  if (P >= 128) {
    int g = 0;
    const int max_blocks = (int)(workspace_bytes / (2 * sizeof(double)));
    const int rc = head_ce_tc_fwd(feat, W, b, target, P, C, ignore_index, partial, max_blocks, &g, pred, conf, st);
    if (rc == VMTL_OK) {
      ce_finalize<<<1, kFinThreads, 0, st>>>(partial, g, out, loss);
      return launch_status();
    }
    if (rc != VMTL_EUNSUPPORTED) return rc;
  }
  const int grid = loss_grid(P, kLossThreads, 4);
  if (workspace_bytes < (size_t)grid * 2 * sizeof(double)) return VMTL_EWORKSPACE;
  const size_t smem = conf ? (size_t)C * C * sizeof(unsigned int) : 0;
  const float4* f4 = reinterpret_cast<const float4*>(feat);
  unsigned long long* cf = reinterpret_cast<unsigned long long*>(conf);
  switch (cpad) {
    case 16: head_ce_fwd_kernel<16><<<grid, kLossThreads, smem, st>>>(f4, W, b, target, P, C, ignore_index, partial, pred, cf); break;
    case 20: head_ce_fwd_kernel<20><<<grid, kLossThreads, smem, st>>>(f4, W, b, target, P, C, ignore_index, partial, pred, cf); break;
    default: head_ce_fwd_kernel<32><<<grid, kLossThreads, smem, st>>>(f4, W, b, target, P, C, ignore_index, partial, pred, cf); break;
  }
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  ce_finalize<<<1, kFinThreads, 0, st>>>(partial, grid, out, loss);
  return launch_status();
}

extern "C" int vmtl_head_ce_bwd(const float* feat, const float* W, const float* b, const int64_t* target,
                                int64_t P, int Cin, int C, int64_t ignore_index, const double* fwd_out,
                                const float* gscale, float* dfeat, float* dW, float* db, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!feat || !W || !b || !target || !fwd_out || !gscale || !dW || !db || !workspace || P < 0 || C < 1)
    return VMTL_EINVAL;
  if (Cin != kHeadCin) return VMTL_EUNSUPPORTED;
  const int cpad = cpad_for(C);
  if (!cpad) return VMTL_EUNSUPPORTED;
  if (!aligned16(feat) || (dfeat && !aligned16(dfeat))) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ROW = cpad * (kHeadCin + 1);
  float* partial = static_cast<float*>(workspace);
  if (P >= 128) {
    int g = 0;
    const int max_blocks = (int)(workspace_bytes / ((size_t)ROW * sizeof(float)));
    const int rc = head_ce_tc_bwd(feat, W, b, target, P, C, ignore_index, fwd_out, gscale, dfeat, partial,
                                  max_blocks, &g, st);
    if (rc == VMTL_OK) {
      const int n = C * (kHeadCin + 1);
      head_ce_bwd_finalize<<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, g, cpad, C, dW, db);
      return launch_status();
    }
    if (rc != VMTL_EUNSUPPORTED) return rc;
  }
  const int grid = loss_grid(P, kLossThreads, 2);
  if (workspace_bytes < (size_t)grid * ROW * sizeof(float)) return VMTL_EWORKSPACE;
  // dynamic smem: max(dl tile [256][cpad+1], reduction [8][ROW]) floats
  size_t a = (size_t)kLossThreads * (cpad + 1), r = (size_t)(kLossThreads / 32) * ROW;
  const size_t smem = (a > r ? a : r) * sizeof(float);
  const float4* f4 = reinterpret_cast<const float4*>(feat);
  float4* d4 = reinterpret_cast<float4*>(dfeat);
#define VMTL_HCB(CP)                                                                              \
  do {                                                                                            \
    cudaFuncSetAttribute(head_ce_bwd_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                         (int)smem);                                                              \
    head_ce_bwd_kernel<CP><<<grid, kLossThreads, smem, st>>>(f4, W, b, target, P, C, ignore_index, \
                                                             fwd_out, gscale, d4, partial);       \
  } while (0)
  switch (cpad) {
    case 16: VMTL_HCB(16); break;
    case 20: VMTL_HCB(20); break;
    default: VMTL_HCB(32); break;
  }
#undef VMTL_HCB
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  const int n = C * (kHeadCin + 1);
  head_ce_bwd_finalize<<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, grid, cpad, C, dW, db);
  return launch_status();
}

extern "C" int vmtl_ce_logits_fwd(const float* logits, const int64_t* target, int64_t P, int64_t HW, int C,
                                  int layout, int64_t ignore_index, double* out, float* loss,
                                  uint8_t* pred, int64_t* conf, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!logits || !target || !out || !workspace || P < 0 || C < 1 || HW < 1) return VMTL_EINVAL;
  if (layout != VMTL_LAYOUT_NCHW && layout != VMTL_LAYOUT_NHWC) return VMTL_EINVAL;
  const int cpad = cpad_for(C);
  if (!cpad) return VMTL_EUNSUPPORTED;
  if (layout == VMTL_LAYOUT_NHWC && !aligned16(logits)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = loss_grid(P, kLossThreads, 8);  // upper bound; the launch uses the resident count
  if (workspace_bytes < (size_t)grid * 2 * sizeof(double)) return VMTL_EWORKSPACE;
  double* partial = static_cast<double*>(workspace);
  size_t smem = conf ? (size_t)C * C * sizeof(unsigned int) : 0;
  if (layout == VMTL_LAYOUT_NHWC) smem += (size_t)kLossThreads * (C | 1) * sizeof(float);
  unsigned long long* cf = reinterpret_cast<unsigned long long*>(conf);
#define VMTL_CEF(CP, NH)                                                                          \
  do {                                                                                            \
    grid = loss_grid(P, kLossThreads, blocks_per_sm(ce_logits_fwd_kernel<CP, NH>, kLossThreads, smem, 8)); \
    ce_logits_fwd_kernel<CP, NH><<<grid, kLossThreads, smem, st>>>(logits, target, P, HW, C,      \
                                                                   ignore_index, partial, pred, cf); \
  } while (0)
  const bool nh = layout == VMTL_LAYOUT_NHWC;
  if (nh) {  // NHWC: tiles arrive by bulk async copy
    const size_t bsmem = 2 * (size_t)kLossThreads * C * 4 + (conf ? (size_t)C * C * 4 : 0) + 32;
#define VMTL_CEFB(CP)                                                                                           \
  do {                                                                                                          \
    auto kern = ce_logits_nhwc_bulk_kernel<CP, false>;                                                          \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);                        \
    grid = loss_grid(P, kLossThreads, blocks_per_sm(kern, kLossThreads, bsmem, 8));                             \
    kern<<<grid, kLossThreads, bsmem, st>>>(logits, target, P, C, ignore_index, partial, pred, cf, nullptr,     \
                                            nullptr, nullptr);                                                  \
  } while (0)
    if (cpad == 16) VMTL_CEFB(16);
    else if (cpad == 20) VMTL_CEFB(20);
    else VMTL_CEFB(32);
#undef VMTL_CEFB
  } else if (cpad == 16) VMTL_CEF(16, false);
  else if (cpad == 20) VMTL_CEF(20, false);
  else VMTL_CEF(32, false);
#undef VMTL_CEF
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  ce_finalize<<<1, kFinThreads, 0, st>>>(partial, grid, out, loss);
  return launch_status();
}

extern "C" int vmtl_ce_logits_bwd(const float* logits, const int64_t* target, int64_t P, int64_t HW, int C,
                                  int layout, int64_t ignore_index, const double* fwd_out,
                                  const float* gscale, float* dlogits, void* stream) {
  if (!logits || !target || !fwd_out || !gscale || !dlogits || P < 0 || C < 1 || HW < 1) return VMTL_EINVAL;
  if (layout != VMTL_LAYOUT_NCHW && layout != VMTL_LAYOUT_NHWC) return VMTL_EINVAL;
  const int cpad = cpad_for(C);
  if (!cpad) return VMTL_EUNSUPPORTED;
  if (layout == VMTL_LAYOUT_NHWC && (!aligned16(logits) || !aligned16(dlogits))) return VMTL_EALIGN;
  if (P == 0) return VMTL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = layout == VMTL_LAYOUT_NHWC ? (size_t)kLossThreads * (C | 1) * sizeof(float) : 0;
#define VMTL_CEB(CP, NH)                                                                          \
  do {                                                                                            \
    const int grid =                                                                              \
        loss_grid(P, kLossThreads, blocks_per_sm(ce_logits_bwd_kernel<CP, NH>, kLossThreads, smem, 8)); \
    ce_logits_bwd_kernel<CP, NH><<<grid, kLossThreads, smem, st>>>(logits, target, P, HW, C,      \
                                                                   ignore_index, fwd_out, gscale, dlogits); \
  } while (0)
  const bool nh = layout == VMTL_LAYOUT_NHWC;
  if (nh) {
    const size_t bsmem = 2 * (size_t)kLossThreads * C * 4 + 32;
#define VMTL_CEBB(CP)                                                                                           \
  do {                                                                                                          \
    auto kern = ce_logits_nhwc_bulk_kernel<CP, true>;                                                           \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);                        \
    const int grid = loss_grid(P, kLossThreads, blocks_per_sm(kern, kLossThreads, bsmem, 8));                   \
    kern<<<grid, kLossThreads, bsmem, st>>>(logits, target, P, C, ignore_index, nullptr, nullptr, nullptr,      \
                                            fwd_out, gscale, dlogits);                                          \
  } while (0)
    if (cpad == 16) VMTL_CEBB(16);
    else if (cpad == 20) VMTL_CEBB(20);
    else VMTL_CEBB(32);
#undef VMTL_CEBB
    return launch_status();
  }
  if (cpad == 16) VMTL_CEB(16, false);
  else if (cpad == 20) VMTL_CEB(20, false);
  else VMTL_CEB(32, false);
#undef VMTL_CEB
  return launch_status();
}

static int silog_lpp(int Cin) {
  switch (Cin) {
    case 1: return 1;
    case 16: return 4;
    case 32: return 8;
    case 64: return 16;
    default: return 0;
  }
}

// kind 0: vmtl_head_ce_* (Cin, C); kind 1: vmtl_head_silog_* (Cin); kind 2: vmtl_ce_logits_* (C).  1 = covered.
extern "C" int vmtl_head_supported(int kind, int Cin, int C) {
  switch (kind) {
    case 0: return Cin == kHeadCin && C >= 1 && cpad_for(C) != 0;
    case 1: return silog_lpp(Cin) != 0;
    case 2: return C >= 1 && cpad_for(C) != 0;
    default: return 0;
  }
}

extern "C" int vmtl_head_silog_fwd(const float* feat, const float* w, const float* b, const float* target,
                                   int64_t P, int Cin, float min_depth, double* out, float* scalars,
                                   float* pred, void* workspace, size_t workspace_bytes, void* stream) {
  if (!feat || !target || !out || !workspace || P < 0) return VMTL_EINVAL;
  const int lpp = silog_lpp(Cin);
  if (!lpp) return VMTL_EUNSUPPORTED;
  if (lpp > 1 && (!w || !b)) return VMTL_EINVAL;
  if (lpp > 1 && (!aligned16(feat) || !aligned16(w))) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = loss_grid(P, kLossThreads / lpp * 4, 4);
  if (workspace_bytes < (size_t)grid * 5 * sizeof(double)) return VMTL_EWORKSPACE;
  double* partial = static_cast<double*>(workspace);
  switch (lpp) {
    case 1: head_silog_fwd_kernel<1><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, partial, pred); break;
    case 4: head_silog_fwd_kernel<4><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, partial, pred); break;
    case 8: head_silog_fwd_kernel<8><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, partial, pred); break;
    default: head_silog_fwd_kernel<16><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, partial, pred); break;
  }
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  silog_finalize<<<1, kFinThreads, 0, st>>>(partial, grid, P, out, scalars);
  return launch_status();
}

extern "C" int vmtl_head_silog_bwd(const float* feat, const float* w, const float* b, const float* target,
                                   int64_t P, int Cin, float min_depth, const double* fwd_out,
                                   const float* gscale, float* dfeat, float* dw, float* db,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (!feat || !target || !fwd_out || !gscale || P < 0) return VMTL_EINVAL;
  const int lpp = silog_lpp(Cin);
  if (!lpp) return VMTL_EUNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (lpp == 1) {
    if (!dfeat) return VMTL_EINVAL;
    const int grid = loss_grid(P, kLossThreads * 4, 4);
    head_silog_bwd_kernel<1><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, fwd_out, gscale, dfeat, nullptr);
    return launch_status();
  }
  if (!w || !b || !dw || !db || !workspace) return VMTL_EINVAL;
  if (!aligned16(feat) || !aligned16(w) || (dfeat && !aligned16(dfeat))) return VMTL_EALIGN;
  const int grid = loss_grid(P, kLossThreads / lpp * 4, 4);
  if (workspace_bytes < (size_t)grid * (Cin + 1) * sizeof(float)) return VMTL_EWORKSPACE;
  float* partial = static_cast<float*>(workspace);
  switch (lpp) {
    case 4: head_silog_bwd_kernel<4><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, fwd_out, gscale, dfeat, partial); break;
    case 8: head_silog_bwd_kernel<8><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, fwd_out, gscale, dfeat, partial); break;
    default: head_silog_bwd_kernel<16><<<grid, kLossThreads, 0, st>>>(feat, w, b, target, P, min_depth, fwd_out, gscale, dfeat, partial); break;
  }
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  silog_bwd_finalize<<<(Cin + 1 + 31) / 32, kFinThreads, 0, st>>>(partial, grid, Cin, dw, db);
  return launch_status();
}

extern "C" int vmtl_silog_finalize(double* out, float* scalars, void* stream) {
  if (!out) return VMTL_EINVAL;
  silog_refinalize<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(out, scalars);
  return launch_status();
}

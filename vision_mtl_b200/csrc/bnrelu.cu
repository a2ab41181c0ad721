// BatchNorm2d (+ ReLU, + 2x2 max-pool) in NHWC as streaming kernels -- SURVEY 8(f) rows 1 and 4.
//
// Reference call sites: every `conv -> BatchNorm2d -> ReLU` chain of the MTAN network
//   DoubleConv                                vision_mtl/utils/model_utils.py:61-80
//   attention conv1 -> bn1 -> relu            vision_mtl/models/mtan_model.py:65-69, :152-156
//   attention conv3 -> bn3 -> relu -> maxpool vision_mtl/models/mtan_model.py:77-81
//   decoder  conv3 / conv_out -> bn -> relu   vision_mtl/models/mtan_model.py:141-142, :165-167
// which the reference runs as separate ATen kernels (batch_norm: statistics + apply, relu, max_pool2d: 3W + 4R
// of the activation forward, 3W + 6R backward).  Here:
//   fwd   statistics pass  (R x)            -> per-block column partials (sum, sum of squares)
//         finalize                           -> mean, invstd, running statistics, folded (A, B) = (gamma*invstd, beta - mean*A)
//         apply pass       (R x, W y)        y = max(A x + B, 0)   [optionally 2x2 max-pooled: W y/4]
//   bwd   statistics pass  (R dy, R x)      g = dy * [A x + B > 0];  sum g, sum g*xhat
//         finalize                           -> dbeta, dgamma, c1 = dbeta/M, c2 = dgamma/M
//         apply pass       (R dy, R x, W dx) dx = A (g - c1 - xhat c2)
// With y == NULL the forward stops after the finalize: the MTAN gate kernels then consume x directly and apply
// max(A x + B, 0) while they convert their operand (the hidden tensor h is never materialised).
// BatchNorm semantics are nn.BatchNorm2d's: biased variance normalises, unbiased variance feeds running_var.
// Every reduction is two-stage with a fixed-order fp64 second stage (deterministic).
#include <math.h>

#include "vmtl_common.cuh"

namespace vmtl {

constexpr int kBnThreads = 256;

struct BnMap {
  int rows, r, g;
  bool active;
};
__device__ __forceinline__ BnMap bn_map(int C4) {
  BnMap m;
  m.rows = kBnThreads / C4;
  m.r = threadIdx.x / C4;
  m.g = threadIdx.x - m.r * C4;
  m.active = m.r < m.rows;
  return m;
}
static int bn_grid(int64_t rows_total, int C, int per_sm) {
  const int rows = kBnThreads / (C / 4);
  int64_t want = (rows_total + rows - 1) / rows;
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

__device__ __forceinline__ float4 bn_ld(const float* p, int64_t i4) {
  return ldg_stream(reinterpret_cast<const float4*>(p) + i4);
}
__device__ __forceinline__ void bn_st(float* p, int64_t i4, const float4& v) {
  stg_stream(reinterpret_cast<float4*>(p) + i4, v);
}
__device__ __forceinline__ float4 bn_c4(const float* p, int g) { return reinterpret_cast<const float4*>(p)[g]; }

// block reduction over the `rows` threads sharing a channel group; writes one partial row [NV][4*C4]
template <int NV>
__device__ __forceinline__ void bn_block_reduce(const float4 (&acc)[NV], const BnMap& m, int C4, float* partial_row) {
  __shared__ float4 s_red[kBnThreads * NV];
  if (m.active) {
#pragma unroll
    for (int i = 0; i < NV; ++i) s_red[(m.r * C4 + m.g) * NV + i] = acc[i];
  }
  __syncthreads();
  if (m.r == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 s = s_red[m.g * NV + i];
      for (int rr = 1; rr < m.rows; ++rr) {
        const float4 v = s_red[(rr * C4 + m.g) * NV + i];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      reinterpret_cast<float4*>(partial_row)[i * C4 + m.g] = s;
    }
  }
}

// ---- forward --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnThreads)
    bn_stats_kernel(const float* __restrict__ x, int64_t M, int C4, float* __restrict__ partial) {
  pdl_trigger();  // the finalize kernel may be scheduled while this one drains
  const BnMap m = bn_map(C4);
  float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (m.active) {
    const int64_t step = (int64_t)gridDim.x * m.rows;
    int64_t p = (int64_t)blockIdx.x * m.rows + m.r;
    auto add = [&](const float4& v) {
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      acc[1].x = fmaf(v.x, v.x, acc[1].x); acc[1].y = fmaf(v.y, v.y, acc[1].y);
      acc[1].z = fmaf(v.z, v.z, acc[1].z); acc[1].w = fmaf(v.w, v.w, acc[1].w);
    };
    for (; p + 3 * step < M; p += 4 * step) {  // four rows in flight (a read-only pass has nothing else to overlap)
      const float4 v0 = bn_ld(x, p * C4 + m.g), v1 = bn_ld(x, (p + step) * C4 + m.g),
                   v2 = bn_ld(x, (p + 2 * step) * C4 + m.g), v3 = bn_ld(x, (p + 3 * step) * C4 + m.g);
      add(v0); add(v1); add(v2); add(v3);
    }
    for (; p < M; p += step) add(bn_ld(x, p * C4 + m.g));
  }
  bn_block_reduce<2>(acc, m, C4, partial + (int64_t)blockIdx.x * 8 * C4);
}

// partial [nparts][2][C] -> mean, invstd, folded coefficients, running statistics
__global__ void __launch_bounds__(kFinThreads)
    bn_fwd_finalize(const float* __restrict__ partial, int nparts, int64_t M, int C, float eps,
                                float momentum, int training, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ running_mean,
                                float* __restrict__ running_var, float* __restrict__ save_mean,
                                float* __restrict__ save_invstd, float* __restrict__ coef /* [2][C] */,
                                const double* __restrict__ gmoments /* NULL, or global [2][C] over M rows */,
                                const float* __restrict__ conv_bias /* NULL, or the un-added bias of the producing conv */) {
  pdl_wait();     // statistics pass complete, its partials visible
  pdl_trigger();  // the apply pass may be scheduled while this block reduces
  const int c = blockIdx.x * kTallCols + (threadIdx.x & (kTallCols - 1));
  const bool owner = threadIdx.x < kTallCols && c < C;
  // the per-channel operands are fetched BEFORE the reduction, so their latency hides behind it
  float g_c = 0.f, b_c = 0.f, rm_c = 0.f, rv_c = 0.f, cb_c = 0.f;
  if (owner) {
    g_c = gamma[c];
    b_c = beta[c];
    if (running_mean) rm_c = running_mean[c];
    if (running_var) rv_c = running_var[c];
    if (conv_bias) cb_c = conv_bias[c];
  }
  double s = 0.0, q = 0.0;
  if (training && !gmoments)  // block-uniform branch: the reduction synchronises
    block_colsum2_tall(partial, nparts, 2 * (int64_t)C, c < C ? c : 0, C + (c < C ? c : 0), c < C, &s, &q);
  if (!owner) return;
  if (training && gmoments) {
    s = gmoments[c];
    q = gmoments[C + c];
  }
  double mean, var;
  if (training) {
    mean = s / (double)M;
    var = q / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    // the batch mean of the tensor the reference normalises (x + conv_bias); y itself does not depend on the shift
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * rm_c + momentum * (mean + (double)cb_c));
    if (running_var) {
      const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_var[c] = (float)((1.0 - momentum) * rv_c + momentum * unb);
    }
  } else {
    mean = rm_c;
    var = rv_c;
  }
  const double inv = 1.0 / sqrt(var + (double)eps);
  const float invf = (float)inv, meanf = (float)mean;
  const float a = g_c * invf;  // fp32, so that every consumer folds exactly the same coefficients
  save_mean[c] = meanf;
  save_invstd[c] = invf;
  coef[c] = a;
  coef[C + c] = b_c - meanf * a;
}

template <bool RELU>
__device__ __forceinline__ float4 bn_act(const float4& x, const float4& A, const float4& B) {
  float4 y = make_float4(fmaf(A.x, x.x, B.x), fmaf(A.y, x.y, B.y), fmaf(A.z, x.z, B.z), fmaf(A.w, x.w, B.w));
  if (RELU) {
    y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
  }
  return y;
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_apply_kernel(const float* __restrict__ x, int64_t M, int C4, const float* __restrict__ coef,
                    float* __restrict__ y) {
  pdl_wait();  // coefficients of the finalize kernel
  const BnMap m = bn_map(C4);
  if (!m.active) return;
  const float4 A = bn_c4(coef, m.g), B = bn_c4(coef + 4 * C4, m.g);
  const int64_t step = (int64_t)gridDim.x * m.rows;
  int64_t p = (int64_t)blockIdx.x * m.rows + m.r;
  for (; p + 3 * step < M; p += 4 * step) {
    int64_t i[4];
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      i[k] = (p + k * step) * C4 + m.g;
      v[k] = bn_ld(x, i[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) bn_st(y, i[k], bn_act<RELU>(v[k], A, B));
  }
  for (; p < M; p += step) {
    const int64_t i0 = p * C4 + m.g;
    bn_st(y, i0, bn_act<RELU>(bn_ld(x, i0), A, B));
  }
}

// 2x2 / stride-2 max-pool fused into the apply pass: one output pixel (4 channels) per thread iteration.
// x [B,H,W,C] -> y [B,H/2,W/2,C] (floor, like nn.MaxPool2d(2)).
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_pool_apply_kernel(const float* __restrict__ x, int B, int H, int W, int C4, const float* __restrict__ coef,
                         float* __restrict__ y) {
  pdl_wait();
  const BnMap m = bn_map(C4);
  if (!m.active) return;
  const float4 A = bn_c4(coef, m.g), Bc = bn_c4(coef + 4 * C4, m.g);
  const int Ho = H / 2, Wo = W / 2;
  const int64_t Mo = (int64_t)B * Ho * Wo;
  const int64_t step = (int64_t)gridDim.x * m.rows;
  for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < Mo; p += step) {
    const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), b = (int)(p / ((int64_t)Wo * Ho));
    const int64_t i00 = (((int64_t)b * H + 2 * ho) * W + 2 * wo) * C4 + m.g;
    const float4 v00 = bn_ld(x, i00), v01 = bn_ld(x, i00 + C4), v10 = bn_ld(x, i00 + (int64_t)W * C4),
                 v11 = bn_ld(x, i00 + (int64_t)W * C4 + C4);
    const float4 r = max4(max4(bn_act<RELU>(v00, A, Bc), bn_act<RELU>(v01, A, Bc)),
                          max4(bn_act<RELU>(v10, A, Bc), bn_act<RELU>(v11, A, Bc)));
    bn_st(y, p * C4 + m.g, r);
  }
}

// ---- backward -------------------------------------------------------------------------------------
template <bool RELU>
__device__ __forceinline__ float bn_gate_grad(float dy, float x, float A, float B) {
  return (!RELU || fmaf(A, x, B) > 0.f) ? dy : 0.f;
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t M, int C4,
                        const float* __restrict__ coef, const float* __restrict__ mean,
                        const float* __restrict__ invstd, float* __restrict__ partial) {
  pdl_trigger();
  const BnMap m = bn_map(C4);
  float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (m.active) {
    const float4 A = bn_c4(coef, m.g), B = bn_c4(coef + 4 * C4, m.g), mu = bn_c4(mean, m.g), rs = bn_c4(invstd, m.g);
    const int64_t step = (int64_t)gridDim.x * m.rows;
    for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < M; p += step) {
      const int64_t i = p * C4 + m.g;
      const float4 g = bn_ld(dy, i), v = bn_ld(x, i);
      const float gx = bn_gate_grad<RELU>(g.x, v.x, A.x, B.x), gy = bn_gate_grad<RELU>(g.y, v.y, A.y, B.y);
      const float gz = bn_gate_grad<RELU>(g.z, v.z, A.z, B.z), gw = bn_gate_grad<RELU>(g.w, v.w, A.w, B.w);
      acc[0].x += gx; acc[0].y += gy; acc[0].z += gz; acc[0].w += gw;
      acc[1].x = fmaf(gx, (v.x - mu.x) * rs.x, acc[1].x); acc[1].y = fmaf(gy, (v.y - mu.y) * rs.y, acc[1].y);
      acc[1].z = fmaf(gz, (v.z - mu.z) * rs.z, acc[1].z); acc[1].w = fmaf(gw, (v.w - mu.w) * rs.w, acc[1].w);
    }
  }
  bn_block_reduce<2>(acc, m, C4, partial + (int64_t)blockIdx.x * 8 * C4);
}

// pooled: dy [B,H/2,W/2,C]; the gradient of an output pixel goes to the FIRST maximum of its 2x2 window (row-major
// scan, like ATen's max_pool2d), then through the ReLU
struct PoolPick {
  float g[4];  // gradient reaching the four window positions (00, 01, 10, 11) for one channel
};
template <bool RELU>
__device__ __forceinline__ PoolPick pool_pick(float dy, float x00, float x01, float x10, float x11, float A, float B) {
  float y[4] = {fmaf(A, x00, B), fmaf(A, x01, B), fmaf(A, x10, B), fmaf(A, x11, B)};
  if (RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fmaxf(y[i], 0.f);
  }
  int arg = 0;
  float best = y[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (y[i] > best) {
      best = y[i];
      arg = i;
    }
  PoolPick r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.g[i] = (i == arg && (!RELU || best > 0.f)) ? dy : 0.f;
  return r;
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_pool_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int H, int W, int C4,
                             const float* __restrict__ coef, const float* __restrict__ mean,
                             const float* __restrict__ invstd, float* __restrict__ partial) {
  pdl_trigger();
  const BnMap m = bn_map(C4);
  float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (m.active) {
    const float4 A = bn_c4(coef, m.g), Bc = bn_c4(coef + 4 * C4, m.g), mu = bn_c4(mean, m.g), rs = bn_c4(invstd, m.g);
    const int Ho = H / 2, Wo = W / 2;
    const int64_t Mo = (int64_t)B * Ho * Wo;
    const int64_t step = (int64_t)gridDim.x * m.rows;
    for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < Mo; p += step) {
      const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), b = (int)(p / ((int64_t)Wo * Ho));
      const int64_t i00 = (((int64_t)b * H + 2 * ho) * W + 2 * wo) * C4 + m.g;
      const float4 g = bn_ld(dy, p * C4 + m.g);
      const float4 v[4] = {bn_ld(x, i00), bn_ld(x, i00 + C4), bn_ld(x, i00 + (int64_t)W * C4),
                           bn_ld(x, i00 + (int64_t)W * C4 + C4)};
      const PoolPick px = pool_pick<RELU>(g.x, v[0].x, v[1].x, v[2].x, v[3].x, A.x, Bc.x);
      const PoolPick py = pool_pick<RELU>(g.y, v[0].y, v[1].y, v[2].y, v[3].y, A.y, Bc.y);
      const PoolPick pz = pool_pick<RELU>(g.z, v[0].z, v[1].z, v[2].z, v[3].z, A.z, Bc.z);
      const PoolPick pw = pool_pick<RELU>(g.w, v[0].w, v[1].w, v[2].w, v[3].w, A.w, Bc.w);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[0].x += px.g[i]; acc[0].y += py.g[i]; acc[0].z += pz.g[i]; acc[0].w += pw.g[i];
        acc[1].x = fmaf(px.g[i], (v[i].x - mu.x) * rs.x, acc[1].x);
        acc[1].y = fmaf(py.g[i], (v[i].y - mu.y) * rs.y, acc[1].y);
        acc[1].z = fmaf(pz.g[i], (v[i].z - mu.z) * rs.z, acc[1].z);
        acc[1].w = fmaf(pw.g[i], (v[i].w - mu.w) * rs.w, acc[1].w);
      }
    }
  }
  bn_block_reduce<2>(acc, m, C4, partial + (int64_t)blockIdx.x * 8 * C4);
}

// partial [nparts][2][C] -> dbeta, dgamma, c1 = dbeta/Mnorm, c2 = dgamma/Mnorm (0 in eval mode).
// gmoments != NULL (global-batch statistics): the two sums come from there (all ranks, Mnorm = global rows) and
// only c1 / c2 are produced -- dbeta / dgamma are the caller's LOCAL moments.
__global__ void __launch_bounds__(kFinThreads)
    bn_bwd_finalize(const float* __restrict__ partial, int nparts, int64_t Mnorm, int C, int training,
                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ c12,
                                const double* __restrict__ gmoments, const float* __restrict__ coef,
                                float* __restrict__ dconv_bias /* NULL, or [C]: sum over pixels of dx */) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x * kTallCols + (threadIdx.x & (kTallCols - 1));
  double sb = 0.0, sg = 0.0;
  if (!gmoments) block_colsum2_tall(partial, nparts, 2 * (int64_t)C, c < C ? c : 0, C + (c < C ? c : 0), c < C, &sb, &sg);
  if (threadIdx.x >= kTallCols || c >= C) return;
  if (gmoments) {
    sb = gmoments[c];
    sg = gmoments[C + c];
  } else {
    dbeta[c] = (float)sb;
    dgamma[c] = (float)sg;
  }
  const double k1 = training ? sb / (double)Mnorm : 0.0;
  c12[c] = (float)k1;
  c12[C + c] = training ? (float)(sg / (double)Mnorm) : 0.f;
  // sum_pixels dx = A (sum g - M c1 - c2 sum xhat): zero up to round-off under batch statistics (sum xhat = 0)
  if (dconv_bias && !gmoments) dconv_bias[c] = (float)((double)coef[c] * (sb - (double)Mnorm * k1));
}

// partial [nparts][2][C] -> fp64 moments [2][C] (fixed-order second stage), what a data-parallel caller all-reduces
__global__ void __launch_bounds__(kFinThreads)
    bn_moments_finalize(const float* __restrict__ partial, int nparts, int C, double* __restrict__ moments) {
  const int c = blockIdx.x * kTallCols + (threadIdx.x & (kTallCols - 1));
  double a, b;
  block_colsum2_tall(partial, nparts, 2 * (int64_t)C, c < C ? c : 0, C + (c < C ? c : 0), c < C, &a, &b);
  if (threadIdx.x >= kTallCols || c >= C) return;
  moments[c] = a;
  moments[C + c] = b;
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t M, int C4,
                        const float* __restrict__ coef, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ c12, float* __restrict__ dx) {
  pdl_wait();
  const BnMap m = bn_map(C4);
  if (!m.active) return;
  const float4 A = bn_c4(coef, m.g), B = bn_c4(coef + 4 * C4, m.g), mu = bn_c4(mean, m.g), rs = bn_c4(invstd, m.g);
  const float4 k1 = bn_c4(c12, m.g), k2 = bn_c4(c12 + 4 * C4, m.g);
  const int64_t step = (int64_t)gridDim.x * m.rows;
  for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < M; p += step) {
    const int64_t i = p * C4 + m.g;
    const float4 g = bn_ld(dy, i), v = bn_ld(x, i);
    float4 d;
    d.x = A.x * (bn_gate_grad<RELU>(g.x, v.x, A.x, B.x) - k1.x - (v.x - mu.x) * rs.x * k2.x);
    d.y = A.y * (bn_gate_grad<RELU>(g.y, v.y, A.y, B.y) - k1.y - (v.y - mu.y) * rs.y * k2.y);
    d.z = A.z * (bn_gate_grad<RELU>(g.z, v.z, A.z, B.z) - k1.z - (v.z - mu.z) * rs.z * k2.z);
    d.w = A.w * (bn_gate_grad<RELU>(g.w, v.w, A.w, B.w) - k1.w - (v.w - mu.w) * rs.w * k2.w);
    bn_st(dx, i, d);
  }
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads)
    bn_pool_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, int B, int H, int W, int C4,
                             const float* __restrict__ coef, const float* __restrict__ mean,
                             const float* __restrict__ invstd, const float* __restrict__ c12,
                             float* __restrict__ dx) {
  pdl_wait();
  const BnMap m = bn_map(C4);
  if (!m.active) return;
  const float4 A = bn_c4(coef, m.g), Bc = bn_c4(coef + 4 * C4, m.g), mu = bn_c4(mean, m.g), rs = bn_c4(invstd, m.g);
  const float4 k1 = bn_c4(c12, m.g), k2 = bn_c4(c12 + 4 * C4, m.g);
  // one 2x2 window per iteration over the CEIL grid, so odd trailing rows / columns (not pooled: no gradient from
  // dy) still get their dx = A (0 - c1 - xhat c2)
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2, Ho = H / 2, Wo = W / 2;
  const int64_t Mc = (int64_t)B * Hc * Wc;
  const int64_t step = (int64_t)gridDim.x * m.rows;
  for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < Mc; p += step) {
    const int wo = (int)(p % Wc), ho = (int)((p / Wc) % Hc), b = (int)(p / ((int64_t)Wc * Hc));
    const bool pooled = ho < Ho && wo < Wo;
    const int64_t i00 = (((int64_t)b * H + 2 * ho) * W + 2 * wo) * C4 + m.g;
    const bool in01 = 2 * wo + 1 < W, in10 = 2 * ho + 1 < H;
    const float4 zero = make_float4(0, 0, 0, 0);
    const float4 g = pooled ? bn_ld(dy, (((int64_t)b * Ho + ho) * Wo + wo) * C4 + m.g) : zero;
    const int64_t idx[4] = {i00, i00 + C4, i00 + (int64_t)W * C4, i00 + (int64_t)W * C4 + C4};
    const bool in[4] = {true, in01, in10, in01 && in10};
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = in[i] ? bn_ld(x, idx[i]) : zero;
    PoolPick px, py, pz, pw;
    if (pooled) {
      px = pool_pick<RELU>(g.x, v[0].x, v[1].x, v[2].x, v[3].x, A.x, Bc.x);
      py = pool_pick<RELU>(g.y, v[0].y, v[1].y, v[2].y, v[3].y, A.y, Bc.y);
      pz = pool_pick<RELU>(g.z, v[0].z, v[1].z, v[2].z, v[3].z, A.z, Bc.z);
      pw = pool_pick<RELU>(g.w, v[0].w, v[1].w, v[2].w, v[3].w, A.w, Bc.w);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) px.g[i] = py.g[i] = pz.g[i] = pw.g[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!in[i]) continue;
      float4 d;
      d.x = A.x * (px.g[i] - k1.x - (v[i].x - mu.x) * rs.x * k2.x);
      d.y = A.y * (py.g[i] - k1.y - (v[i].y - mu.y) * rs.y * k2.y);
      d.z = A.z * (pz.g[i] - k1.z - (v[i].z - mu.z) * rs.z * k2.z);
      d.w = A.w * (pw.g[i] - k1.w - (v[i].w - mu.w) * rs.w * k2.w);
      bn_st(dx, idx[i], d);
    }
  }
}

static int bn_check(int64_t M, int C) {
  if (M < 1 || C < 4) return VMTL_EINVAL;
  if (C % 4 != 0 || C > 4 * kBnThreads) return VMTL_EUNSUPPORTED;
  return VMTL_OK;
}
static int bn_max_blocks() { return sm_count() * 8; }

}  // namespace vmtl

using namespace vmtl;

extern "C" size_t vmtl_bnrelu_workspace_bytes(int64_t M, int C) {
  if (bn_check(M, C) != VMTL_OK) return 0;
  // per-block partials [blocks][2][C] + c1/c2 [2][C]
  return ((size_t)bn_max_blocks() * 2 * C + 2 * (size_t)C) * sizeof(float) + 256;
}

// phase 0: the whole op on local statistics.  Global-batch statistics split it around the caller's
// all-reduce: phase 1 = statistics pass -> fp64 `moments` [2][C]; phase 2 = finalize from the (all-reduced) moments
// over `Mstat` rows + apply pass.
static int bnrelu_fwd_impl(const float* x, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float momentum, float eps, int training, int relu, int64_t M, int C,
                           int B, int H, int W, int pool, float* y, float* save_mean, float* save_invstd, float* coef,
                           void* workspace, size_t workspace_bytes, void* stream, int phase = 0,
                           double* moments = nullptr, int64_t Mstat = 0, const float* conv_bias = nullptr) {
  int rc = bn_check(M, C);
  if (rc != VMTL_OK) return rc;
  if (conv_bias && !training) return VMTL_EINVAL;  // running statistics: the caller keeps the bias in x
  if (phase == 1) {
    if (!x || !moments || !workspace) return VMTL_EINVAL;
    if (!aligned16(x) || !aligned16(workspace)) return VMTL_EALIGN;
    if (workspace_bytes < vmtl_bnrelu_workspace_bytes(M, C)) return VMTL_EWORKSPACE;
    cudaStream_t st1 = static_cast<cudaStream_t>(stream);
    float* part = static_cast<float*>(workspace);
    const int np = bn_grid(M, C, blocks_per_sm(bn_stats_kernel, kBnThreads, 0, 8));
    bn_stats_kernel<<<np, kBnThreads, 0, st1>>>(x, M, C / 4, part);
    if ((rc = launch_status()) != VMTL_OK) return rc;
    bn_moments_finalize<<<(C + kTallCols - 1) / kTallCols, kFinThreads, 0, st1>>>(part, np, C, moments);
    return launch_status();
  }
  if (phase == 2 && (!moments || Mstat < 1 || !training)) return VMTL_EINVAL;
  if (!x || !gamma || !beta || !save_mean || !save_invstd || !coef || !workspace) return VMTL_EINVAL;
  if (!training && (!running_mean || !running_var)) return VMTL_EINVAL;
  if (!aligned16(x) || (y && !aligned16(y)) || !aligned16(coef) || !aligned16(workspace)) return VMTL_EALIGN;
  if (workspace_bytes < vmtl_bnrelu_workspace_bytes(M, C)) return VMTL_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C4 = C / 4;
  float* partial = static_cast<float*>(workspace);
  int nparts = 0;
  if (training && phase == 0) {
    nparts = bn_grid(M, C, blocks_per_sm(bn_stats_kernel, kBnThreads, 0, 8));
    bn_stats_kernel<<<nparts, kBnThreads, 0, st>>>(x, M, C4, partial);
    if ((rc = launch_status()) != VMTL_OK) return rc;
  }
  // finalize and apply are programmatic dependents (vmtl_common.cuh): each is scheduled while its predecessor drains
  launch_pdl(bn_fwd_finalize, (C + kTallCols - 1) / kTallCols, kFinThreads, 0, st, (const float*)partial, nparts, phase == 2 ? Mstat : M, C, eps,
             momentum, training, gamma, beta, running_mean, running_var, save_mean, save_invstd, coef,
             (const double*)(phase == 2 ? moments : nullptr), conv_bias);
  if ((rc = launch_status()) != VMTL_OK) return rc;
  if (!y) return VMTL_OK;
  if (pool) {
    const int64_t Mo = (int64_t)B * (H / 2) * (W / 2);
    if (Mo < 1) return VMTL_EINVAL;
    if (relu) {
      const int grid = bn_grid(Mo, C, blocks_per_sm(bn_pool_apply_kernel<true>, kBnThreads, 0, 8));
      launch_pdl(bn_pool_apply_kernel<true>, grid, kBnThreads, 0, st, x, B, H, W, C4, (const float*)coef, y);
    } else {
      const int grid = bn_grid(Mo, C, blocks_per_sm(bn_pool_apply_kernel<false>, kBnThreads, 0, 8));
      launch_pdl(bn_pool_apply_kernel<false>, grid, kBnThreads, 0, st, x, B, H, W, C4, (const float*)coef, y);
    }
    return launch_status();
  }
  if (relu) {
    const int grid = bn_grid(M, C, blocks_per_sm(bn_apply_kernel<true>, kBnThreads, 0, 8));
    launch_pdl(bn_apply_kernel<true>, grid, kBnThreads, 0, st, x, M, C4, (const float*)coef, y);
  } else {
    const int grid = bn_grid(M, C, blocks_per_sm(bn_apply_kernel<false>, kBnThreads, 0, 8));
    launch_pdl(bn_apply_kernel<false>, grid, kBnThreads, 0, st, x, M, C4, (const float*)coef, y);
  }
  return launch_status();
}

extern "C" int vmtl_bnrelu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float momentum, float eps, int training, int relu, int64_t M,
                               int C, float* y, float* save_mean, float* save_invstd, float* coef,
                               const float* conv_bias, void* workspace, size_t workspace_bytes, void* stream) {
  return bnrelu_fwd_impl(x, gamma, beta, running_mean, running_var, momentum, eps, training, relu, M, C, 0, 0, 0, 0, y,
                         save_mean, save_invstd, coef, workspace, workspace_bytes, stream, 0, nullptr, 0, conv_bias);
}

extern "C" int vmtl_bnrelu_pool_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, float momentum, float eps, int training, int relu, int B,
                                    int H, int W, int C, float* y, float* save_mean, float* save_invstd, float* coef,
                                    const float* conv_bias, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 2 || W < 2 || !y) return VMTL_EINVAL;
  return bnrelu_fwd_impl(x, gamma, beta, running_mean, running_var, momentum, eps, training, relu, (int64_t)B * H * W,
                         C, B, H, W, 1, y, save_mean, save_invstd, coef, workspace, workspace_bytes, stream, 0, nullptr, 0,
                         conv_bias);
}

// phase 0: whole op; phase 1: statistics pass -> LOCAL fp64 moments [2][C] = (sum g, sum g xhat) (also the
// caller's dbeta / dgamma); phase 2: c1 / c2 from the all-reduced moments over `Mstat` rows + apply pass.
static int bnrelu_bwd_impl(const float* dy, const float* x, const float* coef, const float* save_mean,
                           const float* save_invstd, int training, int relu, int64_t M, int C, int B, int H, int W,
                           int pool, float* dx, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                           void* stream, int phase = 0, double* moments = nullptr, int64_t Mstat = 0,
                           float* dconv_bias = nullptr) {
  int rc = bn_check(M, C);
  if (rc != VMTL_OK) return rc;
  if (phase != 0 && !moments) return VMTL_EINVAL;
  if (phase == 2 && (!dx || Mstat < 1)) return VMTL_EINVAL;
  if (!dy || !x || !coef || !save_mean || !save_invstd || (phase == 0 && (!dgamma || !dbeta)) || !workspace)
    return VMTL_EINVAL;
  if (!aligned16(dy) || !aligned16(x) || (dx && !aligned16(dx)) || !aligned16(coef) || !aligned16(workspace))
    return VMTL_EALIGN;
  if (workspace_bytes < vmtl_bnrelu_workspace_bytes(M, C)) return VMTL_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C4 = C / 4;
  float* partial = static_cast<float*>(workspace);
  float* c12 = partial + (size_t)bn_max_blocks() * 2 * C;
  int nparts;
  (void)nparts;
#define VMTL_BN_LAUNCH(KERN, ROWS, ...)                                                    \
  do {                                                                                     \
    nparts = bn_grid(ROWS, C, blocks_per_sm(KERN, kBnThreads, 0, 8));                      \
    KERN<<<nparts, kBnThreads, 0, st>>>(__VA_ARGS__);                                      \
  } while (0)
  nparts = 0;
  if (phase != 2) {
    if (pool) {
      const int64_t Mo = (int64_t)B * (H / 2) * (W / 2);
      if (relu) VMTL_BN_LAUNCH(bn_pool_bwd_stats_kernel<true>, Mo, dy, x, B, H, W, C4, coef, save_mean, save_invstd, partial);
      else VMTL_BN_LAUNCH(bn_pool_bwd_stats_kernel<false>, Mo, dy, x, B, H, W, C4, coef, save_mean, save_invstd, partial);
    } else {
      if (relu) VMTL_BN_LAUNCH(bn_bwd_stats_kernel<true>, M, dy, x, M, C4, coef, save_mean, save_invstd, partial);
      else VMTL_BN_LAUNCH(bn_bwd_stats_kernel<false>, M, dy, x, M, C4, coef, save_mean, save_invstd, partial);
    }
    if ((rc = launch_status()) != VMTL_OK) return rc;
  }
  if (phase == 1) {
    bn_moments_finalize<<<(C + kTallCols - 1) / kTallCols, kFinThreads, 0, st>>>(partial, nparts, C, moments);
    return launch_status();
  }
  launch_pdl(bn_bwd_finalize, (C + kTallCols - 1) / kTallCols, kFinThreads, 0, st, (const float*)partial, nparts, phase == 2 ? Mstat : M, C,
             training, dgamma, dbeta, c12, (const double*)(phase == 2 ? moments : nullptr), coef, dconv_bias);
  if ((rc = launch_status()) != VMTL_OK) return rc;
  if (!dx) return VMTL_OK;
#define VMTL_BN_LAUNCH_PDL(KERN, ROWS, ...)                                                \
  launch_pdl(KERN, bn_grid(ROWS, C, blocks_per_sm(KERN, kBnThreads, 0, 8)), kBnThreads, 0, st, __VA_ARGS__)
  const float* c12c = c12;
  if (pool) {
    const int64_t Mc = (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2);
    if (relu) VMTL_BN_LAUNCH_PDL(bn_pool_bwd_apply_kernel<true>, Mc, dy, x, B, H, W, C4, coef, save_mean, save_invstd, c12c, dx);
    else VMTL_BN_LAUNCH_PDL(bn_pool_bwd_apply_kernel<false>, Mc, dy, x, B, H, W, C4, coef, save_mean, save_invstd, c12c, dx);
  } else {
    if (relu) VMTL_BN_LAUNCH_PDL(bn_bwd_apply_kernel<true>, M, dy, x, M, C4, coef, save_mean, save_invstd, c12c, dx);
    else VMTL_BN_LAUNCH_PDL(bn_bwd_apply_kernel<false>, M, dy, x, M, C4, coef, save_mean, save_invstd, c12c, dx);
  }
#undef VMTL_BN_LAUNCH_PDL
#undef VMTL_BN_LAUNCH
  return launch_status();
}

extern "C" int vmtl_bnrelu_bwd(const float* dy, const float* x, const float* coef, const float* save_mean,
                               const float* save_invstd, int training, int relu, int64_t M, int C, float* dx,
                               float* dgamma, float* dbeta, float* dconv_bias, void* workspace, size_t workspace_bytes,
                               void* stream) {
  return bnrelu_bwd_impl(dy, x, coef, save_mean, save_invstd, training, relu, M, C, 0, 0, 0, 0, dx, dgamma, dbeta,
                         workspace, workspace_bytes, stream, 0, nullptr, 0, dconv_bias);
}

extern "C" int vmtl_bnrelu_pool_bwd(const float* dy, const float* x, const float* coef, const float* save_mean,
                                    const float* save_invstd, int training, int relu, int B, int H, int W, int C,
                                    float* dx, float* dgamma, float* dbeta, float* dconv_bias, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 2 || W < 2) return VMTL_EINVAL;
  return bnrelu_bwd_impl(dy, x, coef, save_mean, save_invstd, training, relu, (int64_t)B * H * W, C, B, H, W, 1, dx,
                         dgamma, dbeta, workspace, workspace_bytes, stream, 0, nullptr, 0, dconv_bias);
}

// ---- global-batch statistics: the two halves of each op around the caller's all-reduce -------------
extern "C" int vmtl_bn_moments(const float* x, int64_t M, int C, double* moments, void* workspace,
                               size_t workspace_bytes, void* stream) {
  return bnrelu_fwd_impl(x, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 1, 0, M, C, 0, 0, 0, 0, nullptr, nullptr,
                         nullptr, nullptr, workspace, workspace_bytes, stream, 1, moments, 0);
}

extern "C" int vmtl_bnrelu_fwd_global(const float* x, const float* gamma, const float* beta, float* running_mean,
                                      float* running_var, float momentum, float eps, int relu, int pool, int B, int H,
                                      int W, int C, const double* moments, int64_t M_global, float* y,
                                      float* save_mean, float* save_invstd, float* coef, const float* conv_bias,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 1 || W < 1 || (pool && (H < 2 || W < 2 || !y))) return VMTL_EINVAL;
  return bnrelu_fwd_impl(x, gamma, beta, running_mean, running_var, momentum, eps, 1, relu, (int64_t)B * H * W, C, B, H,
                         W, pool, y, save_mean, save_invstd, coef, workspace, workspace_bytes, stream, 2,
                         const_cast<double*>(moments), M_global, conv_bias);
}

extern "C" int vmtl_bnrelu_bwd_moments(const float* dy, const float* x, const float* coef, const float* save_mean,
                                       const float* save_invstd, int relu, int pool, int B, int H, int W, int C,
                                       double* moments, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 1 || W < 1 || (pool && (H < 2 || W < 2))) return VMTL_EINVAL;
  return bnrelu_bwd_impl(dy, x, coef, save_mean, save_invstd, 1, relu, (int64_t)B * H * W, C, B, H, W, pool, nullptr,
                         nullptr, nullptr, workspace, workspace_bytes, stream, 1, moments, 0);
}

extern "C" int vmtl_bnrelu_bwd_global(const float* dy, const float* x, const float* coef, const float* save_mean,
                                      const float* save_invstd, int relu, int pool, int B, int H, int W, int C,
                                      const double* moments, int64_t M_global, float* dx, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 1 || W < 1 || (pool && (H < 2 || W < 2))) return VMTL_EINVAL;
  return bnrelu_bwd_impl(dy, x, coef, save_mean, save_invstd, 1, relu, (int64_t)B * H * W, C, B, H, W, pool, dx, nullptr,
                         nullptr, workspace, workspace_bytes, stream, 2, const_cast<double*>(moments), M_global);
}

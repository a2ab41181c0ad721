// MTAN attention gate  y = s * sigmoid(BN(h @ W^T + bias))  -- entry points, the streaming
// (HBM-bound) phases and the fp32 CUDA-core contraction.  The tensor-core contraction lives
// in gate_tc.cu.
//
// Reference: conv2 -> bn2 -> sigmoid -> mul at vision_mtl/models/mtan_model.py:71-75 (encoder)
// and :158-162 (decoder): five ATen kernels and four materialised [M,N] intermediates.
//
// BatchNorm is in TRAINING mode on the measured path, so z = h W^T + b needs its
// per-channel batch statistics before the gate can be emitted:
//   fwd  phase 1 : contraction, writes z, per-CTA column partials (sum z, sum z^2)
//        finalize: fp64 fixed-order reduction -> mean, invstd, running-stat update, (A,B)
//        phase 2 : y = s * sigmoid(A*z + B)                       streams z,s -> y
//   bwd  pass 1  : ds = dy*a ; per-channel sum du, sum du*zhat ; P1 = du^T h, P2 = zhat^T h, hsum
//        finalize: dbeta, dgamma, c1 = dbeta/M, c2 = dgamma/M, dW = A (P1 - c1 hsum - c2 P2), db
//        pass 2  : dz = gamma*invstd*(du - c1 - zhat*c2) per tile in registers ; dh = dz W
//   (the CUDA-core path keeps the direct form: phase A statistics, dz materialised, two sgemms)
// Issued bytes (train, tensor-core path): fwd 4M(K + 4N), bwd 4M(2K + 7N) (K = 128): the backward
// reads (dy, s, z) twice (statistics + dW pass, dh pass), h once, writes ds and dh; dz never exists in HBM.
#include <math.h>

#include "gate_internal.cuh"

namespace vmtl {

constexpr int kEwThreads = 256;

struct EwMap {
  int rows, r, g;
  bool active;
};
__device__ __forceinline__ EwMap ew_map(int C4) {
  EwMap m;
  m.rows = kEwThreads / C4;
  m.r = threadIdx.x / C4;
  m.g = threadIdx.x - m.r * C4;
  m.active = m.r < m.rows;
  return m;
}
static int ew_grid(int64_t M, int N, int per_sm = 8) {
  const int rows = kEwThreads / (N / 4);
  int64_t want = (M + rows - 1) / rows;
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i4) {
  return ldg_stream(reinterpret_cast<const float4*>(p) + i4);
}
__device__ __forceinline__ void st4(float* p, int64_t i4, const float4& v) {
  stg_stream(reinterpret_cast<float4*>(p) + i4, v);
}
__device__ __forceinline__ float4 ldc4(const float* p, int g) {
  return reinterpret_cast<const float4*>(p)[g];
}

// block reduction over the `rows` threads sharing a channel group; writes one partial row
template <int NV>
__device__ __forceinline__ void ew_block_reduce(const float4 (&acc)[NV], const EwMap& m, int C4,
                                                float* partial_row /* [NV][4*C4] */) {
  __shared__ float4 s_red[kEwThreads * NV];
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

// ---- forward, column statistics of z (fp32 FFMA path only; the TC kernel fuses this) -----
__global__ void __launch_bounds__(kEwThreads)
    gate_colstats_kernel(const float* __restrict__ z, int64_t M, int C4, float* __restrict__ partial) {
  const EwMap m = ew_map(C4);
  float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (m.active) {
    for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < M; p += (int64_t)gridDim.x * m.rows) {
      const float4 v = ld4(z, p * C4 + m.g);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      acc[1].x = fmaf(v.x, v.x, acc[1].x); acc[1].y = fmaf(v.y, v.y, acc[1].y);
      acc[1].z = fmaf(v.z, v.z, acc[1].z); acc[1].w = fmaf(v.w, v.w, acc[1].w);
    }
  }
  ew_block_reduce<2>(acc, m, C4, partial + (int64_t)blockIdx.x * 8 * C4);
}

// partial [nparts][2][N] -> mean, invstd, coefA/B, running statistics (nn.BatchNorm2d rules:
// biased variance normalises, unbiased variance feeds running_var)
__global__ void __launch_bounds__(kFinThreads)
    gate_fwd_stats_finalize(const float* __restrict__ partial, int nparts, int64_t M, int N,
                                        float eps, float momentum, int training,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        float* __restrict__ running_mean, float* __restrict__ running_var,
                                        float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                        float* __restrict__ coefA, float* __restrict__ coefB,
                                        float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                        const double* __restrict__ gmoments /* NULL, or global [2][N] over M rows */) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s = 0.0, q = 0.0;
  if (training && !gmoments) {  // block-uniform branch: the reduction synchronises
    block_colsum2(partial, nparts, 2 * (int64_t)N, c < N ? c : 0, N + (c < N ? c : 0), c < N, &s, &q);
  }
  if (threadIdx.x >= 32 || c >= N) return;
  if (training && gmoments) {
    s = gmoments[c];
    q = gmoments[N + c];
  }
  double mean, var;
  if (training) {
    mean = s / (double)M;
    var = q / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    if (running_var) {
      const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const double inv = 1.0 / sqrt(var + (double)eps);
  const double a = (double)gamma[c] * inv;
  mean_out[c] = (float)mean;
  invstd_out[c] = (float)inv;
  coefA[c] = (float)a;
  coefB[c] = (float)((double)beta[c] - mean * a);
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = (float)inv;
}

// partial [nparts][2][N] -> fp64 moments [2][N]: what a data-parallel caller all-reduces between the two phases
__global__ void __launch_bounds__(kFinThreads)
    gate_moments_finalize(const float* __restrict__ partial, int nparts, int N, double* __restrict__ moments) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double a, b;
  block_colsum2(partial, nparts, 2 * (int64_t)N, c < N ? c : 0, N + (c < N ? c : 0), c < N, &a, &b);
  if (threadIdx.x >= 32 || c >= N) return;
  moments[c] = a;
  moments[N + c] = b;
}

// ---- forward phase 2: y = s * sigmoid(A*z + B) ------------------------------------------
__device__ __forceinline__ float4 gate_act(const float4& z, const float4& A, const float4& B) {
  return make_float4(sigmoidf_acc(fmaf(A.x, z.x, B.x)), sigmoidf_acc(fmaf(A.y, z.y, B.y)),
                     sigmoidf_acc(fmaf(A.z, z.z, B.z)), sigmoidf_acc(fmaf(A.w, z.w, B.w)));
}

__global__ void __launch_bounds__(kEwThreads)
    gate_apply_kernel(const float* __restrict__ z, const float* __restrict__ s, int64_t M, int C4,
                      const float* __restrict__ coefA, const float* __restrict__ coefB,
                      float* __restrict__ y) {
  const EwMap m = ew_map(C4);
  if (!m.active) return;
  const float4 A = ldc4(coefA, m.g), B = ldc4(coefB, m.g);
  const int64_t step = (int64_t)gridDim.x * m.rows;
  int64_t p = (int64_t)blockIdx.x * m.rows + m.r;
  for (; p + step < M; p += 2 * step) {
    const int64_t i0 = p * C4 + m.g, i1 = (p + step) * C4 + m.g;
    const float4 z0 = ld4(z, i0), s0 = ld4(s, i0), z1 = ld4(z, i1), s1 = ld4(s, i1);
    const float4 a0 = gate_act(z0, A, B), a1 = gate_act(z1, A, B);
    st4(y, i0, make_float4(s0.x * a0.x, s0.y * a0.y, s0.z * a0.z, s0.w * a0.w));
    st4(y, i1, make_float4(s1.x * a1.x, s1.y * a1.y, s1.z * a1.z, s1.w * a1.w));
  }
  for (; p < M; p += step) {
    const int64_t i0 = p * C4 + m.g;
    const float4 z0 = ld4(z, i0), s0 = ld4(s, i0);
    const float4 a0 = gate_act(z0, A, B);
    st4(y, i0, make_float4(s0.x * a0.x, s0.y * a0.y, s0.z * a0.z, s0.w * a0.w));
  }
}

// ---- backward phase A: ds, sum du, sum du*zhat -------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
    gate_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ s,
                          const float* __restrict__ z, int64_t M, int C4,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ ds, float* __restrict__ partial) {
  const EwMap m = ew_map(C4);
  float4 acc[2] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
  if (m.active) {
    // folded BN coefficients straight from the saved statistics (no separate coefficient launch)
    const float4 ga = ldc4(gamma, m.g), be = ldc4(beta, m.g), mu = ldc4(mean, m.g), rs = ldc4(invstd, m.g);
    const float4 A = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
    const float4 B = make_float4(be.x - mu.x * A.x, be.y - mu.y * A.y, be.z - mu.z * A.z, be.w - mu.w * A.w);
    for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < M; p += (int64_t)gridDim.x * m.rows) {
      const int64_t i = p * C4 + m.g;
      const float4 g = ld4(dy, i), sv = ld4(s, i), zv = ld4(z, i);
      const float4 a = gate_act(zv, A, B);
      if (ds) st4(ds, i, make_float4(g.x * a.x, g.y * a.y, g.z * a.z, g.w * a.w));
      const float dux = g.x * sv.x * a.x * (1.f - a.x), duy = g.y * sv.y * a.y * (1.f - a.y);
      const float duz = g.z * sv.z * a.z * (1.f - a.z), duw = g.w * sv.w * a.w * (1.f - a.w);
      acc[0].x += dux; acc[0].y += duy; acc[0].z += duz; acc[0].w += duw;
      acc[1].x = fmaf(dux, (zv.x - mu.x) * rs.x, acc[1].x);
      acc[1].y = fmaf(duy, (zv.y - mu.y) * rs.y, acc[1].y);
      acc[1].z = fmaf(duz, (zv.z - mu.z) * rs.z, acc[1].z);
      acc[1].w = fmaf(duw, (zv.w - mu.w) * rs.w, acc[1].w);
    }
  }
  ew_block_reduce<2>(acc, m, C4, partial + (int64_t)blockIdx.x * 8 * C4);
}

// also publishes the folded BN coefficients for phase B (same arithmetic as gate_bwd_stats_kernel)
__global__ void __launch_bounds__(kFinThreads)
    gate_bwd_stats_finalize(const float* __restrict__ partial, int nparts, int64_t M, int N,
                                        int training, float* __restrict__ dgamma,
                                        float* __restrict__ dbeta, float* __restrict__ c1,
                                        float* __restrict__ c2, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, float* __restrict__ coefA,
                                        float* __restrict__ coefB, float* __restrict__ mean_out,
                                        float* __restrict__ invstd_out,
                                        const double* __restrict__ gmoments /* NULL, or global [2][N] */,
                                        int64_t Mstat) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double sb, sg;
  block_colsum2(partial, nparts, 2 * (int64_t)N, c < N ? c : 0, N + (c < N ? c : 0), c < N, &sb, &sg);
  if (threadIdx.x >= 32 || c >= N) return;
  dbeta[c] = (float)sb;  // parameter gradients stay LOCAL sums (the data-parallel wrapper averages them)
  dgamma[c] = (float)sg;
  if (gmoments) {
    sb = gmoments[c];
    sg = gmoments[N + c];
    M = Mstat;
  }
  c1[c] = training ? (float)(sb / (double)M) : 0.f;
  c2[c] = training ? (float)(sg / (double)M) : 0.f;
  const float a = gamma[c] * invstd[c];
  coefA[c] = a;
  coefB[c] = beta[c] - mean[c] * a;
  mean_out[c] = mean[c];
  invstd_out[c] = invstd[c];
}

// ---- backward finalize of the tensor-core path ---------------------------------------------------
// Per-CTA partials of pass 1 (gate_tc_bwd_tma.cuh) -> every parameter gradient, in fp64, fixed order:
//   dbeta = sum du, dgamma = sum du zhat, c1 = dbeta/M, c2 = dgamma/M (0 in eval mode),
//   dW[n,k] = A_n (P1[n,k] - c1_n hsum[k] - c2_n P2[n,k]),  dbias_n = A_n (sum du - M c1 - c2 sum zhat)
// (dbias is analytically zero under batch statistics: what is left is the round-off of sum zhat).
// CTA b of pass 1 owned the column chunk b % nch: the partials of chunk c are rows c, c + nch, ... of each buffer.
//
// Blocks [0, N * K/32): one [1 row n x 32 k] patch of dW each.  1024 threads = 32 k-lanes x (1 row x 32 part groups): every thread sums its share
// of the partial rows (a handful of independent loads in flight), the 8 groups are combined in a fixed order.
// The last ceil(N/32) blocks produce the per-channel outputs and the folded coefficients pass 2 needs.
// ROWS rows of dW per block, 32 / ROWS part groups.  Narrow gates (N <= 64: every CTA of pass 1 holds the whole N, 148
// partial slices per dW element) want ROWS = 1: with 4 rows the launch was 33-66 blocks walking 5-step chains (17 us in
// ncu); wide gates (column chunks split the CTAs: 37-74 slices per element) want ROWS = 4, else the launch is a thousand
// mostly idle 1024-thread blocks (26 us at N = 256 against 14).
template <int kFinRows>
__global__ void __launch_bounds__(kFinThreads)
    gate_bwd_tc_finalize(const float* __restrict__ pw_partial, const float* __restrict__ hs_partial,
                         const float* __restrict__ col_partial, int nparts, int nch, int64_t M, int N, int K,
                         int training, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ dW,
                         float* __restrict__ dbias, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ coefA,
                         float* __restrict__ coefB, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                         const double* __restrict__ gmoments /* NULL, or global (sum du, sum du zhat) [2][N] */,
                         int64_t Mstat /* rows behind gmoments */) {
  constexpr int kFinGroups = 32 / kFinRows;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int Nc = N / nch, per_chunk = nparts / nch;
  const int kslabs = K / 32;
  const int nb_w = (N / kFinRows) * kslabs;
  __shared__ double s_acc[4][32][33];
  if ((int)blockIdx.x < nb_w) {  // block-uniform
    const int r = ly % kFinRows, pg = ly / kFinRows;
    const int n = ((int)blockIdx.x / kslabs) * kFinRows + r, k = ((int)blockIdx.x % kslabs) * 32 + lx;
    const int c = n / Nc, nl = n % Nc;  // the 4 rows of a patch share their chunk (Nc is a multiple of 4)
    const float* pw = pw_partial + (int64_t)c * 2 * Nc * K;
    const int64_t pw_stride = (int64_t)nch * 2 * Nc * K;
    const float* hp = hs_partial + (int64_t)c * K;  // any chunk's CTAs cover every unit exactly once: use this one's
    const float* cp = col_partial + (int64_t)c * 3 * Nc;
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    // the kernel is one dependent-latency chain per thread: keep 4 partial rows (up to 16 loads) in flight
    int p = pg;
    for (; p + 3 * kFinGroups < per_chunk; p += 4 * kFinGroups) {
      float v1[4], v2[4], v3[4], v4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int pp = p + q * kFinGroups;
        v1[q] = pw[pp * pw_stride + (int64_t)nl * K + k];
        v2[q] = pw[pp * pw_stride + (int64_t)(Nc + nl) * K + k];
        v3[q] = r == 0 ? hp[(int64_t)pp * nch * K + k] : 0.f;
        v4[q] = lx < 2 ? cp[(int64_t)pp * nch * 3 * Nc + lx * Nc + nl] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a1 += (double)v1[q];
        a2 += (double)v2[q];
        a3 += (double)v3[q];
        a4 += (double)v4[q];
      }
    }
    for (; p < per_chunk; p += kFinGroups) {
      a1 += (double)pw[p * pw_stride + (int64_t)nl * K + k];
      a2 += (double)pw[p * pw_stride + (int64_t)(Nc + nl) * K + k];
      if (r == 0) a3 += (double)hp[(int64_t)p * nch * K + k];
      if (lx < 2) a4 += (double)cp[(int64_t)p * nch * 3 * Nc + lx * Nc + nl];
    }
    s_acc[0][ly][lx] = a1;
    s_acc[1][ly][lx] = a2;
    s_acc[2][ly][lx] = a3;
    s_acc[3][ly][lx] = a4;
    __syncthreads();
    if (pg != 0) return;
    double p1 = 0.0, p2 = 0.0, hs = 0.0, sdu = 0.0, sdz = 0.0;
#pragma unroll
    for (int g = 0; g < kFinGroups; ++g) {
      p1 += s_acc[0][g * kFinRows + r][lx];
      p2 += s_acc[1][g * kFinRows + r][lx];
      hs += s_acc[2][g * kFinRows][lx];
      sdu += s_acc[3][g * kFinRows + r][0];
      sdz += s_acc[3][g * kFinRows + r][1];
    }
    double k1 = training ? sdu / (double)M : 0.0, k2 = training ? sdz / (double)M : 0.0;
    if (gmoments && training) {  // normalisation terms of the GLOBAL batch; P1, P2, hsum stay this replica's
      k1 = gmoments[n] / (double)Mstat;
      k2 = gmoments[N + n] / (double)Mstat;
    }
    const double a = (double)gamma[n] * (double)invstd[n];
    dW[(int64_t)n * K + k] = (float)(a * (p1 - k1 * hs - k2 * p2));
    return;
  }
  const int col = ((int)blockIdx.x - nb_w) * 32 + lx;  // 32 consecutive columns never straddle a chunk (Nc = 32 or 64)
  const bool ok = col < N;
  const int c = (ok ? col : 0) / Nc, nl = (ok ? col : 0) % Nc;
  const float* cp = col_partial + (int64_t)c * 3 * Nc;
  double a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (ok)
    for (int p = ly; p < per_chunk; p += 32) {
      const float* row = cp + (int64_t)p * nch * 3 * Nc;
      a1 += (double)row[nl];
      a2 += (double)row[Nc + nl];
      a3 += (double)row[2 * Nc + nl];
    }
  s_acc[0][ly][lx] = a1;
  s_acc[1][ly][lx] = a2;
  s_acc[2][ly][lx] = a3;
  __syncthreads();
  if (ly != 0 || !ok) return;
  double sdu = 0.0, sdz = 0.0, sz = 0.0;
#pragma unroll
  for (int g = 0; g < 32; ++g) {
    sdu += s_acc[0][g][lx];
    sdz += s_acc[1][g][lx];
    sz += s_acc[2][g][lx];
  }
  dbeta[col] = (float)sdu;  // parameter gradients stay LOCAL sums
  dgamma[col] = (float)sdz;
  double k1 = training ? sdu / (double)M : 0.0, k2 = training ? sdz / (double)M : 0.0;
  if (gmoments && training) {  // normalisation terms of the GLOBAL batch
    k1 = gmoments[col] / (double)Mstat;
    k2 = gmoments[N + col] / (double)Mstat;
  }
  const double a = (double)gamma[col] * (double)invstd[col];
  dbias[col] = (float)(a * (sdu - (double)M * k1 - k2 * sz));
  c1[col] = (float)k1;
  c2[col] = (float)k2;
  const float af = gamma[col] * invstd[col];  // same fp32 arithmetic as pass 1 used for the activation
  coefA[col] = af;
  coefB[col] = beta[col] - mean[col] * af;
  mean_out[col] = mean[col];
  invstd_out[col] = invstd[col];
}

// col_partial of the tensor-core pass 1 ([b][3][Nc], CTA b owns chunk b % nch) -> fp64 moments [2][N] = (sum du,
// sum du zhat) of this replica
__global__ void __launch_bounds__(kFinThreads)
    gate_bwd_tc_moments(const float* __restrict__ col_partial, int nparts, int nch, int N,
                                    double* __restrict__ moments) {
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int Nc = N / nch, per_chunk = nparts / nch;
  const int col = (int)blockIdx.x * 32 + lx;
  const bool ok = col < N;
  const int c = (ok ? col : 0) / Nc, nl = (ok ? col : 0) % Nc;
  const float* cp = col_partial + (int64_t)c * 3 * Nc;
  __shared__ double s_acc[2][32][33];
  double a1 = 0.0, a2 = 0.0;
  if (ok)
    for (int p = ly; p < per_chunk; p += 32) {
      const float* row = cp + (int64_t)p * nch * 3 * Nc;
      a1 += (double)row[nl];
      a2 += (double)row[Nc + nl];
    }
  s_acc[0][ly][lx] = a1;
  s_acc[1][ly][lx] = a2;
  __syncthreads();
  if (ly != 0 || !ok) return;
  double sdu = 0.0, sdz = 0.0;
#pragma unroll
  for (int g = 0; g < 32; ++g) {
    sdu += s_acc[0][g][lx];
    sdz += s_acc[1][g][lx];
  }
  moments[col] = sdu;
  moments[N + col] = sdz;
}

// ---- backward: materialise dz (fp32 FFMA path) + db partials ------------------------------
__global__ void __launch_bounds__(kEwThreads)
    gate_bwd_dz_kernel(const float* __restrict__ dy, const float* __restrict__ s,
                       const float* __restrict__ z, int64_t M, int C4, const float* __restrict__ coefA,
                       const float* __restrict__ coefB, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ c1,
                       const float* __restrict__ c2, float* __restrict__ dz,
                       float* __restrict__ partial) {
  const EwMap m = ew_map(C4);
  float4 acc[1] = {make_float4(0, 0, 0, 0)};
  if (m.active) {
    const float4 A = ldc4(coefA, m.g), B = ldc4(coefB, m.g), mu = ldc4(mean, m.g), rs = ldc4(invstd, m.g);
    const float4 k1 = ldc4(c1, m.g), k2 = ldc4(c2, m.g);
    for (int64_t p = (int64_t)blockIdx.x * m.rows + m.r; p < M; p += (int64_t)gridDim.x * m.rows) {
      const int64_t i = p * C4 + m.g;
      const float4 g = ld4(dy, i), sv = ld4(s, i), zv = ld4(z, i);
      const float4 a = gate_act(zv, A, B);
      float4 d;
      d.x = A.x * (g.x * sv.x * a.x * (1.f - a.x) - k1.x - (zv.x - mu.x) * rs.x * k2.x);
      d.y = A.y * (g.y * sv.y * a.y * (1.f - a.y) - k1.y - (zv.y - mu.y) * rs.y * k2.y);
      d.z = A.z * (g.z * sv.z * a.z * (1.f - a.z) - k1.z - (zv.z - mu.z) * rs.z * k2.z);
      d.w = A.w * (g.w * sv.w * a.w * (1.f - a.w) - k1.w - (zv.w - mu.w) * rs.w * k2.w);
      st4(dz, i, d);
      acc[0].x += d.x; acc[0].y += d.y; acc[0].z += d.z; acc[0].w += d.w;
    }
  }
  ew_block_reduce<1>(acc, m, C4, partial + (int64_t)blockIdx.x * 4 * C4);
}

// out[j] = fixed-order fp64 sum over nparts rows of length `len`
__global__ void __launch_bounds__(kFinThreads)
    rows_sum_finalize(const float* __restrict__ partial, int nparts, int len,
                                  float* __restrict__ out) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const double s = block_colsum(partial, nparts, (int64_t)len, j < len ? j : 0, j < len);
  if (threadIdx.x < 32 && j < len) out[j] = (float)s;
}

// ---- fp32 CUDA-core contraction (VMTL_GATE_FP32_FFMA) --------------------------------------
// C[i][j] = sum_k A(i,k) * B(k,j) (+ bias[j]);  A(i,k) = A[i*lai + k*lak], B(k,j) = B[k*lbk + j*lbj].
// 64x64 tile, 16-deep k slices, 4x4 register tile per thread.  gridDim.z > 1: split-K, each
// z-slice writes its own [Mdim x ldc] partial at C + z*slice_stride.
__global__ void __launch_bounds__(256)
    sgemm64_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                   const float* __restrict__ bias, int64_t Mdim, int Ndim, int64_t Kdim, int64_t lai,
                   int64_t lak, int64_t lbk, int64_t lbj, int ldc, int64_t kchunk,
                   int64_t slice_stride) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = kbeg + kchunk < Kdim ? kbeg + kchunk : Kdim;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // loader mapping: make the unit-stride dimension run across consecutive threads
  const bool a_k_fast = (lak == 1);
  const bool b_j_fast = (lbj == 1);
  for (int64_t k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int lin = threadIdx.x + e * 256;  // 0..1023
      int ai, ak;
      if (a_k_fast) { ak = lin & 15; ai = lin >> 4; } else { ai = lin & 63; ak = lin >> 6; }
      const int64_t gi = i0 + ai, gk = k0 + ak;
      As[ak][ai] = (gi < Mdim && gk < kend) ? A[gi * lai + gk * lak] : 0.f;
      int bj, bk;
      if (b_j_fast) { bj = lin & 63; bk = lin >> 6; } else { bk = lin & 15; bj = lin >> 4; }
      const int gj = j0 + bj;
      const int64_t gk2 = k0 + bk;
      Bs[bk][bj] = (gj < Ndim && gk2 < kend) ? B[gk2 * lbk + (int64_t)gj * lbj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* Cz = C + (int64_t)blockIdx.z * slice_stride;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t gi = i0 + ty * 4 + a;
    if (gi >= Mdim) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + tx * 4 + b;
      if (gj < Ndim) Cz[gi * ldc + gj] = acc[a][b] + (bias ? bias[gj] : 0.f);
    }
  }
}

static int sgemm64(const float* A, const float* B, float* C, const float* bias, int64_t Mdim, int Ndim,
                   int64_t Kdim, int64_t lai, int64_t lak, int64_t lbk, int64_t lbj, int ldc, int splits,
                   int64_t slice_stride, cudaStream_t st) {
  dim3 grid((unsigned)((Mdim + 63) / 64), (unsigned)((Ndim + 63) / 64), (unsigned)splits);
  int64_t kchunk = (Kdim + splits - 1) / splits;
  kchunk = (kchunk + 15) / 16 * 16;
  sgemm64_kernel<<<grid, 256, 0, st>>>(A, B, C, bias, Mdim, Ndim, Kdim, lai, lak, lbk, lbj, ldc, kchunk,
                                       slice_stride);
  return launch_status();
}

static int gate_check(int64_t M, int K, int N, int precision) {
  if (M < 1 || K < 1 || N < 4) return VMTL_EINVAL;
  if (N % 4 != 0 || N > 4 * kEwThreads) return VMTL_EUNSUPPORTED;  // float4 channel groups, one block row
  if (precision != VMTL_GATE_FP32_FFMA && precision != VMTL_GATE_TC_3XTF32 && precision != VMTL_GATE_TC_TF32)
    return VMTL_EINVAL;
  return VMTL_OK;
}

}  // namespace vmtl

using namespace vmtl;

extern "C" size_t vmtl_gate_workspace_bytes(int64_t M, int K, int N, int precision, int backward) {
  if (gate_check(M, K, N, precision) != VMTL_OK) return 0;
  return gate_ws_floats(M, K, N, precision, backward, nullptr, nullptr) * sizeof(float) + 256;
}

extern "C" int vmtl_gate_tc_supported(int K, int N) { return gate_tc_supported(K, N) ? 1 : 0; }

// phase 0: the whole forward on local statistics.  Global-batch statistics: phase 1 = contraction
// (z, per-CTA partials) -> fp64 `moments` [2][N] = (sum z, sum z^2); phase 2 = finalize from the all-reduced moments
// over `Mstat` rows + the gate pass over (z, s).
static int gate_fwd_impl(const float* h, const float* h_coef, const float* s, const float* W, const float* bias,
                         const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, int training, int precision,
                         int64_t M, int K, int N, float* y, float* save_z, float* save_mean,
                         float* save_invstd, void* workspace, size_t workspace_bytes, void* stream, int phase,
                         double* moments, int64_t Mstat) {
  int rc = gate_check(M, K, N, precision);
  if (rc != VMTL_OK) return rc;
  if (phase != 0 && (!training || !moments)) return VMTL_EINVAL;
  if (phase == 2 && Mstat < 1) return VMTL_EINVAL;
  if ((phase != 2 && (!h || !W || !bias)) || (phase != 1 && (!s || !gamma || !beta || !y)) || !workspace)
    return VMTL_EINVAL;
  if (h_coef && (precision == VMTL_GATE_FP32_FFMA || !gate_tc_supported(K, N))) return VMTL_EUNSUPPORTED;
  if (training && !save_z) return VMTL_EINVAL;
  if (!training && (!running_mean || !running_var)) return VMTL_EINVAL;
  if ((h && !aligned16(h)) || (s && !aligned16(s)) || (W && !aligned16(W)) || (y && !aligned16(y)) ||
      !aligned16(workspace) || (save_z && !aligned16(save_z)))
    return VMTL_EALIGN;
  GateWs ws;
  const size_t need = gate_ws_floats(M, K, N, precision, 0, &ws, static_cast<float*>(workspace));
  if (workspace_bytes < need * sizeof(float)) return VMTL_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C4 = N / 4;
  const int split3 = precision == VMTL_GATE_TC_3XTF32;

  if (!training) {
    gate_fwd_stats_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(nullptr, 0, M, N, eps, momentum, 0, gamma,
                                                             beta, running_mean, running_var, ws.mean,
                                                             ws.invstd, ws.coefA, ws.coefB, save_mean,
                                                             save_invstd, nullptr);
    if ((rc = launch_status()) != VMTL_OK) return rc;
    const bool tc = precision != VMTL_GATE_FP32_FFMA && gate_tc_supported(K, N);
    if (tc && !save_z)  // inference: single fused pass, z never stored
      return gate_tc_fwd_eval(h, h_coef, s, W, bias, ws.coefA, ws.coefB, M, K, N, split3, y, st);
    float* zbuf = save_z ? save_z : ws.zbuf;  // the CUDA-core path stages z (workspace when not saved)
    if (!zbuf) return VMTL_EWORKSPACE;
    if (tc) {
      int unused = 0;  // a backward will follow: keep z (batch partials are computed but unused)
      rc = gate_tc_fwd_gemm(h, h_coef, W, bias, M, K, N, split3, zbuf, ws.partial, ws.partial_rows, &unused, st);
    } else {
      rc = sgemm64(h, W, zbuf, bias, M, N, K, K, 1, 1, K, N, 1, 0, st);
    }
    if (rc != VMTL_OK) return rc;
    gate_apply_kernel<<<ew_grid(M, N, blocks_per_sm(gate_apply_kernel, kEwThreads, 0, 8)), kEwThreads, 0, st>>>(zbuf, s, M, C4, ws.coefA, ws.coefB, y);
    return launch_status();
  }

  int nparts = 0;
  if (phase != 2) {
    if (precision != VMTL_GATE_FP32_FFMA && gate_tc_supported(K, N)) {
      rc = gate_tc_fwd_gemm(h, h_coef, W, bias, M, K, N, split3, save_z, ws.partial, ws.partial_rows, &nparts, st);
      if (rc != VMTL_OK) return rc;
    } else {
      rc = sgemm64(h, W, save_z, bias, M, N, K, K, 1, 1, K, N, 1, 0, st);
      if (rc != VMTL_OK) return rc;
      nparts = ew_grid(M, N);
      gate_colstats_kernel<<<nparts, kEwThreads, 0, st>>>(save_z, M, C4, ws.partial);
      if ((rc = launch_status()) != VMTL_OK) return rc;
    }
  }
  if (phase == 1) {
    gate_moments_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, nparts, N, moments);
    return launch_status();
  }
  gate_fwd_stats_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, nparts, phase == 2 ? Mstat : M, N, eps,
                                                           momentum, 1, gamma, beta, running_mean, running_var,
                                                           ws.mean, ws.invstd, ws.coefA, ws.coefB,
                                                           save_mean, save_invstd, phase == 2 ? moments : nullptr);
  if ((rc = launch_status()) != VMTL_OK) return rc;
  gate_apply_kernel<<<ew_grid(M, N, blocks_per_sm(gate_apply_kernel, kEwThreads, 0, 8)), kEwThreads, 0, st>>>(save_z, s, M, C4, ws.coefA, ws.coefB, y);
  return launch_status();
}

extern "C" int vmtl_gate_fwd(const float* h, const float* h_coef, const float* s, const float* W, const float* bias,
                             const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float momentum, float eps, int training, int precision,
                             int64_t M, int K, int N, float* y, float* save_z, float* save_mean,
                             float* save_invstd, void* workspace, size_t workspace_bytes, void* stream) {
  return gate_fwd_impl(h, h_coef, s, W, bias, gamma, beta, running_mean, running_var, momentum, eps, training,
                       precision, M, K, N, y, save_z, save_mean, save_invstd, workspace, workspace_bytes, stream, 0,
                       nullptr, 0);
}

extern "C" int vmtl_gate_fwd_moments(const float* h, const float* h_coef, const float* W, const float* bias,
                                     int precision, int64_t M, int K, int N, float* save_z, double* moments,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  return gate_fwd_impl(h, h_coef, nullptr, W, bias, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, 1, precision, M, K, N,
                       nullptr, save_z, nullptr, nullptr, workspace, workspace_bytes, stream, 1, moments, 0);
}

extern "C" int vmtl_gate_fwd_global(const float* s, const float* z, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, float momentum, float eps,
                                    int precision, int64_t M, int K, int N, const double* moments, int64_t M_global,
                                    float* y, float* save_mean, float* save_invstd, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return gate_fwd_impl(nullptr, nullptr, s, nullptr, nullptr, gamma, beta, running_mean, running_var, momentum, eps, 1,
                       precision, M, K, N, y, const_cast<float*>(z), save_mean, save_invstd, workspace,
                       workspace_bytes, stream, 2, const_cast<double*>(moments), M_global);
}

// coefA/B from saved statistics (backward re-derives them instead of trusting workspace reuse)
// phase 0: the whole backward.  Global-batch statistics: phase 1 = pass 1 (ds, this replica's sums and dW pieces,
// kept in the workspace) -> fp64 `moments` [2][N] = (sum du, sum du zhat); phase 2 = finalize with the all-reduced
// moments over `Mstat` rows (dW / dbias / dgamma / dbeta stay this replica's contributions) + the dh pass.  The
// workspace must be the same, untouched buffer in both phases.
static int gate_bwd_impl(const float* dy, const float* h, const float* h_coef, const float* s, const float* z,
                         const float* W, const float* gamma, const float* beta,
                         const float* save_mean, const float* save_invstd, int training,
                         int precision, int64_t M, int K, int N, float* dh, float* ds, float* dW,
                         float* dbias, float* dgamma, float* dbeta, void* workspace,
                         size_t workspace_bytes, void* stream, int phase, double* moments, int64_t Mstat) {
  int rc = gate_check(M, K, N, precision);
  if (rc != VMTL_OK) return rc;
  if (phase != 0 && (!training || !moments)) return VMTL_EINVAL;
  if (phase == 2 && Mstat < 1) return VMTL_EINVAL;
  if (!dy || !h || !s || !z || !W || !gamma || !beta || !save_mean || !save_invstd ||
      (phase != 1 && (!dW || !dbias || !dgamma || !dbeta)) || !workspace)
    return VMTL_EINVAL;
  if (h_coef && (precision == VMTL_GATE_FP32_FFMA || !gate_tc_supported(K, N))) return VMTL_EUNSUPPORTED;
  if (!aligned16(dy) || !aligned16(h) || !aligned16(s) || !aligned16(z) || !aligned16(W) ||
      !aligned16(workspace) || (dh && !aligned16(dh)) || (ds && !aligned16(ds)))
    return VMTL_EALIGN;
  GateWs ws;
  const size_t need = gate_ws_floats(M, K, N, precision, 1, &ws, static_cast<float*>(workspace));
  if (workspace_bytes < need * sizeof(float)) return VMTL_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C4 = N / 4;
  const int split3 = precision == VMTL_GATE_TC_3XTF32;

  if (precision != VMTL_GATE_FP32_FFMA && gate_tc_supported(K, N)) {
    // tensor-core path: pass 1 (ds + statistics + dW partials), finalize, pass 2 (dh)
    int np1 = gate_tc_bwd_pass1_grid(M, N);
    const int nch = N <= 64 ? 1 : N / 64;
    if (phase != 2) {
      rc = gate_tc_bwd_pass1(dy, h, h_coef, s, z, gamma, beta, save_mean, save_invstd, M, K, N, split3, ds, ws, &np1, st);
      if (rc != VMTL_OK) return rc;
    }
    if (phase == 1) {
      gate_bwd_tc_moments<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, np1, nch, N, moments);
      return launch_status();
    }
    if (N <= 64)
      gate_bwd_tc_finalize<1><<<N * (K / 32) + (N + 31) / 32, kFinThreads, 0, st>>>(
        ws.gemm_partial, ws.hs_partial, ws.partial, np1, nch, M, N, K, training, gamma, beta, save_mean, save_invstd, dW,
        dbias, dgamma, dbeta, ws.c1, ws.c2, ws.coefA, ws.coefB, ws.mean, ws.invstd, phase == 2 ? moments : nullptr,
        Mstat);
    else
      gate_bwd_tc_finalize<4><<<(N / 4) * (K / 32) + (N + 31) / 32, kFinThreads, 0, st>>>(
        ws.gemm_partial, ws.hs_partial, ws.partial, np1, nch, M, N, K, training, gamma, beta, save_mean, save_invstd, dW,
        dbias, dgamma, dbeta, ws.c1, ws.c2, ws.coefA, ws.coefB, ws.mean, ws.invstd, phase == 2 ? moments : nullptr,
        Mstat);
    if ((rc = launch_status()) != VMTL_OK) return rc;
    return dh ? gate_tc_bwd_dh(dy, s, z, W, ws, M, K, N, split3, dh, st) : VMTL_OK;
  }

  // CUDA-core path
  // phase A
  const int nparts = ew_grid(M, N, blocks_per_sm(gate_bwd_stats_kernel, kEwThreads, 0, 8));
  if (phase != 2) {
    gate_bwd_stats_kernel<<<nparts, kEwThreads, 0, st>>>(dy, s, z, M, C4, gamma, beta, save_mean, save_invstd,
                                                         ds, ws.partial);
    if ((rc = launch_status()) != VMTL_OK) return rc;
  }
  if (phase == 1) {
    gate_moments_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, nparts, N, moments);
    return launch_status();
  }
  gate_bwd_stats_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, nparts, M, N, training, dgamma,
                                                           dbeta, ws.c1, ws.c2, gamma, beta, save_mean,
                                                           save_invstd, ws.coefA, ws.coefB, ws.mean, ws.invstd,
                                                           phase == 2 ? moments : nullptr, Mstat);
  if ((rc = launch_status()) != VMTL_OK) return rc;

  // phase B
  gate_bwd_dz_kernel<<<nparts, kEwThreads, 0, st>>>(dy, s, z, M, C4, ws.coefA, ws.coefB, ws.mean,
                                                    ws.invstd, ws.c1, ws.c2, ws.dz, ws.partial);
  if ((rc = launch_status()) != VMTL_OK) return rc;
  rows_sum_finalize<<<(N + 31) / 32, kFinThreads, 0, st>>>(ws.partial, nparts, N, dbias);
  if ((rc = launch_status()) != VMTL_OK) return rc;
  if (dh) {
    // dh[M,K] = dz[M,N] @ W[N,K]
    rc = sgemm64(ws.dz, W, dh, nullptr, M, K, N, N, 1, K, 1, K, 1, 0, st);
    if (rc != VMTL_OK) return rc;
  }
  // dW[N,K] = dz^T[N,M] @ h[M,K]   (split-K over M, fixed-order second stage)
  int splits = ws.gemm_slots;
  const int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int per = (sm_count() * 2) / tiles;
  if (per < 1) per = 1;
  if (splits > per) splits = per;
  rc = sgemm64(ws.dz, h, ws.gemm_partial, nullptr, N, K, M, 1, N, K, 1, K, splits, (int64_t)N * K, st);
  if (rc != VMTL_OK) return rc;
  rows_sum_finalize<<<(N * K + 31) / 32, kFinThreads, 0, st>>>(ws.gemm_partial, splits, N * K, dW);
  return launch_status();
}

extern "C" int vmtl_gate_bwd(const float* dy, const float* h, const float* h_coef, const float* s, const float* z,
                             const float* W, const float* gamma, const float* beta,
                             const float* save_mean, const float* save_invstd, int training,
                             int precision, int64_t M, int K, int N, float* dh, float* ds, float* dW,
                             float* dbias, float* dgamma, float* dbeta, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return gate_bwd_impl(dy, h, h_coef, s, z, W, gamma, beta, save_mean, save_invstd, training, precision, M, K, N, dh, ds,
                       dW, dbias, dgamma, dbeta, workspace, workspace_bytes, stream, 0, nullptr, 0);
}

extern "C" int vmtl_gate_bwd_moments(const float* dy, const float* h, const float* h_coef, const float* s,
                                     const float* z, const float* W, const float* gamma, const float* beta,
                                     const float* save_mean, const float* save_invstd, int precision, int64_t M,
                                     int K, int N, float* ds, double* moments, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return gate_bwd_impl(dy, h, h_coef, s, z, W, gamma, beta, save_mean, save_invstd, 1, precision, M, K, N, nullptr, ds,
                       nullptr, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream, 1, moments, 0);
}

extern "C" int vmtl_gate_bwd_global(const float* dy, const float* h, const float* h_coef, const float* s,
                                    const float* z, const float* W, const float* gamma, const float* beta,
                                    const float* save_mean, const float* save_invstd, int precision, int64_t M,
                                    int K, int N, const double* moments, int64_t M_global, float* dh, float* dW,
                                    float* dbias, float* dgamma, float* dbeta, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return gate_bwd_impl(dy, h, h_coef, s, z, W, gamma, beta, save_mean, save_invstd, 1, precision, M, K, N, dh, nullptr,
                       dW, dbias, dgamma, dbeta, workspace, workspace_bytes, stream, 2, const_cast<double*>(moments),
                       M_global);
}

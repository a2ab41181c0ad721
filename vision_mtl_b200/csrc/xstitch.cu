// Cross-stitch unit, forward and backward, NHWC float4 streaming kernels.
//
// Replaces torch.stack + torch.einsum("aa(c),abcij->abcij") and its autograd backward
// (reference: vision_mtl/models/cross_stitch_model.py:32-37, call site :143-156).
//
// Data layout: T task feature maps, each a row-major [npix, C] fp32 matrix (NHWC), read
// through T separate pointers -- the [T,B,C,H,W] stack copy of the reference never exists.
// Both kernels are pure HBM streams:
//   fwd : 2*4*T*npix*C bytes (read x, write y)
//   bwd : 3*4*T*npix*C bytes (read dy and x, write dx) + the alpha-gradient reduction,
//         which is a per-thread register accumulation (each thread owns one float4 channel
//         group for its whole life), a shared-memory / warp-shuffle block reduction and a
//         fixed-order fp64 second stage -> deterministic.
#include "vmtl_common.cuh"

namespace vmtl {

struct XsFwdArgs {
  const float4* x[VMTL_MAX_TASKS];
  float4* y[VMTL_MAX_TASKS];
};
struct XsBwdArgs {
  const float4* dy[VMTL_MAX_TASKS];
  const float4* x[VMTL_MAX_TASKS];
  float4* dx[VMTL_MAX_TASKS];
};

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_mul(const float4& a, const float4& b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ void f4_fma(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z);
  acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ float4 f4_splat(float v) { return make_float4(v, v, v, v); }

// ------------------------------------------------------------------ forward
template <int T, bool DIAG, bool CW>
__global__ void __launch_bounds__(256)
    xstitch_fwd_kernel(XsFwdArgs a, const float* __restrict__ alpha, int64_t n4, int C4) {
  extern __shared__ float4 s_alpha[];  // CW: [T*T][C4]
  float w[T][T];
  if (CW) {
    const float4* al4 = reinterpret_cast<const float4*>(alpha);
    for (int i = threadIdx.x; i < T * T * C4; i += blockDim.x) s_alpha[i] = al4[i];
    __syncthreads();
  } else {
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
      for (int j = 0; j < T; ++j) w[i][j] = alpha[i * T + j];
  }
  constexpr int U = (T <= 2) ? 4 : (T <= 4 ? 2 : 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int c = (int)(i % C4);
  const int cstep = (int)(stride % C4);

  auto mix = [&](const float4 (&xin)[T], int cg, int64_t idx) {
#pragma unroll
    for (int o = 0; o < T; ++o) {
      float4 r;
      if (DIAG) {
        r = CW ? f4_mul(s_alpha[(o * T + o) * C4 + cg], xin[o])
               : f4_mul(f4_splat(w[o][o]), xin[o]);
      } else {
        r = f4_zero();
#pragma unroll
        for (int b = 0; b < T; ++b) {
          if (CW)
            f4_fma(r, s_alpha[(o * T + b) * C4 + cg], xin[b]);
          else
            f4_fma(r, f4_splat(w[o][b]), xin[b]);
        }
      }
      stg_stream(a.y[o] + idx, r);
    }
  };

  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 xin[U][T];
    int cg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int t = 0; t < T; ++t) xin[u][t] = ldg_stream(a.x[t] + i + u * stride);
      cg[u] = c;
      c += cstep;
      if (c >= C4) c -= C4;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) mix(xin[u], cg[u], i + u * stride);
  }
  for (; i < n4; i += stride) {
    float4 xin[T];
#pragma unroll
    for (int t = 0; t < T; ++t) xin[t] = ldg_stream(a.x[t] + i);
    mix(xin, c, i);
    c += cstep;
    if (c >= C4) c -= C4;
  }
}

// ------------------------------------------------------------------ backward
// Thread layout: threadIdx.x = r*cgw + gl ; gl = float4 channel group inside this block's
// channel slice (blockIdx.y selects the slice), r = pixel row inside one pass.  One pass
// covers `rows` consecutive pixels -> rows*cgw consecutive float4 when the slice is the
// whole channel range, i.e. fully coalesced 128-bit accesses.
template <int T, bool DIAG, bool CW>
__global__ void __launch_bounds__(512)
    xstitch_bwd_kernel(XsBwdArgs a, const float* __restrict__ alpha, float* __restrict__ partial,
                       int64_t npix, int C4, int cgw, int rows, int write_dx) {
  constexpr int NACC = DIAG ? T : T * T;
  extern __shared__ float4 s_red[];
  const int t = threadIdx.x;
  const int r = t / cgw;
  const int gl = t - r * cgw;
  const int g = blockIdx.y * cgw + gl;
  const bool active = (r < rows) && (g < C4);

  float4 w[T][T];
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) {
      if (CW) {
        w[i][j] = active ? reinterpret_cast<const float4*>(alpha)[(i * T + j) * C4 + g] : f4_zero();
      } else {
        w[i][j] = f4_splat(alpha[i * T + j]);
      }
    }

  float4 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = f4_zero();

  if (active) {
    const int64_t pstep = (int64_t)gridDim.x * rows;
    for (int64_t p = (int64_t)blockIdx.x * rows + r; p < npix; p += pstep) {
      const int64_t idx = p * C4 + g;
      float4 dy[T], xv[T];
#pragma unroll
      for (int i = 0; i < T; ++i) dy[i] = ldg_stream(a.dy[i] + idx);
#pragma unroll
      for (int i = 0; i < T; ++i) xv[i] = ldg_stream(a.x[i] + idx);
      if (write_dx) {
#pragma unroll
        for (int b = 0; b < T; ++b) {
          float4 d;
          if (DIAG) {
            d = f4_mul(w[b][b], dy[b]);
          } else {
            d = f4_zero();
#pragma unroll
            for (int o = 0; o < T; ++o) f4_fma(d, w[o][b], dy[o]);
          }
          stg_stream(a.dx[b] + idx, d);
        }
      }
#pragma unroll
      for (int o = 0; o < T; ++o) {
        if (DIAG) {
          f4_fma(acc[o], dy[o], xv[o]);
        } else {
#pragma unroll
          for (int b = 0; b < T; ++b) f4_fma(acc[o * T + b], dy[o], xv[b]);
        }
      }
    }
  }

  if (CW) {
    // block reduction over the `rows` threads that share a channel group
    if (r < rows) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) s_red[(r * cgw + gl) * NACC + i] = acc[i];
    }
    __syncthreads();
    if (r == 0 && g < C4) {
      float4* out = reinterpret_cast<float4*>(partial) + (int64_t)blockIdx.x * NACC * C4;
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float4 s = s_red[gl * NACC + i];
        for (int rr = 1; rr < rows; ++rr) {
          const float4 v = s_red[(rr * cgw + gl) * NACC + i];
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        out[i * C4 + g] = s;
      }
    }
  } else {
    // layer-wise alpha: every lane contributes to the same T*T scalars -> warp shuffles
    float* s_f = reinterpret_cast<float*>(s_red);
    const int warp = t >> 5, lane = t & 31, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float v = acc[i].x + acc[i].y + acc[i].z + acc[i].w;
      v = warp_sum(v);
      if (lane == 0) s_f[warp * NACC + i] = v;
    }
    __syncthreads();
    if (t < NACC) {
      float s = 0.f;
      for (int wi = 0; wi < nwarp; ++wi) s += s_f[wi * NACC + t];
      partial[(int64_t)blockIdx.x * NACC + t] = s;
    }
  }
}

// second stage: fixed-order fp64 sum over the per-block partials
template <bool DIAG>
__global__ void xstitch_dalpha_finalize(const float* __restrict__ partial, float* __restrict__ dalpha,
                                        int nblocks, int T, int Cdim /* C or 1 */) {
  const int n = T * T * Cdim;
  const int lx = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lx;
  const int nacc = DIAG ? T : T * T;
  int src = -1;
  if (j < n) {
    const int c = j % Cdim;
    const int ab = j / Cdim;
    const int o = ab / T, b = ab % T;
    if (DIAG)
      src = (o == b) ? o * Cdim + c : -1;  // off-diagonal: exact zeros, as autograd gives for the reference einsum
    else
      src = ab * Cdim + c;
  }
  const double t = block_colsum(partial, nblocks, (int64_t)nacc * Cdim, src < 0 ? 0 : src, src >= 0);
  if (threadIdx.x < 32 && j < n) dalpha[j] = (float)t;
}

struct XsBwdPlan {
  int cgw, rows, nchunk, threads, gridx;
  int64_t npix_eff;
  int c4_eff;
  size_t smem, partial_bytes;
};

static XsBwdPlan xs_bwd_plan(int T, int64_t npix, int C, int channel_wise, int mode) {
  XsBwdPlan p;
  const int nacc_max = T * T;
  (void)mode;
  if (channel_wise) {
    p.c4_eff = C / 4;
    p.npix_eff = npix;
    const int target = 256;
    if (p.c4_eff <= 512) {
      p.cgw = p.c4_eff;
      p.nchunk = 1;
    } else {
      p.nchunk = (p.c4_eff + 255) / 256;
      p.cgw = (p.c4_eff + p.nchunk - 1) / p.nchunk;
    }
    p.rows = p.cgw >= target ? 1 : target / p.cgw;
  } else {
    // layer-wise alpha: the channel position is irrelevant, stream the flat float4 array
    p.c4_eff = 1;
    p.npix_eff = npix * (C / 4);
    p.cgw = 1;
    p.nchunk = 1;
    p.rows = 256;
  }
  p.threads = ((p.rows * p.cgw + 31) / 32) * 32;
  int64_t want = (p.npix_eff + p.rows - 1) / p.rows;
  int64_t cap = (int64_t)sm_count() * (p.threads <= 256 ? 4 : 2);
  if (T >= 3) cap = (int64_t)sm_count() * 2;
  p.gridx = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  p.smem = channel_wise ? (size_t)p.rows * p.cgw * nacc_max * sizeof(float4)
                        : (size_t)(p.threads / 32) * nacc_max * sizeof(float);
  p.partial_bytes = (size_t)p.gridx * nacc_max * (channel_wise ? C : 1) * sizeof(float);
  return p;
}

template <int T>
static int xs_fwd_launch(const XsFwdArgs& a, const float* alpha, int64_t n4, int C4, int cw,
                         int mode, cudaStream_t st) {
  const int threads = 256;
  constexpr int U = (T <= 2) ? 4 : (T <= 4 ? 2 : 1);
  int64_t want = (n4 + (int64_t)threads * U - 1) / ((int64_t)threads * U);
  size_t smem = cw ? (size_t)T * T * C4 * sizeof(float4) : 0;
  if (smem > 160 * 1024) return VMTL_EUNSUPPORTED;
#define VMTL_XS_FWD(DIAG, CWB)                                                                   \
  do {                                                                                           \
    if (smem > 48 * 1024)                                                                        \
      cudaFuncSetAttribute(xstitch_fwd_kernel<T, DIAG, CWB>,                                     \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    const int64_t cap =                                                                          \
        (int64_t)sm_count() * blocks_per_sm(xstitch_fwd_kernel<T, DIAG, CWB>, threads, smem, 8); \
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);                            \
    xstitch_fwd_kernel<T, DIAG, CWB><<<grid, threads, smem, st>>>(a, alpha, n4, C4);             \
  } while (0)
  const bool diag = (mode == VMTL_XS_REFERENCE_DIAG);
  if (diag && cw) VMTL_XS_FWD(true, true);
  else if (diag && !cw) VMTL_XS_FWD(true, false);
  else if (!diag && cw) VMTL_XS_FWD(false, true);
  else VMTL_XS_FWD(false, false);
#undef VMTL_XS_FWD
  return launch_status();
}

template <int T>
static int xs_bwd_launch(const XsBwdArgs& a, const float* alpha, float* dalpha, int64_t npix, int C,
                         int cw, int mode, float* partial, const XsBwdPlan& p, int write_dx,
                         cudaStream_t st) {
  dim3 grid(p.gridx, p.nchunk);
#define VMTL_XS_BWD(DIAG, CWB)                                                                   \
  do {                                                                                           \
    if (p.smem > 48 * 1024)                                                                      \
      cudaFuncSetAttribute(xstitch_bwd_kernel<T, DIAG, CWB>,                                     \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);            \
    xstitch_bwd_kernel<T, DIAG, CWB><<<grid, p.threads, p.smem, st>>>(                           \
        a, alpha, partial, p.npix_eff, p.c4_eff, p.cgw, p.rows, write_dx);                       \
  } while (0)
  const bool diag = (mode == VMTL_XS_REFERENCE_DIAG);
  if (diag && cw) VMTL_XS_BWD(true, true);
  else if (diag && !cw) VMTL_XS_BWD(true, false);
  else if (!diag && cw) VMTL_XS_BWD(false, true);
  else VMTL_XS_BWD(false, false);
#undef VMTL_XS_BWD
  int rc = launch_status();
  if (rc != VMTL_OK) return rc;
  const int cdim = cw ? C : 1;
  const int n = T * T * cdim;
  if (diag)
    xstitch_dalpha_finalize<true><<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, dalpha, p.gridx, T, cdim);
  else
    xstitch_dalpha_finalize<false><<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, dalpha, p.gridx, T, cdim);
  return launch_status();
}

// ------------------------------------------------------------------ pad + cat + stitch / upsample + stitch
// CSNet's decoder sites (vision_mtl/models/cross_stitch_model.py:121-156, utils/model_utils.py:46-58): before the
// stitch, every task's feature map is either `cat([skip, zero_pad(x)])` (decoder blocks 0..3) or the nearest x2
// up-sampling of x (last block).  The reference materialises that tensor (F.pad, torch.cat / interpolate) and then
// stacks and mixes it; here the stitch kernel GATHERS its operand from (skip, x) and writes the mixed map straight
// into the next convolution's NHWC input -- the concatenated / up-sampled tensor never exists.
//   PAD : domain = output pixels p of [B,Ho,Wo]; channel groups [0,Cs4) come from skip[p], [Cs4,C4) from
//         x[b, ho-py, wo-px] (zero outside), py = (Ho-Hi)/2, px = (Wo-Wi)/2 (the centred F.pad of the reference)
//   UP2 : domain = INPUT pixels q of [B,Hi,Wi] (no skip); each produces / collects its 2x2 output pixels
// Thread layout as in xstitch_bwd_kernel: a thread owns one channel group for life.
struct XsCatGeom {
  int Ho, Wo, Hi, Wi, Cs4, Cx4, py, px;
};
struct XsCatFwdArgs {
  const float4* skip[VMTL_MAX_TASKS];
  const float4* x[VMTL_MAX_TASKS];
  float4* y[VMTL_MAX_TASKS];
};
struct XsCatBwdArgs {
  const float4* dy[VMTL_MAX_TASKS];
  const float4* skip[VMTL_MAX_TASKS];
  const float4* x[VMTL_MAX_TASKS];
  float4* dskip[VMTL_MAX_TASKS];
  float4* dx[VMTL_MAX_TASKS];
};

// index of x's float4 for output pixel p, channel group gx (PAD mode); -1 outside the padded frame
__device__ __forceinline__ int64_t xs_pad_src(const XsCatGeom& gm, uint32_t p, int gx) {
  const uint32_t wo = p % (uint32_t)gm.Wo, t = p / (uint32_t)gm.Wo;
  const uint32_t ho = t % (uint32_t)gm.Ho, b = t / (uint32_t)gm.Ho;
  const int hi = (int)ho - gm.py, wi = (int)wo - gm.px;
  if (hi < 0 || hi >= gm.Hi || wi < 0 || wi >= gm.Wi) return -1;
  return (((int64_t)b * gm.Hi + hi) * gm.Wi + wi) * gm.Cx4 + gx;
}
// first of the four output float4 indices of input pixel q, channel group g (UP2 mode; C4 = Cx4)
__device__ __forceinline__ int64_t xs_up2_dst(const XsCatGeom& gm, uint32_t q, int g) {
  const uint32_t wi = q % (uint32_t)gm.Wi, t = q / (uint32_t)gm.Wi;
  const uint32_t hi = t % (uint32_t)gm.Hi, b = t / (uint32_t)gm.Hi;
  return (((int64_t)b * gm.Ho + 2 * hi) * gm.Wo + 2 * wi) * gm.Cx4 + g;
}

template <int T, bool DIAG, bool CW>
__device__ __forceinline__ void xs_load_alpha(float4 (&w)[T][T], const float* __restrict__ alpha, int C4, int g,
                                              bool active) {
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) {
      if (CW)
        w[i][j] = active ? reinterpret_cast<const float4*>(alpha)[(i * T + j) * C4 + g] : f4_zero();
      else
        w[i][j] = f4_splat(alpha[i * T + j]);
    }
}

template <int T, bool DIAG, bool CW, bool UP2>
__global__ void __launch_bounds__(512)
    xstitch_cat_fwd_kernel(XsCatFwdArgs a, const float* __restrict__ alpha, XsCatGeom gm, int64_t ndom, int C4,
                           int cgw, int rows) {
  const int t = threadIdx.x;
  const int r = t / cgw;
  const int gl = t - r * cgw;
  const int g = blockIdx.y * cgw + gl;
  if (!((r < rows) && (g < C4))) return;
  float4 w[T][T];
  xs_load_alpha<T, DIAG, CW>(w, alpha, C4, g, true);
  const int64_t pstep = (int64_t)gridDim.x * rows;
  for (int64_t p = (int64_t)blockIdx.x * rows + r; p < ndom; p += pstep) {
    float4 xin[T];
    if (UP2) {
#pragma unroll
      for (int i = 0; i < T; ++i) xin[i] = ldg_stream(a.x[i] + p * C4 + g);
    } else if (g < gm.Cs4) {
#pragma unroll
      for (int i = 0; i < T; ++i) xin[i] = ldg_stream(a.skip[i] + p * gm.Cs4 + g);
    } else {
      const int64_t src = xs_pad_src(gm, (uint32_t)p, g - gm.Cs4);
#pragma unroll
      for (int i = 0; i < T; ++i) xin[i] = src >= 0 ? ldg_stream(a.x[i] + src) : f4_zero();
    }
    const int64_t dst = UP2 ? xs_up2_dst(gm, (uint32_t)p, g) : p * C4 + g;
#pragma unroll
    for (int o = 0; o < T; ++o) {
      float4 v;
      if (DIAG) {
        v = f4_mul(w[o][o], xin[o]);
      } else {
        v = f4_zero();
#pragma unroll
        for (int b = 0; b < T; ++b) f4_fma(v, w[o][b], xin[b]);
      }
      stg_stream(a.y[o] + dst, v);
      if (UP2) {
        stg_stream(a.y[o] + dst + C4, v);
        stg_stream(a.y[o] + dst + (int64_t)gm.Wo * C4, v);
        stg_stream(a.y[o] + dst + (int64_t)gm.Wo * C4 + C4, v);
      }
    }
  }
}

template <int T, bool DIAG, bool CW, bool UP2>
__global__ void __launch_bounds__(512)
    xstitch_cat_bwd_kernel(XsCatBwdArgs a, const float* __restrict__ alpha, float* __restrict__ partial,
                           XsCatGeom gm, int64_t ndom, int C4, int cgw, int rows) {
  constexpr int NACC = DIAG ? T : T * T;
  extern __shared__ float4 s_red[];
  const int t = threadIdx.x;
  const int r = t / cgw;
  const int gl = t - r * cgw;
  const int g = blockIdx.y * cgw + gl;
  const bool active = (r < rows) && (g < C4);
  float4 w[T][T];
  xs_load_alpha<T, DIAG, CW>(w, alpha, C4, g, active);
  float4 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = f4_zero();
  if (active) {
    const int64_t pstep = (int64_t)gridDim.x * rows;
    for (int64_t p = (int64_t)blockIdx.x * rows + r; p < ndom; p += pstep) {
      float4 dy[T], xv[T];
      int64_t src = -1;  // where this thread's input gradient goes (PAD: skip or x; -1 = the zero padding)
      if (UP2) {
        const int64_t d0 = xs_up2_dst(gm, (uint32_t)p, g);
#pragma unroll
        for (int i = 0; i < T; ++i) {
          const float4 q0 = ldg_stream(a.dy[i] + d0), q1 = ldg_stream(a.dy[i] + d0 + C4);
          const float4 q2 = ldg_stream(a.dy[i] + d0 + (int64_t)gm.Wo * C4), q3 = ldg_stream(a.dy[i] + d0 + (int64_t)gm.Wo * C4 + C4);
          dy[i] = make_float4((q0.x + q1.x) + (q2.x + q3.x), (q0.y + q1.y) + (q2.y + q3.y),
                              (q0.z + q1.z) + (q2.z + q3.z), (q0.w + q1.w) + (q2.w + q3.w));
          xv[i] = ldg_stream(a.x[i] + p * C4 + g);
        }
        src = p * C4 + g;
      } else {
#pragma unroll
        for (int i = 0; i < T; ++i) dy[i] = ldg_stream(a.dy[i] + p * C4 + g);
        if (g < gm.Cs4) {
          src = p * gm.Cs4 + g;
#pragma unroll
          for (int i = 0; i < T; ++i) xv[i] = ldg_stream(a.skip[i] + src);
        } else {
          src = xs_pad_src(gm, (uint32_t)p, g - gm.Cs4);
#pragma unroll
          for (int i = 0; i < T; ++i) xv[i] = src >= 0 ? ldg_stream(a.x[i] + src) : f4_zero();
        }
      }
      if (src >= 0) {
        const bool to_skip = !UP2 && g < gm.Cs4;
#pragma unroll
        for (int b = 0; b < T; ++b) {
          float4 d;
          if (DIAG) {
            d = f4_mul(w[b][b], dy[b]);
          } else {
            d = f4_zero();
#pragma unroll
            for (int o = 0; o < T; ++o) f4_fma(d, w[o][b], dy[o]);
          }
          float4* dst = to_skip ? a.dskip[b] : a.dx[b];
          if (dst) stg_stream(dst + src, d);
        }
      }
#pragma unroll
      for (int o = 0; o < T; ++o) {
        if (DIAG) {
          f4_fma(acc[o], dy[o], xv[o]);
        } else {
#pragma unroll
          for (int b = 0; b < T; ++b) f4_fma(acc[o * T + b], dy[o], xv[b]);
        }
      }
    }
  }
  if (CW) {
    if (r < rows) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) s_red[(r * cgw + gl) * NACC + i] = acc[i];
    }
    __syncthreads();
    if (r == 0 && g < C4) {
      float4* out = reinterpret_cast<float4*>(partial) + (int64_t)blockIdx.x * NACC * C4;
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float4 s = s_red[gl * NACC + i];
        for (int rr = 1; rr < rows; ++rr) {
          const float4 v = s_red[(rr * cgw + gl) * NACC + i];
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        out[i * C4 + g] = s;
      }
    }
  } else {
    float* s_f = reinterpret_cast<float*>(s_red);
    const int warp = t >> 5, lane = t & 31, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float v = acc[i].x + acc[i].y + acc[i].z + acc[i].w;
      v = warp_sum(v);
      if (lane == 0) s_f[warp * NACC + i] = v;
    }
    __syncthreads();
    if (t < NACC) {
      float s = 0.f;
      for (int wi = 0; wi < nwarp; ++wi) s += s_f[wi * NACC + t];
      partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NACC + t] = s;
    }
  }
}

// thread-layout plan of the gather kernels: channel groups stay channel groups (the source depends on them), so the
// layer-wise case cannot use the flat view of xs_bwd_plan
struct XsCatPlan {
  int cgw, rows, nchunk, threads, gridx;
  size_t smem, partial_bytes;
};
static XsCatPlan xs_cat_plan(int T, int64_t ndom, int C4, int channel_wise) {
  XsCatPlan p;
  if (C4 <= 512) {
    p.cgw = C4;
    p.nchunk = 1;
  } else {
    p.nchunk = (C4 + 255) / 256;
    p.cgw = (C4 + p.nchunk - 1) / p.nchunk;
  }
  p.rows = p.cgw >= 256 ? 1 : 256 / p.cgw;
  p.threads = ((p.rows * p.cgw + 31) / 32) * 32;
  int64_t want = (ndom + p.rows - 1) / p.rows;
  int64_t cap = (int64_t)sm_count() * (p.threads <= 256 ? 4 : 2);
  p.gridx = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  p.smem = channel_wise ? (size_t)p.rows * p.cgw * T * T * sizeof(float4) : (size_t)(p.threads / 32) * T * T * sizeof(float);
  // CW: [gridx][NACC][C] floats; layer-wise: [nchunk * gridx][NACC]
  p.partial_bytes = channel_wise ? (size_t)p.gridx * T * T * C4 * 4 * sizeof(float)
                                 : (size_t)p.gridx * p.nchunk * T * T * sizeof(float);
  return p;
}

static int xs_cat_geom(int B, int Ho, int Wo, int Cs, int Hi, int Wi, int Cx, int up2, XsCatGeom* gm, int64_t* ndom,
                       int* C4) {
  if (B < 1 || Ho < 1 || Wo < 1 || Hi < 1 || Wi < 1 || Cs < 0 || Cx < 4) return VMTL_EINVAL;
  if (Cs % 4 != 0 || Cx % 4 != 0) return VMTL_EALIGN;
  if (up2) {
    if (Cs != 0 || Ho != 2 * Hi || Wo != 2 * Wi) return VMTL_EUNSUPPORTED;
  } else if (Ho < Hi || Wo < Wi) {
    return VMTL_EUNSUPPORTED;  // the reference's F.pad would crop here; CSNet never does
  }
  gm->Ho = Ho; gm->Wo = Wo; gm->Hi = Hi; gm->Wi = Wi; gm->Cs4 = Cs / 4; gm->Cx4 = Cx / 4;
  gm->py = (Ho - Hi) / 2; gm->px = (Wo - Wi) / 2;
  *ndom = up2 ? (int64_t)B * Hi * Wi : (int64_t)B * Ho * Wo;
  *C4 = (Cs + Cx) / 4;
  if ((int64_t)B * Ho * Wo >= (1ll << 31)) return VMTL_EUNSUPPORTED;  // 32-bit pixel arithmetic
  return VMTL_OK;
}

static int xs_check(int T, int64_t npix, int C, int mode) {
  if (T < 1 || T > VMTL_MAX_TASKS || npix < 0 || C < 4) return VMTL_EINVAL;
  if (C % 4 != 0) return VMTL_EALIGN;
  if (mode != VMTL_XS_REFERENCE_DIAG && mode != VMTL_XS_FULL_MIX) return VMTL_EINVAL;
  return VMTL_OK;
}

}  // namespace vmtl

using namespace vmtl;

extern "C" int vmtl_xstitch_fwd(const float* const* x_host, float* const* y_host, const float* alpha,
                                int T, int64_t npix, int C, int channel_wise, int mode,
                                void* stream) {
  int rc = xs_check(T, npix, C, mode);
  if (rc != VMTL_OK) return rc;
  if (!x_host || !y_host || !alpha) return VMTL_EINVAL;
  XsFwdArgs a{};
  for (int t = 0; t < T; ++t) {
    if (!x_host[t] || !y_host[t]) return VMTL_EINVAL;
    if (!aligned16(x_host[t]) || !aligned16(y_host[t])) return VMTL_EALIGN;
    a.x[t] = reinterpret_cast<const float4*>(x_host[t]);
    a.y[t] = reinterpret_cast<float4*>(y_host[t]);
  }
  if (channel_wise && !aligned16(alpha)) return VMTL_EALIGN;
  if (npix == 0) return VMTL_OK;
  const int C4 = C / 4;
  const int64_t n4 = npix * C4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (T) {
    case 1: return xs_fwd_launch<1>(a, alpha, n4, C4, channel_wise, mode, st);
    case 2: return xs_fwd_launch<2>(a, alpha, n4, C4, channel_wise, mode, st);
    case 3: return xs_fwd_launch<3>(a, alpha, n4, C4, channel_wise, mode, st);
    case 4: return xs_fwd_launch<4>(a, alpha, n4, C4, channel_wise, mode, st);
    default: return VMTL_EUNSUPPORTED;
  }
}

extern "C" size_t vmtl_xstitch_bwd_workspace_bytes(int T, int64_t npix, int C, int channel_wise) {
  if (xs_check(T, npix, C, VMTL_XS_FULL_MIX) != VMTL_OK) return 0;
  XsBwdPlan p = xs_bwd_plan(T, npix < 1 ? 1 : npix, C, channel_wise, VMTL_XS_FULL_MIX);
  return p.partial_bytes + 256;
}

extern "C" int vmtl_xstitch_bwd(const float* const* dy_host, const float* const* x_host,
                                float* const* dx_host, const float* alpha, float* dalpha, int T,
                                int64_t npix, int C, int channel_wise, int mode, void* workspace,
                                size_t workspace_bytes, void* stream) {
  int rc = xs_check(T, npix, C, mode);
  if (rc != VMTL_OK) return rc;
  if (!dy_host || !x_host || !alpha || !dalpha || !workspace) return VMTL_EINVAL;
  XsBwdArgs a{};
  for (int t = 0; t < T; ++t) {
    if (!dy_host[t] || !x_host[t]) return VMTL_EINVAL;
    if (!aligned16(dy_host[t]) || !aligned16(x_host[t])) return VMTL_EALIGN;
    a.dy[t] = reinterpret_cast<const float4*>(dy_host[t]);
    a.x[t] = reinterpret_cast<const float4*>(x_host[t]);
    if (dx_host) {
      if (!dx_host[t]) return VMTL_EINVAL;
      if (!aligned16(dx_host[t])) return VMTL_EALIGN;
      a.dx[t] = reinterpret_cast<float4*>(dx_host[t]);
    }
  }
  if ((channel_wise && !aligned16(alpha)) || !aligned16(workspace)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (npix == 0) {
    const size_t n = (size_t)T * T * (channel_wise ? C : 1);
    return cudaMemsetAsync(dalpha, 0, n * sizeof(float), st) == cudaSuccess ? VMTL_OK : VMTL_ECUDA;
  }
  XsBwdPlan p = xs_bwd_plan(T, npix, C, channel_wise, mode);
  if (workspace_bytes < p.partial_bytes) return VMTL_EWORKSPACE;
  if (p.smem > 200 * 1024) return VMTL_EUNSUPPORTED;
  float* partial = static_cast<float*>(workspace);
  const int write_dx = dx_host != nullptr;
  switch (T) {
    case 1: return xs_bwd_launch<1>(a, alpha, dalpha, npix, C, channel_wise, mode, partial, p, write_dx, st);
    case 2: return xs_bwd_launch<2>(a, alpha, dalpha, npix, C, channel_wise, mode, partial, p, write_dx, st);
    case 3: return xs_bwd_launch<3>(a, alpha, dalpha, npix, C, channel_wise, mode, partial, p, write_dx, st);
    case 4: return xs_bwd_launch<4>(a, alpha, dalpha, npix, C, channel_wise, mode, partial, p, write_dx, st);
    default: return VMTL_EUNSUPPORTED;
  }
}

// ---- pad + cat + stitch / upsample + stitch ---------------------------------------------------------------
extern "C" int vmtl_xstitch_cat_fwd(const float* const* skip_host, const float* const* x_host, float* const* y_host,
                                    const float* alpha, int T, int B, int Ho, int Wo, int Cs, int Hi, int Wi, int Cx,
                                    int up2, int channel_wise, int mode, void* stream) {
  XsCatGeom gm;
  int64_t ndom;
  int C4;
  int rc = xs_cat_geom(B, Ho, Wo, Cs, Hi, Wi, Cx, up2, &gm, &ndom, &C4);
  if (rc != VMTL_OK) return rc;
  if ((rc = xs_check(T, ndom, C4 * 4, mode)) != VMTL_OK) return rc;
  if (!x_host || !y_host || !alpha || (Cs > 0 && !skip_host)) return VMTL_EINVAL;
  XsCatFwdArgs a{};
  for (int t = 0; t < T; ++t) {
    if (!x_host[t] || !y_host[t] || (Cs > 0 && !skip_host[t])) return VMTL_EINVAL;
    if (!aligned16(x_host[t]) || !aligned16(y_host[t]) || (Cs > 0 && !aligned16(skip_host[t]))) return VMTL_EALIGN;
    a.x[t] = reinterpret_cast<const float4*>(x_host[t]);
    a.y[t] = reinterpret_cast<float4*>(y_host[t]);
    a.skip[t] = Cs > 0 ? reinterpret_cast<const float4*>(skip_host[t]) : nullptr;
  }
  if (channel_wise && !aligned16(alpha)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const XsCatPlan p = xs_cat_plan(T, ndom, C4, channel_wise);
  dim3 grid(p.gridx, p.nchunk);
#define VMTL_XC_FWD(TT, DIAG, CWB, UP)                                                                                \
  xstitch_cat_fwd_kernel<TT, DIAG, CWB, UP><<<grid, p.threads, 0, st>>>(a, alpha, gm, ndom, C4, p.cgw, p.rows)
#define VMTL_XC_FWD_T(TT)                                                                                             \
  do {                                                                                                                \
    if (diag && cw && up) VMTL_XC_FWD(TT, true, true, true);                                                          \
    else if (diag && cw) VMTL_XC_FWD(TT, true, true, false);                                                          \
    else if (diag && up) VMTL_XC_FWD(TT, true, false, true);                                                          \
    else if (diag) VMTL_XC_FWD(TT, true, false, false);                                                               \
    else if (cw && up) VMTL_XC_FWD(TT, false, true, true);                                                            \
    else if (cw) VMTL_XC_FWD(TT, false, true, false);                                                                 \
    else if (up) VMTL_XC_FWD(TT, false, false, true);                                                                 \
    else VMTL_XC_FWD(TT, false, false, false);                                                                        \
  } while (0)
  const bool diag = mode == VMTL_XS_REFERENCE_DIAG, cw = channel_wise != 0, up = up2 != 0;
  switch (T) {
    case 1: VMTL_XC_FWD_T(1); break;
    case 2: VMTL_XC_FWD_T(2); break;
    case 3: VMTL_XC_FWD_T(3); break;
    case 4: VMTL_XC_FWD_T(4); break;
    default: return VMTL_EUNSUPPORTED;
  }
#undef VMTL_XC_FWD_T
#undef VMTL_XC_FWD
  return launch_status();
}

extern "C" size_t vmtl_xstitch_cat_bwd_workspace_bytes(int T, int B, int Ho, int Wo, int Cs, int Hi, int Wi, int Cx,
                                                       int up2, int channel_wise) {
  XsCatGeom gm;
  int64_t ndom;
  int C4;
  if (xs_cat_geom(B, Ho, Wo, Cs, Hi, Wi, Cx, up2, &gm, &ndom, &C4) != VMTL_OK || T < 1 || T > VMTL_MAX_TASKS) return 0;
  return xs_cat_plan(T, ndom, C4, channel_wise).partial_bytes + 256;
}

extern "C" int vmtl_xstitch_cat_bwd(const float* const* dy_host, const float* const* skip_host,
                                    const float* const* x_host, float* const* dskip_host, float* const* dx_host,
                                    const float* alpha, float* dalpha, int T, int B, int Ho, int Wo, int Cs, int Hi,
                                    int Wi, int Cx, int up2, int channel_wise, int mode, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  XsCatGeom gm;
  int64_t ndom;
  int C4;
  int rc = xs_cat_geom(B, Ho, Wo, Cs, Hi, Wi, Cx, up2, &gm, &ndom, &C4);
  if (rc != VMTL_OK) return rc;
  if ((rc = xs_check(T, ndom, C4 * 4, mode)) != VMTL_OK) return rc;
  if (!dy_host || !x_host || !alpha || !dalpha || !workspace || (Cs > 0 && !skip_host)) return VMTL_EINVAL;
  XsCatBwdArgs a{};
  for (int t = 0; t < T; ++t) {
    if (!dy_host[t] || !x_host[t] || (Cs > 0 && !skip_host[t])) return VMTL_EINVAL;
    if (!aligned16(dy_host[t]) || !aligned16(x_host[t])) return VMTL_EALIGN;
    a.dy[t] = reinterpret_cast<const float4*>(dy_host[t]);
    a.x[t] = reinterpret_cast<const float4*>(x_host[t]);
    a.skip[t] = Cs > 0 ? reinterpret_cast<const float4*>(skip_host[t]) : nullptr;
    a.dskip[t] = (Cs > 0 && dskip_host && dskip_host[t]) ? reinterpret_cast<float4*>(dskip_host[t]) : nullptr;
    a.dx[t] = (dx_host && dx_host[t]) ? reinterpret_cast<float4*>(dx_host[t]) : nullptr;
  }
  if ((channel_wise && !aligned16(alpha)) || !aligned16(workspace)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const XsCatPlan p = xs_cat_plan(T, ndom, C4, channel_wise);
  if (workspace_bytes < p.partial_bytes) return VMTL_EWORKSPACE;
  if (p.smem > 200 * 1024) return VMTL_EUNSUPPORTED;
  float* partial = static_cast<float*>(workspace);
  dim3 grid(p.gridx, p.nchunk);
#define VMTL_XC_BWD(TT, DIAG, CWB, UP)                                                                                \
  do {                                                                                                                \
    if (p.smem > 48 * 1024)                                                                                           \
      cudaFuncSetAttribute(xstitch_cat_bwd_kernel<TT, DIAG, CWB, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                           (int)p.smem);                                                                              \
    xstitch_cat_bwd_kernel<TT, DIAG, CWB, UP><<<grid, p.threads, p.smem, st>>>(a, alpha, partial, gm, ndom, C4, p.cgw, \
                                                                               p.rows);                               \
  } while (0)
#define VMTL_XC_BWD_T(TT)                                                                                             \
  do {                                                                                                                \
    if (diag && cw && up) VMTL_XC_BWD(TT, true, true, true);                                                          \
    else if (diag && cw) VMTL_XC_BWD(TT, true, true, false);                                                          \
    else if (diag && up) VMTL_XC_BWD(TT, true, false, true);                                                          \
    else if (diag) VMTL_XC_BWD(TT, true, false, false);                                                               \
    else if (cw && up) VMTL_XC_BWD(TT, false, true, true);                                                            \
    else if (cw) VMTL_XC_BWD(TT, false, true, false);                                                                 \
    else if (up) VMTL_XC_BWD(TT, false, false, true);                                                                 \
    else VMTL_XC_BWD(TT, false, false, false);                                                                        \
  } while (0)
  const bool diag = mode == VMTL_XS_REFERENCE_DIAG, cw = channel_wise != 0, up = up2 != 0;
  switch (T) {
    case 1: VMTL_XC_BWD_T(1); break;
    case 2: VMTL_XC_BWD_T(2); break;
    case 3: VMTL_XC_BWD_T(3); break;
    case 4: VMTL_XC_BWD_T(4); break;
    default: return VMTL_EUNSUPPORTED;
  }
#undef VMTL_XC_BWD_T
#undef VMTL_XC_BWD
  if ((rc = launch_status()) != VMTL_OK) return rc;
  const int cdim = cw ? C4 * 4 : 1;
  const int n = T * T * cdim;
  const int nparts = cw ? p.gridx : p.gridx * p.nchunk;
  if (diag)
    xstitch_dalpha_finalize<true><<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, dalpha, nparts, T, cdim);
  else
    xstitch_dalpha_finalize<false><<<(n + 31) / 32, kFinThreads, 0, st>>>(partial, dalpha, nparts, T, cdim);
  return launch_status();
}

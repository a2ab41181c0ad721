// Bilinear x2 up-sampling (align_corners = True) of NHWC maps, forward and backward, with a row stride on the
// up-sampled side so the result can be written straight into (and its gradient read straight out of) the channel
// slice of a concatenated tensor.
//
// Reference: AttentionModuleDecoder -- `self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)`
// (vision_mtl/models/mtan_model.py:125), applied to the previous attention output before
// `torch.cat((conv1_shared, prev), dim=1)` (:143-145).  ATen's NHWC kernels for it run at ~10 % of the HBM roofline
// (2.9 + 2.7 ms of the 102 ms MTAN step for 1.8 GB of traffic each way), the cat then re-copies the result and the
// cat's backward hands out strided slices that the next op has to repack.
//
// Arithmetic = ATen's upsample_bilinear2d (UpSampleBilinear2d.cu): r = (in - 1) / (out - 1) in fp32,
//   src = r * dst,  i1 = (int)src,  i1p = i1 < in - 1,  l1 = src - i1,  l0 = 1 - l1,
//   out = l0y (l0x x[i1y][i1x] + l1x x[i1y][i1x + i1px]) + l1y (l0x x[i1y + i1py][i1x] + l1x x[i1y + i1py][i1x + i1px]).
// Forward: one thread per (output pixel, float4 channel group): R 4 neighbours (L1/L2-resident: the input is 4x
// smaller), W one float4.  Backward is a GATHER over the output pixels that reference an input pixel (their indices
// lie in [2i - 2, 2i + 3] for any size; weights recomputed with the forward's own expression): deterministic, no atomics.
// Algorithmic bytes: 20 per input element each way (4 in + 16 out).
#include "vmtl_common.cuh"

namespace vmtl {

constexpr int kUpThreads = 256;
constexpr int kUpGroups = 4;  // channel groups a lane of the backward kernel carries per pixel

struct Src {
  int i1, i1p;
  float l0, l1;
};
__device__ __forceinline__ Src up_src(float r, int dst, int in) {
  Src s;
  const float src = r * (float)dst;
  s.i1 = (int)src;
  s.i1p = s.i1 < in - 1 ? 1 : 0;
  s.l1 = src - (float)s.i1;
  s.l0 = 1.f - s.l1;
  return s;
}

// Thread layout (both kernels): `cpl` lanes walk the float4 channel groups of a pixel (each lane takes every cpl-th
// group: 4 groups per lane for the 128-channel maps, so the per-pixel index / weight arithmetic -- the kernels are
// issue-bound, ncu: 63-72 % issue-slot use at 1 group per thread -- is paid once per 64 bytes), the other 256 / cpl
// thread rows take consecutive pixels; 8 lanes x 16 B = one full 128-byte line.  All index arithmetic is 32-bit.
__host__ __device__ inline unsigned up_lanes(int C4) {
  return C4 >= 32 ? (unsigned)C4 / 4 : (C4 >= 8 ? 8u : (unsigned)C4);
}
__global__ void __launch_bounds__(kUpThreads)
    up2_bilinear_fwd_kernel(const float4* __restrict__ x, float* __restrict__ y, int B, int Hi, int Wi, int C4,
                            int64_t ldy /* floats between consecutive output pixels */, float ry, float rx) {
  const unsigned Ho = 2u * Hi, Wo = 2u * Wi;
  const unsigned npix = (unsigned)B * Ho * Wo;
  const unsigned cpl = up_lanes(C4), rows = kUpThreads / cpl;
  const unsigned lane = threadIdx.x % cpl, r = threadIdx.x / cpl;
  if (r >= rows) return;
  for (unsigned pix = blockIdx.x * rows + r; pix < npix; pix += gridDim.x * rows) {
    const unsigned q = pix / Wo, ox = pix - q * Wo, b = q / Ho, oy = q - b * Ho;
    const Src sy = up_src(ry, (int)oy, Hi), sx = up_src(rx, (int)ox, Wi);
    const float4* row0 = x + (size_t)((b * Hi + sy.i1) * (unsigned)Wi) * C4;
    const float4* row1 = row0 + (size_t)sy.i1p * Wi * C4;
    const size_t c0 = (size_t)sx.i1 * C4, c1 = (size_t)(sx.i1 + sx.i1p) * C4;
    float4* out = reinterpret_cast<float4*>(y + (size_t)pix * ldy);
    for (unsigned c4 = lane; c4 < (unsigned)C4; c4 += cpl) {
      const float4 p00 = __ldg(row0 + c0 + c4), p01 = __ldg(row0 + c1 + c4);
      const float4 p10 = __ldg(row1 + c0 + c4), p11 = __ldg(row1 + c1 + c4);
      float4 o;  // ATen's association: l0y (l0x a + l1x b) + l1y (l0x c + l1x d)
      o.x = sy.l0 * (sx.l0 * p00.x + sx.l1 * p01.x) + sy.l1 * (sx.l0 * p10.x + sx.l1 * p11.x);
      o.y = sy.l0 * (sx.l0 * p00.y + sx.l1 * p01.y) + sy.l1 * (sx.l0 * p10.y + sx.l1 * p11.y);
      o.z = sy.l0 * (sx.l0 * p00.z + sx.l1 * p01.z) + sy.l1 * (sx.l0 * p10.z + sx.l1 * p11.z);
      o.w = sy.l0 * (sx.l0 * p00.w + sx.l1 * p01.w) + sy.l1 * (sx.l0 * p10.w + sx.l1 * p11.w);
      stg_stream(out + c4, o);
    }
  }
}

// weight with which output index `dst` references input index `i` (0 when it does not)
__device__ __forceinline__ float up_weight(float r, int dst, int in, int i) {
  const Src s = up_src(r, dst, in);
  return (s.i1 == i ? s.l0 : 0.f) + (s.i1 + s.i1p == i ? s.l1 : 0.f);
}

__global__ void __launch_bounds__(kUpThreads)
    up2_bilinear_bwd_kernel(const float* __restrict__ dy, int64_t lddy, float4* __restrict__ dx, int B, int Hi, int Wi,
                            int C4, float ry, float rx) {
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  const unsigned npix = (unsigned)B * Hi * Wi;
  const unsigned cpl = up_lanes(C4), rows = kUpThreads / cpl;
  const unsigned lane = threadIdx.x % cpl, r = threadIdx.x / cpl;
  if (r >= rows) return;
  for (unsigned pix = blockIdx.x * rows + r; pix < npix; pix += gridDim.x * rows) {
    const unsigned q = pix / (unsigned)Wi, b = q / (unsigned)Hi;
    const int ix = (int)(pix - q * Wi), iy = (int)(q - b * Hi);
    float wx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int ox = 2 * ix - 2 + k;
      wx[k] = (ox >= 0 && ox < Wo) ? up_weight(rx, ox, Wi, ix) : 0.f;
    }
    // up to kUpGroups channel groups per lane share the row / column weights of the pixel
    for (unsigned cb = lane; cb < (unsigned)C4; cb += kUpGroups * cpl) {
      float4 acc[kUpGroups];
#pragma unroll
      for (int j = 0; j < kUpGroups; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1  // rows one at a time (fully unrolled, the 6 x 6 gather needs ~200 registers: one block per SM)
      for (int ky = 0; ky < 6; ++ky) {
        const int oy = 2 * iy - 2 + ky;
        if (oy < 0 || oy >= Ho) continue;
        const float wyv = up_weight(ry, oy, Hi, iy);
        if (wyv == 0.f) continue;
        const float* row = dy + ((size_t)(b * Ho + oy) * Wo) * lddy;
#pragma unroll
        for (int kx = 0; kx < 6; ++kx) {
          if (wx[kx] == 0.f) continue;
          const int ox = 2 * ix - 2 + kx;
          const float4* src = reinterpret_cast<const float4*>(row + (size_t)ox * lddy);
          const float w = wyv * wx[kx];
#pragma unroll
          for (int j = 0; j < kUpGroups; ++j) {
            const unsigned c4 = cb + j * cpl;
            if (c4 < (unsigned)C4) {
              const float4 g = __ldg(src + c4);
              acc[j].x = fmaf(w, g.x, acc[j].x);
              acc[j].y = fmaf(w, g.y, acc[j].y);
              acc[j].z = fmaf(w, g.z, acc[j].z);
              acc[j].w = fmaf(w, g.w, acc[j].w);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kUpGroups; ++j) {
        const unsigned c4 = cb + j * cpl;
        if (c4 < (unsigned)C4) stg_stream(dx + (size_t)pix * C4 + c4, acc[j]);
      }
    }
  }
}

static int up_check(int B, int Hi, int Wi, int C, int64_t ld) {
  if (B < 1 || Hi < 1 || Wi < 1 || C < 4) return VMTL_EINVAL;
  if (C % 4 != 0 || ld < C || ld % 4 != 0) return VMTL_EUNSUPPORTED;
  if ((int64_t)B * Hi * Wi * 4 >= (1ll << 31)) return VMTL_EUNSUPPORTED;  // 32-bit pixel arithmetic
  return VMTL_OK;
}
static int up_grid(int64_t npix, int C4, int (*occ)(void)) {
  const int cpl = (int)up_lanes(C4), rows = kUpThreads / cpl;
  int64_t want = (npix + rows - 1) / rows;
  const int64_t cap = (int64_t)sm_count() * occ();
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}
static int occ_fwd() { return blocks_per_sm(up2_bilinear_fwd_kernel, kUpThreads, 0, 8); }
static int occ_bwd() { return blocks_per_sm(up2_bilinear_bwd_kernel, kUpThreads, 0, 8); }
static float up_ratio(int in) { return in > 1 ? (float)(in - 1) / (float)(2 * in - 1) : 0.f; }

}  // namespace vmtl

using namespace vmtl;

extern "C" int vmtl_up2_bilinear_fwd(const float* x, float* y, int B, int Hi, int Wi, int C, int64_t ldy, void* stream) {
  if (!x || !y) return VMTL_EINVAL;
  int rc = up_check(B, Hi, Wi, C, ldy);
  if (rc != VMTL_OK) return rc;
  if (!aligned16(x) || !aligned16(y)) return VMTL_EALIGN;
  up2_bilinear_fwd_kernel<<<up_grid((int64_t)B * Hi * Wi * 4, C / 4, occ_fwd), kUpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), y, B, Hi, Wi, C / 4, ldy, up_ratio(Hi), up_ratio(Wi));
  return launch_status();
}

extern "C" int vmtl_up2_bilinear_bwd(const float* dy, int64_t lddy, float* dx, int B, int Hi, int Wi, int C,
                                     void* stream) {
  if (!dy || !dx) return VMTL_EINVAL;
  int rc = up_check(B, Hi, Wi, C, lddy);
  if (rc != VMTL_OK) return rc;
  if (!aligned16(dy) || !aligned16(dx)) return VMTL_EALIGN;
  up2_bilinear_bwd_kernel<<<up_grid((int64_t)B * Hi * Wi, C / 4, occ_bwd), kUpThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, lddy, reinterpret_cast<float4*>(dx), B, Hi, Wi, C / 4, up_ratio(Hi), up_ratio(Wi));
  return launch_status();
}

// Multi-tensor Adam: every parameter tensor of a group in ONE launch -- SURVEY 8(f) row 4.
//
// Reference: torch.optim.Adam(params, lr) built at vision_mtl/training_lit.py:51 (and lit_module.py:194),
// default betas / eps, no weight decay, no amsgrad; stepped once per batch (training_lit.py:87).  torch runs it as
// a foreach / multi_tensor_apply sequence (tens of launches for the ~330 MTAN parameter tensors, each moving
// a few KB); here the step is one HBM-bound pass: read p, g, m, v, write p, m, v = 28 bytes per parameter.
//
// The tensor table (parameter / gradient pointers, element counts, offsets into the optimizer's flat moment buffers)
// travels BY VALUE in the kernel parameter space (up to kAdamMaxTensors tensors, ~11 KB; CUDA >= 12.1 allows 32 KB),
// so a launch captured into a CUDA graph carries its table with it and no host buffer has to outlive the capture.
// The step counter and (optionally) the learning rate live in device memory: a captured launch keeps counting, and
// a scheduler that fills the lr tensor acts on graph replays.  The LAST block to finish bumps the counter (ticket in
// optimizer-owned, zero-initialised device memory; the kernel leaves it zero).
//
// Arithmetic = torch's single-tensor Adam (torch/optim/adam.py, _single_tensor_adam, capturable = False):
//   g += wd * p ;  m += (g - m)(1 - b1) ;  v = b2 v + (1 - b2) g g ;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <math.h>

#include "vmtl_common.cuh"

namespace vmtl {

constexpr int kAdamMaxTensors = 384;
constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 4 * 8;  // elements per (block, iteration): 8 float4 per thread

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  int64_t off[kAdamMaxTensors];       // element offset of the tensor's moments in the flat buffers (multiple of 4)
  int32_t numel[kAdamMaxTensors];
  int32_t chunk0[kAdamMaxTensors + 1];  // first chunk of tensor i; chunk0[n] = number of chunks
  int32_t n;
};

struct AdamCoef {
  float omb1, b2, omb2, eps, wd, step_size, inv_bc2_sqrt;  // omb = 1 - beta, rounded from the double difference
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamCoef& c) {
  g = fmaf(c.wd, p, g);
  m = fmaf(g - m, c.omb1, m);
  v = fmaf(v, c.b2, c.omb2 * g * g);
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  p -= c.step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads)
    adam_multi_kernel(const __grid_constant__ AdamTable tab, float* __restrict__ m_flat, float* __restrict__ v_flat,
                      const float* __restrict__ lr_dev, double lr_host, float* __restrict__ step,
                      unsigned int* __restrict__ ticket, double beta1, double beta2, float eps, float wd, int bump_step) {
  const float t = step[0] + 1.f;
  const double lr = lr_dev ? (double)lr_dev[0] : lr_host;
  AdamCoef c;
  c.omb1 = (float)(1.0 - beta1);  // torch: lerp weight / addcmul value are Python doubles cast to fp32
  c.b2 = (float)beta2;
  c.omb2 = (float)(1.0 - beta2);
  c.eps = eps;
  c.wd = wd;
  const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
  c.step_size = (float)(lr / bc1);
  c.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int nchunks = tab.chunk0[tab.n];
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    // tensor of this chunk: largest i with chunk0[i] <= ch (uniform over the block; the table sits in constant space)
    int lo = 0, hi = tab.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab.chunk0[mid] <= ch) lo = mid; else hi = mid - 1;
    }
    const int64_t e0 = (int64_t)(ch - tab.chunk0[lo]) * kAdamChunk;
    const int cnt = (int)min((int64_t)kAdamChunk, (int64_t)tab.numel[lo] - e0);
    float* p = tab.p[lo] + e0;
    const float* g = tab.g[lo] + e0;
    float* m = m_flat + tab.off[lo] + e0;
    float* v = v_flat + tab.off[lo] + e0;
    const bool vec = (((uintptr_t)p | (uintptr_t)g) & 15u) == 0;  // m, v are 16-byte aligned by construction
    if (vec) {
      const int n4 = cnt >> 2;
      for (int i = threadIdx.x; i < n4; i += 2 * kAdamThreads) {  // two float4 quadruples in flight
        const int j = i + kAdamThreads;
        const bool two = j < n4;
        float4 p0 = reinterpret_cast<float4*>(p)[i], g0 = ldg_stream(reinterpret_cast<const float4*>(g) + i);
        float4 m0 = reinterpret_cast<float4*>(m)[i], v0 = reinterpret_cast<float4*>(v)[i];
        float4 p1, g1, m1, v1;
        if (two) {
          p1 = reinterpret_cast<float4*>(p)[j];
          g1 = ldg_stream(reinterpret_cast<const float4*>(g) + j);
          m1 = reinterpret_cast<float4*>(m)[j];
          v1 = reinterpret_cast<float4*>(v)[j];
        }
        adam_elem(p0.x, g0.x, m0.x, v0.x, c); adam_elem(p0.y, g0.y, m0.y, v0.y, c);
        adam_elem(p0.z, g0.z, m0.z, v0.z, c); adam_elem(p0.w, g0.w, m0.w, v0.w, c);
        reinterpret_cast<float4*>(p)[i] = p0;
        reinterpret_cast<float4*>(m)[i] = m0;
        reinterpret_cast<float4*>(v)[i] = v0;
        if (two) {
          adam_elem(p1.x, g1.x, m1.x, v1.x, c); adam_elem(p1.y, g1.y, m1.y, v1.y, c);
          adam_elem(p1.z, g1.z, m1.z, v1.z, c); adam_elem(p1.w, g1.w, m1.w, v1.w, c);
          reinterpret_cast<float4*>(p)[j] = p1;
          reinterpret_cast<float4*>(m)[j] = m1;
          reinterpret_cast<float4*>(v)[j] = v1;
        }
      }
    }
    for (int i = (vec ? (cnt & ~3) : 0) + threadIdx.x; i < cnt; i += kAdamThreads) {
      float pv = p[i], mv = m[i], vv = v[i];
      adam_elem(pv, g[i], mv, vv, c);
      p[i] = pv;
      m[i] = mv;
      v[i] = vv;
    }
  }
  if (!bump_step) return;
  // every block has read `step` before it gets here, so the last one to arrive may advance it
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      step[0] = t;
    }
  }
}

}  // namespace vmtl

using namespace vmtl;

extern "C" int vmtl_adam_max_tensors_per_launch(void) { return kAdamMaxTensors; }

// params / grads / numel / moment_offset: HOST arrays of n_tensors entries (device pointers, element counts, element
// offsets into exp_avg / exp_avg_sq, each a multiple of 4).  exp_avg, exp_avg_sq, lr_dev (may be NULL: lr is used),
// step_dev (float [1], the number of steps taken so far) and ticket (uint32 [1], zero) are DEVICE pointers.
extern "C" int vmtl_adam_step(const void* const* params, const void* const* grads, const int64_t* numel,
                              const int64_t* moment_offset, int n_tensors, float* exp_avg, float* exp_avg_sq,
                              const float* lr_dev, double lr, float* step_dev, unsigned int* ticket, double beta1,
                              double beta2, double eps, double weight_decay, void* stream) {
  if (n_tensors < 0 || (n_tensors > 0 && (!params || !grads || !numel || !moment_offset)) || !exp_avg || !exp_avg_sq ||
      !step_dev || !ticket)
    return VMTL_EINVAL;
  if (!aligned16(exp_avg) || !aligned16(exp_avg_sq)) return VMTL_EALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int done = 0;
  do {  // at least one launch, so that an empty group still advances the step counter
    AdamTable tab;
    int n = 0, chunks = 0;
    while (done + n < n_tensors && n < kAdamMaxTensors) {
      const int i = done + n;
      if (!params[i] || !grads[i] || numel[i] < 0 || numel[i] > 0x7fffffffll || (moment_offset[i] & 3)) return VMTL_EINVAL;
      if ((((uintptr_t)params[i]) | ((uintptr_t)grads[i])) & 3u) return VMTL_EALIGN;
      tab.p[n] = static_cast<float*>(const_cast<void*>(params[i]));
      tab.g[n] = static_cast<const float*>(grads[i]);
      tab.off[n] = moment_offset[i];
      tab.numel[n] = (int32_t)numel[i];
      tab.chunk0[n] = chunks;
      chunks += (int)((numel[i] + kAdamChunk - 1) / kAdamChunk);
      ++n;
    }
    tab.chunk0[n] = chunks;
    tab.n = n;
    done += n;
    const int last = done >= n_tensors;
    int grid = chunks < 1 ? 1 : chunks;
    const int cap = sm_count() * blocks_per_sm(adam_multi_kernel, kAdamThreads, 0, 8);
    if (grid > cap) grid = cap;
    adam_multi_kernel<<<grid, kAdamThreads, 0, st>>>(tab, exp_avg, exp_avg_sq, lr_dev, lr, step_dev, ticket, beta1, beta2,
                                                     (float)eps, (float)weight_decay, last);
    const int rc = launch_status();
    if (rc != VMTL_OK) return rc;
  } while (done < n_tensors);
  return VMTL_OK;
}

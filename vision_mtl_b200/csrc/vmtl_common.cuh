// Shared device/host helpers for libvmtl_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vmtl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvmtl_b200 is written for sm_100a (B200) only"
#endif

namespace vmtl {

constexpr int kWarp = 32;

inline int sm_count() {
  // Immutable per-device attribute, cached per device id.
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Resident CTAs per SM of `kernel`, clipped to `want`: a persistent / grid-stride launch sized beyond what is
// resident runs a second, partial wave (e.g. 8 x 256 threads requested at 48 registers -> 5 fit -> 1.6 waves).
template <typename KernelT>
inline int blocks_per_sm(KernelT kernel, int threads, size_t smem, int want) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
  return n < want ? n : want;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VMTL_OK : VMTL_ECUDA;
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every op here is a short chain (statistics -> finalize -> apply) of kernels that each wait for the previous one to
// drain completely before the next is even scheduled: ~2-4 us of idle GPU per link, a large share of a 20-50 us call.
// A kernel launched through launch_pdl() may be scheduled as soon as every block of its predecessor has executed
// pdl_trigger() (or exited); it must execute pdl_wait() -- which returns once the predecessor has COMPLETED and its
// writes are visible -- before it reads anything a predecessor produced.  Inside a stream capture the attribute
// becomes a programmatic edge of the graph.  Both instructions are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through launch_status()
}

// ---- streaming loads/stores: the feature maps are read once, keep them out of L1 ----
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ longlong2 ldg_stream(const longlong2* p) {
  longlong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];"
               : "=l"(r.x), "=l"(r.y)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- second stage of every two-stage reduction ---------------------------------------------
// Block of kFinThreads = 32 columns x 32 row groups.  Thread (lx, ly) sums rows ly, ly+32, ... of
// column `col` of partial[nparts][rowlen] in fp64; the 32 sub-sums are then combined in a FIXED
// order, so the result is deterministic and independent of the launch geometry of stage one.
// The result is valid in the threads with ly == 0 (threadIdx.x < 32).
constexpr int kFinThreads = 1024;
template <typename T>
__device__ __forceinline__ double block_colsum(const T* __restrict__ partial, int nparts, int64_t rowlen,
                                               int64_t col, bool valid) {
  __shared__ double s_sub[32][33];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  double s = 0.0;
  if (valid) {
    int k = ly;
    // the kernel is one dependent-latency chain per thread: keep 8 independent loads in flight
    for (; k + 224 < nparts; k += 256) {
      T v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = partial[(int64_t)(k + 32 * q) * rowlen + col];
      s += (((double)v[0] + (double)v[1]) + ((double)v[2] + (double)v[3])) +
           (((double)v[4] + (double)v[5]) + ((double)v[6] + (double)v[7]));
    }
    for (; k < nparts; k += 32) s += (double)partial[(int64_t)k * rowlen + col];
  }
  s_sub[ly][lx] = s;
  __syncthreads();
  double t = 0.0;
  if (ly == 0) {
#pragma unroll
    for (int q = 0; q < 32; ++q) t += s_sub[q][lx];
  }
  __syncthreads();
  return t;
}

// Two columns of the same partial matrix at once (their loads overlap instead of running back to back).
template <typename T>
__device__ __forceinline__ void block_colsum2(const T* __restrict__ partial, int nparts, int64_t rowlen,
                                              int64_t col_a, int64_t col_b, bool valid, double* out_a,
                                              double* out_b) {
  __shared__ double s_sub2[2][32][33];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  double sa = 0.0, sb = 0.0;
  if (valid) {
    int k = ly;
    for (; k + 96 < nparts; k += 128) {
      T va[4], vb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        va[q] = partial[(int64_t)(k + 32 * q) * rowlen + col_a];
        vb[q] = partial[(int64_t)(k + 32 * q) * rowlen + col_b];
      }
      sa += ((double)va[0] + (double)va[1]) + ((double)va[2] + (double)va[3]);
      sb += ((double)vb[0] + (double)vb[1]) + ((double)vb[2] + (double)vb[3]);
    }
    for (; k < nparts; k += 32) {
      sa += (double)partial[(int64_t)k * rowlen + col_a];
      sb += (double)partial[(int64_t)k * rowlen + col_b];
    }
  }
  s_sub2[0][ly][lx] = sa;
  s_sub2[1][ly][lx] = sb;
  __syncthreads();
  double ta = 0.0, tb = 0.0;
  if (ly == 0) {
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      ta += s_sub2[0][q][lx];
      tb += s_sub2[1][q][lx];
    }
  }
  __syncthreads();
  *out_a = ta;
  *out_b = tb;
}

// Tall variant for LONG partial lists (the streaming BatchNorm kernels leave 600-900 partial rows): 8 columns x 128
// row groups per block, so a thread's dependent chain of row loads is 4x shorter than with 32 x 32 (the finalize
// kernels are pure latency: ncu showed 14 us for 888 rows, most of it the 28-step chain).  Fixed summation order:
// rows ly, ly+128, ... per thread, then groups 32q..32q+31 by four threads per column, then those four.
// The result is valid in the threads with threadIdx.x < 8.
constexpr int kTallCols = 8;
template <typename T>
__device__ __forceinline__ void block_colsum2_tall(const T* __restrict__ partial, int nparts, int64_t rowlen,
                                                   int64_t col_a, int64_t col_b, bool valid, double* out_a,
                                                   double* out_b) {
  __shared__ double s_t[2][128][kTallCols + 1];
  __shared__ double s_q[2][4][kTallCols + 1];
  const int lx = threadIdx.x & (kTallCols - 1), ly = threadIdx.x >> 3;
  double sa = 0.0, sb = 0.0;
  if (valid) {
    int k = ly;
    for (; k + 384 < nparts; k += 512) {
      T va[4], vb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        va[q] = partial[(int64_t)(k + 128 * q) * rowlen + col_a];
        vb[q] = partial[(int64_t)(k + 128 * q) * rowlen + col_b];
      }
      sa += ((double)va[0] + (double)va[1]) + ((double)va[2] + (double)va[3]);
      sb += ((double)vb[0] + (double)vb[1]) + ((double)vb[2] + (double)vb[3]);
    }
    for (; k < nparts; k += 128) {
      sa += (double)partial[(int64_t)k * rowlen + col_a];
      sb += (double)partial[(int64_t)k * rowlen + col_b];
    }
  }
  s_t[0][ly][lx] = sa;
  s_t[1][ly][lx] = sb;
  __syncthreads();
  if (ly < 4) {
    double ta = 0.0, tb = 0.0;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      ta += s_t[0][ly * 32 + q][lx];
      tb += s_t[1][ly * 32 + q][lx];
    }
    s_q[0][ly][lx] = ta;
    s_q[1][ly][lx] = tb;
  }
  __syncthreads();
  double ta = 0.0, tb = 0.0;
  if (ly == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ta += s_q[0][q][lx];
      tb += s_q[1][q][lx];
    }
  }
  __syncthreads();
  *out_a = ta;
  *out_b = tb;
}

// MUFU forms (max rel. error 2^-22 on the normal range): the accurate expf/logf cost 10-20 instructions
// each and the per-pixel loss kernels run one per class.
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// MUFU sigmoid: ex2.approx + rcp.approx, ~3e-7 relative error on the value (the gate BACKWARD kernels: both passes
// use it, so du of the statistics pass and dz of the dh pass see the same activation)
__device__ __forceinline__ float sigmoidf_fast(float u) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + fast_ex2(-u * kLog2e)));
  return r;
}

__device__ __forceinline__ float sigmoidf_acc(float u) {
  // 1/(1+exp(-u)) with the accurate expf: the parity bar is 1e-4 relative on gradients.
  return 1.0f / (1.0f + expf(-u));
}

}  // namespace vmtl

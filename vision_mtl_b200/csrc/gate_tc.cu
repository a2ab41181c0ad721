// MTAN attention gate -- tensor-core contraction kernels (tcgen05.mma kind::tf32, TMEM, TMA).
//
// z[M,N] = h[M,128] @ W[N,128]^T + bias with fp32-grade accuracy through a tf32 hi/lo split
// (3xTF32: h = h_hi + h_lo, W = W_hi + W_lo, the dropped h_lo*W_lo term is ~2^-22), fp32
// accumulation in TMEM.  The kernels live in the headers included below:
//   gate_tc_tma.cuh      forward: TRAIN (z + batch-statistics partials) and EVAL (folded BN + sigmoid +
//                        product fused into the epilogue, single pass)
//   gate_tc_bwd_tma.cuh  backward: fused statistics + dW pass, dh pass
// This file holds what they share and the host-side dispatch.  Shapes outside K = 128,
// N in {32, 64k} return VMTL_EUNSUPPORTED and the caller (gate.cu) takes the CUDA-core path.
#include <math.h>

#include "gate_internal.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace vmtl {

using namespace tc;

constexpr int kTileM = 128;

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

}  // namespace vmtl
#include "gate_tc_tma.cuh"      // forward
#include "gate_tc_bwd_tma.cuh"  // backward
namespace vmtl {

bool gate_tc_supported(int K, int N) { return K == 128 && (N == 32 || (N % 64 == 0 && N >= 64 && N <= 1024)); }

int gate_tc_fwd_gemm(const float* h, const float* h_coef, const float* W, const float* bias, int64_t M, int K, int N,
                     int split3, float* z_out, float* partial, int partial_rows, int* nparts,
                     cudaStream_t st) {
  if (!gate_tc_supported(K, N)) return VMTL_EUNSUPPORTED;
  const int grid = fwd_tma_grid(M, N <= 64 ? 1 : N / 64);
  if (grid > partial_rows) return VMTL_EWORKSPACE;
  *nparts = grid;
  return dispatch_fwd_tma<false>(h, h_coef, W, bias, M, N, split3, z_out, partial, nullptr, nullptr, nullptr, grid, st);
}

int gate_tc_fwd_eval(const float* h, const float* h_coef, const float* s, const float* W, const float* bias,
                     const float* coefA, const float* coefB, int64_t M, int K, int N, int split3,
                     float* y, cudaStream_t st) {
  if (!gate_tc_supported(K, N)) return VMTL_EUNSUPPORTED;
  const int grid = fwd_tma_grid(M, N <= 64 ? 1 : N / 64);
  return dispatch_fwd_tma<true>(h, h_coef, W, bias, M, N, split3, y, nullptr, s, coefA, coefB, grid, st);
}

static int tc_grid(int64_t M, int rows) {
  const int64_t n = (M + rows - 1) / rows;
  const int sms = sm_count();
  return (int)(n < sms ? n : sms);
}

int gate_tc_bwd_pass1_grid(int64_t M, int N) {
  // (64-row unit, column chunk) items over the SMs; the grid is a multiple of the chunk count
  const int nch = N <= 64 ? 1 : N / 64;
  const int64_t items = ((M + 63) / 64) * nch;
  int grid = (int)(items < sm_count() ? items : sm_count());
  grid = grid / nch * nch;
  return grid < nch ? nch : grid;
}

int gate_tc_bwd_pass1(const float* dy, const float* h, const float* h_coef, const float* s, const float* z, const float* gamma,
                      const float* beta, const float* mean, const float* invstd, int64_t M, int K, int N,
                      int split3, float* ds, const GateWs& ws, int* nparts, cudaStream_t st) {
  if (!gate_tc_supported(K, N) || !ws.hs_partial) return VMTL_EUNSUPPORTED;
  const int nch = N <= 64 ? 1 : N / 64;
  const int grid = gate_tc_bwd_pass1_grid(M, N);
  // per-CTA partials: [2][Nc][K] in ws.gemm_partial (2 sm_count slots of N*K), [3][Nc] in ws.partial, [K] in ws.hs_partial
  const int Nc = N <= 64 ? N : 64;
  if ((int64_t)grid * 2 * Nc > (int64_t)ws.gemm_slots * N || (int64_t)grid * 3 * Nc > (int64_t)ws.partial_rows * 2 * N ||
      grid > ws.hs_rows)
    return VMTL_EWORKSPACE;
  *nparts = grid;
  return split3 ? launch_sdw_tma<true>(dy, h, h_coef, s, z, gamma, beta, mean, invstd, M, N, ds, ws.gemm_partial,
                                       ws.hs_partial, ws.partial, grid, nch, st)
                : launch_sdw_tma<false>(dy, h, h_coef, s, z, gamma, beta, mean, invstd, M, N, ds, ws.gemm_partial,
                                        ws.hs_partial, ws.partial, grid, nch, st);
}

int gate_tc_bwd_dh(const float* dy, const float* s, const float* z, const float* W, const GateWs& ws, int64_t M,
                   int K, int N, int split3, float* dh, cudaStream_t st) {
  if (!gate_tc_supported(K, N)) return VMTL_EUNSUPPORTED;
  const int grid = tc_grid(M, kTileM);
  return split3 ? launch_dh_passes<true>(dy, s, z, W, ws, M, N, dh, grid, st)
                : launch_dh_passes<false>(dy, s, z, W, ws, M, N, dh, grid, st);
}

}  // namespace vmtl
